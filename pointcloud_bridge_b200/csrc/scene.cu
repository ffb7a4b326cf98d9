// Whole-scene tiling and vote scatter-back: the steps either side of the
// block-sharded inference path, as GPU kernels instead of per-window numpy loops.
//   tiling : ScannetDatasetWholeScene.__getitem__  Highway_bridge/utils/BridgeDataLoader.py:214-277
//            (Partsize-identical/data_prep/BridgeDataLoader.py:168-231): sliding block_size x block_size
//            windows with `stride`, points inside [s - padding, e + padding] in x and y, padded to a multiple
//            of block_points, centred in x/y, extended by xyz / scene extent -> [nb, block_points, 9]
//   voting : add_vote + argmax                     Partsize-identical/test_sem_seg.py:58-65, 162
// Window bounds are computed by the host in float64 exactly as the reference does and compared in
// double, so window membership is bit-identical; a point belongs to at most a few windows, found by
// scanning the candidate range around floor((x - x0) / stride).  The reference pads a window by RANDOM
// re-draws of its own points (np.random.choice) and shuffles members + padding (np.random.shuffle), so every
// block is a uniform random subsample of its window.  Here the same structure comes from a counter-based hash
// instead of a host RNG, deterministic for a given (seed, vote) whatever order the atomics filled the window in:
//   * members of a window are ordered by (hash(seed, point index, window), point index) -- the fill kernel
//     writes 64-bit sort keys, the host side sorts them (one radix sort over all windows);
//   * padding entries re-use the first members of that pseudo-random order (a draw without replacement while
//     the padding is not longer than the window, cyclic beyond, as np.random.choice does with replace=True);
//   * members and padding are interleaved by an affine permutation of the padded window, so that the
//     duplicates are spread over all blocks of the window, as the reference's shuffle does.
// All kernels are HBM/atomic-bound streaming passes: one thread per point (count / fill / argmax) or
// per output entry (blocks / vote).
#include "pcb_common.cuh"

namespace pcb {

constexpr int kScThreads = 256;

struct WindowGrid {
    const double *lo_x, *hi_x, *lo_y, *hi_y;      // [grid_x] / [grid_y], padding included
    int grid_x, grid_y;
    double x0, y0, stride;                        // first window start, window pitch
    int reach;                                    // candidate windows checked on either side of floor((x - x0) / stride)
};

template <typename F>
__device__ __forceinline__ void for_each_window(const WindowGrid &g, double x, double y, F f)
{
    const int cx = (int)floor((x - g.x0) / g.stride), cy = (int)floor((y - g.y0) / g.stride);
    const int x_lo = max(0, cx - g.reach), x_hi = min(g.grid_x - 1, cx + 1);
    const int y_lo = max(0, cy - g.reach), y_hi = min(g.grid_y - 1, cy + 1);
    for (int iy = y_lo; iy <= y_hi; ++iy) {
        if (!(y >= g.lo_y[iy] && y <= g.hi_y[iy])) continue;
        for (int ix = x_lo; ix <= x_hi; ++ix)
            if (x >= g.lo_x[ix] && x <= g.hi_x[ix]) f(iy * g.grid_x + ix);
    }
}

__global__ void __launch_bounds__(kScThreads)
scene_window_count_kernel(const float *__restrict__ pts, int64_t P, int pstride, WindowGrid g, int *__restrict__ counts)
{
    const int64_t i = (int64_t)blockIdx.x * kScThreads + threadIdx.x;
    if (i >= P) return;
    const double x = (double)pts[i * pstride], y = (double)pts[i * pstride + 1];
    for_each_window(g, x, y, [&](int w) { atomicAdd(counts + w, 1); });
}

// lowbias32 mix of (seed, point, window): the per-window pseudo-random order of the members
__host__ __device__ __forceinline__ uint32_t scene_hash(uint32_t seed, uint32_t i, uint32_t w)
{
    uint32_t h = seed ^ (i * 0x9E3779B1u) ^ (w * 0x85EBCA77u);
    h ^= h >> 16;
    h *= 0x7FEB352Du;
    h ^= h >> 15;
    h *= 0x846CA68Bu;
    h ^= h >> 16;
    return h;
}

// members as sort keys: window (16 bits) | hash (16 bits) | point index (31 bits).  Sorting the keys puts every
// window's members into a pseudo-random order that does not depend on the order of the atomics.
__global__ void __launch_bounds__(kScThreads)
scene_window_fill_kernel(const float *__restrict__ pts, int64_t P, int pstride, WindowGrid g,
                         const int64_t *__restrict__ offsets, int *__restrict__ cursor, uint32_t seed,
                         int64_t *__restrict__ keys)
{
    const int64_t i = (int64_t)blockIdx.x * kScThreads + threadIdx.x;
    if (i >= P) return;
    const double x = (double)pts[i * pstride], y = (double)pts[i * pstride + 1];
    for_each_window(g, x, y, [&](int w) {
        const int64_t key = ((int64_t)w << 47) | ((int64_t)(scene_hash(seed, (uint32_t)i, (uint32_t)w) >> 16) << 31) | i;
        keys[offsets[w] + atomicAdd(cursor + w, 1)] = key;
    });
}

// one thread per (block, entry): gather the point, centre x/y on the window, append xyz / extent
__global__ void __launch_bounds__(kScThreads)
scene_blocks_kernel(const float *__restrict__ pts, int pstride, const int *__restrict__ members,
                    const int64_t *__restrict__ blk_off, const int *__restrict__ blk_cnt,
                    const int64_t *__restrict__ blk_first, const double *__restrict__ blk_center, int block_points,
                    int64_t total, double ext_x, double ext_y, double ext_z, uint32_t seed, float *__restrict__ data,
                    int64_t *__restrict__ point_idx)
{
    const int64_t t = (int64_t)blockIdx.x * kScThreads + threadIdx.x;
    if (t >= total) return;
    const int64_t blk = t / block_points;
    const int j = (int)(t - blk * block_points);
    const int n = blk_cnt[blk];
    // position q of the padded window (tot entries, a multiple of block_points = 2^k * odd) -> source slot s by an
    // affine permutation (the multiplier is an odd prime larger than any block count: coprime to tot); slots < n are
    // the members in their pseudo-random order, slots >= n the padding: the first members again
    const int64_t tot = (((int64_t)n + block_points - 1) / block_points) * block_points;
    const int64_t q = blk_first[blk] + j;
    const int64_t s = (int64_t)(((unsigned long long)q * 1000003ull + (unsigned long long)(seed % 999983u)) %
                                (unsigned long long)tot);
    const int64_t e = s < n ? s : (s - n) % n;
    const int src = members[blk_off[blk] + e];
    const float *p = pts + (int64_t)src * pstride;
    const double x = (double)p[0], y = (double)p[1], z = (double)p[2];
    float *o = data + t * 9;
    o[0] = (float)(x - blk_center[2 * blk]);
    o[1] = (float)(y - blk_center[2 * blk + 1]);
    o[2] = p[2];
    o[3] = p[3], o[4] = p[4], o[5] = p[5];
    o[6] = (float)(x / ext_x), o[7] = (float)(y / ext_y), o[8] = (float)(z / ext_z);
    point_idx[t] = (int64_t)src;
}

__global__ void __launch_bounds__(kScThreads)
scene_vote_kernel(const int64_t *__restrict__ point_idx, const unsigned char *__restrict__ pred, int64_t total,
                  int64_t P, int num_classes, int *__restrict__ pool)
{
    const int64_t t = (int64_t)blockIdx.x * kScThreads + threadIdx.x;
    if (t >= total) return;
    const int64_t i = point_idx[t];
    const int c = pred[t];
    if (i >= 0 && i < P && c < num_classes) atomicAdd(pool + i * num_classes + c, 1);
}

// np.argmax(vote_label_pool, 1): first maximum
__global__ void __launch_bounds__(kScThreads)
scene_vote_argmax_kernel(const int *__restrict__ pool, int64_t P, int num_classes, unsigned char *__restrict__ labels)
{
    const int64_t i = (int64_t)blockIdx.x * kScThreads + threadIdx.x;
    if (i >= P) return;
    int best = pool[i * num_classes], bi = 0;
    for (int c = 1; c < num_classes; ++c) {
        const int v = pool[i * num_classes + c];
        if (v > best) best = v, bi = c;
    }
    labels[i] = (unsigned char)bi;
}

static inline WindowGrid make_grid(const double *lo_x, const double *hi_x, const double *lo_y, const double *hi_y,
                                   int grid_x, int grid_y, double x0, double y0, double stride, int reach)
{
    WindowGrid g;
    g.lo_x = lo_x, g.hi_x = hi_x, g.lo_y = lo_y, g.hi_y = hi_y, g.grid_x = grid_x, g.grid_y = grid_y;
    g.x0 = x0, g.y0 = y0, g.stride = stride, g.reach = reach;
    return g;
}

}  // namespace pcb

using namespace pcb;

PCB_API int pcb_scene_window_count_f32(const float *points, int64_t P, int point_stride, int grid_x, int grid_y,
                                       const double *lo_x, const double *hi_x, const double *lo_y, const double *hi_y,
                                       double x0, double y0, double stride, int reach, int *counts, pcb_stream_t stream)
{
    PCB_REQUIRE(points && lo_x && hi_x && lo_y && hi_y && counts, PCB_EINVAL);
    PCB_REQUIRE(P > 0 && point_stride >= 3 && grid_x > 0 && grid_y > 0 && stride > 0 && reach >= 0, PCB_EINVAL);
    scene_window_count_kernel<<<(unsigned)ceil_div(P, kScThreads), kScThreads, 0, (cudaStream_t)stream>>>(
        points, P, point_stride, make_grid(lo_x, hi_x, lo_y, hi_y, grid_x, grid_y, x0, y0, stride, reach), counts);
    PCB_RETURN_LAUNCH_STATUS();
}

PCB_API int pcb_scene_window_fill_f32(const float *points, int64_t P, int point_stride, int grid_x, int grid_y,
                                      const double *lo_x, const double *hi_x, const double *lo_y, const double *hi_y,
                                      double x0, double y0, double stride, int reach, const int64_t *offsets, int *cursor,
                                      unsigned seed, int64_t *keys, pcb_stream_t stream)
{
    PCB_REQUIRE(points && lo_x && hi_x && lo_y && hi_y && offsets && cursor && keys, PCB_EINVAL);
    PCB_REQUIRE(P > 0 && P < (1ll << 31) && point_stride >= 3 && grid_x > 0 && grid_y > 0 && stride > 0 && reach >= 0,
                PCB_EINVAL);
    PCB_REQUIRE((int64_t)grid_x * grid_y < (1ll << 16), PCB_ERANGE);       // window id must fit the key's top bits
    scene_window_fill_kernel<<<(unsigned)ceil_div(P, kScThreads), kScThreads, 0, (cudaStream_t)stream>>>(
        points, P, point_stride, make_grid(lo_x, hi_x, lo_y, hi_y, grid_x, grid_y, x0, y0, stride, reach), offsets, cursor,
        seed, keys);
    PCB_RETURN_LAUNCH_STATUS();
}

PCB_API int pcb_scene_blocks_f32(const float *points, int point_stride, const int *members, const int64_t *blk_off,
                                 const int *blk_cnt, const int64_t *blk_first, const double *blk_center, int64_t nblocks,
                                 int block_points, double ext_x, double ext_y, double ext_z, unsigned seed, float *data,
                                 int64_t *point_idx, pcb_stream_t stream)
{
    PCB_REQUIRE(points && members && blk_off && blk_cnt && blk_first && blk_center && data && point_idx, PCB_EINVAL);
    PCB_REQUIRE(nblocks > 0 && block_points > 0 && point_stride >= 6, PCB_EINVAL);
    const int64_t total = nblocks * block_points;
    PCB_REQUIRE(ceil_div(total, kScThreads) < (1ll << 31), PCB_ERANGE);
    scene_blocks_kernel<<<(unsigned)ceil_div(total, kScThreads), kScThreads, 0, (cudaStream_t)stream>>>(
        points, point_stride, members, blk_off, blk_cnt, blk_first, blk_center, block_points, total, ext_x, ext_y, ext_z,
        seed, data, point_idx);
    PCB_RETURN_LAUNCH_STATUS();
}

PCB_API int pcb_scene_vote(const int64_t *point_idx, const unsigned char *pred, int64_t total, int64_t P, int num_classes,
                           int *pool, pcb_stream_t stream)
{
    PCB_REQUIRE(point_idx && pred && pool, PCB_EINVAL);
    PCB_REQUIRE(total > 0 && P > 0 && num_classes > 0 && num_classes <= 255, PCB_EINVAL);
    PCB_REQUIRE(ceil_div(total, kScThreads) < (1ll << 31), PCB_ERANGE);
    scene_vote_kernel<<<(unsigned)ceil_div(total, kScThreads), kScThreads, 0, (cudaStream_t)stream>>>(point_idx, pred, total,
                                                                                                    P, num_classes, pool);
    PCB_RETURN_LAUNCH_STATUS();
}

PCB_API int pcb_scene_vote_argmax(const int *pool, int64_t P, int num_classes, unsigned char *labels, pcb_stream_t stream)
{
    PCB_REQUIRE(pool && labels, PCB_EINVAL);
    PCB_REQUIRE(P > 0 && num_classes > 0 && num_classes <= 255, PCB_EINVAL);
    scene_vote_argmax_kernel<<<(unsigned)ceil_div(P, kScThreads), kScThreads, 0, (cudaStream_t)stream>>>(pool, P, num_classes,
                                                                                                      labels);
    PCB_RETURN_LAUNCH_STATUS();
}
