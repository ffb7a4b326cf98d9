// Nearest-neighbour searches in coordinate space (3-D points):
//   pcb_three_nn_f32   -- the k=3 (k=4) nearest of xyz2 for every xyz1 point + inverse-distance
//                         weights; replaces square_distance + full sort of
//                         Partsize-identical/models/pointnet_util.py:325-332 and
//                         Highway_bridge/models/pointnet2_utils.py:183-191, 253-262.
//   pcb_knn_f32 (D=3)  -- DGCNN.knn on coordinates, Highway_bridge/models/DGCNN.py:49-70.
//   pcb_knn_cdist_f32  -- torch.cdist + topk(largest=False),
//                         Highway_bridge/models/attention_modules.py:584-586, 736-738.
// None of them materialises the [B,N,M] distance matrix.  Candidates are staged in shared
// memory as float4 (x, y, z, |p|^2); distances use the reference's exact fp32 formula
// (pcb_common.cuh) and selection orders by (distance, index).
#include "pcb_common.cuh"

namespace pcb {

// ------------------------------------------------------------------------------------------
// staging helper: cloud[c0 : c0+len] -> s_pts[0:len] as (x, y, z, |p|^2)
// ------------------------------------------------------------------------------------------
// cf != 0: the cloud is channels-first ([3,N], as DGCNN passes x), else points-major ([N,3]).
__device__ __forceinline__ float3 load_point(const float *cloud, int i, int N, int cf)
{
    if (cf) return make_float3(__ldg(cloud + i), __ldg(cloud + N + i), __ldg(cloud + 2 * (size_t)N + i));
    const float *p = cloud + (size_t)i * 3;
    return make_float3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}
__device__ __forceinline__ void stage_points(float4 *s_pts, const float *cloud, int c0, int len,
                                             int N = 0, int cf = 0)
{
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
        float3 v = load_point(cloud, c0 + i, N, cf);
        s_pts[i] = make_float4(v.x, v.y, v.z, norm3(v.x, v.y, v.z));
    }
}

// ------------------------------------------------------------------------------------------
// three_nn: one thread per query, top-K in registers, candidates broadcast from smem.
// ------------------------------------------------------------------------------------------
constexpr int kNnThreads = 256;
constexpr int kNnChunk = 4096;               // candidates per staging pass (64 KB)

template <int K>
__global__ void __launch_bounds__(kNnThreads)
three_nn_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2, int N, int S,
                float *__restrict__ out_dist, int64_t *__restrict__ out_idx,
                float *__restrict__ out_w, int chunk)
{
    extern __shared__ __align__(16) float4 s_pts[];
    const int b = blockIdx.y;
    const int n = blockIdx.x * kNnThreads + threadIdx.x;
    const bool active = n < N;
    const float *cloud2 = xyz2 + (size_t)b * S * 3;

    float qx = 0.f, qy = 0.f, qz = 0.f, qn = 0.f;
    if (active) {
        const float *q = xyz1 + ((size_t)b * N + n) * 3;
        qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
        qn = norm3(qx, qy, qz);
    }
    float bd[K];
    int bi[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        bd[j] = __int_as_float(0x7f800000);      // +inf
        bi[j] = 0;
    }
    for (int c0 = 0; c0 < S; c0 += chunk) {
        const int len = min(chunk, S - c0);
        if (c0 > 0) __syncthreads();
        stage_points(s_pts, cloud2, c0, len);
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int s = 0; s < len; ++s) {
                float4 p = s_pts[s];
                float d = sqdist3(qx, qy, qz, qn, p.x, p.y, p.z, p.w);   // square_distance(xyz1, xyz2)
                if (d < bd[K - 1]) {
                    bd[K - 1] = d;
                    bi[K - 1] = c0 + s;
#pragma unroll
                    for (int j = K - 1; j > 0; --j) {
                        if (bd[j] < bd[j - 1]) {                         // strict: earlier index stays first
                            float td = bd[j]; bd[j] = bd[j - 1]; bd[j - 1] = td;
                            int ti = bi[j]; bi[j] = bi[j - 1]; bi[j - 1] = ti;
                        }
                    }
                }
            }
        }
    }
    if (!active) return;
    const size_t o = ((size_t)b * N + n) * K;
    float rec[K];
    float norm = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        out_dist[o + j] = bd[j];
        out_idx[o + j] = (int64_t)bi[j];
        // dist_recip = 1.0 / (dists + 1e-8); norm = sum(dist_recip)   (pointnet_util.py:330-331)
        rec[j] = __fdiv_rn(1.0f, __fadd_rn(bd[j], 1e-8f));
        norm = (j == 0) ? rec[0] : __fadd_rn(norm, rec[j]);
    }
    if (out_w) {
#pragma unroll
        for (int j = 0; j < K; ++j) out_w[o + j] = __fdiv_rn(rec[j], norm);
    }
}

template <int K>
static int launch_three_nn(const float *xyz1, const float *xyz2, int B, int N, int S, float *od,
                           int64_t *oi, float *ow, cudaStream_t st)
{
    int chunk = S < kNnChunk ? S : kNnChunk;
    size_t smem = (size_t)chunk * sizeof(float4);
    cudaError_t e = cudaFuncSetAttribute(three_nn_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kNnChunk * (int)sizeof(float4));
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)ceil_div(N, kNnThreads), (unsigned)B);
    three_nn_kernel<K><<<grid, kNnThreads, smem, st>>>(xyz1, xyz2, N, S, od, oi, ow, chunk);
    PCB_RETURN_LAUNCH_STATUS();
}

// ------------------------------------------------------------------------------------------
// kNN with k up to 64 in coordinate space: one THREAD per query.
// Candidates are staged chunk by chunk in shared memory as float4 (x, y, z, |p|^2) and read by
// broadcast; a thread tests 64 candidates against its running k-th distance (a register) into
// a 64-bit mask, then drains the mask into its sorted top-K list, which lives in registers
// (K = k rounded up to a compiled size; the first k entries of a top-K list are the top-k).  An
// insertion is a branch-free compare-exchange pass.  Deferring the (rare, divergent) insertions
// to the drain keeps the scan loop branch-free; the round-1 warp-per-query version spent most
// of its issue slots in shuffle insertions (profiles/r1_*).
// ------------------------------------------------------------------------------------------
enum { MODE_PD = 0, MODE_CDIST = 1 };

constexpr int kKnnThreads = 128;
constexpr int kKnnChunk = 1024;              // candidates per staging pass (16 KB)

template <int MODE>
__device__ __forceinline__ float knn_pair(float qx, float qy, float qz, float qn, float ax, float ay, float az,
                                          const float4 p)
{
    if (MODE == MODE_PD) return sqdist3(qx, qy, qz, qn, p.x, p.y, p.z, p.w);
    float t = cdist3_pre(ax, ay, az, qn, p.x, p.y, p.z, p.w);
    t = t < 0.0f ? 0.0f : t;                 // clamp_min_(0)
    return __fsqrt_rn(t);
}

template <int MODE, int K>
__global__ void __launch_bounds__(kKnnThreads, (K <= 20 ? 4 : (K <= 32 ? 3 : 2)))
knn_xyz_kernel(const float *__restrict__ xyz, int N, int k, int64_t *__restrict__ out_idx,
               float *__restrict__ out_dist, int cf)
{
    __shared__ __align__(16) float4 s_pts[kKnnChunk];  // candidates of the current chunk
    const int b = blockIdx.y, tid = threadIdx.x;
    const int q = blockIdx.x * kKnnThreads + tid;
    const bool active = q < N;
    const float *cloud = xyz + (size_t)b * N * 3;

    float qx = 0.f, qy = 0.f, qz = 0.f, qn = 0.f;
    if (active) {
        float3 v = load_point(cloud, q, N, cf);
        qx = v.x, qy = v.y, qz = v.z;
        qn = norm3(qx, qy, qz);
    }
    float ax = qx, ay = qy, az = qz;
    if (MODE == MODE_CDIST) {                          // x1_ = cat(-2 * x1, |x1|^2, 1)
        ax = __fmul_rn(qx, -2.0f);
        ay = __fmul_rn(qy, -2.0f);
        az = __fmul_rn(qz, -2.0f);
    }
    float ld[K];
    int li[K];
#pragma unroll
    for (int p = 0; p < K; ++p) {
        ld[p] = __int_as_float(0x7f800000);
        li[p] = 0x7fffffff;
    }
    float thr = __int_as_float(0x7f800000);                // K-th smallest distance so far
    const float4 pad = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));   // distance = +inf

    for (int c0 = 0; c0 < N; c0 += kKnnChunk) {
        const int len = min(kKnnChunk, N - c0);
        const int len64 = (len + 63) & ~63;
        if (c0 > 0) __syncthreads();
        stage_points(s_pts, cloud, c0, len, N, cf);
        for (int i = len + tid; i < len64; i += blockDim.x) s_pts[i] = pad;
        __syncthreads();
        if (!active) continue;
        for (int s0 = 0; s0 < len64; s0 += 64) {
            unsigned lo = 0u, hi = 0u;
#pragma unroll
            for (int c = 0; c < 32; ++c)
                if (knn_pair<MODE>(qx, qy, qz, qn, ax, ay, az, s_pts[s0 + c]) < thr) lo |= 1u << c;
#pragma unroll
            for (int c = 0; c < 32; ++c)
                if (knn_pair<MODE>(qx, qy, qz, qn, ax, ay, az, s_pts[s0 + 32 + c]) < thr) hi |= 1u << c;
            unsigned long long mask = ((unsigned long long)hi << 32) | lo;
            while (mask) {
                const int c = __ffsll((long long)mask) - 1;           // ascending candidate index
                mask &= mask - 1;
                const float v = knn_pair<MODE>(qx, qy, qz, qn, ax, ay, az, s_pts[s0 + c]);
                if (v < thr) {
                    float cv = v;
                    int ci = c0 + s0 + c;
                    bool placed = false;
#pragma unroll
                    for (int p = 0; p < K; ++p) {                      // compare-exchange down the list
                        // new element: strict < (equal distance -> the earlier index stays first);
                        // once it is placed we carry old entries, which simply shift down one slot
                        const bool sw = placed || (cv < ld[p]);
                        placed = sw;
                        const float nd = sw ? cv : ld[p], nc = sw ? ld[p] : cv;
                        const int ni = sw ? ci : li[p], nci = sw ? li[p] : ci;
                        ld[p] = nd;
                        li[p] = ni;
                        cv = nc;
                        ci = nci;
                    }
                    thr = ld[K - 1];
                }
            }
        }
    }
    if (active) {
        int64_t *oi = out_idx + ((size_t)b * N + q) * k;
        float *od = out_dist ? out_dist + ((size_t)b * N + q) * k : nullptr;
#pragma unroll
        for (int p = 0; p < K; ++p) {
            if (p < k) {
                oi[p] = (int64_t)li[p];
                if (od) od[p] = ld[p];
            }
        }
    }
}

template <int MODE, int K>
static int launch_knn_xyz_k(const float *xyz, int B, int N, int k, int cf, int64_t *oi, float *od, cudaStream_t st)
{
    dim3 grid((unsigned)ceil_div(N, kKnnThreads), (unsigned)B);
    knn_xyz_kernel<MODE, K><<<grid, kKnnThreads, 0, st>>>(xyz, N, k, oi, od, cf);
    PCB_RETURN_LAUNCH_STATUS();
}

template <int MODE>
static int launch_knn_xyz(const float *xyz, int B, int N, int k, int cf, int64_t *oi, float *od,
                          cudaStream_t st)
{
    if (k <= 8) return launch_knn_xyz_k<MODE, 8>(xyz, B, N, k, cf, oi, od, st);
    if (k <= 16) return launch_knn_xyz_k<MODE, 16>(xyz, B, N, k, cf, oi, od, st);
    if (k <= 20) return launch_knn_xyz_k<MODE, 20>(xyz, B, N, k, cf, oi, od, st);
    if (k <= 32) return launch_knn_xyz_k<MODE, 32>(xyz, B, N, k, cf, oi, od, st);
    return launch_knn_xyz_k<MODE, 64>(xyz, B, N, k, cf, oi, od, st);
}

int knn_xyz_pd(const float *xyz, int B, int N, int k, int cf, int64_t *oi, float *od, cudaStream_t st)
{
    return launch_knn_xyz<MODE_PD>(xyz, B, N, k, cf, oi, od, st);
}

}  // namespace pcb

PCB_API int pcb_three_nn_f32(const float *xyz1, const float *xyz2, int B, int N, int S, int k,
                             float *out_dist, int64_t *out_idx, float *out_weight, pcb_stream_t stream)
{
    using namespace pcb;
    PCB_REQUIRE(xyz1 && xyz2 && out_dist && out_idx, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && S > 0 && k > 0, PCB_EINVAL);
    PCB_REQUIRE(k <= 8 && k <= S && B <= 65535, PCB_ERANGE);
    cudaStream_t st = (cudaStream_t)stream;
    switch (k) {
        case 1: return launch_three_nn<1>(xyz1, xyz2, B, N, S, out_dist, out_idx, out_weight, st);
        case 2: return launch_three_nn<2>(xyz1, xyz2, B, N, S, out_dist, out_idx, out_weight, st);
        case 3: return launch_three_nn<3>(xyz1, xyz2, B, N, S, out_dist, out_idx, out_weight, st);
        case 4: return launch_three_nn<4>(xyz1, xyz2, B, N, S, out_dist, out_idx, out_weight, st);
        case 5: return launch_three_nn<5>(xyz1, xyz2, B, N, S, out_dist, out_idx, out_weight, st);
        case 6: return launch_three_nn<6>(xyz1, xyz2, B, N, S, out_dist, out_idx, out_weight, st);
        case 7: return launch_three_nn<7>(xyz1, xyz2, B, N, S, out_dist, out_idx, out_weight, st);
        default: return launch_three_nn<8>(xyz1, xyz2, B, N, S, out_dist, out_idx, out_weight, st);
    }
}

PCB_API int pcb_knn_cdist_f32(const float *xyz, int B, int N, int k, int64_t *out_idx, float *out_dist,
                              pcb_stream_t stream)
{
    using namespace pcb;
    PCB_REQUIRE(xyz && out_idx, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && k > 0, PCB_EINVAL);
    PCB_REQUIRE(k <= 64 && k <= N && B <= 65535, PCB_ERANGE);
    return launch_knn_xyz<MODE_CDIST>(xyz, B, N, k, 0, out_idx, out_dist, (cudaStream_t)stream);
}
