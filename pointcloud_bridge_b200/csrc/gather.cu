// Gather-type kernels (HBM-bound):
//   pcb_gather_f32 / _bwd          index_points          pointnet_util.py:46-63; pointnet2_utils.py:17-39
//   pcb_group_points_f32 / _bwd    grouping + concat     pointnet_util.py:137-147, 260-267;
//                                                        pointnet2_utils.py:50-58, 342-349
//   pcb_graph_feature_f32 / _bwd   get_graph_feature     Highway_bridge/models/DGCNN.py:72-109
//   pcb_interpolate_f32 / _bwd     inverse-distance interpolation   pointnet_util.py:333-334
//   pcb_square_distance_f32        square_distance       pointnet_util.py:22-43
// The output (the big operand) is written once with streaming stores, fully coalesced, with
// 128-bit accesses wherever the row length allows; the gathered source stays L1/L2 resident.
#include <cuda_bf16.h>

#include "pcb_common.cuh"

namespace pcb {

constexpr int kThreads = 256;

__device__ __forceinline__ void store_out(float *p, float v) { st_stream_f1(p, v); }
__device__ __forceinline__ void store_out(__nv_bfloat16 *p, float v) { *p = __float2bfloat16(v); }
__device__ __forceinline__ float load_grad(const float *p) { return __ldg(p); }
__device__ __forceinline__ float load_grad(const __nv_bfloat16 *p) { return __bfloat162float(*p); }

__device__ __forceinline__ bool resolve_index(long long &i, int N, int clamp)
{
    if (clamp) {
        i = i < 0 ? 0 : (i > N - 1 ? N - 1 : i);
        return true;
    }
    if (i < 0) i += N;                                  // Python negative indexing
    return i >= 0 && i < N;
}

// ------------------------------------------------------------------------------------------
// index_points.  VEC = 4: C % 4 == 0 and 16-byte aligned rows; VEC = 1: anything.
// One thread per VEC consecutive output floats -> coalesced stores.  total < 2^31.
// ------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(kThreads)
gather_kernel(const float *__restrict__ points, const int64_t *__restrict__ idx, int N, int C, FastDiv dCV,
              FastDiv dM, unsigned total, int clamp, float *__restrict__ out, int *__restrict__ err)
{
    const unsigned t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    const unsigned rowg = dCV.div(t);                  // b * M + m
    const int c = (int)(t - rowg * dCV.d) * VEC;
    const unsigned b = dM.div(rowg);
    long long i = idx[rowg];
    const bool ok = resolve_index(i, N, clamp);
    if (VEC == 4) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) v = __ldg(reinterpret_cast<const float4 *>(points + ((size_t)b * N + i) * C + c));
        st_stream_f4(reinterpret_cast<float4 *>(out + (size_t)rowg * C + c), v);
    } else {
        float v = ok ? __ldg(points + ((size_t)b * N + i) * C + c) : 0.f;
        st_stream_f1(out + (size_t)rowg * C + c, v);
    }
    if (!ok && err && c == 0) atomicAdd(err, 1);
}

// 128-bit variant with U rows in flight per thread: thread (rg, cv) serves float4 column cv of rows
// rg, rg + RG, ..., so the U index loads and then the U row loads are independent of each other (one
// load in flight per thread left the kernel latency-bound at ~55 % of the HBM peak).
constexpr int kGatherU = 4;
__global__ void __launch_bounds__(kThreads)
gather_vec_ilp_kernel(const float *__restrict__ points, const int64_t *__restrict__ idx, int N, int C, FastDiv dCV,
                      FastDiv dM, unsigned RG, unsigned rows, int clamp, float *__restrict__ out, int *__restrict__ err)
{
    const unsigned t = blockIdx.x * kThreads + threadIdx.x;
    const unsigned rg = dCV.div(t);
    if (rg >= RG) return;
    const int c = (int)(t - rg * dCV.d) * 4;
    long long i[kGatherU];
    unsigned row[kGatherU];
    bool live[kGatherU], ok[kGatherU];
#pragma unroll
    for (int u = 0; u < kGatherU; ++u) {
        row[u] = rg + u * RG;
        live[u] = row[u] < rows;
        i[u] = live[u] ? idx[row[u]] : 0;
    }
    float4 v[kGatherU];
#pragma unroll
    for (int u = 0; u < kGatherU; ++u) {
        ok[u] = resolve_index(i[u], N, clamp);
        const unsigned b = dM.div(row[u]);
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live[u] && ok[u]) v[u] = __ldg(reinterpret_cast<const float4 *>(points + ((size_t)b * N + i[u]) * C + c));
    }
#pragma unroll
    for (int u = 0; u < kGatherU; ++u)
        if (live[u]) {
            st_stream_f4(reinterpret_cast<float4 *>(out + (size_t)row[u] * C + c), v[u]);
            if (!ok[u] && err && c == 0) atomicAdd(err, 1);
        }
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
gather_bwd_kernel(const float *__restrict__ gout, const int64_t *__restrict__ idx, int N, int C, FastDiv dCV,
                  FastDiv dM, unsigned total, int clamp, float *__restrict__ gpoints)
{
    const unsigned t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    const unsigned rowg = dCV.div(t);
    const int c = (int)(t - rowg * dCV.d) * VEC;
    const unsigned b = dM.div(rowg);
    long long i = idx[rowg];
    if (!resolve_index(i, N, clamp)) return;
    float *dst = gpoints + ((size_t)b * N + i) * C + c;
    if (VEC == 4) {
        float4 g = ld_stream_f4(reinterpret_cast<const float4 *>(gout + (size_t)rowg * C + c));
        atomicAdd(reinterpret_cast<float4 *>(dst), g);
    } else {
        atomicAdd(dst, __ldg(gout + (size_t)rowg * C + c));
    }
}

// ------------------------------------------------------------------------------------------
// grouping + concat.  out[b,s,k,:] = cat(xyz[idx]-new_xyz, points[idx]) (or the MSG order).
// One thread per output float; threads of a row share the index (broadcast load).
// ------------------------------------------------------------------------------------------
template <typename OT>
__global__ void __launch_bounds__(kThreads)
group_points_kernel(const float *__restrict__ xyz, const float *__restrict__ points,
                    const float *__restrict__ new_xyz, const int64_t *__restrict__ idx, int N, int D, FastDiv dC,
                    FastDiv dK, FastDiv dS, int xyz_first, int points_cf, int clamp, unsigned total,
                    OT *__restrict__ out)
{
    // dC.d = row pitch of `out` in elements (>= 3 + D; the extra columns are written as zeros)
    const unsigned t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    const unsigned rowg = dC.div(t);                   // (b*S + s)*K + k
    const int c = (int)(t - rowg * dC.d);
    if (c >= 3 + D) {
        store_out(out + t, 0.f);
        return;
    }
    const unsigned bs = dK.div(rowg);                  // b*S + s
    const unsigned b = dS.div(bs);
    long long i = idx[rowg];
    const bool ok = resolve_index(i, N, clamp);
    const int cx = xyz_first ? c : c - D;              // coordinate channel if in [0,3)
    float v = 0.f;
    if (ok) {
        if (cx >= 0 && cx < 3) {
            v = __fsub_rn(__ldg(xyz + ((size_t)b * N + i) * 3 + cx), __ldg(new_xyz + (size_t)bs * 3 + cx));
        } else {
            const int cf = xyz_first ? c - 3 : c;
            v = points_cf ? __ldg(points + ((size_t)b * D + cf) * N + i)
                          : __ldg(points + ((size_t)b * N + i) * D + cf);
        }
    }
    store_out(out + t, v);
}

template <typename GT>
__global__ void __launch_bounds__(kThreads)
group_points_bwd_kernel(const GT *__restrict__ gout, const int64_t *__restrict__ idx, int N, int D, int C,
                        FastDiv dD, FastDiv dSK, int xyz_first, int points_cf, int clamp, unsigned total,
                        float *__restrict__ gpoints)
{
    // C = row pitch of gout in elements (>= 3 + D)
    // one thread per (row, feature channel)
    const unsigned t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    const unsigned rowg = dD.div(t);
    const int cf = (int)(t - rowg * dD.d);
    const unsigned b = dSK.div(rowg);
    long long i = idx[rowg];
    if (!resolve_index(i, N, clamp)) return;
    const float g = load_grad(gout + (size_t)rowg * C + (xyz_first ? 3 + cf : cf));
    float *dst = points_cf ? gpoints + ((size_t)b * D + cf) * N + i : gpoints + ((size_t)b * N + i) * D + cf;
    atomicAdd(dst, g);
}

// ---- 8 channels per thread (row pitch a multiple of 8: the padded rows handed to the shared-MLP GEMM) ----
// One index load and one address decomposition per 8 outputs, one 16-byte store (bf16) or two (fp32);
// feature chunks of 16-byte aligned point-major rows come in with two 128-bit loads.
__device__ __forceinline__ uint4 pack8_bf16(const float v[8])
{
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<unsigned *>(&a);
    u.y = *reinterpret_cast<unsigned *>(&b);
    u.z = *reinterpret_cast<unsigned *>(&c);
    u.w = *reinterpret_cast<unsigned *>(&d);
    return u;
}
__device__ __forceinline__ void store8(float *p, const float v[8])
{
    st_stream_f4(reinterpret_cast<float4 *>(p), make_float4(v[0], v[1], v[2], v[3]));
    st_stream_f4(reinterpret_cast<float4 *>(p) + 1, make_float4(v[4], v[5], v[6], v[7]));
}
__device__ __forceinline__ void store8(__nv_bfloat16 *p, const float v[8])
{
    *reinterpret_cast<uint4 *>(p) = pack8_bf16(v);
}

__device__ __forceinline__ float ldp(const float *p) { return __ldg(p); }
__device__ __forceinline__ float ldp(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
__device__ __forceinline__ void ldp8(const float *p, float v[8])
{
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), c = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = c.x, v[5] = c.y, v[6] = c.z, v[7] = c.w;
}
__device__ __forceinline__ void ldp8(const __nv_bfloat16 *p, float v[8])
{
    const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p));
    const unsigned w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        v[2 * j] = __uint_as_float(w[j] << 16);
        v[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
    }
}

// PT: type of the feature rows (fp32, or bf16 when the previous layer ran under autocast)
// Thread (rg, q) serves 8-channel chunk q of rows rg, rg + RG, ... (U rows in flight: the index loads, then
// the gathers, then the stores of the U rows are independent -- a single dependent idx -> row -> store
// chain per thread left the kernel latency-bound at a third of the HBM peak).
#ifndef PCB_GROUP_U
#define PCB_GROUP_U 1
#endif
constexpr int kGroupU = PCB_GROUP_U;
template <typename OT, typename PT>
__global__ void __launch_bounds__(kThreads)
group_points_chunk_kernel(const float *__restrict__ xyz, const PT *__restrict__ points,
                          const float *__restrict__ new_xyz, const int64_t *__restrict__ idx, int N, int D, FastDiv dQ,
                          FastDiv dK, FastDiv dS, int xyz_first, int points_cf, int clamp, int vec_ok, unsigned RG,
                          unsigned rows, OT *__restrict__ out)
{
    pdl_wait();
    pdl_trigger();
    // dQ.d = chunks per row = pitch / 8
    const unsigned t = blockIdx.x * kThreads + threadIdx.x;
    const unsigned rg = dQ.div(t);
    if (rg >= RG) return;
    const int q = (int)(t - rg * dQ.d);
    const int c0 = q * 8;
    unsigned row[kGroupU];
    long long i[kGroupU];
    bool live[kGroupU];
#pragma unroll
    for (int u = 0; u < kGroupU; ++u) {
        row[u] = rg + u * RG;                          // (b*S + s)*K + k
        live[u] = row[u] < rows;
        i[u] = live[u] ? idx[row[u]] : 0;
    }
    const int f0 = xyz_first ? c0 - 3 : c0;            // feature channel of the chunk's first element
    float v[kGroupU][8];
    if (vec_ok && f0 >= 0 && f0 + 8 <= D && (f0 & 7) == 0) {       // whole chunk inside the features: same for all rows
#pragma unroll
        for (int u = 0; u < kGroupU; ++u) {
            const bool ok = resolve_index(i[u], N, clamp);
            const unsigned b = dS.div(dK.div(row[u]));
#pragma unroll
            for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
            if (live[u] && ok) ldp8(points + ((size_t)b * N + i[u]) * D + f0, v[u]);
        }
    } else if (!xyz_first && c0 == D) {                // the [dxyz | 0 0 0 0 0] chunk that ends a [feat | dxyz] row
        float px[kGroupU][3], qx[kGroupU][3];
        bool ok[kGroupU];
#pragma unroll
        for (int u = 0; u < kGroupU; ++u) {
            ok[u] = resolve_index(i[u], N, clamp) && live[u];
            const unsigned bs = live[u] ? dK.div(row[u]) : 0u;
            const unsigned b = dS.div(bs);
            const float *p = xyz + ((size_t)b * N + (ok[u] ? i[u] : 0)) * 3;
            const float *c = new_xyz + (size_t)bs * 3;
#pragma unroll
            for (int j = 0; j < 3; ++j) px[u][j] = __ldg(p + j), qx[u][j] = __ldg(c + j);
        }
#pragma unroll
        for (int u = 0; u < kGroupU; ++u) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
            if (ok[u]) {
#pragma unroll
                for (int j = 0; j < 3; ++j) v[u][j] = __fsub_rn(px[u][j], qx[u][j]);
            }
        }
    } else {
#pragma unroll
        for (int u = 0; u < kGroupU; ++u) {
            const bool ok = resolve_index(i[u], N, clamp);
            const unsigned bs = dK.div(row[u]);
            const unsigned b = dS.div(bs);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = c0 + j;
                float x = 0.f;
                if (live[u] && ok && c < 3 + D) {
                    const int cx = xyz_first ? c : c - D;
                    if (cx >= 0 && cx < 3) {
                        x = __fsub_rn(__ldg(xyz + ((size_t)b * N + i[u]) * 3 + cx), __ldg(new_xyz + (size_t)bs * 3 + cx));
                    } else {
                        const int cf = xyz_first ? c - 3 : c;
                        x = points_cf ? ldp(points + ((size_t)b * D + cf) * N + i[u])
                                      : ldp(points + ((size_t)b * N + i[u]) * D + cf);
                    }
                }
                v[u][j] = x;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < kGroupU; ++u)
        if (live[u]) store8(out + ((size_t)row[u] * dQ.d + q) * 8, v[u]);
}

// ---- row-tile variant (point-major features) ----
// A CTA owns TR consecutive output rows.  Their neighbour indices are resolved ONCE into shared memory
// (global row number of the source point, or -1), so the gathers of the main loop do not hang off a
// per-thread index load, every warp runs one code path (feature chunks first, the [dxyz | 0] tail
// chunks of the same rows right after, while their cache lines are still in L2), and four independent
// items are in flight per thread.  VEC8: [feat | dxyz | 0] rows with D % 8 == 0 and pitch D + 8 (the
// training layout), 16-byte loads and stores; otherwise one element per item, any layout / pitch.
constexpr int kTileMaxRows = 128;
constexpr int kTileU = 4;
template <typename OT, typename PT, bool VEC8>
__global__ void __launch_bounds__(kThreads)
group_tile_kernel(const float *__restrict__ xyz, const PT *__restrict__ points, const float *__restrict__ new_xyz,
                  const int64_t *__restrict__ idx, int N, int D, int C, FastDiv dK, FastDiv dS, FastDiv dW,
                  int xyz_first, int clamp, int TR, unsigned rows, OT *__restrict__ out)
{
    pdl_wait();
    pdl_trigger();
    __shared__ int s_src[kTileMaxRows];                // b * N + i of the gathered point, -1: index out of range
    __shared__ unsigned s_bs[kTileMaxRows];            // b * S + s (centroid row)
    const unsigned row0 = blockIdx.x * (unsigned)TR;
    const int nr = rows - row0 < (unsigned)TR ? (int)(rows - row0) : TR;
    for (int w = threadIdx.x; w < nr; w += kThreads) {
        const unsigned row = row0 + w;
        long long i = idx[row];
        const bool ok = resolve_index(i, N, clamp);
        const unsigned bs = dK.div(row);
        s_bs[w] = bs;
        s_src[w] = ok ? (int)(dS.div(bs) * (unsigned)N + (unsigned)i) : -1;
    }
    __syncthreads();
    if (VEC8) {
        const int QF = D >> 3;                         // feature chunks per row; dW.d == QF
        const int items = nr * QF;
        for (int w0 = threadIdx.x; w0 < items; w0 += kThreads * kTileU) {
            float v[kTileU][8];
            int r[kTileU], q[kTileU];
#pragma unroll
            for (int u = 0; u < kTileU; ++u) {
                const int w = w0 + u * kThreads;
                r[u] = w < items ? (int)dW.div((unsigned)w) : -1;
                q[u] = w - r[u] * QF;
                const int src = r[u] >= 0 ? s_src[r[u]] : -1;
#pragma unroll
                for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
                if (src >= 0) ldp8(points + (size_t)src * D + q[u] * 8, v[u]);
            }
#pragma unroll
            for (int u = 0; u < kTileU; ++u)
                if (r[u] >= 0) store8(out + (size_t)(row0 + r[u]) * C + q[u] * 8, v[u]);
        }
        for (int w = threadIdx.x; w < nr; w += kThreads) {          // tail chunk: [dx dy dz 0 0 0 0 0]
            const int src = s_src[w];
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
            if (src >= 0) {
                const float *p = xyz + (size_t)src * 3, *c = new_xyz + (size_t)s_bs[w] * 3;
#pragma unroll
                for (int j = 0; j < 3; ++j) v[j] = __fsub_rn(__ldg(p + j), __ldg(c + j));
            }
            store8(out + (size_t)(row0 + w) * C + D, v);
        }
    } else {
        // one warp per row, lanes stride over the C output columns (no per-element division), kTileU rows in
        // flight per warp; a row is written as contiguous 128-byte segments
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        constexpr int kWarps = kThreads / 32;
        for (int r0 = warp; r0 < nr; r0 += kWarps * kTileU) {
            int src[kTileU];
            unsigned bs[kTileU];
#pragma unroll
            for (int u = 0; u < kTileU; ++u) {
                const int r = r0 + u * kWarps;
                src[u] = r < nr ? s_src[r] : -2;           // -2: no such row, -1: index out of range (row of zeros)
                bs[u] = r < nr ? s_bs[r] : 0u;
            }
            for (int c = lane; c < C; c += 32) {
                const int cx = xyz_first ? c : c - D;
                const bool is_xyz = cx >= 0 && cx < 3;
                const int cf = xyz_first ? c - 3 : c;
                float v[kTileU];
#pragma unroll
                for (int u = 0; u < kTileU; ++u) {
                    v[u] = 0.f;
                    if (src[u] >= 0 && c < 3 + D)
                        v[u] = is_xyz ? __fsub_rn(__ldg(xyz + (size_t)src[u] * 3 + cx), __ldg(new_xyz + (size_t)bs[u] * 3 + cx))
                                      : ldp(points + (size_t)src[u] * D + cf);
                }
#pragma unroll
                for (int u = 0; u < kTileU; ++u)
                    if (src[u] != -2) store_out(out + (size_t)(row0 + r0 + u * kWarps) * C + c, v[u]);
            }
        }
    }
}

// backward, 4 feature channels per thread: one 128-bit reduction (red.global.add.v4.f32) per thread.
// D % 4 == 0, point-major gradient rows, 16-byte aligned grad_points.
__device__ __forceinline__ float4 load_grad4(const float *p) { return ld_stream_f4(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 load_grad4(const __nv_bfloat16 *p)
{
    const uint2 u = *reinterpret_cast<const uint2 *>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

template <typename GT>
__global__ void __launch_bounds__(kThreads)
group_points_bwd_vec_kernel(const GT *__restrict__ gout, const int64_t *__restrict__ idx, int N, int D, int C,
                            FastDiv dDV, FastDiv dSK, int foff, int clamp, int aligned, unsigned total,
                            float *__restrict__ gpoints)
{
    pdl_wait();
    pdl_trigger();
    const unsigned t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    const unsigned rowg = dDV.div(t);
    const int cf = (int)(t - rowg * dDV.d) * 4;
    const unsigned b = dSK.div(rowg);
    long long i = idx[rowg];
    if (!resolve_index(i, N, clamp)) return;
    const GT *g = gout + (size_t)rowg * C + foff + cf;
    float4 gv;
    if (aligned)
        gv = load_grad4(g);
    else
        gv = make_float4(load_grad(g), load_grad(g + 1), load_grad(g + 2), load_grad(g + 3));
    atomicAdd(reinterpret_cast<float4 *>(gpoints + ((size_t)b * N + i) * D + cf), gv);
}

// ------------------------------------------------------------------------------------------
// get_graph_feature.  x [B,D,N], idx [B,N,k] -> out [B,2D,N,k].
// A thread owns EPT consecutive (n, j) edges and walks CPT channels: indices are read once per
// CPT channels, both output planes get 128-bit streaming stores, gathers hit one 4*N-byte
// row of x per channel (L1/L2 resident).
// ------------------------------------------------------------------------------------------
constexpr int kGfCPT = 8;

template <int EPT>
__global__ void __launch_bounds__(kThreads)
graph_feature_kernel(const float *__restrict__ x, const int64_t *__restrict__ idx, int D, int N, FastDiv dk,
                     unsigned NK, float *__restrict__ out)
{
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * kGfCPT;
    const unsigned e = (blockIdx.x * kThreads + threadIdx.x) * EPT;
    if (e >= NK) return;
    int n[EPT], nb[EPT];
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
        n[u] = (int)dk.div(e + u);
        long long v = idx[(size_t)b * NK + e + u];
        nb[u] = (int)(v < 0 ? 0 : (v > N - 1 ? N - 1 : v));
    }
#pragma unroll
    for (int cc = 0; cc < kGfCPT; ++cc) {
        const int c = c0 + cc;
        if (c < D) {
            const float *row = x + ((size_t)b * D + c) * N;
            float ctr[EPT], dif[EPT];
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                ctr[u] = __ldg(row + n[u]);
                dif[u] = __fsub_rn(__ldg(row + nb[u]), ctr[u]);
            }
            float *o1 = out + ((size_t)b * 2 * D + c) * NK + e;
            float *o2 = out + ((size_t)b * 2 * D + D + c) * NK + e;
            if (EPT == 4) {
                st_stream_f4(reinterpret_cast<float4 *>(o1), make_float4(dif[0], dif[1], dif[2], dif[3]));
                st_stream_f4(reinterpret_cast<float4 *>(o2), make_float4(ctr[0], ctr[1], ctr[2], ctr[3]));
            } else {
                st_stream_f1(o1, dif[0]);
                st_stream_f1(o2, ctr[0]);
            }
        }
    }
}

// Variant with the CPT channel rows of x[b] staged in shared memory (CPT * N floats): the neighbour
// gathers then are LDS with a few-way bank conflict instead of 32 scattered L1 sectors per warp
// instruction, which is what held the direct-gather kernel at ~46 % of the HBM peak (L1 wavefronts, not
// DRAM).  grid = (edge splits, channel groups, clouds); a thread owns 4 consecutive (n, j) edges.
constexpr int kGfsCPT = 4;
constexpr int kGfsThreads = 512;
__global__ void __launch_bounds__(kGfsThreads)
graph_feature_smem_kernel(const float *__restrict__ x, const int64_t *__restrict__ idx, int D, int N, FastDiv dk,
                          unsigned NK, unsigned edges_per_split, float *__restrict__ out)
{
    extern __shared__ __align__(16) float s_rows[];    // [kGfsCPT][N]
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * kGfsCPT;
    const int nch = D - c0 < kGfsCPT ? D - c0 : kGfsCPT;
    const float *src = x + ((size_t)b * D + c0) * N;
    for (int i = threadIdx.x; i < nch * N; i += kGfsThreads) s_rows[i] = __ldg(src + i);
    __syncthreads();
    const unsigned e_beg = blockIdx.x * edges_per_split;
    const unsigned e_end = e_beg + edges_per_split < NK ? e_beg + edges_per_split : NK;
    const int64_t *ib = idx + (size_t)b * NK;
    for (unsigned e = e_beg + threadIdx.x * 4; e < e_end; e += kGfsThreads * 4) {
        int n[4], nb[4];
        const longlong2 i01 = *reinterpret_cast<const longlong2 *>(ib + e);
        const longlong2 i23 = *reinterpret_cast<const longlong2 *>(ib + e + 2);
        const long long raw[4] = {i01.x, i01.y, i23.x, i23.y};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            n[u] = (int)dk.div(e + u);
            nb[u] = (int)(raw[u] < 0 ? 0 : (raw[u] > N - 1 ? N - 1 : raw[u]));
        }
#pragma unroll
        for (int cc = 0; cc < kGfsCPT; ++cc) {
            if (cc < nch) {
                const float *row = s_rows + cc * N;
                float ctr[4], dif[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ctr[u] = row[n[u]];
                    dif[u] = __fsub_rn(row[nb[u]], ctr[u]);
                }
                st_stream_f4(reinterpret_cast<float4 *>(out + ((size_t)b * 2 * D + c0 + cc) * NK + e),
                             make_float4(dif[0], dif[1], dif[2], dif[3]));
                st_stream_f4(reinterpret_cast<float4 *>(out + ((size_t)b * 2 * D + D + c0 + cc) * NK + e),
                             make_float4(ctr[0], ctr[1], ctr[2], ctr[3]));
            }
        }
    }
}

// grad_x[b,c,n] += sum_j g2[b,c,n,j] - sum_j g1[b,c,n,j];  grad_x[b,c,idx[n,j]] += g1[b,c,n,j]
__global__ void __launch_bounds__(kThreads)
graph_feature_bwd_kernel(const float *__restrict__ gout, const int64_t *__restrict__ idx, int D, int N, FastDiv dk,
                         unsigned NK, float *__restrict__ gx)
{
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * kGfCPT;
    const unsigned e = blockIdx.x * kThreads + threadIdx.x;
    if (e >= NK) return;
    const int n = (int)dk.div(e);
    long long nb = idx[(size_t)b * NK + e];
    nb = nb < 0 ? 0 : (nb > N - 1 ? N - 1 : nb);
#pragma unroll
    for (int cc = 0; cc < kGfCPT; ++cc) {
        const int c = c0 + cc;
        if (c < D) {
            const float g1 = __ldg(gout + ((size_t)b * 2 * D + c) * NK + e);
            const float g2 = __ldg(gout + ((size_t)b * 2 * D + D + c) * NK + e);
            float *row = gx + ((size_t)b * D + c) * N;
            atomicAdd(row + nb, g1);
            atomicAdd(row + n, g2 - g1);
        }
    }
}

// ------------------------------------------------------------------------------------------
// interpolation: out[b,n,:] = sum_j w[b,n,j] * points2[b, idx[b,n,j], :]   (products rounded,
// summed in j order -- pointnet_util.py:334)
// ------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(kThreads)
interp_rows_kernel(const float *__restrict__ p2, const int64_t *__restrict__ idx, const float *__restrict__ w,
                   int S, int D, int k, FastDiv dDV, FastDiv dN, unsigned total, float *__restrict__ out)
{
    const unsigned t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    const unsigned rowg = dDV.div(t);                  // b*N + n
    const int c = (int)(t - rowg * dDV.d) * VEC;
    const unsigned b = dN.div(rowg);
    float acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
    for (int j = 0; j < k; ++j) {
        long long i = idx[(size_t)rowg * k + j];
        i = i < 0 ? 0 : (i > S - 1 ? S - 1 : i);
        const float wj = __ldg(w + (size_t)rowg * k + j);
        const float *src = p2 + ((size_t)b * S + i) * D + c;
        if (VEC == 4) {
            float4 f = __ldg(reinterpret_cast<const float4 *>(src));
            float pr[4] = {__fmul_rn(f.x, wj), __fmul_rn(f.y, wj), __fmul_rn(f.z, wj), __fmul_rn(f.w, wj)};
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[v] = j == 0 ? pr[v] : __fadd_rn(acc[v], pr[v]);
        } else {
            float pr = __fmul_rn(__ldg(src), wj);
            acc[0] = j == 0 ? pr : __fadd_rn(acc[0], pr);
        }
    }
    if (VEC == 4)
        st_stream_f4(reinterpret_cast<float4 *>(out + (size_t)rowg * D + c),
                     make_float4(acc[0], acc[1], acc[2], acc[3]));
    else
        st_stream_f1(out + (size_t)rowg * D + c, acc[0]);
}

// channels-first: p2 [B,D,S] -> out [B,D,N]; a thread owns one n and walks CPT channels
constexpr int kIpCPT = 8;
constexpr int kIpMaxK = 8;

__global__ void __launch_bounds__(kThreads)
interp_cf_kernel(const float *__restrict__ p2, const int64_t *__restrict__ idx, const float *__restrict__ w,
                 int N, int S, int D, int k, float *__restrict__ out)
{
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * kIpCPT;
    const int n = blockIdx.x * kThreads + threadIdx.x;
    if (n >= N) return;
    int id[kIpMaxK];
    float wt[kIpMaxK];
#pragma unroll
    for (int j = 0; j < kIpMaxK; ++j) {
        if (j < k) {
            long long i = idx[((size_t)b * N + n) * k + j];
            id[j] = (int)(i < 0 ? 0 : (i > S - 1 ? S - 1 : i));
            wt[j] = __ldg(w + ((size_t)b * N + n) * k + j);
        }
    }
#pragma unroll
    for (int cc = 0; cc < kIpCPT; ++cc) {
        const int c = c0 + cc;
        if (c < D) {
            const float *row = p2 + ((size_t)b * D + c) * S;
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < kIpMaxK; ++j) {
                if (j < k) {
                    float pr = __fmul_rn(__ldg(row + id[j]), wt[j]);
                    acc = j == 0 ? pr : __fadd_rn(acc, pr);
                }
            }
            st_stream_f1(out + ((size_t)b * D + c) * N + n, acc);
        }
    }
}

// backward, one thread per element of gout in its own layout (coalesced read), k atomics each.
// cf == 0: gout [B,N,D] (inner = D, mid = N); cf != 0: gout [B,D,N] (inner = N, mid = D).
__global__ void __launch_bounds__(kThreads)
interp_bwd_kernel(const float *__restrict__ gout, const int64_t *__restrict__ idx, const float *__restrict__ w,
                  int N, int S, int D, int k, int cf, FastDiv dInner, FastDiv dMid, unsigned total,
                  float *__restrict__ gp2)
{
    const unsigned t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    const unsigned outer = dInner.div(t);
    const unsigned inner = t - outer * dInner.d;
    const unsigned b = dMid.div(outer);
    const unsigned mid = outer - b * dMid.d;
    const unsigned n = cf ? inner : mid;
    const unsigned c = cf ? mid : inner;
    const float g = __ldg(gout + t);
    for (int j = 0; j < k; ++j) {
        long long i = idx[((size_t)b * N + n) * k + j];
        i = i < 0 ? 0 : (i > S - 1 ? S - 1 : i);
        const float wj = __ldg(w + ((size_t)b * N + n) * k + j);
        float *dst = cf ? gp2 + ((size_t)b * D + c) * S + i : gp2 + ((size_t)b * S + i) * D + c;
        atomicAdd(dst, g * wj);
    }
}

// ------------------------------------------------------------------------------------------
// Input rows of a feature-propagation MLP in one pass (training under bf16 autocast):
//   out[b,n,:] = [ points1[b,n,:D1] | sum_j w[b,n,j] * points2[b, idx[b,n,j], :D2] | 0 ... ]  as bf16,
// i.e. interpolate + concat + cast (pointnet_util.py:325-340) without the fp32 intermediate, the
// torch.cat and the two dtype copies.  One thread per PAIR of output channels (D1, D2, pitch even).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 ld_pair(const float *p) { return __ldg(reinterpret_cast<const float2 *>(p)); }
__device__ __forceinline__ float2 ld_pair(const __nv_bfloat16 *p)
{
    return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(p));
}

template <typename T1, typename T2>
__global__ void __launch_bounds__(kThreads)
fp_concat_kernel(const T1 *__restrict__ p1, const T2 *__restrict__ p2, const int64_t *__restrict__ idx,
                 const float *__restrict__ w, int S, int D1, int D2, int k, FastDiv dHalfPitch, FastDiv dN,
                 unsigned total, __nv_bfloat16 *__restrict__ out)
{
    const unsigned t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    const unsigned rowg = dHalfPitch.div(t);           // b*N + n
    const int c = (int)(t - rowg * dHalfPitch.d) * 2;
    float2 v = make_float2(0.f, 0.f);
    if (c < D1) {
        v = ld_pair(p1 + (size_t)rowg * D1 + c);
    } else if (c < D1 + D2) {
        const unsigned b = dN.div(rowg);
        const int c2 = c - D1;
        for (int j = 0; j < k; ++j) {
            long long i = idx[(size_t)rowg * k + j];
            i = i < 0 ? 0 : (i > S - 1 ? S - 1 : i);
            const float wj = __ldg(w + (size_t)rowg * k + j);
            const float2 f = ld_pair(p2 + ((size_t)b * S + i) * D2 + c2);
            const float px = __fmul_rn(f.x, wj), py = __fmul_rn(f.y, wj);
            v.x = j == 0 ? px : __fadd_rn(v.x, px);
            v.y = j == 0 ? py : __fadd_rn(v.y, py);
        }
    }
    *reinterpret_cast<__nv_bfloat162 *>(out + (size_t)t * 2) = __floats2bfloat162_rn(v.x, v.y);
}

// backward w.r.t. points2: gp2[b, idx, c] += w * gout[b, n, D1 + c]  (gout rows have `pitch` elements)
__global__ void __launch_bounds__(kThreads)
fp_concat_bwd_kernel(const __nv_bfloat16 *__restrict__ gout, const int64_t *__restrict__ idx,
                     const float *__restrict__ w, int S, int D1, int D2, int k, int pitch, FastDiv dD2, FastDiv dN,
                     unsigned total, float *__restrict__ gp2)
{
    const unsigned t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    const unsigned rowg = dD2.div(t);
    const int c = (int)(t - rowg * dD2.d);
    const unsigned b = dN.div(rowg);
    const float g = __bfloat162float(gout[(size_t)rowg * pitch + D1 + c]);
    for (int j = 0; j < k; ++j) {
        long long i = idx[(size_t)rowg * k + j];
        i = i < 0 ? 0 : (i > S - 1 ? S - 1 : i);
        atomicAdd(gp2 + ((size_t)b * S + i) * D2 + c, g * __ldg(w + (size_t)rowg * k + j));
    }
}

// ---- 8 channels per thread (D1, D2, pitch multiples of 8): 16-byte loads and stores ----
__device__ __forceinline__ void ld8(const float *p, float v[8])
{
    const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
}
__device__ __forceinline__ void ld8(const __nv_bfloat16 *p, float v[8])
{
    const uint4 u = *reinterpret_cast<const uint4 *>(p);
    const unsigned w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&w[j]));
        v[2 * j] = f.x, v[2 * j + 1] = f.y;
    }
}

template <typename T1, typename T2>
__global__ void __launch_bounds__(kThreads)
fp_concat_chunk_kernel(const T1 *__restrict__ p1, const T2 *__restrict__ p2, const int64_t *__restrict__ idx,
                       const float *__restrict__ w, int S, int D1, int D2, int k, FastDiv dQ, FastDiv dN,
                       unsigned total, __nv_bfloat16 *__restrict__ out)
{
    pdl_wait();
    pdl_trigger();
    const unsigned t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    const unsigned rowg = dQ.div(t);                   // b*N + n
    const int c = (int)(t - rowg * dQ.d) * 8;
    float v[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = 0.f;
    if (c < D1) {
        ld8(p1 + (size_t)rowg * D1 + c, v);
    } else if (c < D1 + D2) {
        const unsigned b = dN.div(rowg);
        const int c2 = c - D1;
        for (int j = 0; j < k; ++j) {
            long long i = idx[(size_t)rowg * k + j];
            i = i < 0 ? 0 : (i > S - 1 ? S - 1 : i);
            const float wj = __ldg(w + (size_t)rowg * k + j);
            float f[8];
            ld8(p2 + ((size_t)b * S + i) * D2 + c2, f);
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                const float pr = __fmul_rn(f[m], wj);
                v[m] = j == 0 ? pr : __fadd_rn(v[m], pr);
            }
        }
    }
    *reinterpret_cast<uint4 *>(out + (size_t)t * 8) = pack8_bf16(v);
}

// backward, 4 channels per thread (D1, D2, pitch multiples of 4): k 128-bit reductions per thread
__global__ void __launch_bounds__(kThreads)
fp_concat_bwd_vec_kernel(const __nv_bfloat16 *__restrict__ gout, const int64_t *__restrict__ idx,
                         const float *__restrict__ w, int S, int D1, int D2, int k, int pitch, FastDiv dDV, FastDiv dN,
                         unsigned total, float *__restrict__ gp2)
{
    pdl_wait();
    pdl_trigger();
    const unsigned t = blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    const unsigned rowg = dDV.div(t);
    const int c = (int)(t - rowg * dDV.d) * 4;
    const unsigned b = dN.div(rowg);
    const float4 g = load_grad4(gout + (size_t)rowg * pitch + D1 + c);
    for (int j = 0; j < k; ++j) {
        long long i = idx[(size_t)rowg * k + j];
        i = i < 0 ? 0 : (i > S - 1 ? S - 1 : i);
        const float wj = __ldg(w + (size_t)rowg * k + j);
        atomicAdd(reinterpret_cast<float4 *>(gp2 + ((size_t)b * S + i) * D2 + c),
                  make_float4(g.x * wj, g.y * wj, g.z * wj, g.w * wj));
    }
}

// ------------------------------------------------------------------------------------------
// square_distance (materialising; kept for API completeness -- the product path never needs it)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
row_sumsq_kernel(const float *__restrict__ x, int64_t rows, int C, int64_t row_stride, int elem_stride,
                 int64_t rows_per_batch, int64_t batch_stride, float *__restrict__ out)
{
    int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (r >= rows) return;
    int64_t b = r / rows_per_batch, i = r - b * rows_per_batch;
    out[r] = row_sumsq_aten(x + b * batch_stride + i * row_stride, C, elem_stride);
}

__global__ void __launch_bounds__(kThreads)
square_distance_kernel(const float *__restrict__ src, const float *__restrict__ dst, int N, int M, int C,
                       float *__restrict__ out)
{
    extern __shared__ float s_src[];                   // one src row + its norm
    const int b = blockIdx.y, n = blockIdx.x;
    const float *s = src + ((size_t)b * N + n) * C;
    for (int c = threadIdx.x; c < C; c += kThreads) s_src[c] = __ldg(s + c);
    __syncthreads();
    if (threadIdx.x == 0) s_src[C] = row_sumsq_aten(s_src, C, 1);
    __syncthreads();
    const float sn = s_src[C];
    for (int m = threadIdx.x; m < M; m += kThreads) {
        const float *d = dst + ((size_t)b * M + m) * C;
        float acc = 0.f;
        for (int c = 0; c < C; ++c) acc = __fmaf_rn(s_src[c], __ldg(d + c), acc);
        float dn = row_sumsq_aten(d, C, 1);
        float t = __fmul_rn(-2.0f, acc);
        t = __fadd_rn(t, sn);
        out[((size_t)b * N + n) * M + m] = __fadd_rn(t, dn);
    }
}

static inline unsigned blocks_for(int64_t total) { return (unsigned)ceil_div(total, kThreads); }

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Largest number of clouds per launch such that per-launch element counts stay below 2^31
// (32-bit index arithmetic in the kernels); 0 if a single cloud is already too large.
static inline int clouds_per_launch(int B, int64_t elems_per_cloud)
{
    const int64_t lim = (1ll << 31) - 1;
    if (elems_per_cloud <= 0 || elems_per_cloud > lim) return 0;
    int64_t n = lim / elems_per_cloud;
    return (int)(n < B ? n : B);
}

int row_sumsq_launch(const float *x, int64_t rows, int C, int64_t row_stride, int elem_stride,
                     int64_t rows_per_batch, int64_t batch_stride, float *out, cudaStream_t st)
{
    row_sumsq_kernel<<<(unsigned)ceil_div(rows, kThreads), kThreads, 0, st>>>(x, rows, C, row_stride, elem_stride,
                                                                            rows_per_batch, batch_stride, out);
    PCB_RETURN_LAUNCH_STATUS();
}

}  // namespace pcb

using namespace pcb;

PCB_API int pcb_gather_f32(const float *points, const int64_t *idx, int B, int N, int C, int64_t M,
                           int clamp, float *out, int *err_count, pcb_stream_t stream)
{
    PCB_REQUIRE(points && idx && out, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && C > 0 && M > 0, PCB_EINVAL);
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = C % 4 == 0 && aligned16(points) && aligned16(out);
    const int CV = vec ? C / 4 : C;
    const int step = clouds_per_launch(B, M * (int64_t)C);
    PCB_REQUIRE(step > 0, PCB_ERANGE);
    for (int b0 = 0; b0 < B; b0 += step) {
        const int nb = B - b0 < step ? B - b0 : step;
        const unsigned total = (unsigned)((int64_t)nb * M * CV);
        const float *p = points + (size_t)b0 * N * C;
        const int64_t *ix = idx + (size_t)b0 * M;
        float *o = out + (size_t)b0 * M * C;
        if (vec && (int64_t)nb * M >= 4096) {
            const unsigned rows = (unsigned)((int64_t)nb * M);
            const unsigned RG = (rows + kGatherU - 1) / kGatherU;
            gather_vec_ilp_kernel<<<blocks_for((int64_t)RG * CV), kThreads, 0, st>>>(
                p, ix, N, C, make_fastdiv(CV), make_fastdiv((unsigned)M), RG, rows, clamp, o, err_count);
        } else if (vec)
            gather_kernel<4><<<blocks_for(total), kThreads, 0, st>>>(p, ix, N, C, make_fastdiv(CV), make_fastdiv((unsigned)M),
                                                                    total, clamp, o, err_count);
        else
            gather_kernel<1><<<blocks_for(total), kThreads, 0, st>>>(p, ix, N, C, make_fastdiv(CV), make_fastdiv((unsigned)M),
                                                                    total, clamp, o, err_count);
    }
    PCB_RETURN_LAUNCH_STATUS();
}

PCB_API int pcb_gather_bwd_f32(const float *grad_out, const int64_t *idx, int B, int N, int C, int64_t M,
                               int clamp, float *grad_points, pcb_stream_t stream)
{
    PCB_REQUIRE(grad_out && idx && grad_points, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && C > 0 && M > 0, PCB_EINVAL);
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = C % 4 == 0 && aligned16(grad_out) && aligned16(grad_points);
    const int CV = vec ? C / 4 : C;
    const int step = clouds_per_launch(B, M * (int64_t)C);
    PCB_REQUIRE(step > 0, PCB_ERANGE);
    for (int b0 = 0; b0 < B; b0 += step) {
        const int nb = B - b0 < step ? B - b0 : step;
        const unsigned total = (unsigned)((int64_t)nb * M * CV);
        const float *g = grad_out + (size_t)b0 * M * C;
        const int64_t *ix = idx + (size_t)b0 * M;
        float *gp = grad_points + (size_t)b0 * N * C;
        if (vec)
            gather_bwd_kernel<4><<<blocks_for(total), kThreads, 0, st>>>(g, ix, N, C, make_fastdiv(CV),
                                                                        make_fastdiv((unsigned)M), total, clamp, gp);
        else
            gather_bwd_kernel<1><<<blocks_for(total), kThreads, 0, st>>>(g, ix, N, C, make_fastdiv(CV),
                                                                        make_fastdiv((unsigned)M), total, clamp, gp);
    }
    PCB_RETURN_LAUNCH_STATUS();
}

template <typename OT, typename PT>
static int group_points_launch(const float *xyz, const PT *points, const float *new_xyz, const int64_t *idx, int B,
                               int N, int S, int K, int D, int xyz_first, int points_cf, int clamp, int pitch, OT *out,
                               cudaStream_t st)
{
    PCB_REQUIRE(xyz && new_xyz && idx && out, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && S > 0 && K > 0 && D >= 0, PCB_EINVAL);
    PCB_REQUIRE(points || D == 0, PCB_EINVAL);
    PCB_REQUIRE(pitch == 0 || pitch >= 3 + D, PCB_EINVAL);
    const int C = pitch ? pitch : 3 + D;
    const int step = clouds_per_launch(B, (int64_t)S * K * C);
    PCB_REQUIRE(step > 0, PCB_ERANGE);
    for (int b0 = 0; b0 < B; b0 += step) {
        const int nb = B - b0 < step ? B - b0 : step;
        const unsigned total = (unsigned)((int64_t)nb * S * K * C);
        const PT *pts = points ? points + (size_t)b0 * N * D : nullptr;
        OT *o = out + (size_t)b0 * S * K * C;
        const unsigned nrows = (unsigned)((int64_t)nb * S * K);
        const int TR = nrows >= 128u * 2 * PCB_NUM_SMS ? 128 : (nrows >= 64u * 2 * PCB_NUM_SMS ? 64 : 32);
        const unsigned tiles = (nrows + TR - 1) / TR;
        const bool tile_ok = !points_cf && (int64_t)nb * N < (1ll << 31);
        if (tile_ok && pts && C % 8 == 0 && aligned16(o) && D % 8 == 0 && !xyz_first && C == D + 8 && aligned16(pts)) {
            launch_pdl(group_tile_kernel<OT, PT, true>, dim3(tiles), dim3(kThreads), 0, st, 
                xyz + (size_t)b0 * N * 3, pts, new_xyz + (size_t)b0 * S * 3, idx + (size_t)b0 * S * K, N, D, C,
                make_fastdiv(K), make_fastdiv(S), make_fastdiv(D / 8), xyz_first, clamp, TR, nrows, o);
            continue;
        }
        if (tile_ok && !(C % 8 == 0 && aligned16(o))) { // any layout / pitch, one element per item
            launch_pdl(group_tile_kernel<OT, PT, false>, dim3(tiles), dim3(kThreads), 0, st, 
                xyz + (size_t)b0 * N * 3, pts, new_xyz + (size_t)b0 * S * 3, idx + (size_t)b0 * S * K, N, D, C,
                make_fastdiv(K), make_fastdiv(S), make_fastdiv(C), xyz_first, clamp, TR, nrows, o);
            continue;
        }
        if (C % 8 == 0 && aligned16(o)) {               // padded rows: 8 channels per thread
            const int vec_ok = pts && !points_cf && D % 8 == 0 && aligned16(pts);
            const unsigned rows = (unsigned)((int64_t)nb * S * K);
            const unsigned RG = (rows + kGroupU - 1) / kGroupU;
            launch_pdl(group_points_chunk_kernel<OT, PT>, dim3(blocks_for((int64_t)RG * (C / 8))), dim3(kThreads), 0, st, 
                xyz + (size_t)b0 * N * 3, pts, new_xyz + (size_t)b0 * S * 3, idx + (size_t)b0 * S * K, N, D,
                make_fastdiv(C / 8), make_fastdiv(K), make_fastdiv(S), xyz_first, points_cf, clamp, vec_ok, RG, rows, o);
            continue;
        }
        if (sizeof(PT) != 4) return PCB_ERANGE;         // bf16 feature rows need the padded (pitch % 8 == 0) layout
        group_points_kernel<OT><<<blocks_for(total), kThreads, 0, st>>>(
            xyz + (size_t)b0 * N * 3, reinterpret_cast<const float *>(pts), new_xyz + (size_t)b0 * S * 3,
            idx + (size_t)b0 * S * K, N, D, make_fastdiv(C), make_fastdiv(K), make_fastdiv(S), xyz_first, points_cf,
            clamp, total, o);
    }
    PCB_RETURN_LAUNCH_STATUS();
}

template <typename GT>
static int group_points_bwd_launch(const GT *grad_out, const int64_t *idx, int B, int N, int S, int K, int D,
                                   int xyz_first, int points_cf, int clamp, int pitch, float *grad_points,
                                   cudaStream_t st)
{
    PCB_REQUIRE(grad_out && idx && grad_points, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && S > 0 && K > 0 && D > 0, PCB_EINVAL);
    PCB_REQUIRE(pitch == 0 || pitch >= 3 + D, PCB_EINVAL);
    const int C = pitch ? pitch : 3 + D;
    const int step = clouds_per_launch(B, (int64_t)S * K * C);
    PCB_REQUIRE(step > 0, PCB_ERANGE);
    for (int b0 = 0; b0 < B; b0 += step) {
        const int nb = B - b0 < step ? B - b0 : step;
        const unsigned total = (unsigned)((int64_t)nb * S * K * D);
        const GT *go = grad_out + (size_t)b0 * S * K * C;
        float *gp = grad_points + (size_t)b0 * N * D;
        if (D % 4 == 0 && !points_cf && aligned16(gp)) {
            const int foff = xyz_first ? 3 : 0;
            const int al = (sizeof(GT) == 4 ? (C % 4 == 0 && foff % 4 == 0 && aligned16(go))
                                            : (C % 4 == 0 && foff % 4 == 0 && (reinterpret_cast<uintptr_t>(go) & 7) == 0));
            launch_pdl(group_points_bwd_vec_kernel<GT>, dim3(blocks_for(total / 4)), dim3(kThreads), 0, st, 
                go, idx + (size_t)b0 * S * K, N, D, C, make_fastdiv(D / 4), make_fastdiv((unsigned)(S * K)), foff, clamp,
                al, total / 4, gp);
            continue;
        }
        group_points_bwd_kernel<GT><<<blocks_for(total), kThreads, 0, st>>>(
            grad_out + (size_t)b0 * S * K * C, idx + (size_t)b0 * S * K, N, D, C, make_fastdiv(D),
            make_fastdiv((unsigned)(S * K)), xyz_first, points_cf, clamp, total, grad_points + (size_t)b0 * N * D);
    }
    PCB_RETURN_LAUNCH_STATUS();
}

PCB_API int pcb_group_points_f32(const float *xyz, const float *points, const float *new_xyz,
                                 const int64_t *idx, int B, int N, int S, int K, int D, int xyz_first,
                                 int points_cf, int clamp, int pitch, float *out, pcb_stream_t stream)
{
    return group_points_launch<float, float>(xyz, points, new_xyz, idx, B, N, S, K, D, xyz_first, points_cf, clamp, pitch,
                                             out, (cudaStream_t)stream);
}

PCB_API int pcb_group_points_bf16(const float *xyz, const void *points, int points_bf16, const float *new_xyz,
                                  const int64_t *idx, int B, int N, int S, int K, int D, int xyz_first,
                                  int points_cf, int clamp, int pitch, void *out, pcb_stream_t stream)
{
    if (points_bf16)
        return group_points_launch<__nv_bfloat16, __nv_bfloat16>(xyz, (const __nv_bfloat16 *)points, new_xyz, idx, B, N, S,
                                                                 K, D, xyz_first, points_cf, clamp, pitch,
                                                                 (__nv_bfloat16 *)out, (cudaStream_t)stream);
    return group_points_launch<__nv_bfloat16, float>(xyz, (const float *)points, new_xyz, idx, B, N, S, K, D, xyz_first,
                                                     points_cf, clamp, pitch, (__nv_bfloat16 *)out, (cudaStream_t)stream);
}

PCB_API int pcb_group_points_bwd_f32(const float *grad_out, const int64_t *idx, int B, int N, int S, int K,
                                     int D, int xyz_first, int points_cf, int clamp, int pitch, float *grad_points,
                                     pcb_stream_t stream)
{
    return group_points_bwd_launch<float>(grad_out, idx, B, N, S, K, D, xyz_first, points_cf, clamp, pitch, grad_points,
                                          (cudaStream_t)stream);
}

PCB_API int pcb_group_points_bwd_bf16(const void *grad_out, const int64_t *idx, int B, int N, int S, int K,
                                      int D, int xyz_first, int points_cf, int clamp, int pitch, float *grad_points,
                                      pcb_stream_t stream)
{
    return group_points_bwd_launch<__nv_bfloat16>((const __nv_bfloat16 *)grad_out, idx, B, N, S, K, D, xyz_first,
                                                  points_cf, clamp, pitch, grad_points, (cudaStream_t)stream);
}

PCB_API int pcb_graph_feature_f32(const float *x, const int64_t *idx, int B, int D, int N, int k, float *out,
                                  pcb_stream_t stream)
{
    PCB_REQUIRE(x && idx && out, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && D > 0 && N > 0 && k > 0, PCB_EINVAL);
    PCB_REQUIRE(B <= 65535 && ceil_div(D, kGfCPT) <= 65535 && (int64_t)N * k < (1ll << 31), PCB_ERANGE);
    const unsigned NK = (unsigned)N * (unsigned)k;
    const size_t smem = (size_t)kGfsCPT * N * sizeof(float);
    if (NK % 4 == 0 && aligned16(out) && aligned16(idx) && smem <= 96 * 1024 && NK >= 16384) {
        // edge splits: enough CTAs for ~2 waves of 3 resident CTAs per SM, each split a multiple of 4 * threads
        const unsigned groups = (unsigned)ceil_div(D, kGfsCPT);
        unsigned splits = (unsigned)ceil_div(6 * PCB_NUM_SMS, (int64_t)groups * B);
        const unsigned quantum = kGfsThreads * 4;
        unsigned eps = (unsigned)ceil_div(ceil_div(NK, splits), quantum) * quantum;
        splits = (unsigned)ceil_div(NK, eps);
        static bool attr_set[kMaxDevices] = {};
        if (cudaError_t e = smem_optin_once(graph_feature_smem_kernel, 96 * 1024, attr_set)) return (int)e;
        dim3 grid(splits, groups, (unsigned)B);
        graph_feature_smem_kernel<<<grid, kGfsThreads, smem, (cudaStream_t)stream>>>(x, idx, D, N, make_fastdiv(k), NK, eps, out);
        PCB_RETURN_LAUNCH_STATUS();
    }
    if (NK % 4 == 0 && aligned16(out)) {
        dim3 grid((unsigned)ceil_div(NK / 4, kThreads), (unsigned)ceil_div(D, kGfCPT), (unsigned)B);
        graph_feature_kernel<4><<<grid, kThreads, 0, (cudaStream_t)stream>>>(x, idx, D, N, make_fastdiv(k), NK, out);
    } else {
        dim3 grid((unsigned)ceil_div(NK, kThreads), (unsigned)ceil_div(D, kGfCPT), (unsigned)B);
        graph_feature_kernel<1><<<grid, kThreads, 0, (cudaStream_t)stream>>>(x, idx, D, N, make_fastdiv(k), NK, out);
    }
    PCB_RETURN_LAUNCH_STATUS();
}

PCB_API int pcb_graph_feature_bwd_f32(const float *grad_out, const int64_t *idx, int B, int D, int N, int k,
                                      float *grad_x, pcb_stream_t stream)
{
    PCB_REQUIRE(grad_out && idx && grad_x, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && D > 0 && N > 0 && k > 0, PCB_EINVAL);
    PCB_REQUIRE(B <= 65535 && ceil_div(D, kGfCPT) <= 65535 && (int64_t)N * k < (1ll << 31), PCB_ERANGE);
    const unsigned NK = (unsigned)N * (unsigned)k;
    dim3 grid((unsigned)ceil_div(NK, kThreads), (unsigned)ceil_div(D, kGfCPT), (unsigned)B);
    graph_feature_bwd_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(grad_out, idx, D, N, make_fastdiv(k), NK, grad_x);
    PCB_RETURN_LAUNCH_STATUS();
}

PCB_API int pcb_interpolate_f32(const float *points2, const int64_t *idx, const float *weight, int B, int N,
                                int S, int D, int k, int channels_first, float *out, pcb_stream_t stream)
{
    PCB_REQUIRE(points2 && idx && weight && out, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && S > 0 && D > 0 && k > 0, PCB_EINVAL);
    PCB_REQUIRE(k <= kIpMaxK && B <= 65535, PCB_ERANGE);
    cudaStream_t st = (cudaStream_t)stream;
    if (channels_first) {
        dim3 grid((unsigned)ceil_div(N, kThreads), (unsigned)ceil_div(D, kIpCPT), (unsigned)B);
        interp_cf_kernel<<<grid, kThreads, 0, st>>>(points2, idx, weight, N, S, D, k, out);
        PCB_RETURN_LAUNCH_STATUS();
    }
    const bool vec = D % 4 == 0 && aligned16(points2) && aligned16(out);
    const int DV = vec ? D / 4 : D;
    const int step = clouds_per_launch(B, (int64_t)N * D);
    PCB_REQUIRE(step > 0, PCB_ERANGE);
    for (int b0 = 0; b0 < B; b0 += step) {
        const int nb = B - b0 < step ? B - b0 : step;
        const unsigned total = (unsigned)((int64_t)nb * N * DV);
        const float *p = points2 + (size_t)b0 * S * D;
        const int64_t *ix = idx + (size_t)b0 * N * k;
        const float *w = weight + (size_t)b0 * N * k;
        float *o = out + (size_t)b0 * N * D;
        if (vec)
            interp_rows_kernel<4><<<blocks_for(total), kThreads, 0, st>>>(p, ix, w, S, D, k, make_fastdiv(DV),
                                                                         make_fastdiv(N), total, o);
        else
            interp_rows_kernel<1><<<blocks_for(total), kThreads, 0, st>>>(p, ix, w, S, D, k, make_fastdiv(DV),
                                                                         make_fastdiv(N), total, o);
    }
    PCB_RETURN_LAUNCH_STATUS();
}

PCB_API int pcb_interpolate_bwd_f32(const float *grad_out, const int64_t *idx, const float *weight, int B,
                                    int N, int S, int D, int k, int channels_first, float *grad_points2,
                                    pcb_stream_t stream)
{
    PCB_REQUIRE(grad_out && idx && weight && grad_points2, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && S > 0 && D > 0 && k > 0, PCB_EINVAL);
    const int step = clouds_per_launch(B, (int64_t)N * D);
    PCB_REQUIRE(step > 0, PCB_ERANGE);
    for (int b0 = 0; b0 < B; b0 += step) {
        const int nb = B - b0 < step ? B - b0 : step;
        const unsigned total = (unsigned)((int64_t)nb * N * D);
        interp_bwd_kernel<<<blocks_for(total), kThreads, 0, (cudaStream_t)stream>>>(
            grad_out + (size_t)b0 * N * D, idx + (size_t)b0 * N * k, weight + (size_t)b0 * N * k, N, S, D, k,
            channels_first, make_fastdiv(channels_first ? N : D), make_fastdiv(channels_first ? D : N), total,
            grad_points2 + (size_t)b0 * S * D);
    }
    PCB_RETURN_LAUNCH_STATUS();
}

template <typename T1, typename T2>
static int fp_concat_launch(const void *p1, const void *p2, const int64_t *idx, const float *w, int B, int N, int S,
                            int D1, int D2, int k, int pitch, void *out, cudaStream_t st)
{
    const int step = clouds_per_launch(B, (int64_t)N * pitch);
    PCB_REQUIRE(step > 0, PCB_ERANGE);
    for (int b0 = 0; b0 < B; b0 += step) {
        const int nb = B - b0 < step ? B - b0 : step;
        const unsigned total = (unsigned)((int64_t)nb * N * (pitch / 2));
        if (D1 % 8 == 0 && D2 % 8 == 0 && pitch % 8 == 0 && aligned16(out) && aligned16(p2) && (!p1 || aligned16(p1))) {
            launch_pdl(fp_concat_chunk_kernel<T1, T2>, dim3(blocks_for(total / 4)), dim3(kThreads), 0, st, 
                p1 ? (const T1 *)p1 + (size_t)b0 * N * D1 : nullptr, (const T2 *)p2 + (size_t)b0 * S * D2,
                idx + (size_t)b0 * N * k, w + (size_t)b0 * N * k, S, D1, D2, k, make_fastdiv(pitch / 8), make_fastdiv(N),
                total / 4, (__nv_bfloat16 *)out + (size_t)b0 * N * pitch);
            continue;
        }
        fp_concat_kernel<T1, T2><<<blocks_for(total), kThreads, 0, st>>>(
            p1 ? (const T1 *)p1 + (size_t)b0 * N * D1 : nullptr, (const T2 *)p2 + (size_t)b0 * S * D2,
            idx + (size_t)b0 * N * k, w + (size_t)b0 * N * k, S, D1, D2, k, make_fastdiv(pitch / 2), make_fastdiv(N),
            total, (__nv_bfloat16 *)out + (size_t)b0 * N * pitch);
    }
    PCB_RETURN_LAUNCH_STATUS();
}

PCB_API int pcb_fp_concat_bf16(const void *points1, int p1_bf16, const void *points2, int p2_bf16, const int64_t *idx,
                               const float *weight, int B, int N, int S, int D1, int D2, int k, int pitch, void *out,
                               pcb_stream_t stream)
{
    PCB_REQUIRE(points2 && idx && weight && out && (points1 || D1 == 0), PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && S > 0 && D1 >= 0 && D2 > 0 && k > 0, PCB_EINVAL);
    PCB_REQUIRE(D1 % 2 == 0 && D2 % 2 == 0 && pitch % 2 == 0 && pitch >= D1 + D2, PCB_ERANGE);
    cudaStream_t st = (cudaStream_t)stream;
    if (p1_bf16 && p2_bf16)
        return fp_concat_launch<__nv_bfloat16, __nv_bfloat16>(points1, points2, idx, weight, B, N, S, D1, D2, k, pitch, out, st);
    if (p1_bf16)
        return fp_concat_launch<__nv_bfloat16, float>(points1, points2, idx, weight, B, N, S, D1, D2, k, pitch, out, st);
    if (p2_bf16)
        return fp_concat_launch<float, __nv_bfloat16>(points1, points2, idx, weight, B, N, S, D1, D2, k, pitch, out, st);
    return fp_concat_launch<float, float>(points1, points2, idx, weight, B, N, S, D1, D2, k, pitch, out, st);
}

PCB_API int pcb_fp_concat_bwd_bf16(const void *grad_out, const int64_t *idx, const float *weight, int B, int N, int S,
                                   int D1, int D2, int k, int pitch, float *grad_points2, pcb_stream_t stream)
{
    PCB_REQUIRE(grad_out && idx && weight && grad_points2, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && S > 0 && D1 >= 0 && D2 > 0 && k > 0 && pitch >= D1 + D2, PCB_EINVAL);
    const int step = clouds_per_launch(B, (int64_t)N * D2);
    PCB_REQUIRE(step > 0, PCB_ERANGE);
    for (int b0 = 0; b0 < B; b0 += step) {
        const int nb = B - b0 < step ? B - b0 : step;
        const unsigned total = (unsigned)((int64_t)nb * N * D2);
        if (D1 % 4 == 0 && D2 % 4 == 0 && pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(grad_out) & 7) == 0 &&
            aligned16(grad_points2)) {
            launch_pdl(fp_concat_bwd_vec_kernel, dim3(blocks_for(total / 4)), dim3(kThreads), 0, (cudaStream_t)stream, 
                (const __nv_bfloat16 *)grad_out + (size_t)b0 * N * pitch, idx + (size_t)b0 * N * k,
                weight + (size_t)b0 * N * k, S, D1, D2, k, pitch, make_fastdiv(D2 / 4), make_fastdiv(N), total / 4,
                grad_points2 + (size_t)b0 * S * D2);
            continue;
        }
        fp_concat_bwd_kernel<<<blocks_for(total), kThreads, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16 *)grad_out + (size_t)b0 * N * pitch, idx + (size_t)b0 * N * k,
            weight + (size_t)b0 * N * k, S, D1, D2, k, pitch, make_fastdiv(D2), make_fastdiv(N), total,
            grad_points2 + (size_t)b0 * S * D2);
    }
    PCB_RETURN_LAUNCH_STATUS();
}

PCB_API int pcb_square_distance_f32(const float *src, const float *dst, int B, int N, int M, int C, float *out,
                                    pcb_stream_t stream)
{
    PCB_REQUIRE(src && dst && out, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && M > 0 && C > 0, PCB_EINVAL);
    PCB_REQUIRE(C <= 512 && B <= 65535, PCB_ERANGE);
    dim3 grid((unsigned)N, (unsigned)B);
    square_distance_kernel<<<grid, kThreads, (C + 1) * sizeof(float), (cudaStream_t)stream>>>(src, dst, N, M, C, out);
    PCB_RETURN_LAUNCH_STATUS();
}
