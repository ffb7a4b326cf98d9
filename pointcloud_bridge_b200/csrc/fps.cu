// Farthest point sampling -- replaces the npoint-iteration Python loop of
// Partsize-identical/models/pointnet_util.py:66-88 (and its copy
// Highway_bridge/models/pointnet2_utils.py:63-80).
//
// One persistent CTA per cloud.  Every thread keeps the coordinates and the running
// min-distance of its PPT points in registers for the whole call; the cloud is also kept in
// shared memory (SoA) so that the coordinates of the newly selected point are one broadcast
// LDS away.  Each of the npoint serial steps is: distance update (exact fp32: three rounded
// squares, two rounded adds, no FMA), per-thread running argmax, a two-instruction warp
// argmax (REDUX max on the distance bit pattern -- distances are >= +0 so unsigned order ==
// float order -- then REDUX min on the index among the lanes that hold the max, which is the
// reference's first-index tie-break), one __syncthreads, and the same pair of REDUX on the
// per-warp partials, done redundantly by every warp so no second barrier is needed.
//
// Algorithmic traffic: 12*N bytes in + 8*npoint bytes out per cloud; the kernel is bound by
// the latency of the serial chain, not by HBM.
#include "pcb_common.cuh"

namespace pcb {

constexpr unsigned kBigIdx = 0x7fffffffu;

template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS, 1)
fps_reg_kernel(const float *__restrict__ xyz, int N, const int64_t *__restrict__ start, int npoint,
               int64_t *__restrict__ out)
{
    constexpr int NWARPS = THREADS / 32;
    extern __shared__ float s_xyz[];                 // sx[NP] | sy[NP] | sz[NP]
    __shared__ uint2 s_red[2][32];
    const int NP = THREADS * PPT;
    float *sx = s_xyz, *sy = s_xyz + NP, *sz = s_xyz + 2 * NP;

    const int b = blockIdx.x;
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const float *p = xyz + (size_t)b * N * 3;

    float px[PPT], py[PPT], pz[PPT], pd[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        int i = j * THREADS + t;
        bool ok = i < N;
        px[j] = ok ? __ldg(p + 3 * i + 0) : 0.0f;
        py[j] = ok ? __ldg(p + 3 * i + 1) : 0.0f;
        pz[j] = ok ? __ldg(p + 3 * i + 2) : 0.0f;
        pd[j] = ok ? 1e10f : 0.0f;                   // padding can never beat a real point
        sx[i] = px[j];
        sy[i] = py[j];
        sz[i] = pz[j];
    }
    long long s0 = start[b];
    unsigned far = (unsigned)(s0 < 0 ? 0 : (s0 >= N ? N - 1 : s0));
    __syncthreads();

    int64_t *o = out + (size_t)b * npoint;
    int buf = 0;
    for (int it = 0; it < npoint; ++it) {
        if (t == 0) o[it] = (int64_t)far;
        const float cx = sx[far], cy = sy[far], cz = sz[far];
        unsigned best = 0u, besti = kBigIdx;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            float dx = __fsub_rn(px[j], cx);
            float dy = __fsub_rn(py[j], cy);
            float dz = __fsub_rn(pz[j], cz);
            float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            bool valid = (j * THREADS + t) < N;
            float nd = (d < pd[j]) ? d : pd[j];      // distance[mask] = dist[mask], mask = dist < distance
            pd[j] = valid ? nd : 0.0f;
            unsigned bits = __float_as_uint(pd[j]);
            if (bits > best || besti == kBigIdx) {   // strict > keeps the lowest index (j ascending)
                best = bits;
                besti = (unsigned)(j * THREADS + t);
            }
        }
        // warp argmax, lowest index among equals
        unsigned wmax = __reduce_max_sync(PCB_FULL_MASK, best);
        unsigned widx = __reduce_min_sync(PCB_FULL_MASK, best == wmax ? besti : kBigIdx);
        if (NWARPS == 1) {
            far = widx;
        } else {
            if (lane == 0) s_red[buf][warp] = make_uint2(wmax, widx);
            __syncthreads();
            uint2 v = lane < NWARPS ? s_red[buf][lane] : make_uint2(0u, kBigIdx);
            unsigned m = __reduce_max_sync(PCB_FULL_MASK, v.x);
            far = __reduce_min_sync(PCB_FULL_MASK, v.x == m ? v.y : kBigIdx);
            buf ^= 1;
        }
    }
}

// ---- packed-pair variant (PPT even): the distance arithmetic on Blackwell's two-wide fp32 instructions ----
// sub/mul.rn.f32x2 (FADD2 / FMUL2 in SASS) round each half exactly like the scalar instruction, so
// the results are bit-identical, but a pair of points costs 10 issue slots for its 16 flops instead of
// 16 (three packed subtractions, three packed squares, four scalar adds).  The running argmax is split into a max over the thread's distances (FMNMX) and one equality scan
// for the lowest slot that holds it: 4 integer-pipe instructions per point instead of 6.  At
// N = 4096 the kernel is bound by instruction issue on its one SM (16 warps x ~140 instructions per
// step); this is the inner loop that sets the step time.
__device__ __forceinline__ unsigned long long pk2(float lo, float hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(unsigned long long v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b)
{
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS, 1)
fps_pair_kernel(const float *__restrict__ xyz, int N, const int64_t *__restrict__ start, int npoint,
                int64_t *__restrict__ out)
{
    static_assert(PPT % 2 == 0, "pairs of points");
    constexpr int NWARPS = THREADS / 32;
    constexpr int NP2 = PPT / 2;
    extern __shared__ float s_xyz[];                 // sx[NP] | sy[NP] | sz[NP]
    __shared__ uint2 s_red[2][32];
    const int NP = THREADS * PPT;
    float *sx = s_xyz, *sy = s_xyz + NP, *sz = s_xyz + 2 * NP;

    const int b = blockIdx.x;
    const int t = threadIdx.x;
    const int lane = t & 31, warp = t >> 5;
    const float *p = xyz + (size_t)b * N * 3;

    unsigned long long px[NP2], py[NP2], pz[NP2];    // slots (2q, 2q+1) = points (2q*THREADS + t, (2q+1)*THREADS + t)
    float pd[PPT];
#pragma unroll
    for (int q = 0; q < NP2; ++q) {
        float x[2], y[2], z[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = (2 * q + h) * THREADS + t;
            const bool ok = i < N;
            x[h] = ok ? __ldg(p + 3 * i + 0) : 0.0f;
            y[h] = ok ? __ldg(p + 3 * i + 1) : 0.0f;
            z[h] = ok ? __ldg(p + 3 * i + 2) : 0.0f;
            pd[2 * q + h] = ok ? 1e10f : 0.0f;       // padding sits at the origin with distance 0: min(0, d) stays 0,
            sx[i] = x[h], sy[i] = y[h], sz[i] = z[h];   // and on an all-zero tie the lowest (real) index wins
        }
        px[q] = pk2(x[0], x[1]), py[q] = pk2(y[0], y[1]), pz[q] = pk2(z[0], z[1]);
    }
    long long s0 = start[b];
    unsigned far = (unsigned)(s0 < 0 ? 0 : (s0 >= N ? N - 1 : s0));
    __syncthreads();

    int64_t *o = out + (size_t)b * npoint;
    int buf = 0;
    for (int it = 0; it < npoint; ++it) {
        if (t == 0) o[it] = (int64_t)far;
        const float cxs = sx[far], cys = sy[far], czs = sz[far];
        const unsigned long long cx = pk2(cxs, cxs), cy = pk2(cys, cys), cz = pk2(czs, czs);
        float tmax = 0.0f;
#pragma unroll
        for (int q = 0; q < NP2; ++q) {
            const unsigned long long dx = sub2(px[q], cx), dy = sub2(py[q], cy), dz = sub2(pz[q], cz);
            // the two adds stay scalar: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 whatever the
            // rounding modifiers and -fmad say (checked in SASS), which would change the last bit
            float x0, x1, y0, y1, z0, z1;
            upk2(mul2(dx, dx), x0, x1);
            upk2(mul2(dy, dy), y0, y1);
            upk2(mul2(dz, dz), z0, z1);
            const float d0 = __fadd_rn(__fadd_rn(x0, y0), z0), d1 = __fadd_rn(__fadd_rn(x1, y1), z1);
            pd[2 * q] = fminf(pd[2 * q], d0);        // == (d < dist ? d : dist): no NaNs, both >= +0
            pd[2 * q + 1] = fminf(pd[2 * q + 1], d1);
            tmax = fmaxf(tmax, fmaxf(pd[2 * q], pd[2 * q + 1]));
        }
        unsigned besti = kBigIdx;
#pragma unroll
        for (int j = PPT - 1; j >= 0; --j)
            if (pd[j] == tmax) besti = (unsigned)(j * THREADS + t);      // lowest slot = lowest index of this thread
        const unsigned best = __float_as_uint(tmax);
        // warp argmax, lowest index among equals
        unsigned wmax = __reduce_max_sync(PCB_FULL_MASK, best);
        unsigned widx = __reduce_min_sync(PCB_FULL_MASK, best == wmax ? besti : kBigIdx);
        if (NWARPS == 1) {
            far = widx;
        } else {
            if (lane == 0) s_red[buf][warp] = make_uint2(wmax, widx);
            __syncthreads();
            uint2 v = lane < NWARPS ? s_red[buf][lane] : make_uint2(0u, kBigIdx);
            unsigned m = __reduce_max_sync(PCB_FULL_MASK, v.x);
            far = __reduce_min_sync(PCB_FULL_MASK, v.x == m ? v.y : kBigIdx);
            buf ^= 1;
        }
    }
}

// Large clouds (8192 < N <= 49152): running distances live in shared memory, coordinates are
// re-read from global memory (L1/L2 resident) every step.  Same arithmetic and tie-break.
__global__ void __launch_bounds__(1024, 1)
fps_smem_kernel(const float *__restrict__ xyz, int N, const int64_t *__restrict__ start, int npoint,
                int64_t *__restrict__ out)
{
    extern __shared__ float s_dist[];
    __shared__ uint2 s_red[2][32];
    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const float *p = xyz + (size_t)b * N * 3;
    for (int i = t; i < N; i += 1024) s_dist[i] = 1e10f;
    long long s0 = start[b];
    unsigned far = (unsigned)(s0 < 0 ? 0 : (s0 >= N ? N - 1 : s0));
    __syncthreads();
    int64_t *o = out + (size_t)b * npoint;
    int buf = 0;
    for (int it = 0; it < npoint; ++it) {
        if (t == 0) o[it] = (int64_t)far;
        const float cx = __ldg(p + 3 * far), cy = __ldg(p + 3 * far + 1), cz = __ldg(p + 3 * far + 2);
        unsigned best = 0u, besti = kBigIdx;
        for (int i = t; i < N; i += 1024) {
            float dx = __fsub_rn(__ldg(p + 3 * i + 0), cx);
            float dy = __fsub_rn(__ldg(p + 3 * i + 1), cy);
            float dz = __fsub_rn(__ldg(p + 3 * i + 2), cz);
            float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            float od = s_dist[i];
            float nd = (d < od) ? d : od;
            s_dist[i] = nd;
            unsigned bits = __float_as_uint(nd);
            if (bits > best || besti == kBigIdx) {
                best = bits;
                besti = (unsigned)i;
            }
        }
        unsigned wmax = __reduce_max_sync(PCB_FULL_MASK, best);
        unsigned widx = __reduce_min_sync(PCB_FULL_MASK, best == wmax ? besti : kBigIdx);
        if (lane == 0) s_red[buf][warp] = make_uint2(wmax, widx);
        __syncthreads();
        uint2 v = s_red[buf][lane];
        unsigned m = __reduce_max_sync(PCB_FULL_MASK, v.x);
        far = __reduce_min_sync(PCB_FULL_MASK, v.x == m ? v.y : kBigIdx);
        buf ^= 1;
    }
}

template <int THREADS, int PPT>
static int launch_fps_reg(const float *xyz, int B, int N, const int64_t *start, int npoint,
                          int64_t *out, cudaStream_t st)
{
    size_t smem = (size_t)THREADS * PPT * 3 * sizeof(float);
    if (smem + 1024 > 48 * 1024) {               // static smem (s_red) counts against the 48 KB default
        cudaError_t e = cudaFuncSetAttribute(fps_reg_kernel<THREADS, PPT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    fps_reg_kernel<THREADS, PPT><<<B, THREADS, smem, st>>>(xyz, N, start, npoint, out);
    PCB_RETURN_LAUNCH_STATUS();
}

template <int THREADS, int PPT>
static int launch_fps_pair(const float *xyz, int B, int N, const int64_t *start, int npoint,
                           int64_t *out, cudaStream_t st)
{
    size_t smem = (size_t)THREADS * PPT * 3 * sizeof(float);
    if (smem + 1024 > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fps_pair_kernel<THREADS, PPT>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    fps_pair_kernel<THREADS, PPT><<<B, THREADS, smem, st>>>(xyz, N, start, npoint, out);
    PCB_RETURN_LAUNCH_STATUS();
}

}  // namespace pcb

// Tuning hook for experiments (bench/profiling only): PCB_FPS_VARIANT=s selects the scalar kernels,
// b|c other (threads, points-per-thread) splits of the pair kernel for 2048 < N <= 4096.
static int fps_variant()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("PCB_FPS_VARIANT");
        v = e ? (e[0] == 'b' ? 1 : (e[0] == 'c' ? 2 : (e[0] == 's' ? 3 : 0))) : 0;
    }
    return v;
}

PCB_API int pcb_fps_f32(const float *xyz, int B, int N, const int64_t *start, int npoint,
                        int64_t *out_idx, pcb_stream_t stream)
{
    using namespace pcb;
    PCB_REQUIRE(xyz && start && out_idx, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && npoint > 0, PCB_EINVAL);
    PCB_REQUIRE(N <= 49152, PCB_ERANGE);
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 32) return launch_fps_reg<32, 1>(xyz, B, N, start, npoint, out_idx, st);
    if (fps_variant() == 3) {                      // scalar reference kernels
        if (N <= 64) return launch_fps_reg<32, 2>(xyz, B, N, start, npoint, out_idx, st);
        if (N <= 128) return launch_fps_reg<32, 4>(xyz, B, N, start, npoint, out_idx, st);
        if (N <= 256) return launch_fps_reg<32, 8>(xyz, B, N, start, npoint, out_idx, st);
        if (N <= 512) return launch_fps_reg<64, 8>(xyz, B, N, start, npoint, out_idx, st);
        if (N <= 1024) return launch_fps_reg<128, 8>(xyz, B, N, start, npoint, out_idx, st);
        if (N <= 2048) return launch_fps_reg<256, 8>(xyz, B, N, start, npoint, out_idx, st);
        if (N <= 4096) return launch_fps_reg<512, 8>(xyz, B, N, start, npoint, out_idx, st);
        if (N <= 8192) return launch_fps_reg<1024, 8>(xyz, B, N, start, npoint, out_idx, st);
    }
    if (N <= 64) return launch_fps_pair<32, 2>(xyz, B, N, start, npoint, out_idx, st);
    if (N <= 128) return launch_fps_pair<32, 4>(xyz, B, N, start, npoint, out_idx, st);
    if (N <= 256) return launch_fps_pair<32, 8>(xyz, B, N, start, npoint, out_idx, st);
    if (N <= 512) return launch_fps_pair<64, 8>(xyz, B, N, start, npoint, out_idx, st);
    if (N <= 1024) return launch_fps_pair<128, 8>(xyz, B, N, start, npoint, out_idx, st);
    if (N <= 2048) return launch_fps_pair<256, 8>(xyz, B, N, start, npoint, out_idx, st);
    if (N <= 4096) {
        switch (fps_variant()) {
            case 1: return launch_fps_pair<256, 16>(xyz, B, N, start, npoint, out_idx, st);
            case 2: return launch_fps_pair<1024, 4>(xyz, B, N, start, npoint, out_idx, st);
            default: return launch_fps_pair<512, 8>(xyz, B, N, start, npoint, out_idx, st);
        }
    }
    if (N <= 8192) return launch_fps_pair<1024, 8>(xyz, B, N, start, npoint, out_idx, st);
    size_t smem = (size_t)N * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(fps_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
    fps_smem_kernel<<<B, 1024, smem, st>>>(xyz, N, start, npoint, out_idx);
    PCB_RETURN_LAUNCH_STATUS();
}
