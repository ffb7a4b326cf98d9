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

}  // namespace pcb

// Tuning hook for experiments (bench/profiling only): PCB_FPS_VARIANT=a|b|c selects the
// (threads, points-per-thread) split used for 2048 < N <= 4096.
static int fps_variant()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("PCB_FPS_VARIANT");
        v = e ? (e[0] == 'b' ? 1 : (e[0] == 'c' ? 2 : 0)) : 0;
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
    if (N <= 64) return launch_fps_reg<32, 2>(xyz, B, N, start, npoint, out_idx, st);
    if (N <= 128) return launch_fps_reg<32, 4>(xyz, B, N, start, npoint, out_idx, st);
    if (N <= 256) return launch_fps_reg<32, 8>(xyz, B, N, start, npoint, out_idx, st);
    if (N <= 512) return launch_fps_reg<64, 8>(xyz, B, N, start, npoint, out_idx, st);
    if (N <= 1024) return launch_fps_reg<128, 8>(xyz, B, N, start, npoint, out_idx, st);
    if (N <= 2048) return launch_fps_reg<256, 8>(xyz, B, N, start, npoint, out_idx, st);
    if (N <= 4096) {
        switch (fps_variant()) {
            case 1: return launch_fps_reg<256, 16>(xyz, B, N, start, npoint, out_idx, st);
            case 2: return launch_fps_reg<1024, 4>(xyz, B, N, start, npoint, out_idx, st);
            default: return launch_fps_reg<512, 8>(xyz, B, N, start, npoint, out_idx, st);
        }
    }
    if (N <= 8192) return launch_fps_reg<1024, 8>(xyz, B, N, start, npoint, out_idx, st);
    size_t smem = (size_t)N * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(fps_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
    fps_smem_kernel<<<B, 1024, smem, st>>>(xyz, N, start, npoint, out_idx);
    PCB_RETURN_LAUNCH_STATUS();
}
