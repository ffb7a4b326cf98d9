// Shared device helpers for libpcbridge (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/pcbridge.h"

#define PCB_API extern "C" __attribute__((visibility("default")))

#define PCB_FULL_MASK 0xffffffffu
#define PCB_NUM_SMS 148

#define PCB_RETURN_LAUNCH_STATUS() return (int)cudaGetLastError()

#define PCB_REQUIRE(cond, code) \
    do {                        \
        if (!(cond)) return (code); \
    } while (0)

namespace pcb {

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel call site, device): function attributes are per
// device, so a process that drives several GPUs must opt in on each of them.  `flags` is the call site's own
// zero-initialised static array.
constexpr int kMaxDevices = 64;
template <typename K>
static inline cudaError_t smem_optin_once(K kernel, int bytes, bool (&flags)[kMaxDevices])
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < kMaxDevices && flags[dev]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev >= 0 && dev < kMaxDevices) flags[dev] = true;
    return e;
}

__host__ __device__ __forceinline__ int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------
// The training step is ~250 dependent kernels of 5-30 us: the launch latency between two of them and the prologue of
// the second (barrier init, TMEM allocation, descriptor prefetch, per-channel constants) are a measurable part of it.
// Kernels launched through launch_pdl() may start while their predecessor in the stream is still draining:
//   * pdl_wait() blocks until the PREVIOUS kernel has completed and its writes are visible -- it comes before the
//     first access to global memory (everything before it is CTA-local setup);
//   * pdl_trigger(), always AFTER pdl_wait(), lets the NEXT kernel's CTAs be scheduled once every CTA of this one
//     has passed its own wait (they take the SM slots that this kernel's tail leaves idle and run their prologue).
//     Wait-then-trigger keeps the overlap one kernel deep: when a kernel's prologue runs, everything before its
//     immediate predecessor has completed.
// Without the launch attribute (ordinary launch, PCB_NO_PDL=1) both are no-ops and the stream order is the usual one.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

static inline bool pdl_enabled()
{
    static const bool on = [] { const char *e = getenv("PCB_NO_PDL"); return !(e && e[0] == '1'); }();
    return on;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr, cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// |p|^2 exactly as ATen's CPU sum(v**2, -1) evaluates it for three components:
// (x*x + y*y) + z*z, every operation rounded, no FMA.
__device__ __forceinline__ float norm3(float x, float y, float z)
{
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// dot of two 3-vectors as MKL sgemm evaluates it: sequential FMA chain from 0.
__device__ __forceinline__ float dot3_chain(float ax, float ay, float az, float bx, float by, float bz)
{
    float acc = __fmaf_rn(ax, bx, 0.0f);         // the chain starts from +0
    acc = __fmaf_rn(ay, by, acc);
    acc = __fmaf_rn(az, bz, acc);
    return acc;
}

// square_distance(src, dst) element: ((-2 * dot) + |src|^2) + |dst|^2
// (pointnet_util.py:40-42).
__device__ __forceinline__ float sqdist3(float sx, float sy, float sz, float sn, float dx, float dy,
                                         float dz, float dn)
{
    float t = __fmul_rn(-2.0f, dot3_chain(sx, sy, sz, dx, dy, dz));
    t = __fadd_rn(t, sn);
    return __fadd_rn(t, dn);
}

// torch.cdist element before clamp/sqrt: K=5 FMA chain over
// [-2x, |x|^2, 1] . [y, 1, |y|^2]   (ATen _euclidean_dist).
__device__ __forceinline__ float cdist3_pre(float m2x, float m2y, float m2z, float xn, float yx,
                                            float yy, float yz, float yn)
{
    float t = __fmaf_rn(m2x, yx, 0.0f);
    t = __fmaf_rn(m2y, yy, t);
    t = __fmaf_rn(m2z, yz, t);
    t = __fadd_rn(xn, t);                        // fma(|x|^2, 1, t)
    t = __fadd_rn(yn, t);                        // fma(1, |y|^2, t)
    return t;
}

// |v|^2 over C contiguous floats in ATen's CPU order: 4 interleaved accumulators of 8 lanes
// over the leading floor(C/8) vectors, folded ((A0+A1)+A2)+A3, scalar tail summed first from 0,
// then the 8 lane partials left to right.
__device__ __forceinline__ float row_sumsq_aten(const float *v, int C, int stride)
{
    float acc[4][8];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int l = 0; l < 8; ++l) acc[a][l] = 0.f;
    const int nvec = C / 8, nilp = nvec / 4;
    for (int r = 0; r < nilp; ++r)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                float x = v[(size_t)((r * 4 + a) * 8 + l) * stride];
                acc[a][l] = __fadd_rn(acc[a][l], __fmul_rn(x, x));
            }
    for (int i = nilp * 4; i < nvec; ++i)
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            float x = v[(size_t)(i * 8 + l) * stride];
            acc[0][l] = __fadd_rn(acc[0][l], __fmul_rn(x, x));
        }
#pragma unroll
    for (int a = 1; a < 4; ++a)
#pragma unroll
        for (int l = 0; l < 8; ++l) acc[0][l] = __fadd_rn(acc[0][l], acc[a][l]);
    float fin = 0.f;
    for (int c = nvec * 8; c < C; ++c) {
        float x = v[(size_t)c * stride];
        fin = __fadd_rn(fin, __fmul_rn(x, x));
    }
#pragma unroll
    for (int l = 0; l < 8; ++l) fin = __fadd_rn(fin, acc[0][l]);
    return fin;
}

// Chan's pairwise update of (count, mean, M2): (n, mean, m2) <- (n, mean, m2) + (cn, cm, cq)
__device__ __forceinline__ void chan_merge(float &n, float &mean, float &m2, float cn, float cm, float cq)
{
    const float tot = n + cn;
    if (tot > 0.f) {
        const float w = cn / tot, d = cm - mean;
        mean = fmaf(d, w, mean);
        m2 = m2 + cq + d * d * n * w;
        n = tot;
    }
}

// Division by a runtime-constant 32-bit divisor in two instructions (mul.hi + shift); the
// element index -> (row, channel) -> (cloud, ...) decompositions of these kernels were the
// instruction-issue bottleneck with native 64-bit division (round-1 ncu: 70 % issue, 2 % DRAM).
struct FastDiv {
    unsigned d, magic, shift;
    __device__ __forceinline__ unsigned div(unsigned n) const { return (__umulhi(n, magic) + n) >> shift; }   // n < 2^31
};
static inline FastDiv make_fastdiv(unsigned d)
{
    FastDiv f;
    f.d = d;
    unsigned l = 0;
    while ((1ull << l) < d) ++l;
    f.shift = l;
    f.magic = (unsigned)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
    return f;
}

// streaming (read-once) 128-bit load / store that keep L1 for the gathered operand
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_f4(float4 *p, const float4 &v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_stream_f1(float *p, float v)
{
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, no tensor map) -------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                         uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- Ampere-style 16-byte async copy (LDGSTS); src_bytes < 16 zero-fills the remainder ----
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, int src_bytes)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
                 "r"(src_bytes)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace pcb
