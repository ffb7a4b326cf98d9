// Training-mode BatchNorm + ReLU (+ max over the neighbour axis) on point-major rows [M, C] --
// the elementwise half of the shared MLP of the set-abstraction / feature-propagation modules
// (Conv2d 1x1 -> BatchNorm2d -> ReLU -> torch.max over nsample:
//  Partsize-identical/models/pointnet_util.py:213-217, 273-279, 343-345;
//  Highway_bridge/models/pointnet2_utils.py:150-154, 353-356).
// The training GEMM stays a library call in round 1; these kernels replace PyTorch's separate
// statistics / transform / ReLU / max / threshold_backward / bias-sum passes with ONE persistent
// cooperative kernel per direction:
//   forward : column sums of y -> grid sync -> fold + mean/invstd/running stats -> grid sync
//             -> act(BN(y)) [+ max over K]        (y read twice, the second time from L2; 1 write)
//   backward: column sums of dy, dy*yhat, yhat -> grid sync -> fold (+ conv-bias gradient)
//             -> grid sync -> gy                  (y and gz read twice; 1 write)
// A layer's activations (<= 67 MB bf16 here) fit the 126 MB L2, so the second read does not go to
// HBM; one launch instead of seven per layer matters as much as the bytes, because the layers are
// small (13 MB on average in the MSG train step: ~2 us of HBM time against ~4 us per launch).
// Activations may be fp32 or bf16 (autocast); statistics, affine parameters and all arithmetic
// are fp32.  The convolution bias is folded in here (BN(xW + b) only needs b for the running
// mean), so no separate bias-add or bias-gradient pass exists.
// Column reductions are two-stage and deterministic: every CTA stores its partial sums to
// parts[cta][NACC*C] with plain stores; after the grid barrier one warp per channel adds the
// <= 296 partials in a fixed order.
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include <cstdlib>

#include "pcb_common.cuh"

namespace cg = cooperative_groups;

namespace pcb {

#ifndef PCB_BN_U
#define PCB_BN_U 4    // rows in flight per thread, forward phases
#endif
#ifndef PCB_BN_UB
#define PCB_BN_UB 4   // rows in flight per thread, backward phases (two loads per row)
#endif
#ifndef PCB_BN_MINB
#define PCB_BN_MINB 2 // __launch_bounds__ min CTAs per SM
#endif
#ifdef PCB_BN_TRACE   // kernel-tuning aid: per-CTA %globaltimer stamps at the phase boundaries
__device__ unsigned long long g_bn_trace[8 * 1024];
#define BN_STAMP(i)                                                                   \
    do {                                                                              \
        if (threadIdx.x == 0 && blockIdx.x < 1024) {                                  \
            unsigned long long t_;                                                    \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                    \
            g_bn_trace[blockIdx.x * 8 + (i)] = t_;                                    \
        }                                                                             \
    } while (0)
#else
#define BN_STAMP(i)
#endif
constexpr int kBnThreads = 256;
#ifndef PCB_BN_APPLY_MINB
#define PCB_BN_APPLY_MINB 1   // min CTAs per SM of the ordinary-launch elementwise kernels (tuning switch)
#endif
constexpr int kBnMaxParts = PCB_NUM_SMS * 2;      // at most 2 CTAs per SM (register budget) -> <= 296 partials

// VecIO<T, V>: V consecutive channels per thread, one 16-byte access for (float,4) and (bf16,8),
// one 8-byte access for (bf16,4).  ldraw/cvt split the load from its first use so that several
// rows can be in flight per thread.
template <typename T, int V>
struct VecIO;
template <>
struct VecIO<float, 4> {
    typedef float4 Raw;
    static __device__ __forceinline__ Raw ldraw(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
    static __device__ __forceinline__ void cvt(const Raw &t, float v[4]) { v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w; }
    static __device__ __forceinline__ void store(float *p, const float v[4])
    {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <>
struct VecIO<__nv_bfloat16, 4> {
    typedef uint2 Raw;
    static __device__ __forceinline__ Raw ldraw(const __nv_bfloat16 *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
    static __device__ __forceinline__ void cvt(const Raw &t, float v[4])
    {
        // bf16 -> fp32 is a 16-bit shift
        v[0] = __uint_as_float(t.x << 16), v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16), v[3] = __uint_as_float(t.y & 0xffff0000u);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float v[4])
    {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 t;
        t.x = *reinterpret_cast<unsigned *>(&a);
        t.y = *reinterpret_cast<unsigned *>(&b);
        *reinterpret_cast<uint2 *>(p) = t;
    }
};
template <>
struct VecIO<__nv_bfloat16, 8> {
    typedef uint4 Raw;
    static __device__ __forceinline__ Raw ldraw(const __nv_bfloat16 *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
    static __device__ __forceinline__ void cvt(const Raw &t, float v[8])
    {
        const unsigned w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const float v[8])
    {
        unsigned w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 a = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<unsigned *>(&a);
        }
        *reinterpret_cast<uint4 *>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

// V one-byte argmax entries in one 4- or 8-byte access
template <int V>
struct AmIO;
template <>
struct AmIO<8> {
    typedef uint2 Raw;
    static __device__ __forceinline__ Raw ld(const unsigned char *p) { return *reinterpret_cast<const uint2 *>(p); }
    static __device__ __forceinline__ int get(const Raw &t, int i) { return (int)(((i < 4 ? t.x : t.y) >> (8 * (i & 3))) & 0xffu); }
    static __device__ __forceinline__ void st(unsigned char *p, const int bi[8])
    {
        unsigned w[2] = {0u, 0u};
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i >> 2] |= (unsigned)(bi[i] & 0xff) << (8 * (i & 3));
        *reinterpret_cast<uint2 *>(p) = make_uint2(w[0], w[1]);
    }
};
template <>
struct AmIO<4> {
    typedef unsigned Raw;
    static __device__ __forceinline__ Raw ld(const unsigned char *p) { return *reinterpret_cast<const unsigned *>(p); }
    static __device__ __forceinline__ int get(const Raw &t, int i) { return (int)((t >> (8 * i)) & 0xffu); }
    static __device__ __forceinline__ void st(unsigned char *p, const int bi[4])
    {
        unsigned w = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) w |= (unsigned)(bi[i] & 0xff) << (8 * i);
        *reinterpret_cast<unsigned *>(p) = w;
    }
};

// ---------------------------------------------------------------------------------------------
// Skeletons.  threadIdx.x % TX owns one group of V channels (per-channel constants are loaded
// once per group by f.init), threadIdx.x / TX walks "units" (rows, or pooling groups of pool_k
// rows).  U units are loaded (f.load -> Pack) before the first is consumed.
// ---------------------------------------------------------------------------------------------
struct Lanes {
    int CV, TX, TY, tx, ty;
    __device__ __forceinline__ Lanes(int C, int V)
    {
        CV = C / V;
        TX = CV < kBnThreads ? CV : kBnThreads;
        TY = kBnThreads / TX;
        tx = threadIdx.x % TX;
        ty = threadIdx.x / TX;
    }
};

// column sums over the units [blockIdx.x * upc, +upc) -> parts[blockIdx.x][NACC * C]
template <int NACC, int V, int U, typename F>
__device__ __forceinline__ void column_reduce(F &f, int64_t units, int C, int64_t upc, float *__restrict__ parts,
                                              float (*s_part)[kBnThreads][V])
{
    const Lanes L(C, V);
    const int64_t u0 = (int64_t)blockIdx.x * upc;
    const int64_t u1 = u0 + upc < units ? u0 + upc : units;
    // channel groups of one row sit on lanes tx, tx+TX, ...: when TX is a power of two <= 32 the
    // row partials of a warp are folded with shuffles first
    const bool warp_fold = L.TX <= 32 && (L.TX & (L.TX - 1)) == 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *my_parts = parts + (size_t)blockIdx.x * NACC * C;
    for (int cg0 = 0; cg0 < L.CV; cg0 += L.TX) {
        const int cgi = cg0 + L.tx;
        const int c = cgi * V;
        float acc[NACC][V];
#pragma unroll
        for (int a = 0; a < NACC; ++a)
#pragma unroll
            for (int v = 0; v < V; ++v) acc[a][v] = 0.f;
        if (L.ty < L.TY && cgi < L.CV) {
            f.init(c);
            for (int64_t u = u0 + L.ty; u < u1; u += (int64_t)U * L.TY) {
                typename F::Pack p[U];
#pragma unroll
                for (int j = 0; j < U; ++j)
                    if (u + j * L.TY < u1) p[j] = f.load(u + j * L.TY, c);
#pragma unroll
                for (int j = 0; j < U; ++j)
                    if (u + j * L.TY < u1) f.use(p[j], u + j * L.TY, c, acc);
            }
        }
        if (warp_fold) {
            for (int off = L.TX; off < 32; off <<= 1)
#pragma unroll
                for (int a = 0; a < NACC; ++a)
#pragma unroll
                    for (int v = 0; v < V; ++v) acc[a][v] += __shfl_xor_sync(PCB_FULL_MASK, acc[a][v], off);
            if (lane < L.TX) {
#pragma unroll
                for (int a = 0; a < NACC; ++a)
#pragma unroll
                    for (int v = 0; v < V; ++v) s_part[a][warp * 32 + lane][v] = acc[a][v];
            }
            __syncthreads();
            if (threadIdx.x < L.TX && cgi < L.CV) {
#pragma unroll
                for (int a = 0; a < NACC; ++a)
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        float s = 0.f;
#pragma unroll
                        for (int w = 0; w < kBnThreads / 32; ++w) s += s_part[a][w * 32 + threadIdx.x][v];
                        my_parts[(size_t)a * C + c + v] = s;
                    }
            }
        } else {
#pragma unroll
            for (int a = 0; a < NACC; ++a)
#pragma unroll
                for (int v = 0; v < V; ++v) s_part[a][threadIdx.x][v] = acc[a][v];
            __syncthreads();
            if (L.ty == 0 && cgi < L.CV) {
#pragma unroll
                for (int a = 0; a < NACC; ++a)
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        float s = 0.f;
                        for (int y = 0; y < L.TY; ++y) s += s_part[a][y * L.TX + L.tx][v];
                        my_parts[(size_t)a * C + c + v] = s;
                    }
            }
        }
        __syncthreads();
    }
}

// elementwise pass over all units, grid-strided.  (Handing chunks out dynamically through a
// ticket counter balanced the SMs -- the slowest finishes this phase ~25 % after the median --
// but the per-chunk __syncthreads and the smaller number of loads in flight cost more than that.)
template <int V, int U, typename F>
__device__ __forceinline__ void row_stream(F &f, int64_t units, int C)
{
    const Lanes L(C, V);
    if (L.ty >= L.TY) return;
    const int64_t stride = (int64_t)gridDim.x * L.TY;
    for (int cgi = L.tx; cgi < L.CV; cgi += L.TX) {
        const int c = cgi * V;
        f.init(c);
        for (int64_t u = (int64_t)blockIdx.x * L.TY + L.ty; u < units; u += U * stride) {
            typename F::Pack p[U];
#pragma unroll
            for (int j = 0; j < U; ++j)
                if (u + j * stride < units) p[j] = f.load(u + j * stride, c);
#pragma unroll
            for (int j = 0; j < U; ++j)
                if (u + j * stride < units) f.emit(p[j], u + j * stride, c);
        }
    }
}

// sum over the CTA partials of the NACC accumulators of channel c (one warp; fixed order).  All
// loads are issued before the first add: the fold sits between two grid barriers, so its latency
// (one L2 round trip instead of ten) is on every CTA's critical path.
constexpr int kFoldJ = (kBnMaxParts + 31) / 32;
template <int NACC>
__device__ __forceinline__ void fold_channel(const float *parts, int nparts, int C, int c, float s[NACC])
{
    const int lane = threadIdx.x & 31;
    float v[kFoldJ][NACC];
#pragma unroll
    for (int j = 0; j < kFoldJ; ++j) {
        const int i = lane + 32 * j;
#pragma unroll
        for (int a = 0; a < NACC; ++a) v[j][a] = i < nparts ? __ldcg(parts + ((size_t)i * NACC + a) * C + c) : 0.f;
    }
#pragma unroll
    for (int a = 0; a < NACC; ++a) {
        s[a] = 0.f;
#pragma unroll
        for (int j = 0; j < kFoldJ; ++j) s[a] += v[j][a];
#pragma unroll
        for (int off = 16; off; off >>= 1) s[a] += __shfl_xor_sync(PCB_FULL_MASK, s[a], off);
    }
}

// channel -> warp assignment of the fold: consecutive channels go to different CTAs (SMs)
#define PCB_FOLD_LOOP(c, C) \
    for (int c = blockIdx.x + (threadIdx.x >> 5) * gridDim.x; c < (C); c += gridDim.x * (kBnThreads / 32))

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
// sum_rows (y - y[0]) and sum_rows (y - y[0])^2 (shifted: no cancellation when |mean| >> std)
template <typename T, int V>
struct FwdStats {
    typedef typename VecIO<T, V>::Raw Pack;
    const T *y;
    int C;
    float s[V];
    __device__ __forceinline__ void init(int c) { VecIO<T, V>::cvt(VecIO<T, V>::ldraw(y + c), s); }
    __device__ __forceinline__ Pack load(int64_t r, int c) const { return VecIO<T, V>::ldraw(y + r * C + c); }
    __device__ __forceinline__ void use(const Pack &p, int64_t, int, float acc[2][V]) const
    {
        float v[V];
        VecIO<T, V>::cvt(p, v);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float d = v[i] - s[i];
            acc[0][i] += d;
            acc[1][i] = fmaf(d, d, acc[1][i]);
        }
    }
};

// z = act(y * sc + sh), sc = invstd * gamma, sh = beta - mean * sc
template <typename T, int V>
struct FwdApply {
    typedef typename VecIO<T, V>::Raw Pack;
    const T *y;
    T *out;
    const float *mean, *invstd, *gamma, *beta;
    const float *s_const;                                  // shared copy [sc | sh] of all C channels, or nullptr
    int C, relu;
    int64_t opitch;                                        // row pitch of `out` (>= C: a column slice of a wider buffer)
    float sc[V], sh[V];
    __device__ __forceinline__ void init(int c)
    {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            if (s_const) {
                sc[i] = s_const[c + i];
                sh[i] = s_const[C + c + i];
            } else {
                sc[i] = __ldcg(invstd + c + i) * gamma[c + i];
                sh[i] = beta[c + i] - __ldcg(mean + c + i) * sc[i];
            }
        }
    }
    __device__ __forceinline__ Pack load(int64_t r, int c) const { return VecIO<T, V>::ldraw(y + r * C + c); }
    __device__ __forceinline__ void emit(const Pack &p, int64_t r, int c) const
    {
        float v[V];
        VecIO<T, V>::cvt(p, v);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            v[i] = fmaf(v[i], sc[i], sh[i]);
            if (relu) v[i] = fmaxf(v[i], 0.f);
        }
        VecIO<T, V>::store(out + r * opitch + c, v);
    }
};

// out[g] = max_k z[g * pool_k + k]; the winning k (first on ties, as torch.max) goes to argmax
template <typename T, int V, int KB>
struct FwdApplyPooled : FwdApply<T, V> {
    struct Pack {};
    unsigned char *argmax;
    int pool_k;
    __device__ __forceinline__ Pack load(int64_t, int) const { return Pack(); }
    __device__ __forceinline__ void emit(const Pack &, int64_t g, int c) const
    {
        float best[V];
        int bi[V];
#pragma unroll
        for (int i = 0; i < V; ++i) best[i] = 0.f, bi[i] = 0;
        const T *row = this->y + (g * pool_k) * this->C + c;
        for (int k0 = 0; k0 < pool_k; k0 += KB) {
            typename VecIO<T, V>::Raw t[KB];
#pragma unroll
            for (int j = 0; j < KB; ++j)
                if (k0 + j < pool_k) t[j] = VecIO<T, V>::ldraw(row + (size_t)(k0 + j) * this->C);
#pragma unroll
            for (int j = 0; j < KB; ++j)
                if (k0 + j < pool_k) {
                    float v[V];
                    VecIO<T, V>::cvt(t[j], v);
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        float z = fmaf(v[i], this->sc[i], this->sh[i]);
                        if (this->relu) z = fmaxf(z, 0.f);
                        if (k0 + j == 0 || z > best[i]) best[i] = z, bi[i] = k0 + j;
                    }
                }
        }
        VecIO<T, V>::store(this->out + g * this->opitch + c, best);
        if (argmax) AmIO<V>::st(argmax + g * this->C + c, bi);
    }
};

struct BnFwdArgs {
    const void *y;
    void *out;
    unsigned char *argmax;
    const float *bias, *gamma, *beta;
    float *running_mean, *running_var, *mean, *invstd, *work;
    const float *var;                                  // biased batch variance (bn_apply_rows_kernel only)
    int64_t M, upc;
    int C, Cv, pool_k, relu, nparts;                   // C = row pitch (multiple of 4), Cv <= C real channels
    int64_t out_pitch;                                 // row pitch of out (C, or wider when out is a column slice)
    float eps, momentum;
    // bn_apply kernels only: group partials (n, mean, M2) [groups][3][C] left by the statistics GEMM's first fold level
    // (gemm_rows.cu, deferred mode); every CTA merges them for its constants, CTA 0 also writes mean / invstd / var_out
    const float *gparts;
    float *var_out;
    int groups, rs;                                    // rs: row lanes per pooling group (bn_apply_pooled_kernel)
    void *ymax;                                        // pooled: y of the winning row [M / pool_k][C] (may be NULL)
};

template <typename T, int V>
__global__ void __launch_bounds__(kBnThreads, PCB_BN_MINB)
bn_fwd_fused_kernel(const BnFwdArgs a)
{
    cg::grid_group grid = cg::this_grid();
    BN_STAMP(0);
    const T *y = (const T *)a.y;
    const int C = a.C;
    float *parts = a.work + 3 * (size_t)C;
    __shared__ float s_part[2][kBnThreads][V];
    if ((int)blockIdx.x < a.nparts) {
        FwdStats<T, V> f;
        f.y = y, f.C = C;
        column_reduce<2, V, PCB_BN_U>(f, a.M, C, a.upc, parts, s_part);
    }
    BN_STAMP(1);
    grid.sync();
    BN_STAMP(2);
    // mean / invstd of y from the shifted sums; running statistics update (momentum, unbiased
    // variance, conv bias added to the running mean) as torch.nn.functional.batch_norm does
    {
        const int lane = threadIdx.x & 31;
        PCB_FOLD_LOOP(c, a.Cv) {
            // operands of the finalize step first, so that their latency overlaps the fold's
            const float shift = (float)y[c];
            const float b = a.bias ? a.bias[c] : 0.f;
            const float rm = a.running_mean ? a.running_mean[c] : 0.f, rv = a.running_mean ? a.running_var[c] : 0.f;
            float s[2];
            fold_channel<2>(parts, a.nparts, C, c, s);
            if (lane == 0) {
                const float M = (float)a.M;
                const float m1 = s[0] / M;
                const float var = fmaxf(s[1] / M - m1 * m1, 0.f);
                const float mu = shift + m1;                  // mean of the bias-free pre-activation
                a.mean[c] = mu;
                a.invstd[c] = rsqrtf(var + a.eps);
                if (a.running_mean) {
                    a.running_mean[c] = (1.f - a.momentum) * rm + a.momentum * (mu + b);
                    const float unbiased = a.M > 1 ? var * (M / (float)(a.M - 1)) : var;
                    a.running_var[c] = (1.f - a.momentum) * rv + a.momentum * unbiased;
                }
            }
        }
    }
    BN_STAMP(3);
    grid.sync();
    BN_STAMP(4);
    // Per-channel constants of the elementwise phase: computed once per CTA into shared memory with
    // coalesced loads.  (Every thread fetching its own channels' mean / invstd from L2 right after
    // the barrier -- 2368 warps x 16 requests on the same few cache lines -- serialised on one L2
    // slice and cost ~10 us per launch whatever the layer size: round-1 phase trace.)
    float *s_const = nullptr;
    if (C <= kBnThreads * V) {                             // 2*C floats fit the statistics scratch
        s_const = &s_part[0][0][0];
        for (int c = threadIdx.x; c < C; c += kBnThreads) {
            const bool real = c < a.Cv;                     // pad channels (zero columns of y) produce zeros
            const float sc = real ? __ldcg(a.invstd + c) * a.gamma[c] : 0.f;
            s_const[c] = sc;
            s_const[C + c] = real ? a.beta[c] - __ldcg(a.mean + c) * sc : 0.f;
        }
        __syncthreads();
    }
    if (a.pool_k > 1) {
        FwdApplyPooled<T, V, 8> f;
        f.y = y, f.out = (T *)a.out, f.mean = a.mean, f.invstd = a.invstd, f.gamma = a.gamma, f.beta = a.beta;
        f.C = C, f.relu = a.relu, f.argmax = a.argmax, f.pool_k = a.pool_k, f.s_const = s_const, f.opitch = a.out_pitch;
        row_stream<V, 1>(f, a.M / a.pool_k, C);
    } else {
        FwdApply<T, V> f;
        f.y = y, f.out = (T *)a.out, f.mean = a.mean, f.invstd = a.invstd, f.gamma = a.gamma, f.beta = a.beta;
        f.C = C, f.relu = a.relu, f.s_const = s_const, f.opitch = a.out_pitch;
        row_stream<V, PCB_BN_U>(f, a.M, C);
    }
    BN_STAMP(5);
}

// ---------------------------------------------------------------------------------------------
// backward.  dy(row) = gz(row) * [act passes]  (pooled: gz of the group, only for the winning k);
// yhat = (y - mean) * invstd;  gy = gamma * invstd * (dy - sum_dy / M - yhat * sum_dy_yhat / M)
// ---------------------------------------------------------------------------------------------
template <typename T, int V>
struct BwdBase {
    const T *y, *gz;
    const unsigned char *argmax;
    T *gy;
    const float *mean, *invstd, *gamma, *beta, *sums;
    const float *s_const;                                  // shared [nm | is | sc | sh | a0 | a1] x C, or nullptr
    int C, Cv, relu, pool_k;
    int64_t gpitch;                                        // row pitch of gz (C, or wider: column slice of a concatenated gradient)
    float invM;
    float nm[V], is[V], sc[V], sh[V], a0[V], a1[V];      // nm = -mean * invstd: yhat = fma(y, is, nm)
    __device__ __forceinline__ void consts(int c, bool with_sums)
    {
        if (s_const) {
#pragma unroll
            for (int i = 0; i < V; ++i) {
                nm[i] = s_const[c + i], is[i] = s_const[C + c + i], sc[i] = s_const[2 * C + c + i];
                sh[i] = s_const[3 * C + c + i], a0[i] = s_const[4 * C + c + i], a1[i] = s_const[5 * C + c + i];
            }
            return;
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
            if (c + i >= Cv) {                                // pad channel: yhat = 0, z = 0 (masked), gy = 0
                nm[i] = is[i] = sc[i] = sh[i] = a0[i] = a1[i] = 0.f;
                continue;
            }
            const float m = mean[c + i];
            is[i] = invstd[c + i];
            nm[i] = -m * is[i];
            sc[i] = is[i] * gamma[c + i];                     // same expressions as the forward pass
            sh[i] = beta[c + i] - m * sc[i];
            if (with_sums) {
                a0[i] = __ldcg(sums + c + i) * invM;
                a1[i] = __ldcg(sums + C + c + i) * invM;
            }
        }
    }
    // one row: dy and yhat from the raw y / gz values; k < 0: no pooling
    __device__ __forceinline__ void row(const typename VecIO<T, V>::Raw &yr, const float g[V],
                                        const typename AmIO<V>::Raw &am, int k, float dy[V], float yh[V]) const
    {
        float v[V];
        VecIO<T, V>::cvt(yr, v);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            yh[i] = fmaf(v[i], is[i], nm[i]);
            const float z = fmaf(v[i], sc[i], sh[i]);
            bool pass = !relu || z > 0.f;
            if (k >= 0) pass = pass && (k == AmIO<V>::get(am, i));
            dy[i] = pass ? g[i] : 0.f;
        }
    }
    __device__ __forceinline__ void accumulate(const float dy[V], const float yh[V], float acc[3][V]) const
    {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            acc[0][i] += dy[i];
            acc[1][i] = fmaf(dy[i], yh[i], acc[1][i]);
            acc[2][i] += yh[i];
        }
    }
    __device__ __forceinline__ void write(int64_t r, int c, const float dy[V], const float yh[V]) const
    {
        float o[V];
#pragma unroll
        for (int i = 0; i < V; ++i) o[i] = sc[i] * fmaf(-yh[i], a1[i], dy[i] - a0[i]);
        VecIO<T, V>::store(gy + r * C + c, o);
    }
};

// unit = row
template <typename T, int V, bool APPLY>
struct BwdRows : BwdBase<T, V> {
    struct Pack {
        typename VecIO<T, V>::Raw y, g;
    };
    __device__ __forceinline__ void init(int c) { this->consts(c, APPLY); }
    __device__ __forceinline__ Pack load(int64_t r, int c) const
    {
        Pack p;
        p.y = VecIO<T, V>::ldraw(this->y + r * this->C + c);
        p.g = VecIO<T, V>::ldraw(this->gz + r * this->gpitch + c);
        return p;
    }
    __device__ __forceinline__ void use(const Pack &p, int64_t, int, float acc[3][V]) const
    {
        float g[V], dy[V], yh[V];
        VecIO<T, V>::cvt(p.g, g);
        this->row(p.y, g, typename AmIO<V>::Raw(), -1, dy, yh);
        this->accumulate(dy, yh, acc);
    }
    __device__ __forceinline__ void emit(const Pack &p, int64_t r, int c) const
    {
        float g[V], dy[V], yh[V];
        VecIO<T, V>::cvt(p.g, g);
        this->row(p.y, g, typename AmIO<V>::Raw(), -1, dy, yh);
        this->write(r, c, dy, yh);
    }
};

// unit = pooling group of pool_k consecutive rows sharing one gz row and one argmax row
template <typename T, int V, bool APPLY, int KB>
struct BwdGroups : BwdBase<T, V> {
    struct Pack {
        typename VecIO<T, V>::Raw g;
        typename AmIO<V>::Raw am;
    };
    // a unit is one of `rs` contiguous row slices (kp rows each) of a pooling group: unit u = group * rs + slice
    int rs, kp;
    __device__ __forceinline__ void init(int c) { this->consts(c, APPLY); }
    __device__ __forceinline__ Pack load(int64_t u, int c) const
    {
        const int64_t g = rs > 1 ? u / rs : u;
        Pack p;
        p.g = VecIO<T, V>::ldraw(this->gz + g * this->gpitch + c);
        p.am = AmIO<V>::ld(this->argmax + g * this->C + c);
        return p;
    }
    template <typename Sink>
    __device__ __forceinline__ void walk(const Pack &p, int64_t u, int c, Sink sink) const
    {
        float g[V];
        VecIO<T, V>::cvt(p.g, g);
        const int64_t grp = rs > 1 ? u / rs : u;
        const int k_lo = (int)(u - grp * rs) * kp;
        const int k_hi = k_lo + kp < this->pool_k ? k_lo + kp : this->pool_k;
        const int64_t r0 = grp * this->pool_k;
        const T *row = this->y + r0 * this->C + c;
        for (int k0 = k_lo; k0 < k_hi; k0 += KB) {
            typename VecIO<T, V>::Raw t[KB];
#pragma unroll
            for (int j = 0; j < KB; ++j)
                if (k0 + j < k_hi) t[j] = VecIO<T, V>::ldraw(row + (size_t)(k0 + j) * this->C);
#pragma unroll
            for (int j = 0; j < KB; ++j)
                if (k0 + j < k_hi) {
                    float dy[V], yh[V];
                    this->row(t[j], g, p.am, k0 + j, dy, yh);
                    sink(r0 + k0 + j, dy, yh);
                }
        }
    }
    __device__ __forceinline__ void use(const Pack &p, int64_t grp, int c, float acc[3][V]) const
    {
        walk(p, grp, c, [&](int64_t, const float *dy, const float *yh) { this->accumulate(dy, yh, acc); });
    }
    __device__ __forceinline__ void emit(const Pack &p, int64_t grp, int c) const
    {
        walk(p, grp, c, [&](int64_t r, const float *dy, const float *yh) { this->write(r, c, dy, yh); });
    }
};

struct BnBwdArgs {
    const void *gz, *y;
    const unsigned char *argmax;
    void *gy;
    const float *mean, *invstd, *gamma, *beta;
    float *work;
    int64_t M, upc, gz_pitch;
    int C, Cv, pool_k, relu, nparts;
    int rs;                                                // pooled: row slices per pooling group (units = groups * rs)
};

template <typename T, int V, typename F>
__device__ __forceinline__ void bwd_fill(F &f, const BnBwdArgs &a)
{
    f.y = (const T *)a.y, f.gz = (const T *)a.gz, f.argmax = a.argmax, f.gy = (T *)a.gy;
    f.mean = a.mean, f.invstd = a.invstd, f.gamma = a.gamma, f.beta = a.beta, f.sums = a.work;
    f.C = a.C, f.Cv = a.Cv, f.relu = a.relu, f.pool_k = a.pool_k, f.invM = 1.f / (float)a.M, f.gpitch = a.gz_pitch;
    f.s_const = nullptr;
}

template <typename T, int V>
__global__ void __launch_bounds__(kBnThreads, PCB_BN_MINB)
bn_bwd_fused_kernel(const BnBwdArgs a)
{
    cg::grid_group grid = cg::this_grid();
    BN_STAMP(0);
    const int C = a.C;
    float *parts = a.work + 3 * (size_t)C;
    const bool pooled = a.pool_k > 1;
    __shared__ float s_part[3][kBnThreads][V];
    if ((int)blockIdx.x < a.nparts) {
        if (pooled) {
            BwdGroups<T, V, false, 4> f;
            bwd_fill<T, V>(f, a);
            f.rs = a.rs, f.kp = (a.pool_k + a.rs - 1) / a.rs;
            column_reduce<3, V, 1>(f, a.M / a.pool_k * a.rs, C, a.upc, parts, s_part);
        } else {
            BwdRows<T, V, false> f;
            bwd_fill<T, V>(f, a);
            column_reduce<3, V, PCB_BN_UB>(f, a.M, C, a.upc, parts, s_part);
        }
    }
    BN_STAMP(1);
    grid.sync();
    BN_STAMP(2);
    // work[0:C] = sum dy (= grad beta), [C:2C] = sum dy*yhat (= grad gamma), [2C:3C] = gradient of
    // the folded conv bias = sum_rows gy = -gamma*invstd * (sum yhat) * (sum dy*yhat) / M
    {
        const int lane = threadIdx.x & 31;
        PCB_FOLD_LOOP(c, a.Cv) {
            const float gi = a.gamma[c] * a.invstd[c];
            float s[3];
            fold_channel<3>(parts, a.nparts, C, c, s);
            if (lane == 0) {
                a.work[c] = s[0];
                a.work[C + c] = s[1];
                a.work[2 * C + c] = -gi * s[2] * s[1] / (float)a.M;
            }
        }
    }
    BN_STAMP(3);
    grid.sync();
    BN_STAMP(4);
    // per-channel constants of the elementwise phase through shared memory (see the forward kernel)
    float *s_const = nullptr;
    if (2 * C <= kBnThreads * V) {                         // 6*C floats fit the statistics scratch
        s_const = &s_part[0][0][0];
        const float invM = 1.f / (float)a.M;
        for (int c = threadIdx.x; c < C; c += kBnThreads) {
            if (c >= a.Cv) {                                 // pad channel: every constant zero -> gy = 0
#pragma unroll
                for (int j = 0; j < 6; ++j) s_const[j * C + c] = 0.f;
                continue;
            }
            const float m = a.mean[c], is = a.invstd[c];
            const float sc = is * a.gamma[c];                // same expressions as BwdBase::consts
            s_const[c] = -m * is;
            s_const[C + c] = is;
            s_const[2 * C + c] = sc;
            s_const[3 * C + c] = a.beta[c] - m * sc;
            s_const[4 * C + c] = __ldcg(a.work + c) * invM;
            s_const[5 * C + c] = __ldcg(a.work + C + c) * invM;
        }
        __syncthreads();
    }
    if (pooled) {
        BwdGroups<T, V, true, 4> f;
        bwd_fill<T, V>(f, a);
        f.s_const = s_const;
        f.rs = a.rs, f.kp = (a.pool_k + a.rs - 1) / a.rs;
        row_stream<V, 1>(f, a.M / a.pool_k * a.rs, C);
    } else {
        BwdRows<T, V, true> f;
        bwd_fill<T, V>(f, a);
        f.s_const = s_const;
        row_stream<V, PCB_BN_UB>(f, a.M, C);
    }
    BN_STAMP(5);
}

// ---------------------------------------------------------------------------------------------
// Elementwise halves on their own (ordinary launches, no grid barrier): used when the per-channel statistics come
// out of the epilogue of the tensor-core GEMM that produced y (gemm_rows.cu: EPI_STATS forward, EPI_BNBWD backward).
// ---------------------------------------------------------------------------------------------
// group partials of channel c (gemm_rows.cu, deferred fold): every load in flight at once, then four independent merge
// chains over contiguous quarters of the groups (the chain of <= 37 dependent Chan updates, a division each, would sit
// on the start-up path of every CTA), joined in quarter order -- a fixed tree, so the result is deterministic
constexpr int kGroupBatch = 20;                             // 592 CTAs / 32 per group = 19 groups: one round trip
template <bool CHAN>
__device__ __forceinline__ void fold_group_parts(const float *gparts, int groups, int C, int c, float &n, float &m, float &q)
{
    n = m = q = 0.f;
    for (int j0 = 0; j0 < groups; j0 += kGroupBatch) {
        float en[kGroupBatch], em[kGroupBatch], eq[kGroupBatch];
#pragma unroll
        for (int u = 0; u < kGroupBatch; ++u) {
            const bool ok = j0 + u < groups;
            const float *e = gparts + (size_t)(ok ? j0 + u : 0) * 3 * C + c;
            en[u] = (CHAN && ok) ? __ldcg(e) : 0.f;
            em[u] = ok ? __ldcg(e + C) : 0.f;
            eq[u] = ok ? __ldcg(e + 2 * (size_t)C) : 0.f;
        }
        float an[4] = {0.f, 0.f, 0.f, 0.f}, am[4] = {0.f, 0.f, 0.f, 0.f}, aq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int u = 0; u < kGroupBatch; ++u) {
            constexpr int QN = kGroupBatch / 4;
            if (CHAN) chan_merge(an[u / QN], am[u / QN], aq[u / QN], en[u], em[u], eq[u]);
            else am[u / QN] += em[u], aq[u / QN] += eq[u];
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            if (CHAN) chan_merge(n, m, q, an[h], am[h], aq[h]);
            else m += am[h], q += aq[h];
        }
    }
}

// per-channel constants [sc | sh] of the forward elementwise pass into shared memory; CTA 0 owns the running statistics
// (and, with deferred statistics, the mean / invstd / var arrays that the backward pass reads)
__device__ __forceinline__ void fwd_apply_consts(const BnFwdArgs &a, float *s_const)
{
    const int C = a.C;
    const float Mf = (float)a.M;
    for (int c = threadIdx.x; c < C; c += kBnThreads) {
        const bool real = c < a.Cv;                         // pad channels (zero columns of y) produce zeros
        float m = 0.f, is = 0.f, var = 0.f;
        if (a.gparts) {
            if (real) {
                float n, q;
                fold_group_parts<true>(a.gparts, a.groups, C, c, n, m, q);
                var = fmaxf(q / Mf, 0.f);
                is = rsqrtf(var + a.eps);
            }
            if (blockIdx.x == 0) a.mean[c] = m, a.invstd[c] = is, a.var_out[c] = var;
        } else if (real) {
            m = a.mean[c], is = a.invstd[c];
            if (a.var) var = a.var[c];
        }
        const float sc = real ? is * a.gamma[c] : 0.f;
        s_const[c] = sc;
        s_const[C + c] = real ? a.beta[c] - m * sc : 0.f;
        if (blockIdx.x == 0 && a.running_mean && real) {
            // running statistics as torch.nn.functional.batch_norm updates them (momentum, unbiased variance; the conv
            // bias that was folded out of y is added back to the mean)
            const float b = a.bias ? a.bias[c] : 0.f;
            a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * (m + b);
            const float unbiased = a.M > 1 ? var * (Mf / (float)(a.M - 1)) : var;
            a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * unbiased;
        }
    }
    __syncthreads();
}

template <typename T, int V>
__global__ void __launch_bounds__(kBnThreads, PCB_BN_APPLY_MINB)
bn_apply_rows_kernel(const BnFwdArgs a)
{
    pdl_wait();
    pdl_trigger();
    const int C = a.C;
    __shared__ float s_buf[2][kBnThreads][V];
    float *s_const = nullptr;
    if (C <= kBnThreads * V) {
        s_const = &s_buf[0][0][0];
        fwd_apply_consts(a, s_const);
    } else if (blockIdx.x == 0 && a.running_mean) {         // very wide rows (never deferred): constants per thread
        const float Mf = (float)a.M;
        for (int c = threadIdx.x; c < a.Cv; c += kBnThreads) {
            const float b = a.bias ? a.bias[c] : 0.f, var = a.var[c];
            a.running_mean[c] = (1.f - a.momentum) * a.running_mean[c] + a.momentum * (a.mean[c] + b);
            const float unbiased = a.M > 1 ? var * (Mf / (float)(a.M - 1)) : var;
            a.running_var[c] = (1.f - a.momentum) * a.running_var[c] + a.momentum * unbiased;
        }
    }
    if (a.pool_k > 1) {
        FwdApplyPooled<T, V, 8> f;
        f.y = (const T *)a.y, f.out = (T *)a.out, f.mean = a.mean, f.invstd = a.invstd, f.gamma = a.gamma, f.beta = a.beta;
        f.C = C, f.relu = a.relu, f.argmax = a.argmax, f.pool_k = a.pool_k, f.s_const = s_const, f.opitch = a.out_pitch;
        row_stream<V, 1>(f, a.M / a.pool_k, C);
    } else {
        FwdApply<T, V> f;
        f.y = (const T *)a.y, f.out = (T *)a.out, f.mean = a.mean, f.invstd = a.invstd, f.gamma = a.gamma, f.beta = a.beta;
        f.C = C, f.relu = a.relu, f.s_const = s_const, f.opitch = a.out_pitch;
        row_stream<V, PCB_BN_U>(f, a.M, C);
    }
}

// Pooled layers (max over pool_k consecutive rows): a.rs row lanes share one pooling group, each takes a contiguous
// slice of its rows, the lanes' (max, first index) pairs are merged in lane order through shared memory -- same result
// as one thread walking the whole group, a.rs times the threads (the last layers of the set-abstraction MLPs have only
// 256 .. 4096 groups: one thread per (group, 8 channels) left 16 .. 64 CTAs on 148 SMs).  C / V <= kBnThreads.
template <typename T, int V>
__global__ void __launch_bounds__(kBnThreads, PCB_BN_APPLY_MINB)
bn_apply_pooled_kernel(const BnFwdArgs a)
{
    pdl_wait();
    pdl_trigger();
    constexpr int KB = 8;
    const int C = a.C;
    __shared__ float s_const[2 * kBnThreads * V];
    __shared__ float s_best[kBnThreads][V];
    __shared__ float s_yb[kBnThreads][V];
    __shared__ int s_bi[kBnThreads][V];
    fwd_apply_consts(a, s_const);
    const int TX = C / V, TY = kBnThreads / TX, rs = a.rs;
    const int gpi = TY / rs;                                // pooling groups per CTA iteration
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int gl = ty / rs, sl = ty - gl * rs;
    const bool lane_ok = ty < gpi * rs;
    const int c = tx * V;
    const int pool_k = a.pool_k, kp = (pool_k + rs - 1) / rs;
    const int k_lo = sl * kp, k_hi = k_lo + kp < pool_k ? k_lo + kp : pool_k;
    const int64_t groups = a.M / pool_k;
    const T *y = (const T *)a.y;
    T *out = (T *)a.out;
    float sc[V], sh[V];
#pragma unroll
    for (int i = 0; i < V; ++i) sc[i] = s_const[c + i], sh[i] = s_const[C + c + i];
    for (int64_t g0 = (int64_t)blockIdx.x * gpi; g0 < groups; g0 += (int64_t)gridDim.x * gpi) {
        const int64_t g = g0 + gl;
        const bool act = lane_ok && g < groups;
        float best[V], yb[V];
        int bi[V];
#pragma unroll
        for (int i = 0; i < V; ++i) best[i] = -INFINITY, yb[i] = 0.f, bi[i] = k_lo;
        if (act) {
            const T *row = y + (g * pool_k) * C + c;
            for (int k0 = k_lo; k0 < k_hi; k0 += KB) {
                typename VecIO<T, V>::Raw t[KB];
#pragma unroll
                for (int j = 0; j < KB; ++j)
                    if (k0 + j < k_hi) t[j] = VecIO<T, V>::ldraw(row + (size_t)(k0 + j) * C);
#pragma unroll
                for (int j = 0; j < KB; ++j)
                    if (k0 + j < k_hi) {
                        float v[V];
                        VecIO<T, V>::cvt(t[j], v);
#pragma unroll
                        for (int i = 0; i < V; ++i) {
                            float z = fmaf(v[i], sc[i], sh[i]);
                            if (a.relu) z = fmaxf(z, 0.f);
                            if (k0 + j == k_lo || z > best[i]) best[i] = z, yb[i] = v[i], bi[i] = k0 + j;
                        }
                    }
            }
        }
        if (rs > 1) {
            if (act && sl > 0) {
#pragma unroll
                for (int i = 0; i < V; ++i)
                    s_best[threadIdx.x][i] = best[i], s_yb[threadIdx.x][i] = yb[i], s_bi[threadIdx.x][i] = bi[i];
            }
            __syncthreads();
            if (act && sl == 0) {
                for (int j = 1; j < rs; ++j) {
                    const int o = threadIdx.x + j * TX;     // lane sl = j of the same group and channels
                    if (j * kp < pool_k) {
#pragma unroll
                        for (int i = 0; i < V; ++i)
                            if (s_best[o][i] > best[i]) best[i] = s_best[o][i], yb[i] = s_yb[o][i], bi[i] = s_bi[o][i];
                    }
                }
            }
        }
        if (act && sl == 0) {
            VecIO<T, V>::store(out + g * a.out_pitch + c, best);
            if (a.argmax) AmIO<V>::st(a.argmax + g * C + c, bi);
            if (a.ymax) VecIO<T, V>::store((T *)a.ymax + g * C + c, yb);       // exact: yb is the stored value of y
        }
        if (rs > 1) __syncthreads();
    }
}

// gy = gamma * invstd * (dy - sum_dy / M - yhat * sum_dy_yhat / M) from the masked gradient dy (gy may alias dy)
template <typename T, int V>
struct BwdApplyDy {
    struct Pack {
        typename VecIO<T, V>::Raw y, d;
    };
    const T *y, *dy;
    T *gy;
    const float *s_const;                                   // shared [nm | is | g_is | a0 | a1] x C
    int C;
    float nm[V], is[V], gi[V], a0[V], a1[V];
    __device__ __forceinline__ void init(int c)
    {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            nm[i] = s_const[c + i], is[i] = s_const[C + c + i], gi[i] = s_const[2 * C + c + i];
            a0[i] = s_const[3 * C + c + i], a1[i] = s_const[4 * C + c + i];
        }
    }
    __device__ __forceinline__ Pack load(int64_t r, int c) const
    {
        Pack p;
        p.y = VecIO<T, V>::ldraw(y + r * C + c);
        p.d = VecIO<T, V>::ldraw(dy + r * C + c);
        return p;
    }
    __device__ __forceinline__ void emit(const Pack &p, int64_t r, int c) const
    {
        float v[V], d[V], o[V];
        VecIO<T, V>::cvt(p.y, v);
        VecIO<T, V>::cvt(p.d, d);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float yh = fmaf(v[i], is[i], nm[i]);
            o[i] = gi[i] * fmaf(-yh, a1[i], d[i] - a0[i]);
        }
        VecIO<T, V>::store(gy + r * C + c, o);
    }
};

struct BnBwdApplyArgs {
    const void *dy, *y;
    void *gy;
    const float *mean, *invstd, *gamma, *sums;              // sums: [>= 2][C] = sum dy, sum dy*yhat
    int64_t M;
    int C, Cv;
    // deferred: group partials (0, sum dy*yhat, sum dy) [groups][3][C] of the data-gradient GEMM's first fold level;
    // every CTA adds them up in group order, CTA 0 writes sums_out [3][C] = (sum dy, sum dy*yhat, 0)
    const float *gparts;
    float *sums_out;
    int groups;
};

constexpr int kBnApplyMaxC = 1024;

template <typename T, int V>
__global__ void __launch_bounds__(kBnThreads, PCB_BN_APPLY_MINB)
bn_bwd_apply_rows_kernel(const BnBwdApplyArgs a)
{
    pdl_wait();
    pdl_trigger();
    const int C = a.C;
    __shared__ float s_const[5 * kBnApplyMaxC];
    const float invM = 1.f / (float)a.M;
    for (int c = threadIdx.x; c < C; c += kBnThreads) {
        const bool real = c < a.Cv;                         // pad channel: every constant zero -> gy = 0
        const float is = real ? a.invstd[c] : 0.f;
        s_const[c] = real ? -a.mean[c] * is : 0.f;
        s_const[C + c] = is;
        s_const[2 * C + c] = real ? is * a.gamma[c] : 0.f;
        float s0 = 0.f, s1 = 0.f;
        if (a.gparts) {
            float n;
            if (real) fold_group_parts<false>(a.gparts, a.groups, C, c, n, s1, s0);
            if (blockIdx.x == 0) a.sums_out[c] = s0, a.sums_out[C + c] = s1, a.sums_out[2 * C + c] = 0.f;
        } else if (real) {
            s0 = a.sums[c], s1 = a.sums[C + c];
        }
        s_const[3 * C + c] = s0 * invM;
        s_const[4 * C + c] = s1 * invM;
    }
    __syncthreads();
    BwdApplyDy<T, V> f;
    f.y = (const T *)a.y, f.dy = (const T *)a.dy, f.gy = (T *)a.gy, f.s_const = s_const, f.C = C;
    row_stream<V, PCB_BN_UB>(f, a.M, C);
}

// ---------------------------------------------------------------------------------------------
// Backward of a POOLED last layer without a cooperative launch.  Only the winning row of a pooling group receives a
// gradient, so the two BatchNorm sums need gz, argmax and the pre-activation of the winning rows (ymax, written by
// bn_apply_pooled_kernel) -- M / pool_k rows instead of M: a small sums kernel (<= kPoolSumCtas CTAs, each leaves
// its partial (0, sum dy*yhat, sum dy) like a fold group of the statistics GEMM), then ONE streaming pass
// (bn_pool_bwd_apply_kernel) whose CTAs add those partials up while they set up their constants.  y is read once.
// ---------------------------------------------------------------------------------------------
constexpr int kPoolSumCtas = 32;

struct BnPoolSumArgs {
    const void *gz, *ymax;
    const float *mean, *invstd, *gamma, *beta;
    float *gparts;                                         // [gridDim.x][3][C]
    int64_t groups, gz_pitch;
    int C, Cv, relu;
};

template <typename T, int V>
__global__ void __launch_bounds__(kBnThreads)
bn_pool_sums_kernel(const BnPoolSumArgs a)
{
    pdl_wait();
    pdl_trigger();
    const int C = a.C;
    const Lanes L(C, V);
    __shared__ float s_acc[2][kBnThreads][V];
    const T *gz = (const T *)a.gz, *ym = (const T *)a.ymax;
    float *out = a.gparts + (size_t)blockIdx.x * 3 * C;
    for (int cg0 = 0; cg0 < L.CV; cg0 += L.TX) {
        const int cgi = cg0 + L.tx, c = cgi * V;
        float acc[2][V];
#pragma unroll
        for (int i = 0; i < V; ++i) acc[0][i] = acc[1][i] = 0.f;
        if (L.ty < L.TY && cgi < L.CV) {
            float nm[V], is[V], sc[V], sh[V];
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const bool real = c + i < a.Cv;
                const float m = real ? a.mean[c + i] : 0.f;
                is[i] = real ? a.invstd[c + i] : 0.f;
                nm[i] = -m * is[i];
                sc[i] = real ? is[i] * a.gamma[c + i] : 0.f;  // same expressions as the forward pass
                sh[i] = real ? a.beta[c + i] - m * sc[i] : 0.f;
            }
            const int64_t stride = (int64_t)gridDim.x * L.TY;
            for (int64_t g = (int64_t)blockIdx.x * L.TY + L.ty; g < a.groups; g += 4 * stride) {
                typename VecIO<T, V>::Raw rg[4], ry[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (g + j * stride < a.groups) {
                        rg[j] = VecIO<T, V>::ldraw(gz + (g + j * stride) * a.gz_pitch + c);
                        ry[j] = VecIO<T, V>::ldraw(ym + (g + j * stride) * C + c);
                    }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (g + j * stride < a.groups) {
                        float gv[V], yv[V];
                        VecIO<T, V>::cvt(rg[j], gv);
                        VecIO<T, V>::cvt(ry[j], yv);
#pragma unroll
                        for (int i = 0; i < V; ++i) {
                            const float yh = fmaf(yv[i], is[i], nm[i]);
                            const float z = fmaf(yv[i], sc[i], sh[i]);
                            const float dy = (!a.relu || z > 0.f) ? gv[i] : 0.f;
                            acc[0][i] += dy;
                            acc[1][i] = fmaf(dy, yh, acc[1][i]);
                        }
                    }
            }
        }
#pragma unroll
        for (int i = 0; i < V; ++i) s_acc[0][threadIdx.x][i] = acc[0][i], s_acc[1][threadIdx.x][i] = acc[1][i];
        __syncthreads();
        if (L.ty == 0 && cgi < L.CV) {
#pragma unroll
            for (int i = 0; i < V; ++i) {
                float s0 = 0.f, s1 = 0.f;
                for (int y = 0; y < L.TY; ++y) s0 += s_acc[0][y * L.TX + L.tx][i], s1 += s_acc[1][y * L.TX + L.tx][i];
                const bool real = c + i < a.Cv;
                out[c + i] = 0.f, out[C + c + i] = real ? s1 : 0.f, out[2 * C + c + i] = real ? s0 : 0.f;
            }
        }
        __syncthreads();
    }
}

struct BnPoolBwdArgs {
    BnBwdArgs b;                                           // gz, y, argmax, gy, statistics; work -> sums_out [3][C]
    const float *gparts;
    int groups;
};

template <typename T, int V>
__global__ void __launch_bounds__(kBnThreads, PCB_BN_APPLY_MINB)
bn_pool_bwd_apply_kernel(const BnPoolBwdArgs p)
{
    pdl_wait();
    pdl_trigger();
    const BnBwdArgs &a = p.b;
    const int C = a.C;
    __shared__ float s_const[6 * kBnApplyMaxC];
    const float invM = 1.f / (float)a.M;
    for (int c = threadIdx.x; c < C; c += kBnThreads) {
        const bool real = c < a.Cv;                         // pad channel: every constant zero -> gy = 0
        const float m = real ? a.mean[c] : 0.f, is = real ? a.invstd[c] : 0.f;
        const float sc = real ? is * a.gamma[c] : 0.f;
        float n, s0 = 0.f, s1 = 0.f;
        if (real) fold_group_parts<false>(p.gparts, p.groups, C, c, n, s1, s0);
        if (blockIdx.x == 0) a.work[c] = s0, a.work[C + c] = s1, a.work[2 * C + c] = 0.f;
        s_const[c] = -m * is;
        s_const[C + c] = is;
        s_const[2 * C + c] = sc;
        s_const[3 * C + c] = real ? a.beta[c] - m * sc : 0.f;
        s_const[4 * C + c] = s0 * invM;
        s_const[5 * C + c] = s1 * invM;
    }
    __syncthreads();
    BwdGroups<T, V, true, 4> f;
    bwd_fill<T, V>(f, a);
    f.s_const = s_const;
    f.rs = a.rs, f.kp = (a.pool_k + a.rs - 1) / a.rs;
    row_stream<V, 1>(f, a.M / a.pool_k * a.rs, C);
}

// Row slices per pooling group: the smallest power of two that puts >= 512 threads' worth of (unit, channel group)
// pairs on every SM (2 resident CTAs of 256 threads), while a slice keeps >= min_rows rows (loads in flight per thread) -- at most 16.
static int pool_row_slices(int64_t groups, int CV, int pool_k, int min_rows)
{
    static const int forced = [] { const char *e = getenv("PCB_POOL_RS"); return e ? atoi(e) : 0; }();   // experiments
    if (forced > 0) return forced <= pool_k ? forced : 1;
    int rs = 1;
    while (rs < 16 && pool_k / (rs * 2) >= min_rows && groups * rs * CV < (int64_t)PCB_NUM_SMS * 512) rs *= 2;
    return rs;
}

// CTAs of `kernel` that are resident at once (occupancy query cached per device): the elementwise kernels loop with a
// grid stride, and every CTA pays a start-up cost (per-channel constants, with deferred statistics a fold over the
// group partials: ~2 us) -- one wave of persistent CTAs pays it once
template <typename K>
static int resident_ctas(K kernel, int (&occ_cache)[kMaxDevices])
{
    int dev = 0;
    cudaGetDevice(&dev);
    int occ = (dev >= 0 && dev < kMaxDevices) ? occ_cache[dev] : 0;
    if (occ == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kBnThreads, 0) != cudaSuccess || occ < 1) occ = 1;
        if (dev >= 0 && dev < kMaxDevices) occ_cache[dev] = occ;
    }
    return occ * PCB_NUM_SMS;
}

// grid of an elementwise launch: every thread row gets work, at most `cap` CTAs (one resident wave)
static int stream_grid(int64_t units, int C, int V, int cap)
{
    const int CV = C / V;
    const int TX = CV < kBnThreads ? CV : kBnThreads;
    const int TY = kBnThreads / TX;
    const int64_t want = ceil_div(units, (int64_t)TY * 4);
    const int64_t g = want < cap ? want : cap;
    return g < 1 ? 1 : (int)g;
}

template <typename T, int V>
static void bn_apply_launch(const BnFwdArgs &a, int64_t units, cudaStream_t st)
{
    static int occ[kMaxDevices] = {};
    const int cap = resident_ctas(bn_apply_rows_kernel<T, V>, occ);
    launch_pdl(bn_apply_rows_kernel<T, V>, dim3(stream_grid(units, a.C, V, cap)), dim3(kBnThreads), 0, st, a);
}

template <typename T, int V>
static void bn_apply_pooled_launch(BnFwdArgs a, int64_t units, cudaStream_t st)
{
    static int occ[kMaxDevices] = {};
    const int cap = resident_ctas(bn_apply_pooled_kernel<T, V>, occ);
    // row lanes: a power of two that divides into the CTA's thread rows
    const int TY = kBnThreads / (a.C / V);
    int rs = pool_row_slices(units, a.C / V, a.pool_k, 2);
    while (rs > TY) rs >>= 1;
    a.rs = rs;
    const int64_t want = ceil_div(units, TY / rs);
    launch_pdl(bn_apply_pooled_kernel<T, V>, dim3((unsigned)(want < cap ? want : cap)), dim3(kBnThreads), 0, st, a);
}

template <typename T, int V>
static void bn_pool_bwd_launch(const BnPoolSumArgs &sa, const BnPoolBwdArgs &pa, int sgrid, int64_t units, cudaStream_t st)
{
    static int occ[kMaxDevices] = {};
    launch_pdl(bn_pool_sums_kernel<T, V>, dim3((unsigned)sgrid), dim3(kBnThreads), 0, st, sa);
    const int cap = resident_ctas(bn_pool_bwd_apply_kernel<T, V>, occ);
    launch_pdl(bn_pool_bwd_apply_kernel<T, V>, dim3(stream_grid(units * 4, pa.b.C, V, cap)), dim3(kBnThreads), 0, st, pa);
}

template <typename T, int V>
static void bn_bwd_apply_launch(const BnBwdApplyArgs &a, cudaStream_t st)
{
    static int occ[kMaxDevices] = {};
    const int cap = resident_ctas(bn_bwd_apply_rows_kernel<T, V>, occ);
    launch_pdl(bn_bwd_apply_rows_kernel<T, V>, dim3(stream_grid(a.M, a.C, V, cap)), dim3(kBnThreads), 0, st, a);
}

// ---------------------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------------------
// SMs the cooperative grids are sized for (0 = all).  A step runner that overlaps a long-running kernel of another
// stream with the step (the FPS chain of the next batch: one persistent CTA per cloud for ~0.3 ms) leaves that many
// SMs out: a cooperative grid needs ALL of its CTAs resident at once, so a grid sized for every SM would stall each of
// the ~70 cooperative launches of a step until the other stream's kernel has finished.
static int g_coop_sms = 0;

// co-resident CTAs of `kernel` on the current device (cooperative launch limit), capped at
// PCB_BN_CTAS_PER_SM per SM (default 2) and kBnMaxParts in total; the occupancy query is cached per device
template <typename K>
static int coop_capacity(K kernel, int (&occ_cache)[kMaxDevices])
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int occ = (dev >= 0 && dev < kMaxDevices) ? occ_cache[dev] : 0;
    if (occ == 0) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kBnThreads, 0);
        const char *e = getenv("PCB_BN_CTAS_PER_SM");
        const int cap = e ? atoi(e) : 2;
        if (cap >= 1 && occ > cap) occ = cap;
        if (occ < 1) occ = 1;
        if (dev >= 0 && dev < kMaxDevices) occ_cache[dev] = occ;
    }
    const char *es = getenv("PCB_BN_SMS");
    int limit = g_coop_sms > 0 ? g_coop_sms : (es ? atoi(es) : 0);
    if (limit >= 1 && limit < sms) sms = limit;
    int g = occ * sms;
    if (g > kBnMaxParts) g = kBnMaxParts;
    return g < 1 ? 1 : g;
}

struct Plan {
    int64_t upc;      // units per CTA in the reduction phase
    int nparts;       // CTAs that take part in it
    int grid;
};

static Plan make_plan(int64_t units, int C, int V, int grid_cap)
{
    const int CV = C / V;
    const int TX = CV < kBnThreads ? CV : kBnThreads;
    const int TY = kBnThreads / TX;
    Plan p;
    p.upc = ceil_div(units, (int64_t)grid_cap);
    if (p.upc < TY) p.upc = TY;                           // at least one unit per row lane
    p.nparts = (int)ceil_div(units, p.upc);
    // the elementwise phase wants every SM busy even when the reduction has few parts
    const int64_t want = ceil_div(units, (int64_t)TY);
    p.grid = (int)(want < grid_cap ? want : grid_cap);
    if (p.grid < p.nparts) p.grid = p.nparts;
    return p;
}

template <typename K, typename A>
static int coop_launch(K kernel, int grid, const A &args, cudaStream_t st)
{
    void *params[] = {(void *)&args};
    return (int)cudaLaunchCooperativeKernel((const void *)kernel, dim3((unsigned)grid), dim3(kBnThreads), params, 0, st);
}

template <typename T, int V>
static int bn_fwd_launch(BnFwdArgs a, cudaStream_t st)
{
    static int occ_cache[kMaxDevices] = {};
    const int cap = coop_capacity(bn_fwd_fused_kernel<T, V>, occ_cache);
    const Plan p = make_plan(a.M, a.C, V, cap);           // statistics are over rows, pooled or not
    a.upc = p.upc, a.nparts = p.nparts;
    int grid = p.grid;
    if (a.pool_k > 1) {
        const Plan q = make_plan(a.M / a.pool_k, a.C, V, cap);
        grid = q.grid > p.nparts ? q.grid : p.nparts;
    }
    return coop_launch(bn_fwd_fused_kernel<T, V>, grid, a, st);
}

template <typename T, int V>
static int bn_bwd_launch(BnBwdArgs a, cudaStream_t st)
{
    static int occ_cache[kMaxDevices] = {};
    const int cap = coop_capacity(bn_bwd_fused_kernel<T, V>, occ_cache);
    a.rs = pool_row_slices(a.M / a.pool_k, a.C / V, a.pool_k, 4);
    const Plan p = make_plan(a.M / a.pool_k * a.rs, a.C, V, cap);
    a.upc = p.upc, a.nparts = p.nparts;
    return coop_launch(bn_bwd_fused_kernel<T, V>, p.grid, a, st);
}

}  // namespace pcb

using namespace pcb;

#define PCB_BN_CHECK(M, C, pool_k)                                       \
    PCB_REQUIRE((M) > 0 && (C) > 0, PCB_EINVAL);                         \
    PCB_REQUIRE((C) % 4 == 0 && (dtype == 0 || dtype == 1), PCB_ERANGE); \
    PCB_REQUIRE((pool_k) >= 1 && (pool_k) <= 255 && (M) % (pool_k) == 0, PCB_ERANGE)

static inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

#ifdef PCB_BN_TRACE
PCB_API int pcb_bn_debug_trace(unsigned long long *host_out)
{
    return (int)cudaMemcpyFromSymbol(host_out, g_bn_trace, sizeof(g_bn_trace));
}
#endif

PCB_API int64_t pcb_bn_work_floats(int C) { return 3 * (int64_t)C * (1 + kBnMaxParts); }

PCB_API int pcb_bn_set_coop_sms(int sms)
{
    g_coop_sms = sms > 0 ? sms : 0;
    return 0;
}

PCB_API int pcb_bn_fwd_rows(const void *y, int dtype, int64_t M, int C, int Cv, int pool_k, const float *bias,
                            const float *gamma, const float *beta, float eps, float momentum, float *running_mean,
                            float *running_var, int relu, float *mean, float *invstd, void *out, int64_t out_pitch,
                            unsigned char *argmax, float *work, pcb_stream_t stream)
{
    PCB_REQUIRE(y && gamma && beta && mean && invstd && out && work, PCB_EINVAL);
    PCB_BN_CHECK(M, C, pool_k);
    PCB_REQUIRE(Cv > 0 && Cv <= C && (Cv == C || C <= 1024), PCB_ERANGE);   // padded rows use the shared-memory constants
    PCB_REQUIRE(!running_mean || running_var, PCB_EINVAL);
    BnFwdArgs a = {};
    a.y = y, a.out = out, a.argmax = argmax, a.bias = bias, a.gamma = gamma, a.beta = beta;
    a.running_mean = running_mean, a.running_var = running_var;
    a.mean = mean, a.invstd = invstd, a.work = work, a.M = M, a.C = C, a.Cv = Cv, a.pool_k = pool_k, a.relu = relu;
    a.eps = eps, a.momentum = momentum, a.upc = 0, a.nparts = 0, a.out_pitch = out_pitch > 0 ? out_pitch : C;
    PCB_REQUIRE(a.out_pitch >= C && a.out_pitch % 4 == 0, PCB_ERANGE);
    cudaStream_t st = (cudaStream_t)stream;
    if (!dtype) return bn_fwd_launch<float, 4>(a, st);
    if (C % 8 == 0 && a.out_pitch % 8 == 0 && al16(y) && al16(out)) return bn_fwd_launch<__nv_bfloat16, 8>(a, st);
    return bn_fwd_launch<__nv_bfloat16, 4>(a, st);
}

PCB_API int pcb_bn_bwd_rows(const void *gz, int64_t gz_pitch, const void *y, const unsigned char *argmax, int dtype,
                            int64_t M, int C, int Cv, int pool_k, const float *mean, const float *invstd, const float *gamma, const float *beta,
                            int relu, float *work, void *gy, pcb_stream_t stream)
{
    PCB_REQUIRE(gz && y && mean && invstd && gamma && beta && work && gy, PCB_EINVAL);
    PCB_BN_CHECK(M, C, pool_k);
    PCB_REQUIRE(Cv > 0 && Cv <= C, PCB_ERANGE);
    PCB_REQUIRE(pool_k == 1 || argmax, PCB_EINVAL);
    BnBwdArgs a = {};
    a.gz = gz, a.y = y, a.argmax = argmax, a.gy = gy, a.mean = mean, a.invstd = invstd, a.gamma = gamma, a.beta = beta;
    a.work = work, a.M = M, a.C = C, a.Cv = Cv, a.pool_k = pool_k, a.relu = relu, a.upc = 0, a.nparts = 0;
    a.gz_pitch = gz_pitch > 0 ? gz_pitch : C;
    PCB_REQUIRE(a.gz_pitch >= C && a.gz_pitch % 4 == 0, PCB_ERANGE);
    cudaStream_t st = (cudaStream_t)stream;
    if (!dtype) return bn_bwd_launch<float, 4>(a, st);
    if (C % 8 == 0 && a.gz_pitch % 8 == 0 && al16(y) && al16(gz) && al16(gy)) return bn_bwd_launch<__nv_bfloat16, 8>(a, st);
    return bn_bwd_launch<__nv_bfloat16, 4>(a, st);
}

// Elementwise half of the forward pass with given statistics (mean / invstd / var from the GEMM epilogue):
// out = [max over pool_k rows of] act(BN(y)); ordinary launch.  With running_mean != NULL the first CTA also updates the
// running statistics from mean / var (bias: the conv bias that was folded out of y, may be NULL).
// Deferred statistics (gparts != NULL): mean / invstd / var are OUTPUTS -- every CTA merges the `groups` group partials
// (n, mean, M2) [groups][3][C] that pcb_linear_bn_stats_rows_bf16 left, CTA 0 writes the three arrays (eps as given).
// ymax (pooled layers, may be NULL): [M / pool_k][C] pre-activation y of the winning rows, for pcb_bn_pool_bwd_rows.
PCB_API int pcb_bn_apply_rows(const void *y, int dtype, int64_t M, int C, int Cv, int pool_k, float *mean,
                              float *invstd, const float *gamma, const float *beta, int relu, void *out,
                              int64_t out_pitch, unsigned char *argmax, float *var, const float *bias, float momentum,
                              float *running_mean, float *running_var, const float *gparts, int groups, float eps,
                              void *ymax, pcb_stream_t stream)
{
    PCB_REQUIRE(y && gamma && beta && mean && invstd && out, PCB_EINVAL);
    PCB_REQUIRE(!running_mean || (running_var && var), PCB_EINVAL);
    PCB_REQUIRE(!gparts || (var && groups >= 1), PCB_EINVAL);
    PCB_BN_CHECK(M, C, pool_k);
    PCB_REQUIRE(Cv > 0 && Cv <= C && (Cv == C || C <= 1024), PCB_ERANGE);
    BnFwdArgs a = {};
    a.y = y, a.out = out, a.argmax = argmax, a.gamma = gamma, a.beta = beta;
    a.mean = mean, a.invstd = invstd, a.M = M, a.C = C, a.Cv = Cv;
    a.pool_k = pool_k, a.relu = relu, a.out_pitch = out_pitch > 0 ? out_pitch : C;
    a.var = var, a.bias = bias, a.momentum = momentum, a.running_mean = running_mean, a.running_var = running_var;
    a.gparts = gparts, a.var_out = var, a.groups = groups, a.eps = eps, a.rs = 1, a.ymax = ymax;
    PCB_REQUIRE(a.out_pitch >= C && a.out_pitch % 4 == 0, PCB_ERANGE);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t units = M / pool_k;
    const bool v8 = dtype && C % 8 == 0 && a.out_pitch % 8 == 0 && al16(y) && al16(out);
    const int V = v8 ? 8 : 4;
    PCB_REQUIRE(!gparts || C <= kBnThreads * V, PCB_ERANGE);    // deferred statistics go through the shared constants
    PCB_REQUIRE(!ymax || (pool_k > 1 && C / V <= kBnThreads), PCB_ERANGE);   // only the row-lane kernel writes ymax
    if (pool_k > 1 && C / V <= kBnThreads) {
        if (!dtype) bn_apply_pooled_launch<float, 4>(a, units, st);
        else if (v8) bn_apply_pooled_launch<__nv_bfloat16, 8>(a, units, st);
        else bn_apply_pooled_launch<__nv_bfloat16, 4>(a, units, st);
        PCB_RETURN_LAUNCH_STATUS();
    }
    if (!dtype) bn_apply_launch<float, 4>(a, units, st);
    else if (v8) bn_apply_launch<__nv_bfloat16, 8>(a, units, st);
    else bn_apply_launch<__nv_bfloat16, 4>(a, units, st);
    PCB_RETURN_LAUNCH_STATUS();
}

// Elementwise half of the backward pass: gy = gamma * invstd * (dy - sums[0] / M - yhat * sums[1] / M) from the masked
// gradient dy and the column sums that the data-gradient GEMM's epilogue produced (pcb_dgrad_bn_rows_bf16).  gy may be dy.
// Deferred sums (gparts != NULL): `sums` [3][C] is an OUTPUT, folded from the `groups` group partials of that GEMM.
PCB_API int pcb_bn_bwd_apply_rows(const void *dy, const void *y, int dtype, int64_t M, int C, int Cv, const float *mean,
                                  const float *invstd, const float *gamma, float *sums, void *gy, const float *gparts,
                                  int groups, pcb_stream_t stream)
{
    PCB_REQUIRE(dy && y && mean && invstd && gamma && sums && gy, PCB_EINVAL);
    PCB_REQUIRE(M > 0 && C > 0 && C % 4 == 0 && C <= kBnApplyMaxC && (dtype == 0 || dtype == 1), PCB_ERANGE);
    PCB_REQUIRE(Cv > 0 && Cv <= C && (!gparts || groups >= 1), PCB_ERANGE);
    BnBwdApplyArgs a = {};
    a.dy = dy, a.y = y, a.gy = gy, a.mean = mean, a.invstd = invstd, a.gamma = gamma, a.sums = sums, a.M = M, a.C = C, a.Cv = Cv;
    a.gparts = gparts, a.sums_out = sums, a.groups = groups;
    cudaStream_t st = (cudaStream_t)stream;
    if (!dtype) bn_bwd_apply_launch<float, 4>(a, st);
    else if (C % 8 == 0 && al16(y) && al16(dy) && al16(gy)) bn_bwd_apply_launch<__nv_bfloat16, 8>(a, st);
    else bn_bwd_apply_launch<__nv_bfloat16, 4>(a, st);
    PCB_RETURN_LAUNCH_STATUS();
}

// Backward of a pooled last layer (pool_k > 1) in two ordinary launches: the BatchNorm sums from gz / ymax / the
// statistics alone (M / pool_k rows), then gy in one pass over y.  Same results as pcb_bn_bwd_rows up to summation
// order; work: >= 3 * C * (1 + 32) floats, work[0 : 3C] = (sum dy, sum dy * yhat, 0) on return.
PCB_API int pcb_bn_pool_bwd_rows(const void *gz, int64_t gz_pitch, const void *ymax, const void *y,
                                 const unsigned char *argmax, int dtype, int64_t M, int C, int Cv, int pool_k,
                                 const float *mean, const float *invstd, const float *gamma, const float *beta, int relu,
                                 float *work, void *gy, pcb_stream_t stream)
{
    PCB_REQUIRE(gz && ymax && y && argmax && mean && invstd && gamma && beta && work && gy, PCB_EINVAL);
    PCB_BN_CHECK(M, C, pool_k);
    PCB_REQUIRE(pool_k > 1 && Cv > 0 && Cv <= C && C <= kBnApplyMaxC, PCB_ERANGE);
    const int64_t gp = gz_pitch > 0 ? gz_pitch : C;
    PCB_REQUIRE(gp >= C && gp % 4 == 0, PCB_ERANGE);
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t groups = M / pool_k;
    const bool v8 = dtype && C % 8 == 0 && gp % 8 == 0 && al16(y) && al16(gz) && al16(gy) && al16(ymax);
    const int V = v8 ? 8 : 4;
    BnPoolSumArgs sa;
    sa.gz = gz, sa.ymax = ymax, sa.mean = mean, sa.invstd = invstd, sa.gamma = gamma, sa.beta = beta;
    sa.gparts = work + 3 * (size_t)C, sa.groups = groups, sa.gz_pitch = gp, sa.C = C, sa.Cv = Cv, sa.relu = relu;
    const int CV = C / V, TX = CV < kBnThreads ? CV : kBnThreads, TY = kBnThreads / TX;
    int64_t sgrid = ceil_div(groups, (int64_t)TY * 4);
    if (sgrid > kPoolSumCtas) sgrid = kPoolSumCtas;
    if (sgrid < 1) sgrid = 1;
    BnPoolBwdArgs pa = {};
    BnBwdArgs &a = pa.b;
    a.gz = gz, a.y = y, a.argmax = argmax, a.gy = gy, a.mean = mean, a.invstd = invstd, a.gamma = gamma, a.beta = beta;
    a.work = work, a.M = M, a.C = C, a.Cv = Cv, a.pool_k = pool_k, a.relu = relu, a.gz_pitch = gp;
    a.rs = pool_row_slices(groups, C / V, pool_k, 4);
    pa.gparts = sa.gparts, pa.groups = (int)sgrid;
    const int64_t units = groups * a.rs;
    if (!dtype) bn_pool_bwd_launch<float, 4>(sa, pa, (int)sgrid, units, st);
    else if (v8) bn_pool_bwd_launch<__nv_bfloat16, 8>(sa, pa, (int)sgrid, units, st);
    else bn_pool_bwd_launch<__nv_bfloat16, 4>(sa, pa, (int)sgrid, units, st);
    PCB_RETURN_LAUNCH_STATUS();
}
