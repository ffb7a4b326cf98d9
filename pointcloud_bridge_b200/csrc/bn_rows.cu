// Training-mode BatchNorm + ReLU (+ max over the neighbour axis) on point-major rows [M, C] --
// the elementwise half of the shared MLP of the set-abstraction / feature-propagation modules
// (Conv2d 1x1 -> BatchNorm2d -> ReLU -> torch.max over nsample:
//  Partsize-identical/models/pointnet_util.py:213-217, 273-279, 343-345;
//  Highway_bridge/models/pointnet2_utils.py:150-154, 353-356).
// The training GEMM stays a library call in round 1; these kernels replace PyTorch's separate
// statistics / transform / ReLU / max / threshold_backward / bias-sum passes with
//   forward : stats (1 read) -> fold -> finalize (C threads) -> apply [+ReLU] [+max over K] (1 read, 1 write)
//   backward: reduce (2 reads) -> fold -> apply (2 reads, 1 write)
// All HBM-bound.  Activations may be fp32 or bf16 (autocast); statistics, affine parameters and
// all arithmetic are fp32.  The convolution bias is folded in here (BN(xW + b) only needs b for
// the running mean), so no separate bias-add or bias-gradient pass exists.
// Column reductions are two-stage: every CTA stores its partial sums to parts[cta][NACC*C] with
// plain stores and a small second kernel adds the <= 592 partials per column (one atomicAdd per
// CTA per column serialised ~1000 same-address atomics in L2 and cost more than the read).
#include <cuda_bf16.h>

#include "pcb_common.cuh"

namespace pcb {

constexpr int kBnThreads = 256;
constexpr int kBnMaxParts = PCB_NUM_SMS * 4;

// VecIO<T, V>: V consecutive channels per thread, one 16-byte access for (float,4) and (bf16,8),
// one 8-byte access for (bf16,4).
template <typename T, int V>
struct VecIO;
template <>
struct VecIO<float, 4> {
    static __device__ __forceinline__ void load(const float *p, float v[4])
    {
        float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    }
    static __device__ __forceinline__ void store(float *p, const float v[4])
    {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <>
struct VecIO<__nv_bfloat16, 4> {
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float v[4])
    {
        uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
        float2 fa = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&t.x));
        float2 fb = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&t.y));
        v[0] = fa.x, v[1] = fa.y, v[2] = fb.x, v[3] = fb.y;
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
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float v[8])
    {
        uint4 t = __ldg(reinterpret_cast<const uint4 *>(p));
        unsigned w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&w[i]));
            v[2 * i] = f.x, v[2 * i + 1] = f.y;
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

// ---------------------------------------------------------------------------------------------
// Column reduction skeleton.  threadIdx.x % TX walks the C/V channel groups, threadIdx.x / TX
// walks rows; a CTA covers `rows_per_cta` consecutive rows and stores NACC*C partial sums.
// f(row, c, acc[NACC][V]) accumulates the V channels of one row.
// ---------------------------------------------------------------------------------------------
template <int NACC, int V, typename F>
__device__ __forceinline__ void column_reduce(int64_t M, int C, int64_t rows_per_cta, float *__restrict__ parts, F f)
{
    __shared__ float s_part[NACC][kBnThreads][V];
    const int CV = C / V;
    const int TX = CV < kBnThreads ? CV : kBnThreads;
    const int TY = kBnThreads / TX;
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
    // channel groups of one row sit on lanes tx, tx+TX, ...: when TX is a power of two <= 32 the
    // row partials of a warp are folded with shuffles first
    const bool warp_fold = TX <= 32 && (TX & (TX - 1)) == 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *my_parts = parts + (size_t)blockIdx.x * NACC * C;
    for (int cg = tx; cg < CV; cg += TX) {
        float acc[NACC][V];
#pragma unroll
        for (int a = 0; a < NACC; ++a)
#pragma unroll
            for (int v = 0; v < V; ++v) acc[a][v] = 0.f;
        if (ty < TY) {
#pragma unroll 4
            for (int64_t r = r0 + ty; r < r1; r += TY) f(r, cg * V, acc);
        }
        if (warp_fold) {
            for (int off = TX; off < 32; off <<= 1)
#pragma unroll
                for (int a = 0; a < NACC; ++a)
#pragma unroll
                    for (int v = 0; v < V; ++v) acc[a][v] += __shfl_xor_sync(PCB_FULL_MASK, acc[a][v], off);
            if (lane < TX) {
#pragma unroll
                for (int a = 0; a < NACC; ++a)
#pragma unroll
                    for (int v = 0; v < V; ++v) s_part[a][warp * 32 + lane][v] = acc[a][v];
            }
            __syncthreads();
            if (threadIdx.x < TX) {
#pragma unroll
                for (int a = 0; a < NACC; ++a)
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        float s = 0.f;
#pragma unroll
                        for (int w = 0; w < kBnThreads / 32; ++w) s += s_part[a][w * 32 + threadIdx.x][v];
                        my_parts[(size_t)a * C + cg * V + v] = s;
                    }
            }
        } else {
#pragma unroll
            for (int a = 0; a < NACC; ++a)
#pragma unroll
                for (int v = 0; v < V; ++v) s_part[a][threadIdx.x][v] = acc[a][v];
            __syncthreads();
            if (ty == 0) {
#pragma unroll
                for (int a = 0; a < NACC; ++a)
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        float s = 0.f;
                        for (int y = 0; y < TY; ++y) s += s_part[a][y * TX + tx][v];
                        my_parts[(size_t)a * C + cg * V + v] = s;
                    }
            }
        }
        __syncthreads();
    }
}

// sums[c] = sum over parts of parts[part][c]: 32 columns x 8 part-lanes per block (the sum over
// <= 592 partials is latency-bound, so it is spread over 8 threads per column, 4 loads in flight)
__global__ void __launch_bounds__(256)
bn_fold_parts_kernel(const float *__restrict__ parts, int nparts, int ncols, float *__restrict__ sums)
{
    __shared__ float s_acc[8][33];
    const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (c < ncols) {
        int i = pl;
        for (; i + 24 < nparts; i += 32) {
            s0 += parts[(size_t)i * ncols + c];
            s1 += parts[(size_t)(i + 8) * ncols + c];
            s2 += parts[(size_t)(i + 16) * ncols + c];
            s3 += parts[(size_t)(i + 24) * ncols + c];
        }
        for (; i < nparts; i += 8) s0 += parts[(size_t)i * ncols + c];
    }
    s_acc[pl][cl] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (pl == 0 && c < ncols) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s += s_acc[j][cl];
        sums[c] = s;
    }
}

// stats: parts -> sum_rows (y - y[0]) and sum_rows (y - y[0])^2 (shifted: no cancellation when
// |mean| >> std)
template <typename T, int V>
__global__ void __launch_bounds__(kBnThreads)
bn_stats_kernel(const T *__restrict__ y, int64_t M, int C, int64_t rows_per_cta, float *__restrict__ parts)
{
    float s[V];
    int s_c = -1;
    column_reduce<2, V>(M, C, rows_per_cta, parts, [&](int64_t r, int c, float acc[2][V]) {
        float v[V];
        VecIO<T, V>::load(y + r * C + c, v);
        if (c != s_c) {                                   // shift = first row, loaded once per channel group
            VecIO<T, V>::load(y + c, s);
            s_c = c;
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const float d = v[i] - s[i];
            acc[0][i] += d;
            acc[1][i] += d * d;
        }
    });
}

// finalize: mean / invstd of y from the shifted sums; running statistics update (momentum,
// unbiased variance, conv bias added to the running mean) as torch.nn.functional.batch_norm does.
template <typename T>
__global__ void bn_finalize_kernel(const float *__restrict__ sums, const T *__restrict__ y, const float *__restrict__ bias,
                                   int64_t M, int C, float eps, float momentum, float *__restrict__ running_mean,
                                   float *__restrict__ running_var, float *__restrict__ mean, float *__restrict__ invstd)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float shift = (float)y[c];
    const float m1 = sums[c] / (float)M;
    const float var = fmaxf(sums[C + c] / (float)M - m1 * m1, 0.f);
    const float mu = shift + m1;                          // mean of the bias-free pre-activation
    mean[c] = mu;
    invstd[c] = rsqrtf(var + eps);
    if (running_mean) {
        const float b = bias ? bias[c] : 0.f;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (mu + b);
        const float unbiased = M > 1 ? var * ((float)M / (float)(M - 1)) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
}

// apply: z = act((y - mean) * invstd * gamma + beta); pool_k > 1: out[r] = max_k z[r*pool_k + k]
// with the winning k (first on ties, as torch.max) stored for the backward pass.
template <typename T, int V>
__global__ void __launch_bounds__(kBnThreads)
bn_apply_kernel(const T *__restrict__ y, unsigned total, int C, FastDiv dCV, int pool_k, const float *__restrict__ mean,
                const float *__restrict__ invstd, const float *__restrict__ gamma, const float *__restrict__ beta,
                int relu, T *__restrict__ out, unsigned char *__restrict__ argmax)
{
    const unsigned t = blockIdx.x * kBnThreads + threadIdx.x;
    if (t >= total) return;
    const int64_t r = dCV.div(t);
    const int c = (int)(t - (unsigned)r * dCV.d) * V;
    float sc[V], sh[V], best[V];
    int bi[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        sc[i] = invstd[c + i] * gamma[c + i];
        sh[i] = beta[c + i] - mean[c + i] * sc[i];
        bi[i] = 0;
        best[i] = 0.f;
    }
#pragma unroll 4
    for (int k = 0; k < pool_k; ++k) {
        float v[V];
        VecIO<T, V>::load(y + (r * pool_k + k) * C + c, v);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            float z = fmaf(v[i], sc[i], sh[i]);
            if (relu) z = fmaxf(z, 0.f);
            if (k == 0 || z > best[i]) {
                best[i] = z;
                bi[i] = k;
            }
        }
    }
    VecIO<T, V>::store(out + r * C + c, best);
    if (argmax) {
        unsigned w[2] = {0u, 0u};
#pragma unroll
        for (int i = 0; i < V; ++i) w[i >> 2] |= (unsigned)(bi[i] & 0xff) << (8 * (i & 3));
        if (V == 8)
            *reinterpret_cast<uint2 *>(argmax + r * C + c) = make_uint2(w[0], w[1]);
        else
            *reinterpret_cast<unsigned *>(argmax + r * C + c) = w[0];
    }
}

// backward.  dy(row) = gz(row) * [act passes]  (pooled: gz of the group, only for the winning k)
template <typename T, int V>
__device__ __forceinline__ void load_dy(const T *__restrict__ gz, const T *__restrict__ y,
                                        const unsigned char *__restrict__ argmax, int64_t r, int c, int C, FastDiv dK,
                                        const float *__restrict__ mean, const float *__restrict__ invstd,
                                        const float *__restrict__ gamma, const float *__restrict__ beta, int relu,
                                        float dy[V], float yh[V])
{
    float v[V], g[V];
    VecIO<T, V>::load(y + r * C + c, v);
    const int pool_k = (int)dK.d;
    const int64_t rg = pool_k > 1 ? (int64_t)dK.div((unsigned)r) : r;
    VecIO<T, V>::load(gz + rg * C + c, g);
    const int k = (int)(r - rg * pool_k);
    unsigned char am[V];
    if (pool_k > 1) {                                      // V winners in one 4- or 8-byte load
        if (V == 8) {
            const uint2 t = __ldg(reinterpret_cast<const uint2 *>(argmax + rg * C + c));
            const unsigned w[2] = {t.x, t.y};
#pragma unroll
            for (int i = 0; i < V; ++i) am[i] = (unsigned char)(w[i >> 2] >> (8 * (i & 3)));
        } else {
            const unsigned t = __ldg(reinterpret_cast<const unsigned *>(argmax + rg * C + c));
#pragma unroll
            for (int i = 0; i < V; ++i) am[i] = (unsigned char)(t >> (8 * (i & 3)));
        }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
        yh[i] = (v[i] - mean[c + i]) * invstd[c + i];
        const float sc = invstd[c + i] * gamma[c + i];     // same expression as the forward pass
        const float z = fmaf(v[i], sc, beta[c + i] - mean[c + i] * sc);
        bool pass = !relu || z > 0.f;
        if (pool_k > 1) pass = pass && (k == (int)am[i]);
        dy[i] = pass ? g[i] : 0.f;
    }
}

// reduce: parts -> sum dy, sum dy * yhat, sum yhat
template <typename T, int V>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_reduce_kernel(const T *__restrict__ gz, const T *__restrict__ y, const unsigned char *__restrict__ argmax,
                     int64_t M, int C, FastDiv dK, int64_t rows_per_cta, const float *__restrict__ mean,
                     const float *__restrict__ invstd, const float *__restrict__ gamma, const float *__restrict__ beta,
                     int relu, float *__restrict__ parts)
{
    column_reduce<3, V>(M, C, rows_per_cta, parts, [&](int64_t r, int c, float acc[3][V]) {
        float dy[V], yh[V];
        load_dy<T, V>(gz, y, argmax, r, c, C, dK, mean, invstd, gamma, beta, relu, dy, yh);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            acc[0][i] += dy[i];
            acc[1][i] += dy[i] * yh[i];
            acc[2][i] += yh[i];
        }
    });
}

// apply: gy = gamma * invstd * (dy - sum_dy / M - yhat * sum_dy_yhat / M)
template <typename T, int V>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_apply_kernel(const T *__restrict__ gz, const T *__restrict__ y, const unsigned char *__restrict__ argmax,
                    int64_t M, unsigned total, int C, FastDiv dCV, FastDiv dK, const float *__restrict__ mean,
                    const float *__restrict__ invstd, const float *__restrict__ gamma, const float *__restrict__ beta,
                    int relu, const float *__restrict__ sums, T *__restrict__ gy)
{
    const unsigned t = blockIdx.x * kBnThreads + threadIdx.x;
    if (t >= total) return;
    const int64_t r = dCV.div(t);
    const int c = (int)(t - (unsigned)r * dCV.d) * V;
    float dy[V], yh[V], o[V];
    load_dy<T, V>(gz, y, argmax, r, c, C, dK, mean, invstd, gamma, beta, relu, dy, yh);
    const float invM = 1.f / (float)M;
#pragma unroll
    for (int i = 0; i < V; ++i)
        o[i] = gamma[c + i] * invstd[c + i] * (dy[i] - sums[c + i] * invM - yh[i] * sums[C + c + i] * invM);
    VecIO<T, V>::store(gy + r * C + c, o);
}

// rows per CTA such that at most kBnMaxParts CTAs (and partial rows) exist
static inline int64_t rows_per_cta_for(int64_t M)
{
    int64_t rpc = ceil_div(M, (int64_t)kBnMaxParts);
    return rpc < 64 ? 64 : rpc;
}

template <typename T, int V>
static int bn_stats_launch(const void *y, int64_t M, int C, float *sums, cudaStream_t st)
{
    const int64_t rpc = rows_per_cta_for(M);
    const int nparts = (int)ceil_div(M, rpc);
    float *parts = sums + 3 * (size_t)C;                  // scratch after the 3C result slots
    bn_stats_kernel<T, V><<<nparts, kBnThreads, 0, st>>>((const T *)y, M, C, rpc, parts);
    bn_fold_parts_kernel<<<(unsigned)ceil_div(2 * C, 32), 256, 0, st>>>(parts, nparts, 2 * C, sums);
    PCB_RETURN_LAUNCH_STATUS();
}

template <typename T, int V>
static int bn_apply_launch(const void *y, int64_t Mout, int C, int pool_k, const float *mean, const float *invstd,
                           const float *gamma, const float *beta, int relu, void *out, unsigned char *argmax,
                           cudaStream_t st)
{
    const unsigned total = (unsigned)(Mout * (C / V));
    bn_apply_kernel<T, V><<<(unsigned)ceil_div(total, kBnThreads), kBnThreads, 0, st>>>(
        (const T *)y, total, C, make_fastdiv(C / V), pool_k, mean, invstd, gamma, beta, relu, (T *)out, argmax);
    PCB_RETURN_LAUNCH_STATUS();
}

template <typename T, int V>
static int bn_bwd_launch(const void *gz, const void *y, const unsigned char *argmax, int64_t M, int C, int pool_k,
                         const float *mean, const float *invstd, const float *gamma, const float *beta, int relu,
                         float *sums, void *gy, cudaStream_t st)
{
    const int64_t rpc = rows_per_cta_for(M);
    const unsigned rblocks = (unsigned)ceil_div(M, rpc);
    float *parts = sums + 3 * (size_t)C;
    const unsigned total = (unsigned)(M * (C / V));
    const FastDiv dK = make_fastdiv(pool_k), dCV = make_fastdiv(C / V);
    bn_bwd_reduce_kernel<T, V><<<rblocks, kBnThreads, 0, st>>>((const T *)gz, (const T *)y, argmax, M, C, dK, rpc, mean,
                                                              invstd, gamma, beta, relu, parts);
    bn_fold_parts_kernel<<<(unsigned)ceil_div(3 * C, 32), 256, 0, st>>>(parts, (int)rblocks, 3 * C, sums);
    bn_bwd_apply_kernel<T, V><<<(unsigned)ceil_div(total, kBnThreads), kBnThreads, 0, st>>>(
        (const T *)gz, (const T *)y, argmax, M, total, C, dCV, dK, mean, invstd, gamma, beta, relu, sums, (T *)gy);
    PCB_RETURN_LAUNCH_STATUS();
}

}  // namespace pcb

using namespace pcb;

#define PCB_BN_CHECK(M, C)                                   \
    PCB_REQUIRE((M) > 0 && (C) > 0, PCB_EINVAL);             \
    PCB_REQUIRE((C) % 4 == 0 && (dtype == 0 || dtype == 1), PCB_ERANGE)

static inline bool al16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

PCB_API int pcb_bn_stats_rows(const void *y, int dtype, int64_t M, int C, float *sums, pcb_stream_t stream)
{
    PCB_REQUIRE(y && sums, PCB_EINVAL);
    PCB_BN_CHECK(M, C);
    cudaStream_t st = (cudaStream_t)stream;
    if (!dtype) return bn_stats_launch<float, 4>(y, M, C, sums, st);
    if (C % 8 == 0 && al16(y)) return bn_stats_launch<__nv_bfloat16, 8>(y, M, C, sums, st);
    return bn_stats_launch<__nv_bfloat16, 4>(y, M, C, sums, st);
}

PCB_API int pcb_bn_finalize(const float *sums, const void *y, int dtype, const float *bias, int64_t M, int C,
                            float eps, float momentum, float *running_mean, float *running_var, float *mean,
                            float *invstd, pcb_stream_t stream)
{
    PCB_REQUIRE(sums && y && mean && invstd, PCB_EINVAL);
    PCB_BN_CHECK(M, C);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned blocks = (unsigned)ceil_div(C, 128);
    if (dtype)
        bn_finalize_kernel<__nv_bfloat16><<<blocks, 128, 0, st>>>(sums, (const __nv_bfloat16 *)y, bias, M, C, eps, momentum,
                                                                running_mean, running_var, mean, invstd);
    else
        bn_finalize_kernel<float><<<blocks, 128, 0, st>>>(sums, (const float *)y, bias, M, C, eps, momentum, running_mean,
                                                        running_var, mean, invstd);
    PCB_RETURN_LAUNCH_STATUS();
}

PCB_API int pcb_bn_apply_rows(const void *y, int dtype, int64_t Mout, int C, int pool_k, const float *mean,
                              const float *invstd, const float *gamma, const float *beta, int relu, void *out,
                              unsigned char *argmax, pcb_stream_t stream)
{
    PCB_REQUIRE(y && mean && invstd && gamma && beta && out, PCB_EINVAL);
    PCB_BN_CHECK(Mout, C);
    PCB_REQUIRE(pool_k >= 1 && pool_k <= 255, PCB_ERANGE);
    PCB_REQUIRE(Mout * (C / 4) < (1ll << 31) && Mout * pool_k < (1ll << 31), PCB_ERANGE);
    cudaStream_t st = (cudaStream_t)stream;
    if (!dtype) return bn_apply_launch<float, 4>(y, Mout, C, pool_k, mean, invstd, gamma, beta, relu, out, argmax, st);
    if (C % 8 == 0 && al16(y) && al16(out))
        return bn_apply_launch<__nv_bfloat16, 8>(y, Mout, C, pool_k, mean, invstd, gamma, beta, relu, out, argmax, st);
    return bn_apply_launch<__nv_bfloat16, 4>(y, Mout, C, pool_k, mean, invstd, gamma, beta, relu, out, argmax, st);
}

PCB_API int pcb_bn_bwd_rows(const void *gz, const void *y, const unsigned char *argmax, int dtype, int64_t M, int C,
                            int pool_k, const float *mean, const float *invstd, const float *gamma, const float *beta,
                            int relu, float *sums, void *gy, pcb_stream_t stream)
{
    PCB_REQUIRE(gz && y && mean && invstd && gamma && beta && sums && gy, PCB_EINVAL);
    PCB_BN_CHECK(M, C);
    PCB_REQUIRE(pool_k >= 1 && pool_k <= 255 && (pool_k == 1 || argmax), PCB_ERANGE);
    PCB_REQUIRE(M * (C / 4) < (1ll << 31), PCB_ERANGE);
    cudaStream_t st = (cudaStream_t)stream;
    if (!dtype)
        return bn_bwd_launch<float, 4>(gz, y, argmax, M, C, pool_k, mean, invstd, gamma, beta, relu, sums, gy, st);
    if (C % 8 == 0 && al16(y) && al16(gz) && al16(gy))
        return bn_bwd_launch<__nv_bfloat16, 8>(gz, y, argmax, M, C, pool_k, mean, invstd, gamma, beta, relu, sums, gy, st);
    return bn_bwd_launch<__nv_bfloat16, 4>(gz, y, argmax, M, C, pool_k, mean, invstd, gamma, beta, relu, sums, gy, st);
}
