// Training-mode BatchNorm + ReLU (+ max over the neighbour axis) on point-major rows [M, C] --
// the elementwise half of the shared MLP of the set-abstraction / feature-propagation modules
// (Conv2d 1x1 -> BatchNorm2d -> ReLU -> torch.max over nsample:
//  Partsize-identical/models/pointnet_util.py:213-217, 273-279, 343-345;
//  Highway_bridge/models/pointnet2_utils.py:150-154, 353-356).
// The GEMM stays a library call in round 1; these kernels replace PyTorch's separate
// statistics / transform / ReLU / max / threshold_backward / bias-sum passes with
//   forward : stats (1 read)  ->  finalize (C threads)  ->  apply [+ReLU] [+max over K] (1 read, 1 write)
//   backward: reduce (2 reads) ->  apply (2 reads, 1 write)
// All HBM-bound: bytes per element are listed next to each entry point in include/pcbridge.h.
// Activations may be fp32 or bf16 (autocast); statistics, affine parameters and all arithmetic
// are fp32.  The convolution bias is folded in here (BN(xW + b) only needs b for the running
// mean), so no separate bias-add or bias-gradient pass exists.
#include <cuda_bf16.h>

#include "pcb_common.cuh"

namespace pcb {

constexpr int kBnThreads = 256;

template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
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
struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16 *p, float v[4])
    {
        uint2 t = __ldg(reinterpret_cast<const uint2 *>(p));
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162 *>(&t.x), b = *reinterpret_cast<__nv_bfloat162 *>(&t.y);
        float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
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

// Thread layout shared by the column-reduction kernels: threadIdx.x % TX walks the C/4 channel
// quads, threadIdx.x / TX walks rows; a CTA covers `rows_per_cta` consecutive rows.
struct Tile {
    int tx, ty, TX, TY;
};
__device__ __forceinline__ Tile make_tile(int C4)
{
    Tile t;
    t.TX = C4 < kBnThreads ? C4 : kBnThreads;
    t.TY = kBnThreads / t.TX;
    t.tx = threadIdx.x % t.TX;
    t.ty = threadIdx.x / t.TX;
    return t;
}

// ---------------------------------------------------------------------------------------------
// stats: sums[0:C] = sum_rows (y - y[0]), sums[C:2C] = sum_rows (y - y[0])^2   (shifted sums:
// no cancellation when |mean| >> std).  sums must be zero on entry.
// ---------------------------------------------------------------------------------------------
template <typename T, int NACC, typename F>
__device__ __forceinline__ void column_reduce(int64_t M, int C, int64_t rows_per_cta, float *__restrict__ sums, F f)
{
    // f(row, c, acc[NACC][4]) accumulates one row's 4 channels
    __shared__ float s_part[NACC][kBnThreads][4];
    const int C4 = C / 4;
    Tile t = make_tile(C4);
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = r0 + rows_per_cta < M ? r0 + rows_per_cta : M;
    for (int cq = t.tx; cq < C4; cq += t.TX) {
        float acc[NACC][4];
#pragma unroll
        for (int a = 0; a < NACC; ++a)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[a][v] = 0.f;
        if (t.ty < t.TY) {
#pragma unroll 4
            for (int64_t r = r0 + t.ty; r < r1; r += t.TY) f(r, cq * 4, acc);
        }
#pragma unroll
        for (int a = 0; a < NACC; ++a)
#pragma unroll
            for (int v = 0; v < 4; ++v) s_part[a][threadIdx.x][v] = acc[a][v];
        __syncthreads();
        if (t.ty == 0) {
#pragma unroll
            for (int a = 0; a < NACC; ++a)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    float s = 0.f;
                    for (int y = 0; y < t.TY; ++y) s += s_part[a][y * t.TX + t.tx][v];
                    atomicAdd(sums + (size_t)a * C + cq * 4 + v, s);
                }
        }
        __syncthreads();
    }
}

template <typename T>
__global__ void __launch_bounds__(kBnThreads)
bn_stats_kernel(const T *__restrict__ y, int64_t M, int C, int64_t rows_per_cta, float *__restrict__ sums)
{
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    int s_c = -1;
    column_reduce<T, 2>(M, C, rows_per_cta, sums, [&](int64_t r, int c, float acc[2][4]) {
        float v[4];
        Vec4<T>::load(y + r * C + c, v);
        if (c != s_c) {                                   // shift = first row, loaded once per channel quad
            Vec4<T>::load(y + c, s);
            s_c = c;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float d = v[i] - s[i];
            acc[0][i] += d;
            acc[1][i] += d * d;
        }
    });
}

// finalize: mean / invstd of (y + bias) from the shifted sums; running statistics update
// (momentum, unbiased variance) exactly as torch.nn.functional.batch_norm does.
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

// ---------------------------------------------------------------------------------------------
// apply: z = act((y - mean) * invstd * gamma + beta); pool_k > 1: out[r] = max_k z[r*pool_k + k]
// with the winning k (first on ties, as torch.max) stored for the backward pass.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kBnThreads)
bn_apply_kernel(const T *__restrict__ y, unsigned total, int C, FastDiv dC4, int pool_k, const float *__restrict__ mean,
                const float *__restrict__ invstd, const float *__restrict__ gamma, const float *__restrict__ beta,
                int relu, T *__restrict__ out, unsigned char *__restrict__ argmax)
{
    const unsigned t = blockIdx.x * kBnThreads + threadIdx.x;
    if (t >= total) return;
    const int64_t r = dC4.div(t);
    const int c = (int)(t - (unsigned)r * dC4.d) * 4;
    float sc[4], sh[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sc[i] = invstd[c + i] * gamma[c + i];
        sh[i] = beta[c + i] - mean[c + i] * sc[i];
    }
    float best[4];
    int bi[4] = {0, 0, 0, 0};
    for (int k = 0; k < pool_k; ++k) {
        float v[4];
        Vec4<T>::load(y + (r * pool_k + k) * C + c, v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float z = fmaf(v[i], sc[i], sh[i]);
            if (relu) z = fmaxf(z, 0.f);
            if (k == 0 || z > best[i]) {
                best[i] = z;
                bi[i] = k;
            }
        }
    }
    Vec4<T>::store(out + r * C + c, best);
    if (argmax) {
        uchar4 a = make_uchar4((unsigned char)bi[0], (unsigned char)bi[1], (unsigned char)bi[2], (unsigned char)bi[3]);
        *reinterpret_cast<uchar4 *>(argmax + r * C + c) = a;
    }
}

// ---------------------------------------------------------------------------------------------
// backward.  dy(row) = gz(row) * [act passes]  (pooled: gz of the group, only for the winning k)
//   reduce: sums[0:C] = sum dy, sums[C:2C] = sum dy * yhat, sums[2C:3C] = sum yhat
//   apply : gy = gamma * invstd * (dy - sum_dy / M - yhat * sum_dy_yhat / M)
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ void load_dy(const T *__restrict__ gz, const T *__restrict__ y,
                                        const unsigned char *__restrict__ argmax, int64_t r, int c, int C, FastDiv dK,
                                        const float *__restrict__ mean, const float *__restrict__ invstd,
                                        const float *__restrict__ gamma, const float *__restrict__ beta, int relu,
                                        float dy[4], float yh[4])
{
    float v[4], g[4];
    Vec4<T>::load(y + r * C + c, v);
    const int pool_k = (int)dK.d;
    const int64_t rg = pool_k > 1 ? (int64_t)dK.div((unsigned)r) : r;
    Vec4<T>::load(gz + rg * C + c, g);
    uchar4 am = make_uchar4(0, 0, 0, 0);
    int k = 0;
    if (pool_k > 1) {
        am = *reinterpret_cast<const uchar4 *>(argmax + rg * C + c);
        k = (int)(r - rg * pool_k);
    }
    const unsigned char amv[4] = {am.x, am.y, am.z, am.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        yh[i] = (v[i] - mean[c + i]) * invstd[c + i];
        const float sc = invstd[c + i] * gamma[c + i];     // same expression as the forward pass
        const float z = fmaf(v[i], sc, beta[c + i] - mean[c + i] * sc);
        bool pass = !relu || z > 0.f;
        if (pool_k > 1) pass = pass && (k == (int)amv[i]);
        dy[i] = pass ? g[i] : 0.f;
    }
}

template <typename T>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_reduce_kernel(const T *__restrict__ gz, const T *__restrict__ y, const unsigned char *__restrict__ argmax,
                     int64_t M, int C, FastDiv dK, int64_t rows_per_cta, const float *__restrict__ mean,
                     const float *__restrict__ invstd, const float *__restrict__ gamma, const float *__restrict__ beta,
                     int relu, float *__restrict__ sums)
{
    column_reduce<T, 3>(M, C, rows_per_cta, sums, [&](int64_t r, int c, float acc[3][4]) {
        float dy[4], yh[4];
        load_dy<T>(gz, y, argmax, r, c, C, dK, mean, invstd, gamma, beta, relu, dy, yh);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            acc[0][i] += dy[i];
            acc[1][i] += dy[i] * yh[i];
            acc[2][i] += yh[i];
        }
    });
}

template <typename T>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_apply_kernel(const T *__restrict__ gz, const T *__restrict__ y, const unsigned char *__restrict__ argmax,
                    int64_t M, unsigned total, int C, FastDiv dC4, FastDiv dK, const float *__restrict__ mean,
                    const float *__restrict__ invstd, const float *__restrict__ gamma, const float *__restrict__ beta,
                    int relu, const float *__restrict__ sums, T *__restrict__ gy)
{
    const unsigned t = blockIdx.x * kBnThreads + threadIdx.x;
    if (t >= total) return;
    const int64_t r = dC4.div(t);
    const int c = (int)(t - (unsigned)r * dC4.d) * 4;
    float dy[4], yh[4], o[4];
    load_dy<T>(gz, y, argmax, r, c, C, dK, mean, invstd, gamma, beta, relu, dy, yh);
    const float invM = 1.f / (float)M;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        o[i] = gamma[c + i] * invstd[c + i] * (dy[i] - sums[c + i] * invM - yh[i] * sums[C + c + i] * invM);
    Vec4<T>::store(gy + r * C + c, o);
}

static inline int64_t rows_per_cta_for(int64_t M)
{
    int64_t ctas = PCB_NUM_SMS * 8;
    int64_t rpc = ceil_div(M, ctas);
    return rpc < 32 ? 32 : rpc;
}

template <typename T>
static int bn_stats_launch(const void *y, int64_t M, int C, float *sums, cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(float) * 2 * (size_t)C, st);
    if (e != cudaSuccess) return (int)e;
    int64_t rpc = rows_per_cta_for(M);
    bn_stats_kernel<T><<<(unsigned)ceil_div(M, rpc), kBnThreads, 0, st>>>((const T *)y, M, C, rpc, sums);
    PCB_RETURN_LAUNCH_STATUS();
}

}  // namespace pcb

using namespace pcb;

#define PCB_BN_CHECK(M, C)                                   \
    PCB_REQUIRE((M) > 0 && (C) > 0, PCB_EINVAL);             \
    PCB_REQUIRE((C) % 4 == 0 && (dtype == 0 || dtype == 1), PCB_ERANGE)

PCB_API int pcb_bn_stats_rows(const void *y, int dtype, int64_t M, int C, float *sums, pcb_stream_t stream)
{
    PCB_REQUIRE(y && sums, PCB_EINVAL);
    PCB_BN_CHECK(M, C);
    return dtype ? bn_stats_launch<__nv_bfloat16>(y, M, C, sums, (cudaStream_t)stream)
                 : bn_stats_launch<float>(y, M, C, sums, (cudaStream_t)stream);
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
    const unsigned total = (unsigned)(Mout * (C / 4));
    unsigned blocks = (unsigned)ceil_div(total, kBnThreads);
    if (dtype)
        bn_apply_kernel<__nv_bfloat16><<<blocks, kBnThreads, 0, st>>>((const __nv_bfloat16 *)y, total, C, make_fastdiv(C / 4),
                                                                     pool_k, mean, invstd, gamma, beta, relu,
                                                                     (__nv_bfloat16 *)out, argmax);
    else
        bn_apply_kernel<float><<<blocks, kBnThreads, 0, st>>>((const float *)y, total, C, make_fastdiv(C / 4), pool_k, mean,
                                                             invstd, gamma, beta, relu, (float *)out, argmax);
    PCB_RETURN_LAUNCH_STATUS();
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
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(float) * 3 * (size_t)C, st);
    if (e != cudaSuccess) return (int)e;
    int64_t rpc = rows_per_cta_for(M);
    unsigned rblocks = (unsigned)ceil_div(M, rpc);
    const unsigned total = (unsigned)(M * (C / 4));
    unsigned ablocks = (unsigned)ceil_div(total, kBnThreads);
    const FastDiv dK = make_fastdiv(pool_k), dC4 = make_fastdiv(C / 4);
    if (dtype) {
        typedef __nv_bfloat16 T;
        bn_bwd_reduce_kernel<T><<<rblocks, kBnThreads, 0, st>>>((const T *)gz, (const T *)y, argmax, M, C, dK, rpc, mean,
                                                               invstd, gamma, beta, relu, sums);
        bn_bwd_apply_kernel<T><<<ablocks, kBnThreads, 0, st>>>((const T *)gz, (const T *)y, argmax, M, total, C, dC4, dK, mean,
                                                              invstd, gamma, beta, relu, sums, (T *)gy);
    } else {
        typedef float T;
        bn_bwd_reduce_kernel<T><<<rblocks, kBnThreads, 0, st>>>((const T *)gz, (const T *)y, argmax, M, C, dK, rpc, mean,
                                                               invstd, gamma, beta, relu, sums);
        bn_bwd_apply_kernel<T><<<ablocks, kBnThreads, 0, st>>>((const T *)gz, (const T *)y, argmax, M, total, C, dC4, dK, mean,
                                                              invstd, gamma, beta, relu, sums, (T *)gy);
    }
    PCB_RETURN_LAUNCH_STATUS();
}
