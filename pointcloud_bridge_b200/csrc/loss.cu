// Mean negative log-likelihood of the segmentation head on point-major logits rows, forward and backward in
// one pass each -- the `F.log_softmax(...)` + `F.nll_loss(...)` pair of
// Partsize-identical/models/pointnet2_sem_seg.py:46-47, 56 (and pointnet2_sem_seg_msg.py) as the training step
// uses it (no class weights, mean over all points).  logits [M, pitch] fp32 or bf16 (the classifier GEMM's padded,
// bias-free rows, `classes` <= pitch real columns; the conv bias [classes] is added here), labels [M] int64.
//   forward : partial[b] = sum over the rows of CTA b of (logsumexp(x) - x[label]);  loss = sum(partial) / M
//   backward: dx[r, c] = (softmax(x[r])[c] - [c == label]) * gscale / M, pad columns 0
// One thread per row (<= 32 classes live in registers); the partial sums are reduced in a fixed order by the
// caller, so the loss is deterministic.  Replaces 14 ATen launches (pad, slice copy, softmax, gather, a
// single-block mean over 65 536 rows, their backward twins) of the round-1 step.
#include <cuda_bf16.h>

#include "pcb_common.cuh"

namespace pcb {

constexpr int kNllThreads = 256;
constexpr int kNllMaxClasses = 32;

__device__ __forceinline__ float ld_logit(const float *p) { return __ldg(p); }
__device__ __forceinline__ float ld_logit(const __nv_bfloat16 *p) { return __bfloat162float(*p); }
__device__ __forceinline__ void st_logit(float *p, float v) { *p = v; }
__device__ __forceinline__ void st_logit(__nv_bfloat16 *p, float v) { *p = __float2bfloat16(v); }

template <typename T>
__global__ void __launch_bounds__(kNllThreads)
nll_rows_fwd_kernel(const T *__restrict__ x, const float *__restrict__ bias, const int64_t *__restrict__ labels, int64_t M,
                    int classes, int pitch, float *__restrict__ partial)
{
    pdl_wait();
    pdl_trigger();
    __shared__ float s_red[kNllThreads / 32];
    float acc = 0.f;
    for (int64_t r = (int64_t)blockIdx.x * kNllThreads + threadIdx.x; r < M; r += (int64_t)gridDim.x * kNllThreads) {
        const T *row = x + r * pitch;
        float v[kNllMaxClasses], mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < kNllMaxClasses; ++c)
            if (c < classes) {
                v[c] = ld_logit(row + c) + (bias ? __ldg(bias + c) : 0.f);
                mx = fmaxf(mx, v[c]);
            }
        float se = 0.f;
#pragma unroll
        for (int c = 0; c < kNllMaxClasses; ++c)
            if (c < classes) se += expf(v[c] - mx);
        long long l = labels[r];
        l = l < 0 ? 0 : (l >= classes ? classes - 1 : l);
        float xl = 0.f;
#pragma unroll
        for (int c = 0; c < kNllMaxClasses; ++c)
            if (c == (int)l) xl = v[c];
        acc += (mx + logf(se)) - xl;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(PCB_FULL_MASK, acc, off);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < kNllThreads / 32; ++w) s += s_red[w];
        partial[blockIdx.x] = s;
    }
}

template <typename T>
__global__ void __launch_bounds__(kNllThreads)
nll_rows_bwd_kernel(const T *__restrict__ x, const float *__restrict__ bias, const int64_t *__restrict__ labels, int64_t M,
                    int classes, int pitch, const float *__restrict__ gscale, T *__restrict__ dx, float *__restrict__ gbias)
{
    pdl_wait();
    pdl_trigger();
    __shared__ float s_gb[kNllThreads / 32][kNllMaxClasses];
    float gb[kNllMaxClasses];
#pragma unroll
    for (int c = 0; c < kNllMaxClasses; ++c) gb[c] = 0.f;
    const float g = gscale[0] / (float)M;
    for (int64_t r = (int64_t)blockIdx.x * kNllThreads + threadIdx.x; r < M; r += (int64_t)gridDim.x * kNllThreads) {
        const T *row = x + r * pitch;
        float v[kNllMaxClasses], mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < kNllMaxClasses; ++c)
            if (c < classes) {
                v[c] = ld_logit(row + c) + (bias ? __ldg(bias + c) : 0.f);
                mx = fmaxf(mx, v[c]);
            }
        float se = 0.f;
#pragma unroll
        for (int c = 0; c < kNllMaxClasses; ++c)
            if (c < classes) {
                v[c] = expf(v[c] - mx);
                se += v[c];
            }
        long long l = labels[r];
        l = l < 0 ? 0 : (l >= classes ? classes - 1 : l);
        const float inv = 1.f / se;
        T *o = dx + r * pitch;
#pragma unroll
        for (int c = 0; c < kNllMaxClasses; ++c)
            if (c < classes) {
                const float d = (v[c] * inv - (c == (int)l ? 1.f : 0.f)) * g;
                st_logit(o + c, d);
                gb[c] += d;
            }
        for (int c = classes; c < pitch; ++c) st_logit(o + c, 0.f);
    }
    if (gbias) {                                           // gradient of the classifier bias: column sums of dx
#pragma unroll
        for (int c = 0; c < kNllMaxClasses; ++c)
            if (c < classes) {
                float t = gb[c];
#pragma unroll
                for (int off = 16; off; off >>= 1) t += __shfl_xor_sync(PCB_FULL_MASK, t, off);
                if ((threadIdx.x & 31) == 0) s_gb[threadIdx.x >> 5][c] = t;
            }
        __syncthreads();
        if (threadIdx.x < classes) {
            float t = 0.f;
            for (int w = 0; w < kNllThreads / 32; ++w) t += s_gb[w][threadIdx.x];
            atomicAdd(gbias + threadIdx.x, t);
        }
    }
}

}  // namespace pcb

using namespace pcb;

PCB_API int pcb_nll_rows_blocks(int64_t M)
{
    int64_t b = ceil_div(M, kNllThreads);
    return (int)(b < 1 ? 1 : (b > PCB_NUM_SMS * 4 ? PCB_NUM_SMS * 4 : b));
}

PCB_API int pcb_nll_rows_fwd(const void *logits, int dtype, const float *bias, const int64_t *labels, int64_t M, int classes,
                             int pitch, float *partial, pcb_stream_t stream)
{
    PCB_REQUIRE(logits && labels && partial, PCB_EINVAL);
    PCB_REQUIRE(M > 0 && classes > 0 && classes <= kNllMaxClasses && pitch >= classes && (dtype == 0 || dtype == 1), PCB_ERANGE);
    const int blocks = pcb_nll_rows_blocks(M);
    if (dtype)
        launch_pdl(nll_rows_fwd_kernel<__nv_bfloat16>, dim3(blocks), dim3(kNllThreads), 0, (cudaStream_t)stream, (const __nv_bfloat16 *)logits, bias, labels, M,
                                                                            classes, pitch, partial);
    else
        launch_pdl(nll_rows_fwd_kernel<float>, dim3(blocks), dim3(kNllThreads), 0, (cudaStream_t)stream, (const float *)logits, bias, labels, M, classes,
                                                                            pitch, partial);
    PCB_RETURN_LAUNCH_STATUS();
}

PCB_API int pcb_nll_rows_bwd(const void *logits, int dtype, const float *bias, const int64_t *labels, int64_t M, int classes,
                             int pitch, const float *grad_loss, void *grad_logits, float *grad_bias, pcb_stream_t stream)
{
    PCB_REQUIRE(logits && labels && grad_loss && grad_logits, PCB_EINVAL);
    PCB_REQUIRE(M > 0 && classes > 0 && classes <= kNllMaxClasses && pitch >= classes && (dtype == 0 || dtype == 1), PCB_ERANGE);
    const int blocks = pcb_nll_rows_blocks(M);
    if (dtype)
        launch_pdl(nll_rows_bwd_kernel<__nv_bfloat16>, dim3(blocks), dim3(kNllThreads), 0, (cudaStream_t)stream, (const __nv_bfloat16 *)logits, bias, labels, M,
                                                                            classes, pitch, grad_loss,
                                                                            (__nv_bfloat16 *)grad_logits, grad_bias);
    else
        launch_pdl(nll_rows_bwd_kernel<float>, dim3(blocks), dim3(kNllThreads), 0, (cudaStream_t)stream, (const float *)logits, bias, labels, M, classes,
                                                                            pitch, grad_loss, (float *)grad_logits, grad_bias);
    PCB_RETURN_LAUNCH_STATUS();
}
