// Optimizer step of the data-parallel training runner (the Adam update of
// Highway_bridge/train_MulSca_BriStruNet_CB.py:158-190 / Partsize-identical/train.py, torch.optim.Adam with
// L2 weight decay) over ONE flat fp32 parameter buffer, fused with the refresh of the bf16 weight shadows
// that the next step's shared-MLP GEMMs read: p, g, m, v are streamed once, the updated parameter is
// written back in fp32 and -- where `shadow_index[i] >= 0` -- also as bf16 into its (zero-padded) GEMM
// operand.  Same arithmetic as torch's fused Adam kernel (lerp form of the first moment, bias
// corrections from pow(beta, step)); `step` and `lr` live in device memory so that a captured CUDA graph
// sees their updates.  HBM-bound: 32 bytes per parameter (+2 for a shadowed one).
#include <cuda_bf16.h>

#include "pcb_common.cuh"

namespace pcb {

__global__ void __launch_bounds__(256)
adam_flat_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                 int64_t n, const float *__restrict__ lr_ptr, float beta1, float beta2, float eps, float weight_decay,
                 const int64_t *__restrict__ step_ptr, const int *__restrict__ shadow_index,
                 const int *__restrict__ shadow_index_t, __nv_bfloat16 *__restrict__ shadow,
                 const unsigned char *__restrict__ skip)
{
    pdl_wait();
    pdl_trigger();
    const float step = (float)step_ptr[0];
    const float lr = lr_ptr[0];
    const float bc1 = 1.f - powf(beta1, step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, step));
    const float step_size = lr / bc1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        // parameters that never receive a gradient (torch.optim.Adam skips grad = None: no weight decay, no state)
        if (skip && skip[i]) continue;
        float pi = p[i];
        float gi = g[i];
        if (weight_decay != 0.f) gi += pi * weight_decay;
        float mi = m[i], vi = v[i];
        mi = mi + (1.f - beta1) * (gi - mi);                        // lerp(m, g, 1 - beta1)
        vi = beta2 * vi + (1.f - beta2) * gi * gi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        pi -= step_size * mi / denom;
        p[i] = pi, m[i] = mi, v[i] = vi;
        if (shadow_index) {
            const int s = shadow_index[i];
            if (s >= 0) {
                const __nv_bfloat16 b = __float2bfloat16(pi);
                shadow[s] = b;
                if (shadow_index_t) shadow[shadow_index_t[i]] = b;            // transposed copy (dgrad operand)
            }
        }
    }
}

}  // namespace pcb

PCB_API int pcb_adam_flat_f32(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n,
                              const float *lr, float beta1, float beta2, float eps, float weight_decay,
                              const int64_t *step, const int *shadow_index, const int *shadow_index_t, void *shadow_bf16,
                              const unsigned char *skip, pcb_stream_t stream)
{
    using namespace pcb;
    PCB_REQUIRE(param && grad && exp_avg && exp_avg_sq && lr && step, PCB_EINVAL);
    PCB_REQUIRE(n > 0 && (!shadow_index || shadow_bf16), PCB_EINVAL);
    int64_t blocks = ceil_div(n, 256);
    const int64_t cap = (int64_t)PCB_NUM_SMS * 16;
    if (blocks > cap) blocks = cap;
    launch_pdl(adam_flat_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2,
                                                                        eps, weight_decay, step, shadow_index,
                                                                        shadow_index_t, (__nv_bfloat16 *)shadow_bf16, skip);
    PCB_RETURN_LAUNCH_STATUS();
}
