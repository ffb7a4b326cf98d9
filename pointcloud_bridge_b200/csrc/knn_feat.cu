// Feature-space kNN -- DGCNN.knn for D > 3 (Highway_bridge/models/DGCNN.py:49-70), where the
// reference builds -2*x^T x + |x|^2 + |x|^2^T as a [B,N,N] matrix (1 GiB at B=16, N=4096) and
// runs topk over it.
//
// Exactness contract (pinned in tests/golden): the inner product is one
// sequential fp32 FMA chain over the channel index, row norms follow ATen's 4x8-lane order,
// pd = (|xi|^2 + (-2*dot)) + |xj|^2, neighbours ordered by (pd, index).  That rules out
// tensor cores for the distances: this is FFMA work.
//
// knn_feat_generic_kernel: any D, any N.  A warp serves QB queries at once; for every batch of
// 32 candidates each lane streams its candidate's channels from global memory (coalesced
// along N in the [B,D,N] layout) while the QB query vectors are broadcast from shared memory,
// then the QB distances go through the same threshold/ballot/shuffle-insert selection as the
// coordinate-space kernel.  knn_feat_tiled_kernel (below) is the register-tiled version used
// when the shape allows.
#include "pcb_common.cuh"

namespace pcb {

int knn_xyz_pd(const float *xyz, int B, int N, int k, int cf, int64_t *oi, float *od, cudaStream_t st);

constexpr int kGenWarps = 8;
constexpr int kGenQB = 4;

struct LaneList {                              // top list of up to 32 entries, one per lane
    float d;
    int i;
    __device__ __forceinline__ void init()
    {
        d = __int_as_float(0x7f800000);
        i = 0x7fffffff;
    }
    __device__ __forceinline__ void insert(float v, int vi, int lane)
    {
        const bool gt = (d > v) || (d == v && i > vi);
        float up_d = __shfl_up_sync(PCB_FULL_MASK, d, 1);
        int up_i = __shfl_up_sync(PCB_FULL_MASK, i, 1);
        bool up_gt = __shfl_up_sync(PCB_FULL_MASK, (int)gt, 1) != 0;
        if (lane == 0) up_gt = false;
        if (gt) {
            d = up_gt ? up_d : v;
            i = up_gt ? up_i : vi;
        }
    }
};

__global__ void __launch_bounds__(kGenWarps * 32)
knn_feat_generic_kernel(const float *__restrict__ x, int D, int N, int k, int64_t *__restrict__ out_idx,
                        float *__restrict__ out_dist)
{
    extern __shared__ __align__(16) float smem[];
    float *s_xx = smem;                                    // [N] row norms of the whole cloud
    float *s_q = smem + ((N + 3) & ~3);                    // [warps][D][QB]
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *xb = x + (size_t)b * D * N;
    const int q0 = (blockIdx.x * kGenWarps + warp) * kGenQB;

    for (int i = threadIdx.x; i < N; i += blockDim.x) s_xx[i] = row_sumsq_aten(xb + i, D, N);
    float *myq = s_q + (size_t)warp * D * kGenQB;
    for (int t = lane; t < D * kGenQB; t += 32) {
        int c = t / kGenQB, qq = t % kGenQB;
        int q = q0 + qq;
        myq[t] = q < N ? __ldg(xb + (size_t)c * N + q) : 0.f;
    }
    __syncthreads();
    if (q0 >= N) return;

    LaneList list[kGenQB];
    float thr[kGenQB], qn[kGenQB];
#pragma unroll
    for (int qq = 0; qq < kGenQB; ++qq) {
        list[qq].init();
        thr[qq] = __int_as_float(0x7f800000);
        qn[qq] = (q0 + qq < N) ? s_xx[q0 + qq] : 0.f;
    }
    for (int base = 0; base < N; base += 32) {
        const int j = base + lane;
        const bool valid = j < N;
        const float *col = xb + (valid ? j : 0);
        float acc[kGenQB];
#pragma unroll
        for (int qq = 0; qq < kGenQB; ++qq) acc[qq] = 0.f;
#pragma unroll 4
        for (int c = 0; c < D; ++c) {
            const float v = __ldg(col + (size_t)c * N);
            const float4 qv = *reinterpret_cast<const float4 *>(myq + c * kGenQB);
            acc[0] = __fmaf_rn(qv.x, v, acc[0]);
            acc[1] = __fmaf_rn(qv.y, v, acc[1]);
            acc[2] = __fmaf_rn(qv.z, v, acc[2]);
            acc[3] = __fmaf_rn(qv.w, v, acc[3]);
        }
        const float xj = valid ? s_xx[j] : 0.f;
#pragma unroll
        for (int qq = 0; qq < kGenQB; ++qq) {
            // pairwise_distance = xx + inner + xx^T   (DGCNN.py:65)
            float d = __fadd_rn(__fadd_rn(qn[qq], __fmul_rn(-2.0f, acc[qq])), xj);
            if (!valid) d = __int_as_float(0x7f800000);
            unsigned m = __ballot_sync(PCB_FULL_MASK, d < thr[qq]);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const float v = __shfl_sync(PCB_FULL_MASK, d, src);
                if (v < thr[qq]) {
                    list[qq].insert(v, base + src, lane);
                    thr[qq] = __shfl_sync(PCB_FULL_MASK, list[qq].d, k - 1);
                }
            }
        }
    }
#pragma unroll
    for (int qq = 0; qq < kGenQB; ++qq) {
        const int q = q0 + qq;
        if (q < N && lane < k) {
            out_idx[((size_t)b * N + q) * k + lane] = (int64_t)list[qq].i;
            if (out_dist) out_dist[((size_t)b * N + q) * k + lane] = list[qq].d;
        }
    }
}

}  // namespace pcb

using namespace pcb;

namespace pcb {
int knn_feat_tiled(const float *x, int B, int N, int D, int k, int64_t *oi, float *od, cudaStream_t st);
}

PCB_API int pcb_knn_f32(const float *x, int B, int N, int D, int k, int channels_first, int64_t *out_idx,
                        float *out_dist, pcb_stream_t stream)
{
    PCB_REQUIRE(x && out_idx, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && D > 0 && k > 0, PCB_EINVAL);
    PCB_REQUIRE(k <= 64 && k <= N && D <= 512 && B <= 65535, PCB_ERANGE);
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 3) return knn_xyz_pd(x, B, N, k, channels_first, out_idx, out_dist, st);
    // feature space: channels-first only (what DGCNN passes); k <= 32
    PCB_REQUIRE(channels_first, PCB_ERANGE);
    PCB_REQUIRE(k <= 32, PCB_ERANGE);
    int rc = knn_feat_tiled(x, B, N, D, k, out_idx, out_dist, st);
    if (rc != PCB_ERANGE) return rc;
    size_t smem = ((size_t)((N + 3) & ~3) + (size_t)kGenWarps * D * kGenQB) * sizeof(float);
    PCB_REQUIRE(smem <= 200 * 1024, PCB_ERANGE);
    cudaError_t e = cudaFuncSetAttribute(knn_feat_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)ceil_div(N, kGenWarps * kGenQB), (unsigned)B);
    knn_feat_generic_kernel<<<grid, kGenWarps * 32, smem, st>>>(x, D, N, k, out_idx, out_dist);
    PCB_RETURN_LAUNCH_STATUS();
}
