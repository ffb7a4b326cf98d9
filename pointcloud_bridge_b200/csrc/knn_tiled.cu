// Register-tiled feature-space kNN (DGCNN.knn for D > 3, Highway_bridge/models/DGCNN.py:49-70).
//
// The distance "matrix" is an SGEMM whose accumulation order is pinned (one sequential FMA
// chain over the channel index per output element), so it runs on the FP32 pipe with the
// classic 128x128 CTA tile / 8x8 register tile: each output element accumulates over k in
// ascending order in its own register, which is exactly the oracle's chain.  The [B,D,N]
// channels-first layout DGCNN uses is already K-major for both operands.
//   * the 128 query columns of the CTA (all D channels) stay resident in shared memory;
//   * candidate tiles stream through a 2-stage cp.async pipeline in 16-channel chunks;
//   * the 128x128 distance tile goes to shared memory, never to HBM; each warp then scans its
//     16 rows: threshold test + ballot (4 per row) and, rarely, a shuffle insertion into the
//     row's sorted top-k list (kept in shared memory between tiles);
//   * row norms come from a pre-pass in ATen's summation order (row_sumsq_aten).
// FLOPs: B*N*N*(2D+3); algorithmic bytes: B*(4*D*N + 8*N*k)  (SURVEY.md section 8d).
#include "pcb_common.cuh"

namespace pcb {

int row_sumsq_launch(const float *x, int64_t rows, int C, int64_t row_stride, int elem_stride,
                     int64_t rows_per_batch, int64_t batch_stride, float *out, cudaStream_t st);

constexpr int TM = 128, TN = 128, KC = 16, NT = 256, DPITCH = TN + 4;
constexpr int kMaxTiledD = 128;

__device__ __forceinline__ float finf() { return __int_as_float(0x7f800000); }

__global__ void __launch_bounds__(NT, 1)
knn_feat_tiled_kernel(const float *__restrict__ x, const float *__restrict__ xx, int D, int N, int k,
                      int64_t *__restrict__ out_idx, float *__restrict__ out_dist)
{
    extern __shared__ __align__(16) float smem[];
    float *sA = smem;                                  // [D][TM]   queries, resident
    float *sB = sA + (size_t)D * TM;                   // [2][KC][TN] candidate chunks
    float *sD = sB + 2 * KC * TN;                      // [TM][DPITCH] distance tile
    float *sLd = sD + TM * DPITCH;                     // [TM][32] top-k distances
    int *sLi = reinterpret_cast<int *>(sLd + TM * 32); // [TM][32] top-k indices
    float *sThr = reinterpret_cast<float *>(sLi + TM * 32);   // [TM] current k-th distance
    float *sXi = sThr + TM;                            // [TM] query norms
    float *sXj = sXi + TM;                             // [TN] candidate norms

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    const int b = blockIdx.y, i0 = blockIdx.x * TM;
    const float *xb = x + (size_t)b * D * N;
    const float *xxb = xx + (size_t)b * N;

    const int nchunks = (D + KC - 1) / KC;
    const int ntiles = (N + TN - 1) / TN;
    const int total = ntiles * nchunks;

    auto issue_B = [&](int step) {
        const int tile = step / nchunks, ch = step - tile * nchunks;
        const int j0 = tile * TN, k0 = ch * KC;
        const int kmax = min(KC, D - k0);
        float *dst = sB + (step & 1) * KC * TN;
        for (int t = tid; t < kmax * (TN / 4); t += NT) {
            const int kk = t / (TN / 4), j4 = (t - kk * (TN / 4)) * 4;
            const int gj = j0 + j4;
            const bool ok = gj < N;
            cp_async16(dst + kk * TN + j4, xb + (size_t)(k0 + kk) * N + (ok ? gj : 0), ok ? 16 : 0);
        }
    };

    // resident query tile + first candidate chunk
    for (int t = tid; t < D * (TM / 4); t += NT) {
        const int c = t / (TM / 4), i4 = (t - c * (TM / 4)) * 4;
        const int gi = i0 + i4;
        const bool ok = gi < N;
        cp_async16(sA + c * TM + i4, xb + (size_t)c * N + (ok ? gi : 0), ok ? 16 : 0);
    }
    issue_B(0);
    cp_async_commit();
    for (int t = tid; t < TM * 32; t += NT) {
        sLd[t] = finf();
        sLi[t] = 0x7fffffff;
    }
    if (tid < TM) {
        sThr[tid] = finf();
        sXi[tid] = (i0 + tid < N) ? __ldg(xxb + i0 + tid) : 0.f;
    }

    float acc[8][8];
    for (int step = 0; step < total; ++step) {
        const int tile = step / nchunks, ch = step - tile * nchunks;
        const int j0 = tile * TN, k0 = ch * KC;
        cp_async_wait<0>();
        __syncthreads();                               // chunk `step` landed; everyone left step-1
        if (step + 1 < total) issue_B(step + 1);
        cp_async_commit();
        if (ch == 0) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
            if (tid < TN) sXj[tid] = (j0 + tid < N) ? __ldg(xxb + j0 + tid) : 0.f;
        }
        const float *bufB = sB + (step & 1) * KC * TN;
        const int kmax = min(KC, D - k0);
#pragma unroll 4
        for (int kk = 0; kk < kmax; ++kk) {
            const float *ap = sA + (size_t)(k0 + kk) * TM;
            const float *bp = bufB + kk * TN;
            const float4 a0 = *reinterpret_cast<const float4 *>(ap + ty * 4);
            const float4 a1 = *reinterpret_cast<const float4 *>(ap + 64 + ty * 4);
            const float4 b0 = *reinterpret_cast<const float4 *>(bp + tx * 4);
            const float4 b1 = *reinterpret_cast<const float4 *>(bp + 64 + tx * 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = __fmaf_rn(a[r], bb[c], acc[r][c]);
        }
        if (ch != nchunks - 1) continue;

        // ---- tile finished: distances -> shared memory ----
        if (nchunks == 1) __syncthreads();             // sXj was written in this very step
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int row = (r < 4) ? (ty * 4 + r) : (64 + ty * 4 + r - 4);
            const float xi = sXi[row];
            float dv[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int col = (c < 4) ? (tx * 4 + c) : (64 + tx * 4 + c - 4);
                // pairwise_distance = xx + inner + xx^T, inner = -2 * dot   (DGCNN.py:63-65)
                float d = __fadd_rn(__fadd_rn(xi, __fmul_rn(-2.0f, acc[r][c])), sXj[col]);
                dv[c] = (j0 + col < N) ? d : finf();
            }
            *reinterpret_cast<float4 *>(sD + row * DPITCH + tx * 4) = make_float4(dv[0], dv[1], dv[2], dv[3]);
            *reinterpret_cast<float4 *>(sD + row * DPITCH + 64 + tx * 4) = make_float4(dv[4], dv[5], dv[6], dv[7]);
        }
        __syncthreads();

        // ---- selection: warp w owns rows [w*16, w*16+16) ----
        for (int rr = 0; rr < TM / 8; ++rr) {
            const int row = warp * (TM / 8) + rr;
            if (i0 + row >= N) break;
            float thr = sThr[row];
            float v[TN / 32];
            unsigned m[TN / 32], any = 0u;
#pragma unroll
            for (int g = 0; g < TN / 32; ++g) {
                v[g] = sD[row * DPITCH + g * 32 + lane];
                m[g] = __ballot_sync(PCB_FULL_MASK, v[g] < thr);
                any |= m[g];
            }
            if (!any) continue;
            float ld = sLd[row * 32 + lane];
            int li = sLi[row * 32 + lane];
#pragma unroll
            for (int g = 0; g < TN / 32; ++g) {
                unsigned mm = m[g];
                while (mm) {
                    const int src = __ffs(mm) - 1;
                    mm &= mm - 1;
                    const float val = __shfl_sync(PCB_FULL_MASK, v[g], src);
                    if (val < thr) {
                        const int vi = j0 + g * 32 + src;
                        const bool gt = (ld > val) || (ld == val && li > vi);
                        const float up_d = __shfl_up_sync(PCB_FULL_MASK, ld, 1);
                        const int up_i = __shfl_up_sync(PCB_FULL_MASK, li, 1);
                        bool up_gt = __shfl_up_sync(PCB_FULL_MASK, (int)gt, 1) != 0;
                        if (lane == 0) up_gt = false;
                        if (gt) {
                            ld = up_gt ? up_d : val;
                            li = up_gt ? up_i : vi;
                        }
                        thr = __shfl_sync(PCB_FULL_MASK, ld, k - 1);
                    }
                }
            }
            sLd[row * 32 + lane] = ld;
            sLi[row * 32 + lane] = li;
            if (lane == 0) sThr[row] = thr;
        }
        // the next step's leading __syncthreads orders this selection before sD / sXj are reused
    }

    __syncthreads();
    for (int rr = 0; rr < TM / 8; ++rr) {
        const int row = warp * (TM / 8) + rr;
        const int q = i0 + row;
        if (q >= N) break;
        if (lane < k) {
            out_idx[((size_t)b * N + q) * k + lane] = (int64_t)sLi[row * 32 + lane];
            if (out_dist) out_dist[((size_t)b * N + q) * k + lane] = sLd[row * 32 + lane];
        }
    }
}

// Returns PCB_ERANGE when the shape is outside the tiled kernel's envelope (caller falls back).
int knn_feat_tiled(const float *x, int B, int N, int D, int k, int64_t *oi, float *od, cudaStream_t st)
{
    static int disabled = -1;
    if (disabled < 0) disabled = getenv("PCB_KNN_GENERIC") ? 1 : 0;
    if (disabled) return PCB_ERANGE;
    if (D > kMaxTiledD || k > 32 || (N % 4) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0) return PCB_ERANGE;
    size_t smem = ((size_t)D * TM + 2 * KC * TN + TM * DPITCH + TM * 32 * 2 + TM + TM + TN) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(knn_feat_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    float *xx = nullptr;
    e = cudaMallocAsync(reinterpret_cast<void **>(&xx), (size_t)B * N * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    // row norms of x[b, :, i]: element stride N, consecutive rows 1 apart
    int rc = row_sumsq_launch(x, (int64_t)B * N, D, 1, N, N, (int64_t)D * N, xx, st);
    if (rc == 0) {
        dim3 grid((unsigned)ceil_div(N, TM), (unsigned)B);
        knn_feat_tiled_kernel<<<grid, NT, smem, st>>>(x, xx, D, N, k, oi, od);
        rc = (int)cudaGetLastError();
    }
    cudaError_t e2 = cudaFreeAsync(xx, st);
    return rc != 0 ? rc : (int)e2;
}

}  // namespace pcb
