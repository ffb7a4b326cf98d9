// Register-tiled feature-space kNN (DGCNN.knn for D > 3, Highway_bridge/models/DGCNN.py:49-70).
//
// The distance "matrix" is an SGEMM whose accumulation order is pinned (one sequential FMA
// chain over the channel index per output element), so it runs on the FP32 pipe with a
// 128(query) x 64(candidate) CTA tile and an 8x8 register tile per thread: every output element
// accumulates over k in ascending order in its own register -- exactly the oracle's chain.  The
// [B,D,N] channels-first layout DGCNN uses is already K-major for both operands.
//   * 128 threads per CTA, two CTAs per SM: while one CTA selects, the other multiplies;
//   * the 128 query columns of the CTA (all D channels) stay resident in shared memory;
//   * candidate tiles stream through a 2-stage cp.async pipeline in 16-channel chunks;
//   * the 128x64 distance tile goes to shared memory (transposed: [candidate][query]), never to
//     HBM; then thread t owns query row t: it scans the 64 distances of its row against its
//     running k-th distance into a 64-bit mask and drains the mask into its sorted top-k list,
//     which lives in REGISTERS (the 64 accumulators are dead during selection); an insertion is
//     a branch-free compare-exchange pass over the K entries.  All 128 threads select
//     concurrently: no shuffles, no per-row warp loop, no shared-memory shifting (round-1
//     profiles: the warp-serial loop took 70 % of the kernel, the smem-shifting variant 35 %);
//   * row norms come from a pre-pass in ATen's summation order (row_sumsq_aten).
// FLOPs: B*N*N*(2D+3); algorithmic bytes: B*(4*D*N + 8*N*k)  (SURVEY.md section 8d).
#include "pcb_common.cuh"

namespace pcb {

int row_sumsq_launch(const float *x, int64_t rows, int C, int64_t row_stride, int elem_stride,
                     int64_t rows_per_batch, int64_t batch_stride, float *out, cudaStream_t st);

constexpr int TM = 128, TN = 64, KC = 16, NT = 128, TMP = TM + 4;
constexpr int kMaxTiledD = 128;

__device__ __forceinline__ float finf() { return __int_as_float(0x7f800000); }

template <int K>
__global__ void __launch_bounds__(NT, 2)
knn_feat_tiled_kernel(const float *__restrict__ x, const float *__restrict__ xx, int D, int N, int k,
                      int64_t *__restrict__ out_idx, float *__restrict__ out_dist)
{
    extern __shared__ __align__(16) float smem[];
    float *sA = smem;                                  // [D][TM]    queries, resident
    float *sB = sA + (size_t)D * TM;                   // [2][KC][TN] candidate chunks
    float *sD = sB + 2 * KC * TN;                      // [TN][TMP]  distance tile, candidate-major
    float *sXi = sD + TN * TMP;                        // [TM] query norms
    float *sXj = sXi + TM;                             // [TN] candidate norms

    const int tid = threadIdx.x;
    const int ty = tid >> 3, tx = tid & 7;             // 16 x 8 threads, 8 x 8 outputs each
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
    sXi[tid] = (i0 + tid < N) ? __ldg(xxb + i0 + tid) : 0.f;

    // sorted top list of query row `tid` (ascending by (distance, index)); entries >= k are slack
    float ld[K];
    int li[K];
#pragma unroll
    for (int p = 0; p < K; ++p) {
        ld[p] = finf();
        li[p] = 0x7fffffff;
    }
    float thr = finf();                                // K-th smallest distance so far (K >= k)
    const bool row_valid = (i0 + tid) < N;
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
            const float4 b1 = *reinterpret_cast<const float4 *>(bp + 32 + tx * 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = __fmaf_rn(a[r], bb[c], acc[r][c]);
        }
        if (ch != nchunks - 1) continue;

        // ---- tile finished: distances -> shared memory, transposed ----
        if (nchunks == 1) __syncthreads();             // sXj was written in this very step
        {
            float xi[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) xi[r] = sXi[(r < 4) ? (ty * 4 + r) : (64 + ty * 4 + r - 4)];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int col = (c < 4) ? (tx * 4 + c) : (32 + tx * 4 + c - 4);
                const float xj = sXj[col];
                const bool ok = (j0 + col) < N;
                float dv[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    // pairwise_distance = xx + inner + xx^T, inner = -2 * dot   (DGCNN.py:63-65)
                    const float d = __fadd_rn(__fadd_rn(xi[r], __fmul_rn(-2.0f, acc[r][c])), xj);
                    dv[r] = ok ? d : finf();
                }
                *reinterpret_cast<float4 *>(sD + col * TMP + ty * 4) = make_float4(dv[0], dv[1], dv[2], dv[3]);
                *reinterpret_cast<float4 *>(sD + col * TMP + 64 + ty * 4) = make_float4(dv[4], dv[5], dv[6], dv[7]);
            }
        }
        __syncthreads();

        // ---- selection: thread `tid` owns query row `tid` ----
        if (row_valid) {
            unsigned long long mask = 0ull;
#pragma unroll 16
            for (int c = 0; c < TN; ++c)
                mask |= (unsigned long long)(sD[c * TMP + tid] < thr) << c;
            while (mask) {
                const int c = __ffsll((long long)mask) - 1;       // ascending candidate index
                mask &= mask - 1;
                const float v = sD[c * TMP + tid];
                if (v < thr) {                                     // thr may have dropped meanwhile
                    float cv = v;
                    int ci = j0 + c;
                    bool placed = false;
#pragma unroll
                    for (int p = 0; p < K; ++p) {                  // compare-exchange down the list
                        // new element: strict < (equal distance -> the earlier index stays first);
                        // once it is placed we carry old entries, which simply shift down one slot
                        const bool sw = placed || (cv < ld[p]);
                        placed = sw;
                        const float nd = sw ? cv : ld[p], nc = sw ? ld[p] : cv;
                        const int ni = sw ? ci : li[p], nci = sw ? li[p] : ci;
                        ld[p] = nd;
                        li[p] = ni;
                        cv = nc;
                        ci = nci;
                    }
                    thr = ld[K - 1];                               // top-K list (K >= k): first k are exact
                }
            }
        }
        // the next step's leading __syncthreads orders this selection before sD / sXj are reused
    }

    if (row_valid) {
        int64_t *oi = out_idx + ((size_t)b * N + i0 + tid) * k;
        float *od = out_dist ? out_dist + ((size_t)b * N + i0 + tid) * k : nullptr;
#pragma unroll
        for (int p = 0; p < K; ++p) {
            if (p < k) {
                oi[p] = (int64_t)li[p];
                if (od) od[p] = ld[p];
            }
        }
    }
}

template <int K>
static int launch_tiled(const float *x, const float *xx, int B, int N, int D, int k, int64_t *oi, float *od,
                        cudaStream_t st)
{
    size_t smem = ((size_t)D * TM + 2 * KC * TN + TN * TMP + TM + TN) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(knn_feat_tiled_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)ceil_div(N, TM), (unsigned)B);
    knn_feat_tiled_kernel<K><<<grid, NT, smem, st>>>(x, xx, D, N, k, oi, od);
    return (int)cudaGetLastError();
}

// Returns PCB_ERANGE when the shape is outside the tiled kernel's envelope (caller falls back).
int knn_feat_tiled(const float *x, int B, int N, int D, int k, int64_t *oi, float *od, cudaStream_t st)
{
    static int disabled = -1;
    if (disabled < 0) disabled = getenv("PCB_KNN_GENERIC") ? 1 : 0;
    if (disabled) return PCB_ERANGE;
    if (D > kMaxTiledD || k > 32 || (N % 4) != 0 || (reinterpret_cast<uintptr_t>(x) & 15) != 0) return PCB_ERANGE;
    // scratch for the row norms from the stream-ordered pool; keep freed blocks cached in the pool
    // (the default release threshold of 0 hands them back to the driver at every sync)
    static bool pool_set = false;
    if (!pool_set) {
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        pool_set = true;
    }
    float *xx = nullptr;
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&xx), (size_t)B * N * sizeof(float), st);
    if (e != cudaSuccess) return (int)e;
    // row norms of x[b, :, i]: element stride N, consecutive rows 1 apart
    int rc = row_sumsq_launch(x, (int64_t)B * N, D, 1, N, N, (int64_t)D * N, xx, st);
    if (rc == 0) {
        if (k <= 8) rc = launch_tiled<8>(x, xx, B, N, D, k, oi, od, st);
        else if (k <= 16) rc = launch_tiled<16>(x, xx, B, N, D, k, oi, od, st);
        else if (k <= 20) rc = launch_tiled<20>(x, xx, B, N, D, k, oi, od, st);
        else if (k <= 24) rc = launch_tiled<24>(x, xx, B, N, D, k, oi, od, st);
        else rc = launch_tiled<32>(x, xx, B, N, D, k, oi, od, st);
    }
    cudaError_t e2 = cudaFreeAsync(xx, st);
    return rc != 0 ? rc : (int)e2;
}

}  // namespace pcb
