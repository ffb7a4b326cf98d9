// Shared-MLP contractions of the TRAINING path on the 5th-generation tensor cores:
//   C[M, N] = A[M, K] . B[N, K]^T      bf16 operands, fp32 accumulation in TMEM, bf16 result
// with M = rows of points (10^3 .. 5*10^5), K and N = channel counts of the layer (8 .. 1536).  It replaces the
// library GEMMs behind the 1x1 Conv2d / Conv1d layers of
//   Partsize-identical/models/pointnet_util.py:213-217 (set abstraction), :275-277 (multi-scale grouping),
//   :343-345 (feature propagation), pointnet2_sem_seg.py:43-46 (head), Highway_bridge/models/pointnet2_utils.py:150-154,
//   353-356 and Highway_bridge/models/DGCNN.py:134-148 (EdgeConv)
// in the forward pass (A = activations, B = weight [out, in]) and in the data-gradient pass (A = gradient rows,
// B = weight transposed [in, out]).  Two epilogues fold the BatchNorm passes that used to follow into the GEMM:
//   EPI_STATS  forward: per-channel sums of (y - s) and (y - s)^2 of the bf16 result, folded to mean / invstd / running
//              statistics by the last CTA (training-mode BatchNorm statistics, torch.nn.functional.batch_norm);
//   EPI_BNBWD  backward: the accumulator is d loss / d z of the PREVIOUS layer's BN+ReLU output; the epilogue reads that
//              layer's pre-activation tile y, applies the ReLU mask and emits dy together with the per-channel sums of
//              dy and dy * yhat that the BatchNorm backward needs (yhat = (y - mean) * invstd).
//
// Persistent, warp-specialised CTAs of 288 threads:
//   warps 0-3  epilogue   tcgen05.ld (warp w owns TMEM lanes 32w..32w+31 = rows of the tile) -> bf16 -> shared-memory
//                         stage -> coalesced 16-byte global stores; column sums by a 16-value warp butterfly
//   warp  4    MMA        one elected thread issues tcgen05.mma (M = 128, N = BN <= 256, K = 16 per instruction) and
//                         tcgen05.commit; owns the TMEM allocation (two accumulator buffers: the epilogue of tile i
//                         overlaps the MMAs of tile i + 1)
//   warps 5-8  producers  cp.async (16 B) global -> shared memory straight into the canonical K-major no-swizzle UMMA
//                         layout, zero fill for the K / M / N tails, D stages in flight per thread; completion:
//                         cp.async.wait_group -> fence.proxy.async -> mbarrier arrive
// Pipelines: full/empty mbarriers per ring stage (producers <-> MMA), full/empty per accumulator buffer (MMA <->
// epilogue).  The operands are small-K / small-N matrices streamed once: the kernel is HBM-bound (algorithmic bytes
// 2 * M * (K + N) [+ 2 * M * N for the y tile of EPI_BNBWD]); the tensor pipe idles most of the time by construction.
#include <cuda_bf16.h>

#include "pcb_common.cuh"
#include "umma.cuh"

namespace pcb {

constexpr int kGemmThreads = 288;
constexpr int kGemmMaxStages = 8;
constexpr int kGemmLboPad = 16;            // bytes added to the K-chunk plane stride: spreads the 16-byte units of one
                                           // row (consecutive K chunks) over different banks for the cp.async stores
enum { EPI_STORE = 0, EPI_STATS = 1, EPI_BNBWD = 2 };

struct GemmParams {
    const __nv_bfloat16 *A;      // [M, lda]
    const __nv_bfloat16 *B;      // [Nb, ldb]: row n = output column n
    __nv_bfloat16 *C;            // [M, ldc]
    int64_t lda, ldb, ldc, M;
    int N;                       // output columns written (multiple of 8)
    int Nb;                      // rows of B that exist (others are zero)
    int K;                       // contraction length (multiple of 8)
    int BN, BK, stages, mtiles, ntiles;
    int Cv;                      // real channels among the N columns (statistics epilogues)
    float *parts;                // [ntiles][gridDim.x][2][BN] partial column sums
    unsigned *tickets;           // [ntiles], zero on entry, zero on exit
    // EPI_STATS
    const float *bias;           // conv bias [Cv] or nullptr (enters the running mean only)
    float eps, momentum;
    float *running_mean, *running_var, *mean, *invstd;
    // EPI_BNBWD
    const __nv_bfloat16 *Y;      // [M, ldy] pre-activation of the layer whose output gradient this GEMM produces
    int64_t ldy;
    const float *bn_mean, *bn_invstd, *gamma, *beta;   // [Cv]
    int relu;
    float *sums;                 // [3][N]: sum dy, sum dy*yhat, 0 (gradient of the folded conv bias)
};

// sum over the 32 lanes of each of 16 per-lane values in 16 shuffles; lane L receives the total of v[L >> 1]
__device__ __forceinline__ float warp_colsum16(const float v[16], int lane)
{
    float a[8], b[4], c[2], d;
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float keep = h4 ? v[i + 8] : v[i], send = h4 ? v[i] : v[i + 8];
        a[i] = keep + __shfl_xor_sync(PCB_FULL_MASK, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = h3 ? a[i + 4] : a[i], send = h3 ? a[i] : a[i + 4];
        b[i] = keep + __shfl_xor_sync(PCB_FULL_MASK, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float keep = h2 ? b[i + 2] : b[i], send = h2 ? b[i] : b[i + 2];
        c[i] = keep + __shfl_xor_sync(PCB_FULL_MASK, send, 4);
    }
    {
        const float keep = h1 ? c[1] : c[0], send = h1 ? c[0] : c[1];
        d = keep + __shfl_xor_sync(PCB_FULL_MASK, send, 2);
    }
    d += __shfl_xor_sync(PCB_FULL_MASK, d, 1);
    return d;
}

__device__ __forceinline__ void bar_sync_named(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void unpack8(const uint4 &t, float v[8])
{
    const unsigned w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

struct GemmSmem {
    int lboA, lboB, a_bytes, stage_bytes, pitchC;
    int off_stage, off_const, off_acc, total;
};

__host__ __device__ __forceinline__ GemmSmem gemm_smem_layout(int BN, int BK, int stages, int epi)
{
    GemmSmem s;
    s.lboA = 128 * 16 + kGemmLboPad;
    s.lboB = BN * 16 + kGemmLboPad;
    const int nch = BK / 8;
    s.a_bytes = nch * s.lboA;
    s.stage_bytes = (nch * (s.lboA + s.lboB) + 127) & ~127;
    s.pitchC = (BN + 8) * 2;
    s.off_stage = stages * s.stage_bytes;
    s.off_const = s.off_stage + ((128 * s.pitchC + 127) & ~127);
    const int nconst = epi == EPI_BNBWD ? 4 : (epi == EPI_STATS ? 1 : 0);
    s.off_acc = s.off_const + nconst * BN * 4;
    s.total = s.off_acc + (epi == EPI_STORE ? 0 : 4 * 2 * BN * 4);
    return s;
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_rows_kernel(const GemmParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_full[kGemmMaxStages], s_empty[kGemmMaxStages], s_accfull[2], s_accempty[2];
    __shared__ uint32_t s_tmem;
    __shared__ int s_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int BN = p.BN, BK = p.BK, S = p.stages;
    const GemmSmem L = gemm_smem_layout(BN, BK, S, EPI);
    const int nt = blockIdx.y, n0 = nt * BN;
    const int nslabs = (p.K + BK - 1) / BK;
    unsigned char *stageC = smem + L.off_stage;
    float *s_const = reinterpret_cast<float *>(smem + L.off_const);
    float *s_acc = reinterpret_cast<float *>(smem + L.off_acc);          // [4 warps][2][BN]

    int tmem_cols = 32;
    while (tmem_cols < 2 * BN) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&s_full[s], 128);
            mbar_init(&s_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&s_accfull[b], 1);
            mbar_init(&s_accempty[b], 128);
        }
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc(smem_u32(&s_tmem), (uint32_t)tmem_cols);
    if (EPI != EPI_STORE && warp < 4) {
        for (int c = tid; c < BN; c += 128) {
            const int gc = n0 + c;
            const bool real = gc < p.Cv;
            if (EPI == EPI_STATS) {
                // common shift of the column sums: last step's running mean of the bias-free output (any value is
                // correct; one near the batch mean avoids cancellation in sum (y-s)^2 - (sum (y-s))^2 / M).  The
                // running mean is rewritten by the last CTA only after every CTA has taken its ticket.
                s_const[c] = (real && p.running_mean) ? p.running_mean[gc] - (p.bias ? p.bias[gc] : 0.f) : 0.f;
            } else {
                const float m = real ? p.bn_mean[gc] : 0.f, is = real ? p.bn_invstd[gc] : 0.f;
                const float sc = real ? is * p.gamma[gc] : 0.f;              // same expressions as bn_rows.cu
                s_const[c] = -m * is;                                        // nm
                s_const[BN + c] = is;
                s_const[2 * BN + c] = sc;
                s_const[3 * BN + c] = real ? p.beta[gc] - m * sc : 0.f;
            }
        }
        for (int i = tid; i < 4 * 2 * BN; i += 128) s_acc[i] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp >= 5) {
        // =========================== producers ===========================
        const int pt = tid - 160;
        const int D = S - 1 < 3 ? S - 1 : 3;                                 // slabs in flight per thread
        int it = 0;
        for (int mt = blockIdx.x; mt < p.mtiles; mt += gridDim.x) {
            const int64_t m0 = (int64_t)mt * 128;
            for (int ks = 0; ks < nslabs; ++ks, ++it) {
                const int s = it % S;
                mbar_wait(&s_empty[s], (((uint32_t)(it / S)) & 1u) ^ 1u);
                const int k0 = ks * BK;
                const int kw = p.K - k0 < BK ? p.K - k0 : BK;                // real columns of this slab (multiple of 8)
                const int nch = ((kw + 15) >> 4) << 1;                       // 16-byte K chunks incl. zero fill to K % 16 == 0
                const int vch = kw >> 3;
                unsigned char *sa = smem + (size_t)s * L.stage_bytes, *sb = sa + L.a_bytes;
                for (int idx = pt; idx < 128 * nch; idx += 128) {
                    const int row = nch == 8 ? idx >> 3 : idx / nch;
                    const int kc = idx - row * nch;
                    const int64_t gr = m0 + row;
                    const bool ok = gr < p.M && kc < vch;
                    const __nv_bfloat16 *src = ok ? p.A + gr * p.lda + k0 + kc * 8 : p.A;
                    cp_async16(sa + kc * L.lboA + row * 16, src, ok ? 16 : 0);
                }
                for (int idx = pt; idx < BN * nch; idx += 128) {
                    const int row = nch == 8 ? idx >> 3 : idx / nch;
                    const int kc = idx - row * nch;
                    const int gn = n0 + row;
                    const bool ok = gn < p.Nb && kc < vch;
                    const __nv_bfloat16 *src = ok ? p.B + (int64_t)gn * p.ldb + k0 + kc * 8 : p.B;
                    cp_async16(sb + kc * L.lboB + row * 16, src, ok ? 16 : 0);
                }
                cp_async_commit();
                if (it >= D) {
                    if (D == 3) cp_async_wait<3>();
                    else if (D == 2) cp_async_wait<2>();
                    else cp_async_wait<1>();
                    proxy_fence();
                    mbar_arrive(&s_full[(it - D) % S]);
                }
            }
        }
        cp_async_wait<0>();
        proxy_fence();
        for (int j = it - D < 0 ? 0 : it - D; j < it; ++j) mbar_arrive(&s_full[j % S]);
        // the MMA warp's last tcgen05.commit arrivals must land before this CTA's shared memory is released
        for (int j = it - S < 0 ? 0 : it - S; j < it; ++j) mbar_wait(&s_empty[j % S], ((uint32_t)(j / S)) & 1u);
    } else if (warp == 4) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16_m128(BN);
            int it = 0, i = 0;
            for (int mt = blockIdx.x; mt < p.mtiles; mt += gridDim.x, ++i) {
                const int buf = i & 1;
                mbar_wait(&s_accempty[buf], (((uint32_t)(i >> 1)) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                for (int ks = 0; ks < nslabs; ++ks, ++it) {
                    const int s = it % S;
                    mbar_wait(&s_full[s], ((uint32_t)(it / S)) & 1u);
                    tc_fence_after();
                    const int k0 = ks * BK;
                    const int kw = p.K - k0 < BK ? p.K - k0 : BK;
                    const int ksteps = (kw + 15) >> 4;
                    const uint32_t a_base = smem_u32(smem + (size_t)s * L.stage_bytes), b_base = a_base + L.a_bytes;
                    for (int kk = 0; kk < ksteps; ++kk) {
                        const uint64_t da = umma_desc(a_base + (uint32_t)(kk * 2 * L.lboA), (uint32_t)L.lboA, 128);
                        const uint64_t db = umma_desc(b_base + (uint32_t)(kk * 2 * L.lboB), (uint32_t)L.lboB, 128);
                        umma_bf16(d_tmem, da, db, idesc, (ks | kk) ? 1u : 0u);
                    }
                    umma_commit(&s_empty[s]);                                // smem slot free when these MMAs have read it
                }
                umma_commit(&s_accfull[buf]);                                // accumulator complete
            }
        }
        __syncwarp();
    } else {
        // =========================== epilogue ===========================
        const int nchunks = BN >> 4;
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        unsigned char *myrow = stageC + (size_t)tid * L.pitchC;
        float *acc0 = s_acc + (size_t)(warp * 2) * BN, *acc1 = acc0 + BN;
        const int cols_out = p.N - n0 < BN ? p.N - n0 : BN;                  // columns of this tile that exist (multiple of 8)
        const int cpr = cols_out >> 3;

        auto prefetch_y = [&](int mt) {                                      // EPI_BNBWD: y tile -> stage (coalesced cp.async)
            const int64_t m0 = (int64_t)mt * 128;
            const int cpf = BN >> 3;                                         // all BN columns: the tail is zero-filled
            for (int q = tid; q < 128 * cpf; q += 128) {
                const int row = q / cpf, cc = q - row * cpf;
                const int64_t gr = m0 + row;
                const bool ok = gr < p.M && cc < cpr;
                const __nv_bfloat16 *src = ok ? p.Y + gr * p.ldy + n0 + cc * 8 : p.Y;
                cp_async16(stageC + (size_t)row * L.pitchC + cc * 16, src, ok ? 16 : 0);
            }
            cp_async_commit();
        };
        if (EPI == EPI_BNBWD && (int)blockIdx.x < p.mtiles) prefetch_y(blockIdx.x);

        int i = 0;
        for (int mt = blockIdx.x; mt < p.mtiles; mt += gridDim.x, ++i) {
            const int buf = i & 1;
            const int64_t m0 = (int64_t)mt * 128;
            const bool valid = m0 + tid < p.M;
            mbar_wait(&s_accfull[buf], ((uint32_t)(i >> 1)) & 1u);
            tc_fence_after();
            if (EPI == EPI_BNBWD) {
                cp_async_wait<0>();
                bar_sync_named(1, 128);                                      // y tile visible to every epilogue thread
            }
            for (int ch = 0; ch < nchunks; ++ch) {
                const int c0 = ch << 4;
                float v[16];
                tmem_ld16(trow + (uint32_t)(buf * BN + c0), v);
                uint4 *dst = reinterpret_cast<uint4 *>(myrow + c0 * 2);
                if (EPI == EPI_STORE) {
                    dst[0] = pack8(v);
                    dst[1] = pack8(v + 8);
                } else if (EPI == EPI_STATS) {
                    const uint4 u0 = pack8(v), u1 = pack8(v + 8);
                    dst[0] = u0;
                    dst[1] = u1;
                    float r[16], d1[16], d2[16];
                    unpack8(u0, r);
                    unpack8(u1, r + 8);                                      // statistics of the values as stored
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float d = valid ? r[j] - s_const[c0 + j] : 0.f;
                        d1[j] = d;
                        d2[j] = d * d;
                    }
                    const float t1 = warp_colsum16(d1, lane), t2 = warp_colsum16(d2, lane);
                    if (!(lane & 1)) {
                        acc0[c0 + (lane >> 1)] += t1;
                        acc1[c0 + (lane >> 1)] += t2;
                    }
                } else {
                    float y[16], d1[16], d2[16];
                    unpack8(dst[0], y);
                    unpack8(dst[1], y + 8);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float z = fmaf(y[j], s_const[2 * BN + c0 + j], s_const[3 * BN + c0 + j]);
                        const bool pass = !p.relu || z > 0.f;
                        v[j] = pass ? v[j] : 0.f;
                    }
                    const uint4 u0 = pack8(v), u1 = pack8(v + 8);
                    dst[0] = u0;
                    dst[1] = u1;
                    float r[16];
                    unpack8(u0, r);
                    unpack8(u1, r + 8);                                      // sums of dy as stored
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float yh = fmaf(y[j], s_const[BN + c0 + j], s_const[c0 + j]);
                        d1[j] = r[j];                                        // rows >= M: A rows are zero -> dy = 0
                        d2[j] = r[j] * yh;
                    }
                    const float t1 = warp_colsum16(d1, lane), t2 = warp_colsum16(d2, lane);
                    if (!(lane & 1)) {
                        acc0[c0 + (lane >> 1)] += t1;
                        acc1[c0 + (lane >> 1)] += t2;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&s_accempty[buf]);                                   // TMEM buffer may be overwritten
            bar_sync_named(1, 128);                                          // stage tile complete
            {
                const int rows = p.M - m0 < 128 ? (int)(p.M - m0) : 128;
                for (int q = tid; q < rows * cpr; q += 128) {
                    const int row = q / cpr, cc = q - row * cpr;
                    const uint4 t = *reinterpret_cast<const uint4 *>(stageC + (size_t)row * L.pitchC + cc * 16);
                    *reinterpret_cast<uint4 *>(p.C + (m0 + row) * p.ldc + n0 + cc * 8) = t;
                }
            }
            bar_sync_named(1, 128);                                          // stage tile free again
            if (EPI == EPI_BNBWD && mt + (int)gridDim.x < p.mtiles) prefetch_y(mt + gridDim.x);
        }
        if (EPI != EPI_STORE) {
            // per-CTA partial sums: the four row quarters (warps) added in a fixed order
            float *my = p.parts + ((size_t)nt * gridDim.x + blockIdx.x) * 2 * BN;
            for (int c = tid; c < 2 * BN; c += 128) {
                const int a = c / BN, cc = c - a * BN;
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) s += s_acc[(size_t)(w * 2 + a) * BN + cc];
                my[c] = s;
            }
            __threadfence();
        }
    }

    // =========================== teardown (+ fold of the column sums by the last CTA of this column tile) ============
    tc_fence_before();
    __syncthreads();
    if (EPI != EPI_STORE) {
        if (tid == 0) {
            __threadfence();
            const unsigned prev = atomicAdd(p.tickets + nt, 1u);
            s_last = prev == gridDim.x - 1;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            const int P = gridDim.x;
            const float *base = p.parts + (size_t)nt * P * 2 * BN;
            const float Mf = (float)p.M;
            // one warp per column, lanes over the CTA partials (fixed order: lane-strided, then a shuffle tree)
            for (int c = warp; c < BN; c += kGemmThreads / 32) {
                const int gc = n0 + c;
                float s0 = 0.f, s1 = 0.f;
                for (int j = lane; j < P; j += 32) {
                    s0 += __ldcg(base + (size_t)j * 2 * BN + c);
                    s1 += __ldcg(base + (size_t)j * 2 * BN + BN + c);
                }
#pragma unroll
                for (int off = 16; off; off >>= 1) {
                    s0 += __shfl_xor_sync(PCB_FULL_MASK, s0, off);
                    s1 += __shfl_xor_sync(PCB_FULL_MASK, s1, off);
                }
                if (lane == 0 && gc < p.N) {
                    if (EPI == EPI_STATS) {
                        if (gc < p.Cv) {
                            const float m1 = s0 / Mf;
                            const float var = fmaxf(s1 / Mf - m1 * m1, 0.f);
                            const float mu = s_const[c] + m1;                // mean of the bias-free pre-activation
                            p.mean[gc] = mu;
                            p.invstd[gc] = rsqrtf(var + p.eps);
                            if (p.running_mean) {
                                const float b = p.bias ? p.bias[gc] : 0.f;
                                p.running_mean[gc] = (1.f - p.momentum) * p.running_mean[gc] + p.momentum * (mu + b);
                                const float unbiased = p.M > 1 ? var * (Mf / (float)(p.M - 1)) : var;
                                p.running_var[gc] = (1.f - p.momentum) * p.running_var[gc] + p.momentum * unbiased;
                            }
                        } else {
                            p.mean[gc] = 0.f;
                            p.invstd[gc] = 0.f;
                        }
                    } else {
                        p.sums[gc] = gc < p.Cv ? s0 : 0.f;
                        p.sums[p.N + gc] = gc < p.Cv ? s1 : 0.f;
                        p.sums[2 * p.N + gc] = 0.f;
                    }
                }
            }
            if (tid == 0) p.tickets[nt] = 0u;                                // ready for the next launch
        }
    }
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

struct GemmPlan {
    int BN, BK, stages, ntiles, mtiles, grid_x;
    size_t smem;
};

static bool gemm_plan(int64_t M, int N, int K, int epi, GemmPlan &g)
{
    g.mtiles = (int)ceil_div(M, 128);
    g.BK = K > 32 ? 64 : (K > 16 ? 32 : 16);
    // column tiles: as few as possible (<= 256 columns each), more when the row tiles alone would leave SMs idle
    int ntiles = (N + 255) / 256;
    const int want = PCB_NUM_SMS / (g.mtiles > 0 ? g.mtiles : 1);
    const int max_split = (N + 31) / 32;
    if (want > ntiles) ntiles = want < max_split ? want : max_split;
    if (ntiles < 1) ntiles = 1;
    int bn = (int)ceil_div(N, ntiles);
    bn = (bn + 15) & ~15;
    g.BN = bn;
    g.ntiles = (int)ceil_div(N, bn);
    const size_t budget = 220 * 1024;
    int stages = kGemmMaxStages;
    for (; stages >= 2; --stages)
        if ((size_t)gemm_smem_layout(g.BN, g.BK, stages, epi).total <= budget) break;
    if (stages < 2) return false;
    // no more stages than a CTA can use: slabs per CTA
    g.grid_x = g.mtiles < PCB_NUM_SMS / g.ntiles ? g.mtiles : PCB_NUM_SMS / g.ntiles;
    if (g.grid_x < 1) g.grid_x = 1;
    g.stages = stages;
    g.smem = (size_t)gemm_smem_layout(g.BN, g.BK, stages, epi).total;
    return true;
}

template <int EPI>
static int gemm_launch(GemmParams &p, cudaStream_t st)
{
    GemmPlan g;
    if (!gemm_plan(p.M, p.N, p.K, EPI, g)) return PCB_ERANGE;
    p.BN = g.BN, p.BK = g.BK, p.stages = g.stages, p.mtiles = g.mtiles, p.ntiles = g.ntiles;
    static int attr_dev[16] = {0};                                           // per device: opt-in to > 48 KB of shared memory
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16 || !attr_dev[dev]) {
        cudaError_t e = cudaFuncSetAttribute(gemm_rows_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
        if (e != cudaSuccess) return (int)e;
        if (dev >= 0 && dev < 16) attr_dev[dev] = 1;
    }
    gemm_rows_kernel<EPI><<<dim3((unsigned)g.grid_x, (unsigned)g.ntiles), kGemmThreads, g.smem, st>>>(p);
    PCB_RETURN_LAUNCH_STATUS();
}

static inline bool al16(const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

static int gemm_check(const GemmParams &p)
{
    PCB_REQUIRE(p.A && p.B && p.C, PCB_EINVAL);
    PCB_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0 && p.Nb > 0, PCB_EINVAL);
    PCB_REQUIRE(p.N % 8 == 0 && p.K % 8 == 0 && p.lda % 8 == 0 && p.ldb % 8 == 0 && p.ldc % 8 == 0, PCB_ERANGE);
    PCB_REQUIRE(p.lda >= p.K && p.ldb >= p.K && p.ldc >= p.N && p.N <= 4096 && p.K <= 8192, PCB_ERANGE);
    PCB_REQUIRE(al16(p.A) && al16(p.B) && al16(p.C), PCB_EALIGN);
    return 0;
}

}  // namespace pcb

using namespace pcb;

// scratch floats of one statistics GEMM (partial column sums of every CTA) / number of ticket words
PCB_API int64_t pcb_gemm_work_floats(int64_t M, int N, int K)
{
    GemmPlan g;
    if (!gemm_plan(M, N, K, EPI_STATS, g)) return -1;
    return (int64_t)g.ntiles * g.grid_x * 2 * g.BN;
}

PCB_API int pcb_gemm_tickets(void) { return 128; }      // upper bound of column tiles per launch

// y[M, N] = x[M, K] . w[Nw, K]^T (rows >= Nw of the result are zero columns), bf16
PCB_API int pcb_linear_rows_bf16(const void *x, int64_t ldx, const void *w, int64_t ldw, int64_t M, int N, int Nw, int K,
                                 void *y, int64_t ldy, pcb_stream_t stream)
{
    GemmParams p = {};
    p.A = (const __nv_bfloat16 *)x, p.B = (const __nv_bfloat16 *)w, p.C = (__nv_bfloat16 *)y;
    p.lda = ldx, p.ldb = ldw, p.ldc = ldy, p.M = M, p.N = N, p.Nb = Nw, p.K = K;
    const int rc = gemm_check(p);
    if (rc) return rc;
    return gemm_launch<EPI_STORE>(p, (cudaStream_t)stream);
}

// same + training-mode BatchNorm statistics of y: mean / invstd of the bias-free output, running statistics updated
// with `momentum` (conv bias added to the running mean).  work: pcb_gemm_work_floats floats; tickets: zeroed words.
PCB_API int pcb_linear_bn_stats_rows_bf16(const void *x, int64_t ldx, const void *w, int64_t ldw, int64_t M, int N, int Nw,
                                          int K, void *y, int64_t ldy, int Cv, const float *bias, float eps, float momentum,
                                          float *running_mean, float *running_var, float *mean, float *invstd,
                                          float *work, unsigned *tickets, pcb_stream_t stream)
{
    GemmParams p = {};
    p.A = (const __nv_bfloat16 *)x, p.B = (const __nv_bfloat16 *)w, p.C = (__nv_bfloat16 *)y;
    p.lda = ldx, p.ldb = ldw, p.ldc = ldy, p.M = M, p.N = N, p.Nb = Nw, p.K = K;
    const int rc = gemm_check(p);
    if (rc) return rc;
    PCB_REQUIRE(mean && invstd && work && tickets && Cv > 0 && Cv <= N, PCB_EINVAL);
    PCB_REQUIRE(!running_mean || running_var, PCB_EINVAL);
    p.Cv = Cv, p.bias = bias, p.eps = eps, p.momentum = momentum;
    p.running_mean = running_mean, p.running_var = running_var, p.mean = mean, p.invstd = invstd;
    p.parts = work, p.tickets = tickets;
    return gemm_launch<EPI_STATS>(p, (cudaStream_t)stream);
}

// data gradient of a layer whose INPUT was z = relu(BN(y)) of the previous layer:
//   gz = gy[M, K] . wt[Nw, K]^T (wt = weight transposed: [in, out]);  dy = gz * [z > 0];  sums = (sum dy, sum dy*yhat, 0)
// dy is written to `dy` ([M, lddy], N columns), the BatchNorm backward finishes with pcb_bn_bwd_apply_rows.
PCB_API int pcb_dgrad_bn_rows_bf16(const void *gy, int64_t ldg, const void *wt, int64_t ldwt, int64_t M, int N, int Nw, int K,
                                   const void *yprev, int64_t ldyp, const float *mean, const float *invstd,
                                   const float *gamma, const float *beta, int Cv, int relu, void *dy, int64_t lddy,
                                   float *sums, float *work, unsigned *tickets, pcb_stream_t stream)
{
    GemmParams p = {};
    p.A = (const __nv_bfloat16 *)gy, p.B = (const __nv_bfloat16 *)wt, p.C = (__nv_bfloat16 *)dy;
    p.lda = ldg, p.ldb = ldwt, p.ldc = lddy, p.M = M, p.N = N, p.Nb = Nw, p.K = K;
    const int rc = gemm_check(p);
    if (rc) return rc;
    PCB_REQUIRE(yprev && mean && invstd && gamma && beta && sums && work && tickets, PCB_EINVAL);
    PCB_REQUIRE(Cv > 0 && Cv <= N && ldyp >= N && ldyp % 8 == 0 && al16(yprev), PCB_ERANGE);
    p.Cv = Cv, p.Y = (const __nv_bfloat16 *)yprev, p.ldy = ldyp, p.bn_mean = mean, p.bn_invstd = invstd, p.gamma = gamma;
    p.beta = beta, p.relu = relu, p.sums = sums, p.parts = work, p.tickets = tickets;
    return gemm_launch<EPI_BNBWD>(p, (cudaStream_t)stream);
}
