// Shared-MLP contractions of the TRAINING path on the 5th-generation tensor cores:
//   C[M, N] = A[M, K] . B[N, K]^T      bf16 operands, fp32 accumulation in TMEM, bf16 result
// with M = rows of points (10^3 .. 5*10^5), K and N = channel counts of the layer (8 .. 1536).  It replaces the
// library GEMMs behind the 1x1 Conv2d / Conv1d layers of
//   Partsize-identical/models/pointnet_util.py:213-217 (set abstraction), :275-277 (multi-scale grouping),
//   :343-345 (feature propagation), pointnet2_sem_seg.py:43-46 (head), Highway_bridge/models/pointnet2_utils.py:150-154,
//   353-356 and Highway_bridge/models/DGCNN.py:134-148 (EdgeConv)
// in the forward pass (A = activations, B = weight [out, in]) and in the data-gradient pass (A = gradient rows,
// B = weight transposed [in, out]).  Two epilogues fold the BatchNorm passes that used to follow into the GEMM:
//   EPI_STATS  forward: per-channel (count, mean, M2) of the bf16 result -- exact two-pass blocks of 16 rows merged with
//              Chan's update per thread, per CTA, per group of 16 CTAs and finally over the groups, always in a fixed
//              order -- folded to mean / invstd / variance (training-mode BatchNorm statistics);
//   EPI_BNBWD  backward: the accumulator is d loss / d z of the PREVIOUS layer's BN+ReLU output; the epilogue reads that
//              layer's pre-activation tile y, applies the ReLU mask and emits dy together with the per-channel sums of
//              dy and dy * yhat that the BatchNorm backward needs (yhat = (y - mean) * invstd).
//
// Design (second version; the first was a persistent warp-specialised kernel of 13 warps whose single-warp roles made
// every phase instruction-latency bound: 2.5-3x slower than the library on these shapes, profiles/r2_gemm_rows.md).
// The layers are memory-bound (K, N of a few dozen to a few hundred) and consist of thousands of tiny 128-row tiles, so
// the kernel is built like a classic occupancy-driven CUDA kernel around ONE tcgen05.mma chain per tile:
//   * a CTA is 128 threads = the 128 rows of a tile = the 128 TMEM lanes; 2-6 CTAs are resident per SM (shared memory
//     and <= 512 TMEM columns permitting) and hide each other's load / MMA / epilogue latencies;
//   * every thread copies operand chunks with cp.async (16 B) straight into the canonical K-major no-swizzle UMMA
//     layout (zero fill for the K / M / N tails), two stages: the slab of the NEXT step is in flight while thread 0
//     issues the MMAs of the current one; the weight slab stays resident when K fits one slab;
//   * epilogue, thread = row: tcgen05.ld -> bf16 -> 16-byte global stores straight from registers (+ the row of a
//     shared-memory tile for the statistics); then one thread per column PAIR sums its rows with plain LDS.32;
//   * CTAs loop over row tiles with stride gridDim.x, keeping their statistics in registers; the per-CTA partials are
//     folded in two deterministic levels by "last arriver" CTAs (group of 16, then all groups).
// Algorithmic bytes: 2 * M * (K + N) [+ 2 * M * N for the y tile of EPI_BNBWD]; the tensor pipe idles by construction.
#include <cuda_bf16.h>

#include "pcb_common.cuh"
#include "umma.cuh"

namespace pcb {

constexpr int kGemmThreads = 128;
constexpr int kGemmLboPad = 16;            // bytes added to the K-chunk plane stride: spreads the 16-byte units of one
                                           // row (consecutive K chunks) over different banks for the cp.async stores
constexpr int kFoldGroup = 16;             // CTAs per first-level fold group
enum { EPI_STORE = 0, EPI_STATS = 1, EPI_BNBWD = 2 };

struct GemmParams {
    const __nv_bfloat16 *A;      // [M, lda]
    const __nv_bfloat16 *B;      // [Nb, ldb]: row n = output column n
    __nv_bfloat16 *C;            // [M, ldc]
    int64_t lda, ldb, ldc, M;
    int N;                       // output columns written (multiple of 8)
    int Nb;                      // rows of B that exist (others are zero)
    int K;                       // contraction length (multiple of 8)
    int BN, BK, mtiles, ntiles, stages;
    int Cv;                      // real channels among the N columns (statistics epilogues)
    float *parts;                // [ntiles][gridDim.x + groups][3][BN] partial column statistics (CTAs, then groups)
    unsigned *tickets;           // [ntiles][1 + groups], zero on entry, zero on exit
    // EPI_STATS
    float eps;
    float *mean, *invstd, *var;  // [N] results (var: biased batch variance)
    // EPI_BNBWD
    const __nv_bfloat16 *Y;      // [M, ldy] pre-activation of the layer whose output gradient this GEMM produces
    int64_t ldy;
    const float *bn_mean, *bn_invstd, *gamma, *beta;   // [Cv]
    int relu;
    float *sums;                 // [3][N]: sum dy, sum dy*yhat, 0 (gradient of the folded conv bias)
};

__device__ __forceinline__ void unpack8(const uint4 &t, float v[8])
{
    const unsigned w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// Chan's pairwise update: (n, mean, m2) <- (n, mean, m2) + (cn, cm, cq)
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

struct GemmSmem {
    int lboA, lboB, a_bytes, stage_bytes, pitchC;
    int off_tile0, off_tile1, off_const, off_comb, total;
};

__host__ __device__ __forceinline__ GemmSmem gemm_smem_layout(int BN, int BK, int epi, int stages)
{
    GemmSmem s;
    s.lboA = 128 * 16 + kGemmLboPad;
    s.lboB = BN * 16 + kGemmLboPad;
    const int nch = BK / 8;
    s.a_bytes = nch * s.lboA;
    s.stage_bytes = (nch * (s.lboA + s.lboB) + 127) & ~127;
    s.pitchC = (BN + 8) * 2;
    const int tile = (128 * s.pitchC + 127) & ~127;
    s.off_tile0 = stages * s.stage_bytes;                                   // operand stages
    s.off_tile1 = s.off_tile0 + (epi == EPI_STORE ? 0 : tile);              // tile0: bf16 result (y tile for EPI_BNBWD)
    s.off_const = s.off_tile1 + (epi == EPI_BNBWD ? tile : 0);              // tile1: dy (EPI_BNBWD)
    s.off_comb = s.off_const + (epi == EPI_BNBWD ? 4 * BN * 4 : 0);
    s.total = s.off_comb + (epi == EPI_STORE ? 0 : 8 * BN * 3 * 4);          // row-group combine: [<= 8][BN][3]
    return s;
}

template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 4)
gemm_rows_kernel(const GemmParams p)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t s_mma[4];                               // MMAs of a stage complete
    __shared__ uint32_t s_tmem;
    __shared__ int s_flag;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int BN = p.BN, BK = p.BK;
    const int S = p.stages;                                                  // 2..4 operand stages
    const GemmSmem L = gemm_smem_layout(BN, BK, EPI, S);
    const int nt = blockIdx.y, n0 = nt * BN;
    const int nslabs = (p.K + BK - 1) / BK;
    unsigned char *tile0 = smem + L.off_tile0, *tile1 = smem + L.off_tile1;
    float *s_const = reinterpret_cast<float *>(smem + L.off_const);         // EPI_BNBWD: [nm | is | sc | sh] x BN
    float *s_comb = reinterpret_cast<float *>(smem + L.off_comb);

    int tmem_cols = 32;
    while (tmem_cols < BN) tmem_cols <<= 1;

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&s_mma[i], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), (uint32_t)tmem_cols);
    if (EPI == EPI_BNBWD) {
        for (int c = tid; c < BN; c += kGemmThreads) {
            const int gc = n0 + c;
            const bool real = gc < p.Cv;
            const float m = real ? p.bn_mean[gc] : 0.f, is = real ? p.bn_invstd[gc] : 0.f;
            const float sc = real ? is * p.gamma[gc] : 0.f;                  // same expressions as bn_rows.cu
            s_const[c] = -m * is;                                            // nm
            s_const[BN + c] = is;
            s_const[2 * BN + c] = sc;
            s_const[3 * BN + c] = real ? p.beta[gc] - m * sc : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t idesc = umma_idesc_bf16_m128(BN);
    const uint32_t smem_base = smem_u32(smem);

    // ---- operand copies: K chunk kc = tid & 7 of rows (tid >> 3) + 16 j; eight consecutive threads copy the 128
    //      contiguous bytes of one row's slab, the copies of one slab are independent instructions
    const int kc = tid & 7, r0 = tid >> 3;
    const uint32_t sa_off = (uint32_t)(kc * L.lboA + r0 * 16), sb_off = (uint32_t)(L.a_bytes + kc * L.lboB + r0 * 16);
    const int64_t a_step = 16 * p.lda * 2, b_step = 16 * p.ldb * 2;          // bytes between the rows of consecutive j
    const char *b_row = reinterpret_cast<const char *>(p.B) + ((int64_t)(n0 + r0) * p.ldb + kc * 8) * 2;
    const int b_valid = p.Nb - n0 - r0;                                      // rows n0 + r0 + 16 j exist while 16 j < b_valid
    const int b_iters = BN >> 4;
    const bool b_resident = nslabs == 1;                                     // the whole weight slab is loaded once

    auto load_slab = [&](int mt, int ks, int stage, bool with_b) {
        const int64_t m0 = (int64_t)mt * 128;
        const int k0 = ks * BK;
        const int kw = p.K - k0 < BK ? p.K - k0 : BK;                        // real columns of this slab (multiple of 8)
        const int nch = ((kw + 15) >> 4) << 1;                               // 16-byte K chunks incl. zero fill to K % 16 == 0
        if (kc < nch) {
            const bool kreal = kc < (kw >> 3);                               // else: zero fill of the K tail
            const uint32_t sdst = smem_base + (uint32_t)stage * (uint32_t)L.stage_bytes;
            const char *asrc = reinterpret_cast<const char *>(p.A) + ((m0 + r0) * p.lda + k0 + kc * 8) * 2;
            const int64_t a_left = p.M - m0 - r0;
            const int a_valid = a_left > 128 ? 128 : (int)a_left;            // row m0 + r0 + 16 j exists while 16 j < a_valid
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool ok = kreal && 16 * j < a_valid;
                cp_async16_s(sdst + sa_off + j * 256, ok ? asrc + j * a_step : reinterpret_cast<const char *>(p.A), ok ? 16 : 0);
            }
            if (with_b) {
                const char *bsrc = b_row + (int64_t)k0 * 2;
#pragma unroll 4
                for (int j = 0; j < b_iters; ++j) {
                    const bool ok = kreal && 16 * j < b_valid;
                    cp_async16_s(sdst + sb_off + j * 256, ok ? bsrc + j * b_step : reinterpret_cast<const char *>(p.B),
                                 ok ? 16 : 0);
                }
            }
        }
    };

    // ---- epilogue bookkeeping
    const int cols_out = p.N - n0 < BN ? p.N - n0 : BN;                      // columns of this tile that exist (multiple of 8)
    unsigned char *out_tile = EPI == EPI_BNBWD ? tile1 : tile0;
    unsigned char *myrow = out_tile + (size_t)tid * L.pitchC;
    // y tile prefetch (EPI_BNBWD): 16-byte chunk cc = tid % cp2 of rows tid / cp2 + rpp j
    int cp2 = 1, cp2_log = 0;
    while (cp2 < (BN >> 3)) cp2 <<= 1, ++cp2_log;
    const int cp_cc = tid & (cp2 - 1), cp_r0 = tid >> cp2_log, rpp = 128 >> cp2_log;
    auto prefetch_y = [&](int mt) {
        const int64_t m0 = (int64_t)mt * 128;
        const int rows = p.M - m0 < 128 ? (int)(p.M - m0) : 128;
        if (cp_cc < (BN >> 3)) {                                             // all BN columns: the tail is zero-filled
            const char *src = reinterpret_cast<const char *>(p.Y) + ((m0 + cp_r0) * p.ldy + n0 + cp_cc * 8) * 2;
            const int64_t step = (int64_t)rpp * p.ldy * 2;
            const uint32_t dst = smem_u32(tile0) + (uint32_t)(cp_r0 * L.pitchC + cp_cc * 16);
#pragma unroll 4
            for (int j = 0; j < cp2; ++j) {
                const bool ok = cp_cc * 8 < cols_out && cp_r0 + j * rpp < rows;
                cp_async16_s(dst + (uint32_t)(j * rpp * L.pitchC), ok ? src + j * step : reinterpret_cast<const char *>(p.Y),
                             ok ? 16 : 0);
            }
        }
    };
    // column statistics: thread = (row group rg, column pair cp)
    const int ncp = BN >> 1;
    int rgroups = 1;
    while (rgroups * 2 * ncp <= 128 && rgroups < 8) rgroups <<= 1;
    const int rpg = 128 / rgroups;                                           // 16 .. 128 rows per thread, blocks of 16
    const bool st_active = EPI != EPI_STORE && tid < ncp * rgroups;
    const int st_rg = tid / ncp, st_cp = tid - st_rg * ncp;
    float rn = 0.f, rm0 = 0.f, rq0 = 0.f, rm1 = 0.f, rq1 = 0.f;              // EPI_STATS: count, mean / M2 of the two columns
    float c_nm0 = 0.f, c_is0 = 0.f, c_nm1 = 0.f, c_is1 = 0.f;                // EPI_BNBWD: rq = sum dy, rm = sum dy*yhat
    if (EPI == EPI_BNBWD && st_active) {
        c_nm0 = s_const[2 * st_cp], c_is0 = s_const[BN + 2 * st_cp];
        c_nm1 = s_const[2 * st_cp + 1], c_is1 = s_const[BN + 2 * st_cp + 1];
    }

    // ---- pipeline over the steps (tile, K slab) of this CTA: the copies of the next S - 1 steps are in flight while the
    //      MMAs of a step run; stage = step % S; s_mma[stage] completes when the MMAs that read the stage have finished.
    //      One cp.async group per step (possibly empty), committed in step order.
    const int my_tiles = ((int)blockIdx.x < p.mtiles) ? (p.mtiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int nsteps = my_tiles * nslabs;
    auto issue_step = [&](int st) {                                          // copies of step `st` into its stage
        if (st < nsteps) {
            const int t = st / nslabs, ks = st - t * nslabs;
            load_slab(blockIdx.x + t * gridDim.x, ks, st % S, st == 0 || !b_resident);
            if (EPI == EPI_BNBWD && st == 0) prefetch_y(blockIdx.x);
        }
        cp_async_commit();
    };
    for (int st = 0; st < S - 1; ++st) issue_step(st);
    int step = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
        const int mt = blockIdx.x + ti * gridDim.x;
        const int64_t m0 = (int64_t)mt * 128;
        const int rows = p.M - m0 < 128 ? (int)(p.M - m0) : 128;
        for (int ks = 0; ks < nslabs; ++ks, ++step) {
            const int stage = step % S;
            // copies of step + S - 1 into the stage that step - 1 used, once its MMAs are done
            {
                const int nxt = step + S - 1;
                if (nxt < nsteps && step >= 1) mbar_wait(&s_mma[nxt % S], ((uint32_t)((step - 1) / S)) & 1u);
                issue_step(nxt);
            }
            switch (S) {                                                     // this step's copies (and y tile) have landed
            case 2: cp_async_wait<1>(); break;
            case 3: cp_async_wait<2>(); break;
            default: cp_async_wait<3>(); break;
            }
            proxy_fence();                                                   // generic-proxy writes -> tensor-core reads
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                const int kw = p.K - ks * BK < BK ? p.K - ks * BK : BK;
                const int ksteps = (kw + 15) >> 4;
                const uint32_t a_base = smem_base + (uint32_t)stage * (uint32_t)L.stage_bytes;
                // resident weights live in stage 0 (loaded with the first step)
                const uint32_t b_base = smem_base + (uint32_t)(b_resident ? 0 : stage) * (uint32_t)L.stage_bytes + L.a_bytes;
                for (int kk = 0; kk < ksteps; ++kk) {
                    const uint64_t da = umma_desc(a_base + (uint32_t)(kk * 2 * L.lboA), (uint32_t)L.lboA, 128);
                    const uint64_t db = umma_desc(b_base + (uint32_t)(kk * 2 * L.lboB), (uint32_t)L.lboB, 128);
                    umma_bf16(tmem_base, da, db, idesc, (ks | kk) ? 1u : 0u);
                }
                umma_commit(&s_mma[stage]);
            }
        }
        // ---- epilogue of the tile: wait for the MMAs of its last slab
        {
            const int last = step - 1;
            mbar_wait(&s_mma[last % S], ((uint32_t)(last / S)) & 1u);
            tc_fence_after();
        }
        __nv_bfloat16 *grow = p.C + (m0 + tid) * p.ldc + n0;
        const bool row_ok = tid < rows;
        for (int c0 = 0; c0 < BN; c0 += 16) {
            float v[16];
            tmem_ld16(trow + (uint32_t)c0, v);
            if (EPI == EPI_BNBWD && p.relu) {
                const uint4 *ysrc = reinterpret_cast<const uint4 *>(tile0 + (size_t)tid * L.pitchC + c0 * 2);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float y[8];
                    unpack8(ysrc[h], y);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = c0 + 8 * h + j;
                        const float z = fmaf(y[j], s_const[2 * BN + c], s_const[3 * BN + c]);
                        v[8 * h + j] = z > 0.f ? v[8 * h + j] : 0.f;
                    }
                }
            }
            const uint4 u0 = pack8(v), u1 = pack8(v + 8);
            if (row_ok && c0 < cols_out) *reinterpret_cast<uint4 *>(grow + c0) = u0;
            if (row_ok && c0 + 8 < cols_out) *reinterpret_cast<uint4 *>(grow + c0 + 8) = u1;
            if (EPI != EPI_STORE) {
                uint4 *dst = reinterpret_cast<uint4 *>(myrow + c0 * 2);
                dst[0] = u0;
                dst[1] = u1;
            }
        }
        tc_fence_before();                                                   // TMEM reads done before the next tile's MMAs
        if (EPI != EPI_STORE) {
            __syncthreads();                                                 // tile complete
            if (st_active) {
                const unsigned char *col = out_tile + st_cp * 4;
                const int r_beg = st_rg * rpg;
                const int r_end = r_beg + rpg < rows ? r_beg + rpg : rows;
                if (EPI == EPI_STATS) {
                    for (int rb = r_beg; rb < r_end; rb += 16) {
                        const int cnt = r_end - rb < 16 ? r_end - rb : 16;
                        unsigned w[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j)                         // 16 independent loads, then arithmetic
                            w[j] = j < cnt ? *reinterpret_cast<const unsigned *>(col + (size_t)(rb + j) * L.pitchC) : 0u;
                        float sa0 = 0.f, sb0 = 0.f, sa1 = 0.f, sb1 = 0.f;
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {                    // rows beyond cnt hold 0
                            sa0 += __uint_as_float(w[j] << 16);
                            sa1 += __uint_as_float(w[j] & 0xffff0000u);
                            sb0 += __uint_as_float(w[j + 1] << 16);
                            sb1 += __uint_as_float(w[j + 1] & 0xffff0000u);
                        }
                        const float cn = (float)cnt, inv = 1.f / cn;
                        const float m0c = (sa0 + sb0) * inv, m1c = (sa1 + sb1) * inv;
                        float qa0 = 0.f, qb0 = 0.f, qa1 = 0.f, qb1 = 0.f;
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const float d0 = __uint_as_float(w[j] << 16) - m0c, d1 = __uint_as_float(w[j] & 0xffff0000u) - m1c;
                            const float e0 = __uint_as_float(w[j + 1] << 16) - m0c;
                            const float e1 = __uint_as_float(w[j + 1] & 0xffff0000u) - m1c;
                            qa0 = fmaf(d0, d0, qa0);
                            qa1 = fmaf(d1, d1, qa1);
                            qb0 = fmaf(e0, e0, qb0);
                            qb1 = fmaf(e1, e1, qb1);
                        }
                        float q0 = qa0 + qb0, q1 = qa1 + qb1;
                        if (cnt < 16) {                                      // the 16 - cnt padding zeros each added mean^2
                            const float pad = (float)(16 - cnt);
                            q0 -= pad * m0c * m0c;
                            q1 -= pad * m1c * m1c;
                        }
                        float n1 = rn;
                        chan_merge(rn, rm0, rq0, cn, m0c, q0);
                        chan_merge(n1, rm1, rq1, cn, m1c, q1);
                    }
                } else if (EPI == EPI_BNBWD) {
                    const unsigned char *ycol = tile0 + st_cp * 4;
                    for (int rb = r_beg; rb < r_end; rb += 8) {
                        const int cnt = r_end - rb < 8 ? r_end - rb : 8;
                        unsigned w[8], yw[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            w[j] = j < cnt ? *reinterpret_cast<const unsigned *>(col + (size_t)(rb + j) * L.pitchC) : 0u;
                            yw[j] = j < cnt ? *reinterpret_cast<const unsigned *>(ycol + (size_t)(rb + j) * L.pitchC) : 0u;
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float d0 = __uint_as_float(w[j] << 16), d1 = __uint_as_float(w[j] & 0xffff0000u);
                            const float h0 = fmaf(__uint_as_float(yw[j] << 16), c_is0, c_nm0);
                            const float h1 = fmaf(__uint_as_float(yw[j] & 0xffff0000u), c_is1, c_nm1);
                            rq0 += d0;                                       // rows beyond cnt: dy = 0
                            rq1 += d1;
                            rm0 = fmaf(d0, h0, rm0);
                            rm1 = fmaf(d1, h1, rm1);
                        }
                    }
                }
            }
            __syncthreads();                                                 // tiles free again
            if (EPI == EPI_BNBWD && ti + 1 < my_tiles) {                     // y tile of the next row tile
                prefetch_y(mt + gridDim.x);
                cp_async_commit();
                cp_async_wait<0>();                                          // (also drains the operand copies in flight;
            }                                                                //  the wait<S-1> of later steps stays valid)
        }
    }

    if (EPI != EPI_STORE) {
        // ---- per-CTA partial statistics: the row groups merged in a fixed order through shared memory
        if (st_active) {
            float *c0p = s_comb + ((size_t)st_rg * BN + 2 * st_cp) * 3;
            c0p[0] = rn, c0p[1] = rm0, c0p[2] = rq0;
            c0p[3] = rn, c0p[4] = rm1, c0p[5] = rq1;
        }
        __syncthreads();
        const int P = gridDim.x, G = (P + kFoldGroup - 1) / kFoldGroup;
        float *lvl0 = p.parts + (size_t)nt * (P + G) * 3 * BN, *lvl1 = lvl0 + (size_t)P * 3 * BN;
        unsigned *tick = p.tickets + (size_t)nt * (1 + G);
        for (int c = tid; c < BN; c += kGemmThreads) {
            float n = 0.f, m = 0.f, q = 0.f;
            for (int g = 0; g < rgroups; ++g) {
                const float *e = s_comb + ((size_t)g * BN + c) * 3;
                if (EPI == EPI_STATS) chan_merge(n, m, q, e[0], e[1], e[2]);
                else m += e[1], q += e[2];
            }
            float *my = lvl0 + (size_t)blockIdx.x * 3 * BN;
            my[c] = n, my[BN + c] = m, my[2 * BN + c] = q;
        }
        // ---- two-level deterministic fold: the last CTA of each group of 16 merges the group (fixed order), the last
        //      group to finish merges the groups and writes the results
        __threadfence();
        __syncthreads();
        const int grp = blockIdx.x / kFoldGroup;
        const int g_lo = grp * kFoldGroup, g_hi = g_lo + kFoldGroup < P ? g_lo + kFoldGroup : P;
        if (tid == 0) s_flag = atomicAdd(tick + 1 + grp, 1u) == (unsigned)(g_hi - g_lo - 1);
        __syncthreads();
        if (s_flag) {
            __threadfence();
            for (int c = tid; c < BN; c += kGemmThreads) {
                float n = 0.f, m = 0.f, q = 0.f;
                for (int j = g_lo; j < g_hi; ++j) {
                    const float *e = lvl0 + (size_t)j * 3 * BN + c;
                    const float en = __ldcg(e), em = __ldcg(e + BN), eq = __ldcg(e + 2 * BN);
                    if (EPI == EPI_STATS) chan_merge(n, m, q, en, em, eq);
                    else m += em, q += eq;
                }
                    if (G == 1) {                                                // a single group: these are the totals
                    const int gc = n0 + c;
                    if (gc < p.N) {
                        if (EPI == EPI_STATS) {
                            const bool real = gc < p.Cv;
                            const float var = real ? fmaxf(q / (float)p.M, 0.f) : 0.f;
                            p.mean[gc] = real ? m : 0.f;
                            p.invstd[gc] = real ? rsqrtf(var + p.eps) : 0.f;
                            p.var[gc] = var;
                        } else {
                            p.sums[gc] = gc < p.Cv ? q : 0.f;
                            p.sums[p.N + gc] = gc < p.Cv ? m : 0.f;
                            p.sums[2 * p.N + gc] = 0.f;
                        }
                    }
                } else {
                    float *o = lvl1 + (size_t)grp * 3 * BN;
                    o[c] = n, o[BN + c] = m, o[2 * BN + c] = q;
                }
            }
            if (G == 1) {
                if (tid == 0) tick[1 + grp] = 0u;
                goto fold_done;
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                tick[1 + grp] = 0u;                                          // ready for the next launch
                s_flag = atomicAdd(tick, 1u) == (unsigned)(G - 1);
            }
            __syncthreads();
            if (s_flag) {
                __threadfence();
                for (int c = tid; c < BN; c += kGemmThreads) {
                    const int gc = n0 + c;
                    float n = 0.f, m = 0.f, q = 0.f;
                    for (int j = 0; j < G; ++j) {
                        const float *e = lvl1 + (size_t)j * 3 * BN + c;
                        const float en = __ldcg(e), em = __ldcg(e + BN), eq = __ldcg(e + 2 * BN);
                        if (EPI == EPI_STATS) chan_merge(n, m, q, en, em, eq);
                        else m += em, q += eq;
                    }
                    if (gc < p.N) {
                        if (EPI == EPI_STATS) {
                            const bool real = gc < p.Cv;
                            const float var = real ? fmaxf(q / (float)p.M, 0.f) : 0.f;
                            p.mean[gc] = real ? m : 0.f;                     // mean of the bias-free pre-activation
                            p.invstd[gc] = real ? rsqrtf(var + p.eps) : 0.f;
                            p.var[gc] = var;                                 // running statistics: pcb_bn_apply_rows
                        } else {
                            p.sums[gc] = gc < p.Cv ? q : 0.f;                // sum dy
                            p.sums[p.N + gc] = gc < p.Cv ? m : 0.f;          // sum dy * yhat
                            p.sums[2 * p.N + gc] = 0.f;
                        }
                    }
                }
                if (tid == 0) tick[0] = 0u;
            }
        }
    }

fold_done:
    // ---- teardown: every MMA has completed (each tile's epilogue waited for its last commit)
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

struct GemmPlan {
    int BN, BK, ntiles, mtiles, grid_x, ctas_per_sm, groups, stages;
    size_t smem;
};

static bool gemm_plan(int64_t M, int N, int K, int epi, GemmPlan &g)
{
    g.mtiles = (int)ceil_div(M, 128);
    g.BK = K > 32 ? 64 : (K > 16 ? 32 : 16);
    // column tiles of <= 128 columns (several CTAs per SM must fit: shared memory, TMEM), more of them when the row
    // tiles alone would leave SMs idle
    int ntiles = (N + 127) / 128;
    const int want = 2 * PCB_NUM_SMS / (g.mtiles > 0 ? g.mtiles : 1);
    const int max_split = (N + 31) / 32;
    if (want > ntiles) ntiles = want < max_split ? want : max_split;
    if (ntiles < 1) ntiles = 1;
    int bn = (int)ceil_div(N, ntiles);
    bn = (bn + 15) & ~15;
    g.BN = bn;
    g.ntiles = (int)ceil_div(N, bn);
    int tmem_cols = 32;
    while (tmem_cols < bn) tmem_cols <<= 1;
    const int nslabs = (K + g.BK - 1) / g.BK;
    // operand stages: two when many CTAs per SM hide the latency for each other; deeper (up to 4, never more than the
    // K slabs + 1) when the tiles are few, so that a CTA has several slabs in flight on its own
    auto occupancy = [&](int stages) {
        const size_t smem = (size_t)gemm_smem_layout(g.BN, g.BK, epi, stages).total;
        int c = (int)((224 * 1024) / (smem + 1024));
        if (c > 512 / tmem_cols) c = 512 / tmem_cols;
        return c > 6 ? 6 : c;
    };
    int stages = 2;
    if (occupancy(2) < 1) return false;
    const int64_t tiles = (int64_t)g.mtiles * g.ntiles;
    while (stages < 4 && stages < nslabs + 1 && occupancy(stages + 1) >= 1 &&
           (int64_t)PCB_NUM_SMS * occupancy(stages + 1) >= tiles)
        ++stages;                                                            // deeper only while every tile still gets its own CTA
    g.stages = stages;
    g.smem = (size_t)gemm_smem_layout(g.BN, g.BK, epi, stages).total;
    g.ctas_per_sm = occupancy(stages);
    const int slots = PCB_NUM_SMS * g.ctas_per_sm / g.ntiles;
    g.grid_x = g.mtiles < slots ? g.mtiles : slots;
    if (g.grid_x < 1) g.grid_x = 1;
    g.groups = (g.grid_x + kFoldGroup - 1) / kFoldGroup;
    return true;
}

template <int EPI>
static int gemm_launch(GemmParams &p, cudaStream_t st)
{
    GemmPlan g;
    if (!gemm_plan(p.M, p.N, p.K, EPI, g)) return PCB_ERANGE;
    p.BN = g.BN, p.BK = g.BK, p.mtiles = g.mtiles, p.ntiles = g.ntiles, p.stages = g.stages;
    static bool attr_set[kMaxDevices] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const bool first = dev >= 0 && dev < kMaxDevices && !attr_set[dev];
    if (cudaError_t e = smem_optin_once(gemm_rows_kernel<EPI>, 224 * 1024, attr_set)) return (int)e;
    if (first)      // several CTAs per SM each want tens of KB: ask for the largest shared-memory carve-out
        cudaFuncSetAttribute(gemm_rows_kernel<EPI>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    // the grid is one wave: never more CTAs than are resident at once (registers may allow fewer than the plan assumed)
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gemm_rows_kernel<EPI>, kGemmThreads, g.smem);
    if (occ >= 1 && occ < g.ctas_per_sm) {
        const int slots = PCB_NUM_SMS * occ / g.ntiles;
        if (slots >= 1 && g.grid_x > slots) g.grid_x = slots;
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

// scratch floats of one statistics GEMM (partial column statistics of every CTA and fold group), either epilogue
PCB_API int64_t pcb_gemm_work_floats(int64_t M, int N, int K)
{
    GemmPlan g;
    int64_t need = 0;
    for (int epi = EPI_STATS; epi <= EPI_BNBWD; ++epi) {
        if (!gemm_plan(M, N, K, epi, g)) return -1;
        const int64_t n = (int64_t)g.ntiles * (g.grid_x + g.groups) * 3 * g.BN;
        if (n > need) need = n;
    }
    return need;
}

// ticket words a statistics GEMM may use: ntiles * (1 + groups) <= this
PCB_API int pcb_gemm_tickets(void) { return 1024; }

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

// same + training-mode BatchNorm statistics of y: mean / invstd / biased variance of the bias-free output (pcb_bn_apply_rows
// turns them into the running statistics).  work: pcb_gemm_work_floats floats; tickets: zeroed words.
PCB_API int pcb_linear_bn_stats_rows_bf16(const void *x, int64_t ldx, const void *w, int64_t ldw, int64_t M, int N, int Nw,
                                          int K, void *y, int64_t ldy, int Cv, float eps, float *mean, float *invstd,
                                          float *var, float *work, unsigned *tickets, pcb_stream_t stream)
{
    GemmParams p = {};
    p.A = (const __nv_bfloat16 *)x, p.B = (const __nv_bfloat16 *)w, p.C = (__nv_bfloat16 *)y;
    p.lda = ldx, p.ldb = ldw, p.ldc = ldy, p.M = M, p.N = N, p.Nb = Nw, p.K = K;
    const int rc = gemm_check(p);
    if (rc) return rc;
    PCB_REQUIRE(mean && invstd && var && work && tickets && Cv > 0 && Cv <= N, PCB_EINVAL);
    p.Cv = Cv, p.eps = eps, p.mean = mean, p.invstd = invstd, p.var = var;
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
