// Shared-MLP contractions of the TRAINING path on the 5th-generation tensor cores:
//   C[M, N] = A[M, K] . B[N, K]^T      bf16 operands, fp32 accumulation in TMEM, bf16 result
// with M = rows of points (10^3 .. 5*10^5), K and N = channel counts of the layer (8 .. 1536).  It replaces the
// library GEMMs behind the 1x1 Conv2d / Conv1d layers of
//   Partsize-identical/models/pointnet_util.py:213-217 (set abstraction), :275-277 (multi-scale grouping),
//   :343-345 (feature propagation), pointnet2_sem_seg.py:43-46 (head), Highway_bridge/models/pointnet2_utils.py:150-154,
//   353-356 and Highway_bridge/models/DGCNN.py:134-148 (EdgeConv)
// in the forward pass (A = activations, B = weight [out, in]) and in the data-gradient pass (A = gradient rows,
// B = weight transposed [in, out]).  Two epilogues fold the BatchNorm passes that used to follow into the GEMM:
//   EPI_STATS  forward: per-channel (count, mean, M2) of the bf16 result -- exact two-pass blocks of <= 16 rows merged with
//              Chan's update, per thread, per CTA and finally by the last CTA, in a fixed order -- folded to mean / invstd /
//              variance (training-mode BatchNorm statistics, torch.nn.functional.batch_norm);
//   EPI_BNBWD  backward: the accumulator is d loss / d z of the PREVIOUS layer's BN+ReLU output; the epilogue reads that
//              layer's pre-activation tile y, applies the ReLU mask and emits dy together with the per-channel sums of
//              dy and dy * yhat that the BatchNorm backward needs (yhat = (y - mean) * invstd).
//
// Persistent, warp-specialised CTAs of 416 threads:
//   warps 0-7  epilogue   two groups of four warps, each owning half of the tile's columns: tcgen05.ld (warp w reads TMEM
//                         lanes 32 (w % 4) .. +31 = rows of the tile) -> bf16 -> 16-byte global stores from registers (+
//                         shared-memory tile for the column statistics, one thread per column pair)
//   warp  8    MMA        one elected thread issues tcgen05.mma (M = 128, N = BN <= 256, K = 16 per instruction) and
//                         tcgen05.commit; owns the TMEM allocation (two accumulator buffers: the epilogue of tile i
//                         overlaps the MMAs of tile i + 1)
//   warps 9-12 producers  cp.async (16 B) global -> shared memory straight into the canonical K-major no-swizzle UMMA
//                         layout, zero fill for the K / M / N tails; each thread's arrival on the stage's mbarrier is
//                         triggered by the completion of its own copies (cp.async.mbarrier.arrive.noinc), so up to
//                         `stages` slabs are in flight and the producers only ever wait for a free stage
// Pipelines: full/empty mbarriers per ring stage (producers <-> MMA), full/empty per accumulator buffer (MMA <->
// epilogue).  The operands are small-K / small-N matrices streamed once: the kernel is HBM-bound (algorithmic bytes
// 2 * M * (K + N) [+ 2 * M * N for the y tile of EPI_BNBWD]); the tensor pipe idles most of the time by construction.
#include <cuda_bf16.h>

#include "pcb_common.cuh"
#include "umma.cuh"

namespace pcb {

#ifdef PCB_GEMM_TRACE   // kernel-tuning aid: clock64 stamps of CTA (0, 0): [role][event] (role 0 producer, 1 MMA, 2 epilogue)
__device__ long long g_gemm_trace[3][512];
#define GEMM_STAMP(role, slot)                                                              \
    do {                                                                                    \
        if (blockIdx.x == 0 && blockIdx.y == 0 && (slot) < 512) g_gemm_trace[role][slot] = clock64(); \
    } while (0)
#else
#define GEMM_STAMP(role, slot)
#endif
constexpr int kGemmThreads = 416;            // 8 epilogue warps + MMA warp + 4 producer warps
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
    float *parts;                // [ntiles][gridDim.x][3][BN] per-CTA partial column statistics
    unsigned *tickets;           // [ntiles], zero on entry, zero on exit
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

__host__ __device__ __forceinline__ GemmSmem gemm_smem_layout(int BN, int BK, int stages, int epi)
{
    GemmSmem s;
    s.lboA = 128 * 16 + kGemmLboPad;
    s.lboB = BN * 16 + kGemmLboPad;
    const int nch = BK / 8;
    s.a_bytes = nch * s.lboA;
    s.stage_bytes = (nch * (s.lboA + s.lboB) + 127) & ~127;
    s.pitchC = (BN + 8) * 2;
    const int tile = (128 * s.pitchC + 127) & ~127;                         // >= 6 KB: also the end-of-kernel combine scratch
    s.off_tile0 = stages * s.stage_bytes;
    s.off_tile1 = s.off_tile0 + tile;
    s.off_const = s.off_tile1 + (epi == EPI_BNBWD ? tile : 0);
    s.off_comb = s.off_const + (epi == EPI_BNBWD ? 4 * BN * 4 : 0);
    s.total = s.off_comb + (epi == EPI_STORE ? 0 : 2 * 768 * 4);
    return s;
}

// One persistent CTA: see the header comment.  Epilogue of one 128 x BN tile: two groups of four warps, group g owning
// the column range [cb, ce) (units of 16 columns split in two), thread = (group, row): row = TMEM lane.
//   A. tcgen05.ld 16 columns at a time -> bf16 -> 16-byte global stores of the thread's row straight from registers and,
//      for the statistics epilogues, the thread's row of the shared-memory tile (EPI_BNBWD: the y tile was prefetched
//      into tile0 with cp.async; dy = acc * [BN(y) > 0] goes to tile1); the TMEM buffer is released right after;
//   B. after a barrier among the 128 threads of the group: column statistics with one thread per column PAIR and row
//      group (independent LDS.32, then arithmetic):
//        EPI_STATS  per block of <= 16 rows an exact two-pass (count, mean, M2), merged into the thread's running
//                   triple with Chan's update -- no cancellation whatever the channel mean, no dependence on any state;
//        EPI_BNBWD  running sums of dy and dy * yhat, yhat = fma(y, invstd, -mean * invstd) from the y tile.
template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 2)
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
    unsigned char *tile0 = smem + L.off_tile0, *tile1 = smem + L.off_tile1;
    float *s_const = reinterpret_cast<float *>(smem + L.off_const);       // EPI_BNBWD: [nm | is | sc | sh] x BN
    float *s_comb = reinterpret_cast<float *>(smem + L.off_comb);         // end-of-kernel combine: [2 groups][768]

    int tmem_cols = 32;
    while (tmem_cols < 2 * BN) tmem_cols <<= 1;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&s_full[s], 128);
            mbar_init(&s_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&s_accfull[b], 1);
            mbar_init(&s_accempty[b], 256);
        }
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc(smem_u32(&s_tmem), (uint32_t)tmem_cols);
    if (EPI == EPI_BNBWD && warp < 8) {
        for (int c = tid; c < BN; c += 256) {
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

    if (warp >= 9) {
        // =========================== producers ===========================
        // Fixed thread -> chunk mapping: K chunk kc = pt & 7 of rows (pt >> 3) + 16 j.  Eight consecutive threads copy the
        // 128 contiguous bytes of one row's slab; no index arithmetic beyond pointer increments inside the loops, and
        // the copies of one slab are independent instructions (the loops are unrolled).
        const int pt = tid - 288;
        const int kc = pt & 7, r0 = pt >> 3;
        const uint32_t sa_off = (uint32_t)(kc * L.lboA + r0 * 16), sb_off = (uint32_t)(L.a_bytes + kc * L.lboB + r0 * 16);
        const uint32_t smem_base = smem_u32(smem);
        const int64_t a_step = 16 * p.lda * 2, b_step = 16 * p.ldb * 2;      // bytes between the rows of consecutive j
        const char *b_row = reinterpret_cast<const char *>(p.B) + ((int64_t)(n0 + r0) * p.ldb + kc * 8) * 2;
        const int b_valid = p.Nb - n0 - r0;                                  // rows n0 + r0 + 16 j exist while 16 j < b_valid
        const int b_iters = BN >> 4;
        int it = 0;
        for (int mt = blockIdx.x; mt < p.mtiles; mt += gridDim.x) {
            const int64_t m0 = (int64_t)mt * 128;
            const char *a_row = reinterpret_cast<const char *>(p.A) + ((m0 + r0) * p.lda + kc * 8) * 2;
            const int64_t a_left = p.M - m0 - r0;                            // row m0 + r0 + 16 j exists while 16 j < a_left
            const int a_valid = a_left > 128 ? 128 : (int)a_left;
            for (int ks = 0; ks < nslabs; ++ks, ++it) {
                const int s = it % S;
                mbar_wait(&s_empty[s], (((uint32_t)(it / S)) & 1u) ^ 1u);
                if (pt == 0) GEMM_STAMP(0, 2 * it);
                const int k0 = ks * BK;
                const int kw = p.K - k0 < BK ? p.K - k0 : BK;                // real columns of this slab (multiple of 8)
                const int nch = ((kw + 15) >> 4) << 1;                       // 16-byte K chunks incl. zero fill to K % 16 == 0
                const int vch = kw >> 3;
                if (kc < nch) {
                    const bool kreal = kc < vch;                             // else: zero fill of the K tail
                    const uint32_t sdst = smem_base + (uint32_t)s * (uint32_t)L.stage_bytes;
                    const char *asrc = a_row + (int64_t)k0 * 2;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const bool ok = kreal && 16 * j < a_valid;
                        cp_async16_s(sdst + sa_off + j * 256, ok ? asrc + j * a_step : reinterpret_cast<const char *>(p.A),
                                     ok ? 16 : 0);
                    }
                    const char *bsrc = b_row + (int64_t)k0 * 2;
#pragma unroll 4
                    for (int j = 0; j < b_iters; ++j) {
                        const bool ok = kreal && 16 * j < b_valid;
                        cp_async16_s(sdst + sb_off + j * 256, ok ? bsrc + j * b_step : reinterpret_cast<const char *>(p.B),
                                     ok ? 16 : 0);
                    }
                }
                // this thread's arrival on the stage's `full` barrier fires when its copies above have landed
                // (cp.async.mbarrier.arrive.noinc): the producers never wait for data, only for free stages
                cp_async_mbar_arrive_noinc(&s_full[s]);
                if (pt == 0) GEMM_STAMP(0, 2 * it + 1);
            }
        }
        // the MMA warp's last tcgen05.commit arrivals must land before this CTA's shared memory is released
        for (int j = it - S < 0 ? 0 : it - S; j < it; ++j) mbar_wait(&s_empty[j % S], ((uint32_t)(j / S)) & 1u);
    } else if (warp == 8) {
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
                    GEMM_STAMP(1, 2 * it);
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
                    GEMM_STAMP(1, 2 * it + 1);
                }
                umma_commit(&s_accfull[buf]);                                // accumulator complete
            }
        }
        __syncwarp();
    } else {
        // =========================== epilogue ===========================
        const int grp = warp >> 2, gtid = tid & 127;                         // group, row of the tile (= TMEM lane)
        const int bar_id = 1 + grp;
        const int units = BN >> 4, u0 = (units + 1) >> 1;
        const int cb = grp == 0 ? 0 : u0 * 16, ce = grp == 0 ? u0 * 16 : BN;  // this group's columns
        const int gcols = ce - cb;                                           // may be 0 (BN == 16: group 1 idles)
        const uint32_t trow = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        unsigned char *out_tile = EPI == EPI_BNBWD ? tile1 : tile0;
        unsigned char *myrow = out_tile + (size_t)gtid * L.pitchC;
        const int cols_out = p.N - n0 < BN ? p.N - n0 : BN;                  // columns of this tile that exist (multiple of 8)
        const int st_cols = (cols_out < ce ? cols_out : ce) - cb;            // columns of this group that are stored
        // coalesced global -> tile mapping of the y prefetch: 16-byte chunk cc = gtid % cp2 of rows gtid / cp2 + rpp j
        int cp2 = 1, cp2_log = 0;
        while (cp2 < (gcols >> 3)) cp2 <<= 1, ++cp2_log;
        const int cp_cc = gtid & (cp2 - 1), cp_r0 = gtid >> cp2_log, rpp = 128 >> cp2_log, passes = cp2;
        // column statistics: thread = (row group rg, column pair cp) over this group's columns
        const int ncp = gcols >> 1;
        int rgroups = 1;
        while (ncp > 0 && rgroups * 2 * ncp <= 128 && rgroups < 8) rgroups <<= 1;
        const int rpg = 128 / rgroups;                                       // 16 .. 128 rows per thread, blocks of 16
        const bool st_active = EPI != EPI_STORE && gtid < ncp * rgroups;
        const int st_rg = ncp > 0 ? gtid / ncp : 0, st_cp = gtid - st_rg * ncp;
        float rn = 0.f, rm0 = 0.f, rq0 = 0.f, rm1 = 0.f, rq1 = 0.f;          // EPI_STATS: count, mean / M2 of the two columns
        float c_nm0 = 0.f, c_is0 = 0.f, c_nm1 = 0.f, c_is1 = 0.f;            // EPI_BNBWD: rq0/rq1 = sum dy, rm0/rm1 = sum dy*yhat
        if (EPI == EPI_BNBWD && st_active) {
            c_nm0 = s_const[cb + 2 * st_cp], c_is0 = s_const[BN + cb + 2 * st_cp];
            c_nm1 = s_const[cb + 2 * st_cp + 1], c_is1 = s_const[BN + cb + 2 * st_cp + 1];
        }

        auto prefetch_y = [&](int mt) {                                      // EPI_BNBWD: y tile -> tile0 (coalesced cp.async)
            const int64_t m0 = (int64_t)mt * 128;
            const int rows = p.M - m0 < 128 ? (int)(p.M - m0) : 128;
            if (cp_cc < (gcols >> 3)) {                                      // all of the group's columns: the tail is zero-filled
                const char *src = reinterpret_cast<const char *>(p.Y) + ((m0 + cp_r0) * p.ldy + n0 + cb + cp_cc * 8) * 2;
                const int64_t step = (int64_t)rpp * p.ldy * 2;
                const uint32_t dst = smem_u32(tile0) + (uint32_t)(cp_r0 * L.pitchC + cb * 2 + cp_cc * 16);
#pragma unroll 4
                for (int j = 0; j < passes; ++j) {
                    const bool ok = cp_cc * 8 < st_cols && cp_r0 + j * rpp < rows;
                    cp_async16_s(dst + (uint32_t)(j * rpp * L.pitchC), ok ? src + j * step : reinterpret_cast<const char *>(p.Y),
                                 ok ? 16 : 0);
                }
            }
            cp_async_commit();
        };
        if (EPI == EPI_BNBWD && gcols > 0) prefetch_y(blockIdx.x);

        int i = 0;
        for (int mt = blockIdx.x; mt < p.mtiles; mt += gridDim.x, ++i) {
            const int buf = i & 1;
            const int64_t m0 = (int64_t)mt * 128;
            const int rows = p.M - m0 < 128 ? (int)(p.M - m0) : 128;
            if (tid == 0) GEMM_STAMP(2, 4 * i);
            mbar_wait(&s_accfull[buf], ((uint32_t)(i >> 1)) & 1u);
            if (tid == 0) GEMM_STAMP(2, 4 * i + 1);
            tc_fence_after();
            if (gcols > 0) {
                if (EPI == EPI_BNBWD) {
                    cp_async_wait<0>();
                    bar_sync_named(bar_id, 128);                             // y tile visible to the whole group
                }
                // ---- A: accumulator row -> bf16 -> global row (16-byte stores straight from registers: the two halves
                //      of a 32-byte sector are written by consecutive instructions of the same thread and merge in L2) and,
                //      for the statistics epilogues, the shared-memory tile; 16 columns per TMEM load
                __nv_bfloat16 *grow = p.C + (m0 + gtid) * p.ldc + n0;
                const bool row_ok = gtid < rows;
                for (int c0 = cb; c0 < ce; c0 += 16) {
                    float v[16];
                    tmem_ld16(trow + (uint32_t)(buf * BN + c0), v);
                    if (EPI == EPI_BNBWD && p.relu) {
                        const uint4 *ysrc = reinterpret_cast<const uint4 *>(tile0 + (size_t)gtid * L.pitchC + c0 * 2);
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
            }
            tc_fence_before();
            mbar_arrive(&s_accempty[buf]);                                   // TMEM buffer may be overwritten
            if (gcols > 0) {
                if (EPI != EPI_STORE) {
                    bar_sync_named(bar_id, 128);                             // tile complete for the group
                    if (tid == 0) GEMM_STAMP(2, 4 * i + 2);
                    // ---- B: column statistics
                    if (st_active) {
                        const unsigned char *col = out_tile + cb * 2 + st_cp * 4;
                        const int r_beg = st_rg * rpg;
                        const int r_end = r_beg + rpg < rows ? r_beg + rpg : rows;
                        if (EPI == EPI_STATS) {
                            for (int rb = r_beg; rb < r_end; rb += 16) {
                                const int cnt = r_end - rb < 16 ? r_end - rb : 16;
                                unsigned w[16];
#pragma unroll
                                for (int j = 0; j < 16; ++j)                 // 16 independent loads, then arithmetic
                                    w[j] = j < cnt ? *reinterpret_cast<const unsigned *>(col + (size_t)(rb + j) * L.pitchC) : 0u;
                                float sa0 = 0.f, sb0 = 0.f, sa1 = 0.f, sb1 = 0.f;
#pragma unroll
                                for (int j = 0; j < 16; j += 2) {
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
                                    const float d0 = j < cnt ? __uint_as_float(w[j] << 16) - m0c : 0.f;
                                    const float d1 = j < cnt ? __uint_as_float(w[j] & 0xffff0000u) - m1c : 0.f;
                                    const float e0 = j + 1 < cnt ? __uint_as_float(w[j + 1] << 16) - m0c : 0.f;
                                    const float e1 = j + 1 < cnt ? __uint_as_float(w[j + 1] & 0xffff0000u) - m1c : 0.f;
                                    qa0 = fmaf(d0, d0, qa0);
                                    qa1 = fmaf(d1, d1, qa1);
                                    qb0 = fmaf(e0, e0, qb0);
                                    qb1 = fmaf(e1, e1, qb1);
                                }
                                float n1 = rn;
                                chan_merge(rn, rm0, rq0, cn, m0c, qa0 + qb0);
                                chan_merge(n1, rm1, rq1, cn, m1c, qa1 + qb1);
                            }
                        } else if (EPI == EPI_BNBWD) {
                            const unsigned char *ycol = tile0 + cb * 2 + st_cp * 4;
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
                                    rq0 += d0;                               // rows beyond cnt: dy = 0
                                    rq1 += d1;
                                    rm0 = fmaf(d0, h0, rm0);
                                    rm1 = fmaf(d1, h1, rm1);
                                }
                            }
                        }
                    }
                    bar_sync_named(bar_id, 128);                             // tiles free again
                    if (tid == 0) GEMM_STAMP(2, 4 * i + 3);
                    if (EPI == EPI_BNBWD && mt + (int)gridDim.x < p.mtiles) prefetch_y(mt + gridDim.x);
                }
            }
        }
        if (EPI != EPI_STORE && gcols > 0) {
            // per-CTA partial statistics: the row groups merged in a fixed order through shared memory
            float *comb = s_comb + grp * 768;                                // [rgroups][gcols][3], rgroups * gcols <= 256
            if (st_active) {
                float *c0p = comb + ((size_t)st_rg * gcols + 2 * st_cp) * 3;
                c0p[0] = rn, c0p[1] = rm0, c0p[2] = rq0;
                c0p[3] = rn, c0p[4] = rm1, c0p[5] = rq1;
            }
            bar_sync_named(bar_id, 128);
            float *my = p.parts + ((size_t)nt * gridDim.x + blockIdx.x) * 3 * BN;
            for (int c = gtid; c < gcols; c += 128) {
                float n = 0.f, m = 0.f, q = 0.f;
                for (int g = 0; g < rgroups; ++g) {
                    const float *e = comb + ((size_t)g * gcols + c) * 3;
                    if (EPI == EPI_STATS) chan_merge(n, m, q, e[0], e[1], e[2]);
                    else m += e[1], q += e[2];
                }
                my[cb + c] = n, my[BN + cb + c] = m, my[2 * BN + cb + c] = q;
            }
            __threadfence();
        }
    }

    // =========================== teardown (+ fold of the column statistics by the last CTA of this column tile) =====
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
            const float *base = p.parts + (size_t)nt * P * 3 * BN;
            // one warp per column, lanes over the CTA partials (fixed order: lane-strided, then a shuffle tree)
            for (int c = warp; c < BN; c += kGemmThreads / 32) {
                const int gc = n0 + c;
                float n = 0.f, m = 0.f, q = 0.f;
                for (int j = lane; j < P; j += 32) {
                    const float *e = base + (size_t)j * 3 * BN + c;
                    const float en = __ldcg(e), em = __ldcg(e + BN), eq = __ldcg(e + 2 * BN);
                    if (EPI == EPI_STATS) chan_merge(n, m, q, en, em, eq);
                    else m += em, q += eq;
                }
#pragma unroll
                for (int off = 16; off; off >>= 1) {
                    const float on = __shfl_xor_sync(PCB_FULL_MASK, n, off), om = __shfl_xor_sync(PCB_FULL_MASK, m, off);
                    const float oq = __shfl_xor_sync(PCB_FULL_MASK, q, off);
                    if (EPI == EPI_STATS) chan_merge(n, m, q, on, om, oq);
                    else m += om, q += oq;
                }
                if (lane == 0 && gc < p.N) {
                    if (EPI == EPI_STATS) {
                        const bool real = gc < p.Cv;
                        const float var = real ? fmaxf(q / (float)p.M, 0.f) : 0.f;
                        p.mean[gc] = real ? m : 0.f;                         // mean of the bias-free pre-activation
                        p.invstd[gc] = real ? rsqrtf(var + p.eps) : 0.f;
                        p.var[gc] = var;                                     // running statistics: pcb_bn_apply_rows
                    } else {
                        p.sums[gc] = gc < p.Cv ? q : 0.f;                    // sum dy
                        p.sums[p.N + gc] = gc < p.Cv ? m : 0.f;              // sum dy * yhat
                        p.sums[2 * p.N + gc] = 0.f;
                    }
                }
            }
            if (tid == 0) p.tickets[nt] = 0u;                                // ready for the next launch
        }
    }
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

struct GemmPlan {
    int BN, BK, stages, ntiles, mtiles, grid_x, ctas_per_sm;
    size_t smem;
};

static bool gemm_plan(int64_t M, int N, int K, int epi, GemmPlan &g)
{
    g.mtiles = (int)ceil_div(M, 128);
    g.BK = K > 32 ? 64 : (K > 16 ? 32 : 16);
    // column tiles: as few as possible (<= 256 columns each; <= 128 for the BN-backward epilogue, which keeps two
    // tiles in shared memory), more when the row tiles alone would leave SMs idle
    const int bn_max = epi == EPI_BNBWD ? 128 : 256;
    int ntiles = (N + bn_max - 1) / bn_max;
    const int want = PCB_NUM_SMS / (g.mtiles > 0 ? g.mtiles : 1);
    const int max_split = (N + 31) / 32;
    if (want > ntiles) ntiles = want < max_split ? want : max_split;
    if (ntiles < 1) ntiles = 1;
    int bn = (int)ceil_div(N, ntiles);
    bn = (bn + 15) & ~15;
    g.BN = bn;
    g.ntiles = (int)ceil_div(N, bn);
    // two CTAs per SM (their epilogues overlap) when TMEM (2 * BN columns each) and shared memory allow >= 4 stages
    int tmem_cols = 32;
    while (tmem_cols < 2 * bn) tmem_cols <<= 1;
    g.ctas_per_sm = 1;
    int stages = 0;
    if (tmem_cols <= 256) {
        for (stages = kGemmMaxStages; stages >= 4; --stages)
            if ((size_t)gemm_smem_layout(g.BN, g.BK, stages, epi).total <= 108 * 1024) break;
        if (stages >= 4) g.ctas_per_sm = 2;
    }
    if (g.ctas_per_sm == 1) {
        for (stages = kGemmMaxStages; stages >= 2; --stages)
            if ((size_t)gemm_smem_layout(g.BN, g.BK, stages, epi).total <= 220 * 1024) break;
        if (stages < 2) return false;
    }
    const int slots = PCB_NUM_SMS * g.ctas_per_sm / g.ntiles;
    g.grid_x = g.mtiles < slots ? g.mtiles : slots;
    if (g.grid_x < 1) g.grid_x = 1;
    g.stages = stages;
    g.smem = (size_t)gemm_smem_layout(g.BN, g.BK, stages, epi).total;
    // a CTA that needs all 512 TMEM columns must not share an SM with another one (its allocation would block)
    if (tmem_cols > 256 && g.smem < 116 * 1024) g.smem = 116 * 1024;
    return true;
}

template <int EPI>
static int gemm_launch(GemmParams &p, cudaStream_t st)
{
    GemmPlan g;
    if (!gemm_plan(p.M, p.N, p.K, EPI, g)) return PCB_ERANGE;
    p.BN = g.BN, p.BK = g.BK, p.stages = g.stages, p.mtiles = g.mtiles, p.ntiles = g.ntiles;
    static bool attr_set[kMaxDevices] = {};
    if (cudaError_t e = smem_optin_once(gemm_rows_kernel<EPI>, 224 * 1024, attr_set)) return (int)e;
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

#ifdef PCB_GEMM_TRACE
PCB_API int pcb_gemm_debug_trace(long long *host_out)
{
    return (int)cudaMemcpyFromSymbol(host_out, g_gemm_trace, sizeof(g_gemm_trace));
}
#endif

// scratch floats of one statistics GEMM (partial column sums of every CTA) / number of ticket words
PCB_API int64_t pcb_gemm_work_floats(int64_t M, int N, int K)
{
    GemmPlan g;
    // sized for either statistics epilogue (the BN-backward one uses narrower column tiles)
    int64_t need = 0;
    for (int epi = EPI_STATS; epi <= EPI_BNBWD; ++epi) {
        if (!gemm_plan(M, N, K, epi, g)) return -1;
        const int64_t n = (int64_t)g.ntiles * g.grid_x * 3 * g.BN;
        if (n > need) need = n;
    }
    return need;
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
