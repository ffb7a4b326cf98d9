// Shared-MLP contractions of the TRAINING path on the 5th-generation tensor cores:
//   C[M, N] = A[M, K] . B[N, K]^T      bf16 operands, fp32 accumulation in TMEM, bf16 result
// with M = rows of points (10^3 .. 5*10^5), K and N = channel counts of the layer (8 .. 1536).  It replaces the
// library GEMMs behind the 1x1 Conv2d / Conv1d layers of
//   Partsize-identical/models/pointnet_util.py:213-217 (set abstraction), :275-277 (multi-scale grouping),
//   :343-345 (feature propagation), pointnet2_sem_seg.py:43-46 (head), Highway_bridge/models/pointnet2_utils.py:150-154,
//   353-356 and Highway_bridge/models/DGCNN.py:134-148 (EdgeConv)
// in the forward pass (A = activations, B = weight [out, in]) and in the data-gradient pass (A = gradient rows,
// B = weight transposed [in, out]).  Two epilogues fold the BatchNorm passes that used to follow into the GEMM:
//   EPI_STATS  forward: per-channel (count, mean, M2) of the bf16 result -- shifted sums per thread (shift = the first
//              value the thread sees, so no cancellation), merged with Chan's update per CTA, per group of 16 CTAs and
//              finally over the groups, always in a fixed order -- folded to mean / invstd / variance;
//   EPI_BNBWD  backward: the accumulator is d loss / d z of the PREVIOUS layer's BN+ReLU output; the epilogue reads that
//              layer's pre-activation tile y, applies the ReLU mask and emits dy together with the per-channel sums of
//              dy and dy * yhat that the BatchNorm backward needs (yhat = (y - mean) * invstd).
//
// Design (third version; profiles/r2_gemm_rows.md has the measurements that led here).  The layers are memory-bound
// (K, N of a few dozen to a few hundred) and consist of thousands of 128-row tiles, so the kernel keeps every per-tile
// cost that is not a byte of HBM traffic off the instruction stream:
//   * operand tiles arrive by TMA tensor copies (cp.async.bulk.tensor.2d, ONE instruction per 128 x BK slab) straight
//     into the swizzled K-major UMMA layout -- 128-byte swizzle for K > 32, 64-byte for K <= 32, 32-byte for K <= 16, so
//     that narrow layers do not pay shared memory for zero fill; K / M / N tails are zero-filled by the TMA unit;
//   * a CTA is 128 threads = the 128 rows of a tile = the 128 TMEM lanes; 2-6 CTAs are resident per SM (shared memory
//     and <= 512 TMEM columns permitting) and hide each other's load / MMA / epilogue latencies; inside a CTA warp 0
//     runs the copy ring (2-4 stages) and lane 0 issues the tcgen05.mma chain of a slab;
//   * epilogue, thread = row: tcgen05.ld -> bf16 -> swizzled shared-memory tile (conflict-free 16-byte stores) -> ONE TMA
//     tensor store per 64-column sub-tile (clips the M / N tails); the statistics pass re-reads the same tile with
//     thread = (16-byte column unit, row group): LDS.128, 8 columns per thread;
//   * EPI_BNBWD double-buffers the y tiles (TMA loads, same swizzled geometry as the output tile);
//   * CTAs loop over row tiles with stride gridDim.x, keeping their statistics in registers; the per-CTA partials are
//     folded in two deterministic levels by "last arriver" CTAs (group of 16, then all groups).
// Algorithmic bytes: 2 * M * (K + N) [+ 2 * M * N for the y tile of EPI_BNBWD]; the tensor pipe idles by construction.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "pcb_common.cuh"
#include "umma.cuh"

namespace pcb {

constexpr int kGemmThreads = 128;
constexpr int kFoldGroup = 32;             // CTAs per first-level fold group
constexpr int kCombFloats = 3 * 1024;      // row-group combine: rgroups * BN <= 1024 columns x (n, mean, M2)
enum { EPI_STORE = 0, EPI_STATS = 1, EPI_BNBWD = 2, EPI_BIAS = 3 };
__host__ __device__ constexpr bool epi_has_stats(int epi) { return epi == EPI_STATS || epi == EPI_BNBWD; }

struct GemmParams {
    int64_t M;
    int N;                       // output columns written (multiple of 8)
    int K;                       // contraction length (multiple of 8)
    int Cv;                      // real channels among the N columns (statistics epilogues)
    int BN, BK, swzA;            // column tile; operand slab = BK columns = swzA bytes per row (swizzle span 32 / 64 / 128)
    int wsub, nsub, wshift;      // output sub-tiles: nsub x [128 rows x wsub columns] (wsub = 16 / 32 / 64 = 1 << wshift)
    int mtiles, ntiles, stages, sshift, nslabs;   // stages = 1 << sshift (2 or 4)
    float *parts;                // [ntiles][gridDim.x + groups][3][BN] partial column statistics (CTAs, then groups)
    unsigned *tickets;           // [ntiles][1 + groups], zero on entry, zero on exit
    // EPI_STATS
    float eps;
    float *mean, *invstd, *var;  // [N] results (var: biased batch variance)
    // EPI_BNBWD
    const float *bn_mean, *bn_invstd, *gamma, *beta;   // [Cv]
    int relu;
    float *sums;                 // [3][N]: sum dy, sum dy*yhat, 0 (gradient of the folded conv bias)
    // deferred final fold (either statistics epilogue): when set, the kernel stops after the first fold level and leaves
    // the group partials [groups][3][N] here -- (n, mean, M2) or (0, sum dy*yhat, sum dy) -- for the elementwise kernel
    // that consumes the statistics anyway (pcb_bn_apply_rows / pcb_bn_bwd_apply_rows fold them at their start): the
    // second "last arriver" level (fence + ticket + fold, ~4 us of serial latency per launch) disappears
    float *gparts;
    // EPI_BIAS (inference: conv with the BatchNorm folded in): out = act(acc + bias), act = identity / ReLU / leaky ReLU;
    // pool_k > 1: only the max over every pool_k consecutive rows is written, to pooled [M / pool_k, ldp]
    const float *bias;           // [Cv] or NULL
    float slope;                 // act == 2: leaky slope
    int act, pool_k;
    void *pooled;
    int64_t ldp;
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

// ---- TMA tensor copies (2-D tiled tensor maps) and swizzled UMMA descriptors ----------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_dst),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, int c0, int c1, uint32_t smem_src)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                     reinterpret_cast<uint64_t>(tm)),
                 "r"(c0), "r"(c1), "r"(smem_src)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
// K-major operand tile whose rows are one swizzle span wide (span = 32 / 64 / 128 bytes): 8-row core groups are
// 8 * span bytes apart (SBO), the leading-dimension offset is unused, layout code 6 / 4 / 2 (sm_100 encoding).
__device__ __forceinline__ uint64_t umma_desc_swz(uint32_t smem_addr, uint32_t span)
{
    const uint32_t layout = span == 128 ? 2u : (span == 64 ? 4u : 6u);
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(((8 * span) >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                                // descriptor version 1 (sm_100)
    d |= (uint64_t)layout << 61;
    return d;
}
// byte offset inside a swizzled tile (base aligned to 1024): 16-byte unit index XOR (row-of-128-bytes index & mask)
__device__ __forceinline__ uint32_t swz(uint32_t off, uint32_t mask) { return off ^ (((off >> 7) & mask) << 4); }

__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4 &v)
{
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// mbarrier wait with a watchdog: a protocol error traps (launch failure) instead of hanging the device
__device__ __forceinline__ void mbar_wait_wd(uint64_t *bar, uint32_t parity)
{
    const uint32_t a = smem_u32(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
        if (ok) return;
        if (spin > (1u << 26)) __trap();
    }
}

struct GemmSmem {
    int a_bytes, b_bytes, nb, off_b, off_out, sub_bytes, out_bytes, off_y, off_const, total;
};

__host__ __device__ __forceinline__ GemmSmem gemm_smem_layout(int BN, int swzA, int nslabs, int stages, int epi, int wsub)
{
    GemmSmem s;
    s.a_bytes = 128 * swzA;                                                  // 4 / 8 / 16 KB
    s.b_bytes = (BN * swzA + 1023) & ~1023;
    s.nb = nslabs == 1 ? 1 : stages;                                         // resident weights when K fits one slab
    s.off_b = stages * s.a_bytes;
    int opnd = s.off_b + s.nb * s.b_bytes;
    if (epi_has_stats(epi) && opnd < kCombFloats * 4) opnd = kCombFloats * 4;  // the combine scratch aliases the operand ring
    s.sub_bytes = 128 * wsub * 2;
    s.out_bytes = ((BN + wsub - 1) / wsub) * s.sub_bytes;
    s.off_out = (opnd + 1023) & ~1023;
    s.off_y = s.off_out + s.out_bytes;
    s.off_const = s.off_y + (epi == EPI_BNBWD ? 2 * s.out_bytes : 0);        // y tiles: double buffered
    s.total = s.off_const + (epi == EPI_BNBWD ? 4 * BN * 4 : (epi == EPI_BIAS ? BN * 4 : 0)) + 1024;   // + alignment slack
    return s;
}

// packed fp32 pairs (FADD2 / FFMA2): two columns per instruction in the statistics passes
__device__ __forceinline__ uint64_t pk2(float lo, float hi)
{
    return (uint64_t)__float_as_uint(lo) | ((uint64_t)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ float pk_lo(uint64_t v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float pk_hi(uint64_t v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// the two bf16 of a 32-bit word as a packed fp32 pair (low half = even column)
__device__ __forceinline__ uint64_t bf2_to_f2(unsigned w) { return (uint64_t)(w << 16) | ((uint64_t)(w & 0xffff0000u) << 32); }

// `cnt` partial triples (n, mean | sum, M2 | sum) of column c, `stride` floats apart, merged in index order; the loads
// of sixteen partials are in flight together (the fold is a chain of L2 round trips otherwise)
template <int EPI>
__device__ __forceinline__ void fold_range(const float *base, int cnt, size_t stride, int BN, float &n, float &m, float &q)
{
    constexpr int FB = 16;                                                   // partials in flight per thread
    for (int j0 = 0; j0 < cnt; j0 += FB) {
        float en[FB], em[FB], eq[FB];
#pragma unroll
        for (int u = 0; u < FB; ++u) {
            const bool ok = j0 + u < cnt;
            const float *e = base + (size_t)(ok ? j0 + u : 0) * stride;
            en[u] = (EPI == EPI_STATS && ok) ? __ldcg(e) : 0.f;
            em[u] = ok ? __ldcg(e + BN) : 0.f;
            eq[u] = ok ? __ldcg(e + 2 * BN) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < FB; ++u) {
            if (EPI == EPI_STATS) chan_merge(n, m, q, en[u], em[u], eq[u]);
            else m += em[u], q += eq[u];
        }
    }
}

// all threads: fold `cnt` partials per column; 128 / BN threads share a column (contiguous slices of the partials,
// merged in slice order through `scratch`); the totals of column tid land in (n, m, q) of the threads tid < BN
template <int EPI>
__device__ __forceinline__ void fold_cols(const float *base, int cnt, int BN, float *scratch, float &n, float &m, float &q)
{
    const int tid = threadIdx.x;
    int slices = kGemmThreads / BN;
    if (slices < 1) slices = 1;
    const int per = (cnt + slices - 1) / slices;
    const int s = tid / BN, c = tid - s * BN;
    n = m = q = 0.f;
    if (s < slices) {
        const int lo = s * per, hi = lo + per < cnt ? lo + per : cnt;
        if (hi > lo) fold_range<EPI>(base + (size_t)lo * 3 * BN + c, hi - lo, (size_t)3 * BN, BN, n, m, q);
        float *e = scratch + ((size_t)s * BN + c) * 3;
        e[0] = n, e[1] = m, e[2] = q;
    }
    __syncthreads();
    if (tid < BN) {
        n = m = q = 0.f;
        for (int g = 0; g < slices; ++g) {
            const float *e = scratch + ((size_t)g * BN + tid) * 3;
            if (EPI == EPI_STATS) chan_merge(n, m, q, e[0], e[1], e[2]);
            else m += e[1], q += e[2];
        }
    }
    __syncthreads();
}

// minimum CTAs per SM: the statistics epilogues are issue-latency bound with 4 warps per scheduler (ncu: 10.2 M warp
// instructions per launch against 4.5 M of the plain GEMM, issue slots 31 % busy, occupancy limited by 122 registers):
// forcing 5 CTAs per SM (96 registers, 4-byte spill) measured SLOWER (statistics 392 -> 396 us, data gradient 460 -> 496 us over the 20 layer shapes): kept at 4
template <int EPI>
__global__ void __launch_bounds__(kGemmThreads, 4)
gemm_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmY, const GemmParams p)
{
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_full[4];                              // operand slab of a stage has landed
    __shared__ __align__(8) uint64_t s_mma[4];                               // MMAs reading a stage have completed
    __shared__ __align__(8) uint64_t s_yfull[2];                             // y tile buffer has landed (EPI_BNBWD)
    __shared__ uint32_t s_tmem;
    __shared__ int s_flag;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int BN = p.BN, BK = p.BK, S = p.stages, sshift = p.sshift, nslabs = p.nslabs;
    const GemmSmem L = gemm_smem_layout(BN, p.swzA, nslabs, S, EPI, p.wsub);
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char *smem = smem_raw + (sbase - smem_u32(smem_raw));
    const uint32_t sA = sbase, sB = sbase + L.off_b, sOut = sbase + L.off_out, sY = sbase + L.off_y;
    float *s_const = reinterpret_cast<float *>(smem + L.off_const);         // EPI_BNBWD: [nm | is | sc | sh] x BN
    float *s_comb = reinterpret_cast<float *>(smem);                        // after the tile loop
    const int nt = blockIdx.y, n0 = nt * BN;
    const bool b_resident = nslabs == 1;

    int tmem_cols = 32;
    while (tmem_cols < BN) tmem_cols <<= 1;

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(&s_full[i], 1), mbar_init(&s_mma[i], 1);
        mbar_init(&s_yfull[0], 1), mbar_init(&s_yfull[1], 1);
        fence_mbar_init();
        tma_prefetch_desc(&tmA), tma_prefetch_desc(&tmB), tma_prefetch_desc(&tmC);
        if (EPI == EPI_BNBWD) tma_prefetch_desc(&tmY);
    }
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), (uint32_t)tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // everything above is CTA-local (barriers, TMEM columns, descriptor prefetch): it may overlap the predecessor's
    // tail.  From here on the kernel reads global memory; its own successor may start its prologue now.
    pdl_wait();
    pdl_trigger();
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
    if (EPI == EPI_BIAS) {
        for (int c = tid; c < BN; c += kGemmThreads) s_const[c] = (p.bias && n0 + c < p.Cv) ? p.bias[n0 + c] : 0.f;
    }
    if (EPI == EPI_BNBWD || EPI == EPI_BIAS) __syncthreads();
    const uint32_t tmem_base = s_tmem;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t idesc = umma_idesc_bf16_m128(BN);

    const int my_tiles = ((int)blockIdx.x < p.mtiles) ? (p.mtiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int nsteps = my_tiles * nslabs;
    const uint32_t a_tx = (uint32_t)L.a_bytes, b_tx = (uint32_t)(BN * p.swzA);

    // operand copies of a step (row tile mt, K slab ks) into `stage`: one thread
    auto issue_loads = [&](int mt, int ks, int stage, bool with_b) {
        mbar_expect_tx(&s_full[stage], a_tx + (with_b ? b_tx : 0u));
        tma_load_2d(sA + (uint32_t)(stage * L.a_bytes), &tmA, ks * BK, mt * 128, &s_full[stage]);
        if (with_b) tma_load_2d(sB + (uint32_t)((b_resident ? 0 : stage) * L.b_bytes), &tmB, ks * BK, n0, &s_full[stage]);
    };
    // y tile of row tile `mt` into buffer `buf` (EPI_BNBWD): nsub boxes of [128 x wsub]
    auto issue_y = [&](int mt, int buf) {
        int boxes = 0;
        for (int j = 0; j < p.nsub; ++j) boxes += (n0 + j * p.wsub < p.N) ? 1 : 0;
        mbar_expect_tx(&s_yfull[buf], (uint32_t)(boxes * L.sub_bytes));
        for (int j = 0; j < p.nsub; ++j)
            if (n0 + j * p.wsub < p.N)
                tma_load_2d(sY + (uint32_t)(buf * L.out_bytes + j * L.sub_bytes), &tmY, n0 + j * p.wsub, mt * 128, &s_yfull[buf]);
    };

    // ---- epilogue bookkeeping: output tile geometry, statistics roles
    const int wsub = p.wsub, wshift = p.wshift, pitchO = wsub * 2;
    const uint32_t omask = (uint32_t)(pitchO >> 4) - 1u;                     // swizzle mask of the output / y tiles
    const int U = BN >> 3;                                                   // 16-byte column units of the tile
    const int rgroups = kGemmThreads / U;                                    // >= 4 (BN <= 256)
    const bool st_active = epi_has_stats(EPI) && tid < U * rgroups;
    const int st_rg = tid / U, st_u = tid - st_rg * U;
    const int st_j = (st_u << 3) >> wshift, st_uu = st_u - ((st_j << wshift) >> 3);
    const uint32_t st_base = (uint32_t)(st_j * L.sub_bytes);
    const uint32_t st_col = (uint32_t)(st_uu * 16);
    // EPI_STATS: count, shift, sum (x - shift), sum (x - shift)^2;  EPI_BNBWD: a = sum dy, b = sum dy * yhat
    float rn = 0.f;
    uint64_t sh2[4], ra2[4], rb2[4], is2[4], nm2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) sh2[i] = ra2[i] = rb2[i] = is2[i] = nm2[i] = 0ull;
    if (EPI == EPI_BNBWD && st_active) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            nm2[i] = pk2(s_const[8 * st_u + 2 * i], s_const[8 * st_u + 2 * i + 1]);
            is2[i] = pk2(s_const[BN + 8 * st_u + 2 * i], s_const[BN + 8 * st_u + 2 * i + 1]);
        }
    }

    // ---- pipeline over the steps (tile, K slab) of this CTA: the copies of the next S - 1 steps are in flight while the
    //      MMAs of a step run; stage = step & (S - 1)
    int ld_t = 0, ld_ks = 0, ld_step = 0;                                    // next step to load (warp 0)
    auto load_next = [&]() {                                                 // warp 0, all lanes keep the counters
        if (lane == 0) issue_loads(blockIdx.x + ld_t * gridDim.x, ld_ks, ld_step & (S - 1), !b_resident || ld_step == 0);
        ++ld_step;
        if (++ld_ks == nslabs) ld_ks = 0, ++ld_t;
    };
    if (warp == 0) {
        for (int st = 0; st < S - 1 && st < nsteps; ++st) load_next();
        if (EPI == EPI_BNBWD && lane == 0 && my_tiles > 0) issue_y(blockIdx.x, 0);
    }
    int step = 0;
    for (int ti = 0; ti < my_tiles; ++ti) {
        const int mt = blockIdx.x + ti * gridDim.x;
        const int64_t m0 = (int64_t)mt * 128;
        const int rows = p.M - m0 < 128 ? (int)(p.M - m0) : 128;
        if (warp == 0) {
            for (int ks = 0; ks < nslabs; ++ks, ++step) {
                const int stage = step & (S - 1);
                if (ld_step < nsteps) {                                      // refill the stage that step - 1 used
                    if (step >= 1) mbar_wait_wd(&s_mma[(step - 1) & (S - 1)], ((uint32_t)((step - 1) >> sshift)) & 1u);
                    load_next();
                }
                mbar_wait_wd(&s_full[stage], ((uint32_t)(step >> sshift)) & 1u);
                tc_fence_after();
                if (lane == 0) {
                    const int kw = p.K - ks * BK < BK ? p.K - ks * BK : BK;
                    const int ksteps = (kw + 15) >> 4;
                    const uint64_t da = umma_desc_swz(sA + (uint32_t)(stage * L.a_bytes), (uint32_t)p.swzA);
                    const uint64_t db = umma_desc_swz(sB + (uint32_t)((b_resident ? 0 : stage) * L.b_bytes), (uint32_t)p.swzA);
                    for (int kk = 0; kk < ksteps; ++kk)                      // + 32 bytes along K = + 2 in the address field
                        umma_bf16(tmem_base, da + (uint64_t)(2 * kk), db + (uint64_t)(2 * kk), idesc, (ks | kk) ? 1u : 0u);
                    umma_commit(&s_mma[stage]);
                }
                __syncwarp();
            }
        } else {
            step += nslabs;
        }
        // ---- epilogue of the tile: its last slab's MMAs have completed; the previous tile's store has read the out tile
        mbar_wait_wd(&s_mma[(step - 1) & (S - 1)], ((uint32_t)((step - 1) >> sshift)) & 1u);
        tc_fence_after();
        if (tid == 0) bulk_wait_read0();
        __syncthreads();                                                     // (A) out tile + the other y buffer are free
        const uint32_t ybuf = sY + (uint32_t)((ti & 1) * L.out_bytes);
        if (EPI == EPI_BNBWD) {
            if (tid == 0 && ti + 1 < my_tiles) issue_y(mt + gridDim.x, (ti + 1) & 1);
            mbar_wait_wd(&s_yfull[ti & 1], ((uint32_t)(ti >> 1)) & 1u);
        }
        for (int c0 = 0; c0 < BN; c0 += 16) {
            float v[16];
            tmem_ld16(trow + (uint32_t)c0, v);
            const int j = c0 >> wshift, cs = c0 - (j << wshift);
            const uint32_t off0 = (uint32_t)(j * L.sub_bytes) + swz((uint32_t)(tid * pitchO + cs * 2), omask);
            const uint32_t off1 = (uint32_t)(j * L.sub_bytes) + swz((uint32_t)(tid * pitchO + cs * 2 + 16), omask);
            if (EPI == EPI_BNBWD && p.relu) {
                float y[8];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    unpack8(lds128(ybuf + (h ? off1 : off0)), y);
                    const float4 *sc4 = reinterpret_cast<const float4 *>(s_const + 2 * BN + c0 + 8 * h);
                    const float4 *sh4 = reinterpret_cast<const float4 *>(s_const + 3 * BN + c0 + 8 * h);
                    const float4 s0 = sc4[0], s1 = sc4[1], t0 = sh4[0], t1 = sh4[1];
                    const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                    const float sf[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float z = fmaf(y[q], sc[q], sf[q]);
                        v[8 * h + q] = z > 0.f ? v[8 * h + q] : 0.f;
                    }
                }
            }
            if (EPI == EPI_BIAS) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    float t = v[q] + s_const[c0 + q];
                    if (p.act == 1) t = fmaxf(t, 0.f);
                    else if (p.act == 2) t = t > 0.f ? t : t * p.slope;
                    v[q] = t;
                }
            }
            sts128(sOut + off0, pack8(v));
            sts128(sOut + off1, pack8(v + 8));
        }
        tc_fence_before();                                                   // TMEM reads done before the next tile's MMAs
        proxy_fence();                                                       // generic-proxy writes -> TMA store reads
        __syncthreads();                                                     // (B) tile complete
        if (EPI == EPI_BIAS && p.pool_k > 1) {
            // max over every pool_k consecutive rows of the tile (pool_k divides 128 and M): thread = (16-byte column
            // unit, pooling group), packed bf16 max, one 16-byte store per (group, unit) straight to the pooled rows
            const int G = 128 / p.pool_k;
            for (int it = tid; it < U * G; it += kGemmThreads) {
                const int gq = it / U, u = it - gq * U;
                const int r0 = gq * p.pool_k;
                if (r0 >= rows || n0 + 8 * u >= p.N) continue;
                const int j = (u << 3) >> wshift, uu = u - ((j << wshift) >> 3);
                const uint32_t base = sOut + (uint32_t)(j * L.sub_bytes), col = (uint32_t)(uu * 16);
                uint4 best = lds128(base + swz((uint32_t)(r0 * pitchO) + col, omask));
                for (int r = 1; r < p.pool_k; ++r) {
                    const uint4 w = lds128(base + swz((uint32_t)((r0 + r) * pitchO) + col, omask));
                    __nv_bfloat162 *b2 = reinterpret_cast<__nv_bfloat162 *>(&best);
                    const __nv_bfloat162 *w2 = reinterpret_cast<const __nv_bfloat162 *>(&w);
#pragma unroll
                    for (int i = 0; i < 4; ++i) b2[i] = __hmax2(b2[i], w2[i]);
                }
                __nv_bfloat16 *o = (__nv_bfloat16 *)p.pooled + ((m0 + r0) / p.pool_k) * p.ldp + n0 + 8 * u;
                *reinterpret_cast<uint4 *>(o) = best;
            }
        } else if (tid == 0) {
            for (int j = 0; j < p.nsub; ++j)
                if (n0 + j * wsub < p.N) tma_store_2d(&tmC, n0 + j * wsub, mt * 128, sOut + (uint32_t)(j * L.sub_bytes));
            bulk_commit();
        }
        if (epi_has_stats(EPI) && st_active) {
            if (EPI == EPI_STATS) {
                if (ti == 0 && st_rg < rows) {
                    const uint4 w = lds128(sOut + st_base + swz((uint32_t)(st_rg * pitchO) + st_col, omask));
                    sh2[0] = bf2_to_f2(w.x), sh2[1] = bf2_to_f2(w.y), sh2[2] = bf2_to_f2(w.z), sh2[3] = bf2_to_f2(w.w);
                }
                for (int r = st_rg; r < rows; r += 4 * rgroups) {
                    uint4 w[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int rr = r + q * rgroups;
                        if (rr < rows) w[q] = lds128(sOut + st_base + swz((uint32_t)(rr * pitchO) + st_col, omask));
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (r + q * rgroups < rows) {
                            const unsigned ww[4] = {w[q].x, w[q].y, w[q].z, w[q].w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint64_t d = sub2(bf2_to_f2(ww[i]), sh2[i]);
                                ra2[i] = add2(ra2[i], d);
                                rb2[i] = fma2(d, d, rb2[i]);
                            }
                            rn += 1.f;
                        }
                    }
                }
            } else {
                for (int r = st_rg; r < rows; r += 2 * rgroups) {
                    uint4 w[2], yw[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int rr = r + q * rgroups;
                        const uint32_t o = st_base + swz((uint32_t)(rr * pitchO) + st_col, omask);
                        w[q] = rr < rows ? lds128(sOut + o) : make_uint4(0, 0, 0, 0);
                        yw[q] = rr < rows ? lds128(ybuf + o) : make_uint4(0, 0, 0, 0);
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {                            // rows beyond `rows`: dy = 0
                        const unsigned dw[4] = {w[q].x, w[q].y, w[q].z, w[q].w}, yy[4] = {yw[q].x, yw[q].y, yw[q].z, yw[q].w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint64_t d = bf2_to_f2(dw[i]);
                            ra2[i] = add2(ra2[i], d);
                            rb2[i] = fma2(d, fma2(bf2_to_f2(yy[i]), is2[i], nm2[i]), rb2[i]);
                        }
                    }
                }
            }
        }
    }
    if (tid == 0) bulk_wait_all0();                                          // the last store has left shared memory

    if (epi_has_stats(EPI)) {
        // ---- per-CTA partial statistics: the row groups merged in a fixed order through shared memory (aliases the
        //      operand ring: every copy has been consumed and every MMA has completed)
        __syncthreads();
        if (st_active) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float *e = s_comb + ((size_t)st_rg * BN + 8 * st_u + i) * 3;
                const float a = (i & 1) ? pk_hi(ra2[i >> 1]) : pk_lo(ra2[i >> 1]);
                const float b = (i & 1) ? pk_hi(rb2[i >> 1]) : pk_lo(rb2[i >> 1]);
                if (EPI == EPI_STATS) {
                    const float s0 = (i & 1) ? pk_hi(sh2[i >> 1]) : pk_lo(sh2[i >> 1]);
                    const float inv = rn > 0.f ? 1.f / rn : 0.f;
                    const float md = a * inv;
                    e[0] = rn, e[1] = s0 + md, e[2] = fmaxf(b - a * md, 0.f);
                } else {
                    e[0] = 0.f, e[1] = b, e[2] = a;
                }
            }
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
        // ---- two-level deterministic fold: the last CTA of each group of kFoldGroup merges the group (fixed order),
        //      the last group to finish merges the groups and writes the results
        __threadfence();
        __syncthreads();
        const int grp = blockIdx.x / kFoldGroup;
        const int g_lo = grp * kFoldGroup, g_hi = g_lo + kFoldGroup < P ? g_lo + kFoldGroup : P;
        if (tid == 0) s_flag = atomicAdd(tick + 1 + grp, 1u) == (unsigned)(g_hi - g_lo - 1);
        __syncthreads();
        if (s_flag) {
            __threadfence();
            float n, m, q;
            fold_cols<EPI>(lvl0 + (size_t)g_lo * 3 * BN, g_hi - g_lo, BN, s_comb, n, m, q);
            bool final_level = G == 1;
            if (p.gparts) {                                                  // deferred: the consumer merges the groups
                if (tid < BN && n0 + tid < p.N) {
                    float *o = p.gparts + (size_t)grp * 3 * p.N + n0 + tid;
                    const bool real = n0 + tid < p.Cv;
                    o[0] = real ? n : 0.f, o[p.N] = real ? m : 0.f, o[2 * (size_t)p.N] = real ? q : 0.f;
                }
                if (tid == 0) tick[1 + grp] = 0u;
                final_level = false;
            } else if (!final_level) {
                if (tid < BN) {
                    float *o = lvl1 + (size_t)grp * 3 * BN;
                    o[tid] = n, o[BN + tid] = m, o[2 * BN + tid] = q;
                }
                __threadfence();
                __syncthreads();
                if (tid == 0) {
                    tick[1 + grp] = 0u;                                      // ready for the next launch
                    s_flag = atomicAdd(tick, 1u) == (unsigned)(G - 1);
                }
                __syncthreads();
                if (s_flag) {
                    __threadfence();
                    fold_cols<EPI>(lvl1, G, BN, s_comb, n, m, q);
                    final_level = true;
                    if (tid == 0) tick[0] = 0u;
                }
            } else if (tid == 0) {
                tick[1 + grp] = 0u;
            }
            if (final_level && tid < BN) {
                const int gc = n0 + tid;
                if (gc < p.N) {
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
        }
    }

    // ---- teardown: every MMA has completed (each tile's epilogue waited for its last commit)
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

struct GemmPlan {
    int BN, BK, swzA, wsub, nsub, nslabs, ntiles, mtiles, grid_x, ctas_per_sm, groups, stages;
    size_t smem;
};

// CTAs per SM the registers of each epilogue variant allow (cudaFuncGetAttributes once per device; the occupancy API
// answered 1 in most launches -- measured with ncu, round 2 -- which serialised the whole design on one CTA per SM)
static int g_reg_limit[4][kMaxDevices];

static int gemm_force_span()                                                 // PCB_GEMM_SPAN=128: one swizzle mode for every K
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("PCB_GEMM_SPAN");
        v = e ? atoi(e) : 0;
    }
    return v;
}

static bool gemm_plan(int64_t M, int N, int K, int epi, GemmPlan &g)
{
    g.mtiles = (int)ceil_div(M, 128);
    g.swzA = K > 32 ? 128 : (K > 16 ? 64 : 32);
    if (gemm_force_span() == 128 || gemm_force_span() == 64) g.swzA = g.swzA > gemm_force_span() ? g.swzA : gemm_force_span();
    g.BK = g.swzA / 2;
    g.nslabs = (K + g.BK - 1) / g.BK;
    // column tiles of <= 128 columns (several CTAs per SM must fit: shared memory, TMEM), more of them when the row
    // tiles alone would leave SMs idle; several column tiles are multiples of 64 columns (whole output sub-tiles)
    int ntiles = (N + 127) / 128;
    const int want = 2 * PCB_NUM_SMS / (g.mtiles > 0 ? g.mtiles : 1);
    const int max_split = (N + 63) / 64;
    if (want > ntiles) ntiles = want < max_split ? want : max_split;
    if (ntiles <= 1) {
        g.BN = (N + 15) & ~15;
        g.ntiles = 1;
    } else {
        int bn = (int)ceil_div(N, ntiles);
        bn = (bn + 63) & ~63;
        g.BN = bn;
        g.ntiles = (int)ceil_div(N, bn);
    }
    g.wsub = g.BN >= 64 ? 64 : (g.BN >= 32 ? 32 : 16);
    g.nsub = (g.BN + g.wsub - 1) / g.wsub;
    int tmem_cols = 32;
    while (tmem_cols < g.BN) tmem_cols <<= 1;
    // operand stages: two when many CTAs per SM hide the latency for each other; deeper (up to 4, never more than the
    // K slabs + 1) when the tiles are few, so that a CTA has several slabs in flight on its own
    int dev = 0;
    cudaGetDevice(&dev);
    const int reg_limit = (dev >= 0 && dev < kMaxDevices && g_reg_limit[epi][dev] > 0) ? g_reg_limit[epi][dev] : 4;
    auto occupancy = [&](int stages) {
        const size_t smem = (size_t)gemm_smem_layout(g.BN, g.swzA, g.nslabs, stages, epi, g.wsub).total;
        int c = (int)((227 * 1024) / (smem + 1024 + 128));                      // + reserved KB + static shared memory
        if (c > 512 / tmem_cols) c = 512 / tmem_cols;
        return c > reg_limit ? reg_limit : c;
    };
    int stages = 2;
    if (occupancy(2) < 1) return false;
    const int64_t tiles = (int64_t)g.mtiles * g.ntiles;
    if (g.nslabs >= 3 && occupancy(4) >= 1 && (int64_t)PCB_NUM_SMS * occupancy(4) >= tiles)
        stages = 4;                                                          // deeper only while every tile still gets its own CTA
    g.stages = stages;
    g.smem = (size_t)gemm_smem_layout(g.BN, g.swzA, g.nslabs, stages, epi, g.wsub).total;
    g.ctas_per_sm = occupancy(stages);
    const int slots = PCB_NUM_SMS * g.ctas_per_sm / g.ntiles;
    g.grid_x = g.mtiles < slots ? g.mtiles : slots;
    if (g.grid_x < 1) g.grid_x = 1;
    g.groups = (g.grid_x + kFoldGroup - 1) / kFoldGroup;
    return true;
}

// ---- tensor maps: cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}

// rows x cols bf16 matrix with leading dimension ld (elements); box = box_rows x box_cols, box_cols * 2 == span bytes
static int make_map(CUtensorMap *tm, const void *base, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return PCB_EINVAL;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const int span = box_cols * 2;
    const CUtensorMapSwizzle sw =
        span == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (span == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : PCB_EINVAL;
}

struct GemmOperands {
    const void *A, *B, *Y;
    void *C;
    int64_t lda, ldb, ldc, ldy;
    int Nb;                      // rows of B that exist (others are zero)
};

template <int EPI>
static int gemm_launch(GemmParams &p, const GemmOperands &o, cudaStream_t st, int *groups_out = nullptr)
{
    static bool attr_set[kMaxDevices] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < kMaxDevices && !attr_set[dev]) {
        if (cudaError_t e = smem_optin_once(gemm_rows_kernel<EPI>, 226 * 1024, attr_set)) return (int)e;
        // several CTAs per SM each want tens of KB: ask for the largest shared-memory carve-out
        cudaFuncSetAttribute(gemm_rows_kernel<EPI>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, gemm_rows_kernel<EPI>) == cudaSuccess && fa.numRegs > 0) {
            const int regs = (fa.numRegs + 7) & ~7;                          // allocation granularity: 8 registers per thread
            int c = 65536 / (regs * kGemmThreads);
            g_reg_limit[EPI][dev] = c < 1 ? 1 : (c > 8 ? 8 : c);
        }
    }
    GemmPlan g;
    if (!gemm_plan(p.M, p.N, p.K, EPI, g)) return PCB_ERANGE;
    p.BN = g.BN, p.BK = g.BK, p.swzA = g.swzA, p.wsub = g.wsub, p.nsub = g.nsub, p.nslabs = g.nslabs;
    p.mtiles = g.mtiles, p.ntiles = g.ntiles, p.stages = g.stages;
    p.sshift = g.stages == 4 ? 2 : 1;
    p.wshift = g.wsub == 64 ? 6 : (g.wsub == 32 ? 5 : 4);
    if (groups_out) *groups_out = g.groups;
    CUtensorMap tmA, tmB, tmC, tmY;
    if (int rc = make_map(&tmA, o.A, p.M, p.K, o.lda, 128, g.BK)) return rc;
    if (int rc = make_map(&tmB, o.B, o.Nb, p.K, o.ldb, g.BN, g.BK)) return rc;
    if (EPI == EPI_BIAS && p.pool_k > 1) {
        tmC = tmA;                                                           // never stored through: the epilogue pools
    } else if (int rc = make_map(&tmC, o.C, p.M, p.N, o.ldc, 128, g.wsub)) {
        return rc;
    }
    if (EPI == EPI_BNBWD) {
        if (int rc = make_map(&tmY, o.Y, p.M, p.N, o.ldy, 128, g.wsub)) return rc;
    } else {
        tmY = tmC;
    }
    return (int)launch_pdl(gemm_rows_kernel<EPI>, dim3((unsigned)g.grid_x, (unsigned)g.ntiles), dim3(kGemmThreads), g.smem, st,
                           tmA, tmB, tmC, tmY, p);
}

static inline bool al16(const void *q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

static int gemm_check(const GemmParams &p, const GemmOperands &o)
{
    PCB_REQUIRE(o.A && o.B && o.C, PCB_EINVAL);
    PCB_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0 && o.Nb > 0, PCB_EINVAL);
    PCB_REQUIRE(p.N % 8 == 0 && p.K % 8 == 0 && o.lda % 8 == 0 && o.ldb % 8 == 0 && o.ldc % 8 == 0, PCB_ERANGE);
    PCB_REQUIRE(o.lda >= p.K && o.ldb >= p.K && o.ldc >= p.N && p.N <= 4096 && p.K <= 8192, PCB_ERANGE);
    PCB_REQUIRE(p.M < ((int64_t)1 << 31) - 128, PCB_ERANGE);                 // 32-bit TMA row coordinates
    PCB_REQUIRE(al16(o.A) && al16(o.B) && al16(o.C), PCB_EALIGN);
    return 0;
}

}  // namespace pcb

using namespace pcb;

// scratch floats of one statistics GEMM (partial column statistics of every CTA and fold group), either epilogue
PCB_API int64_t pcb_gemm_work_floats(int64_t M, int N, int K)
{
    GemmPlan g;
    if (!gemm_plan(M, N, K, EPI_STATS, g)) return -1;                        // BN / ntiles do not depend on the epilogue
    // upper bound over every occupancy the launch may pick (<= 8 CTAs per SM): the scratch never depends on register counts
    const int64_t pmax = g.mtiles < PCB_NUM_SMS * 8 ? g.mtiles : PCB_NUM_SMS * 8;
    return (int64_t)g.ntiles * (pmax + (pmax + kFoldGroup - 1) / kFoldGroup) * 3 * g.BN;
}

// ticket words a statistics GEMM may use: ntiles * (1 + groups) <= this
PCB_API int pcb_gemm_tickets(void) { return 1024; }

// upper bound of the fold groups of one statistics GEMM (<= 8 CTAs per SM): rows of a deferred `gparts` buffer
PCB_API int pcb_gemm_max_groups(void) { return (PCB_NUM_SMS * 8 + kFoldGroup - 1) / kFoldGroup; }

// y[M, N] = x[M, K] . w[Nw, K]^T (rows >= Nw of the result are zero columns), bf16
PCB_API int pcb_linear_rows_bf16(const void *x, int64_t ldx, const void *w, int64_t ldw, int64_t M, int N, int Nw, int K,
                                 void *y, int64_t ldy, pcb_stream_t stream)
{
    GemmParams p = {};
    GemmOperands o = {};
    o.A = x, o.B = w, o.C = y, o.lda = ldx, o.ldb = ldw, o.ldc = ldy, o.Nb = Nw;
    p.M = M, p.N = N, p.K = K;
    const int rc = gemm_check(p, o);
    if (rc) return rc;
    return gemm_launch<EPI_STORE>(p, o, (cudaStream_t)stream);
}

// same + training-mode BatchNorm statistics of y: mean / invstd / biased variance of the bias-free output (pcb_bn_apply_rows
// turns them into the running statistics).  work: pcb_gemm_work_floats floats; tickets: zeroed words.
PCB_API int pcb_linear_bn_stats_rows_bf16(const void *x, int64_t ldx, const void *w, int64_t ldw, int64_t M, int N, int Nw,
                                          int K, void *y, int64_t ldy, int Cv, float eps, float *mean, float *invstd,
                                          float *var, float *work, unsigned *tickets, float *gparts, int *groups_out,
                                          pcb_stream_t stream)
{
    GemmParams p = {};
    GemmOperands o = {};
    o.A = x, o.B = w, o.C = y, o.lda = ldx, o.ldb = ldw, o.ldc = ldy, o.Nb = Nw;
    p.M = M, p.N = N, p.K = K;
    const int rc = gemm_check(p, o);
    if (rc) return rc;
    PCB_REQUIRE(work && tickets && Cv > 0 && Cv <= N, PCB_EINVAL);
    PCB_REQUIRE(gparts ? groups_out != nullptr : (mean && invstd && var), PCB_EINVAL);
    p.Cv = Cv, p.eps = eps, p.mean = mean, p.invstd = invstd, p.var = var;
    p.parts = work, p.tickets = tickets, p.gparts = gparts;
    return gemm_launch<EPI_STATS>(p, o, (cudaStream_t)stream, groups_out);
}

// data gradient of a layer whose INPUT was z = relu(BN(y)) of the previous layer:
//   gz = gy[M, K] . wt[Nw, K]^T (wt = weight transposed: [in, out]);  dy = gz * [z > 0];  sums = (sum dy, sum dy*yhat, 0)
// dy is written to `dy` ([M, lddy], N columns), the BatchNorm backward finishes with pcb_bn_bwd_apply_rows.
PCB_API int pcb_dgrad_bn_rows_bf16(const void *gy, int64_t ldg, const void *wt, int64_t ldwt, int64_t M, int N, int Nw, int K,
                                   const void *yprev, int64_t ldyp, const float *mean, const float *invstd,
                                   const float *gamma, const float *beta, int Cv, int relu, void *dy, int64_t lddy,
                                   float *sums, float *work, unsigned *tickets, float *gparts, int *groups_out,
                                   pcb_stream_t stream)
{
    GemmParams p = {};
    GemmOperands o = {};
    o.A = gy, o.B = wt, o.C = dy, o.Y = yprev, o.lda = ldg, o.ldb = ldwt, o.ldc = lddy, o.ldy = ldyp, o.Nb = Nw;
    p.M = M, p.N = N, p.K = K;
    const int rc = gemm_check(p, o);
    if (rc) return rc;
    PCB_REQUIRE(yprev && mean && invstd && gamma && beta && work && tickets, PCB_EINVAL);
    PCB_REQUIRE(gparts ? groups_out != nullptr : sums != nullptr, PCB_EINVAL);
    PCB_REQUIRE(Cv > 0 && Cv <= N && ldyp >= N && ldyp % 8 == 0 && al16(yprev), PCB_ERANGE);
    p.Cv = Cv, p.bn_mean = mean, p.bn_invstd = invstd, p.gamma = gamma;
    p.beta = beta, p.relu = relu, p.sums = sums, p.parts = work, p.tickets = tickets, p.gparts = gparts;
    return gemm_launch<EPI_BNBWD>(p, o, (cudaStream_t)stream, groups_out);
}

// Inference layer (1x1 conv with the BatchNorm folded into w / bias): y = act(x . w^T + bias), act 0 identity /
// 1 ReLU / 2 leaky ReLU(slope), bf16 in / out, bias fp32 [Cv] or NULL.  pool_k > 1 (must divide 128 and M): only
// the max over every pool_k consecutive rows is written, y = [M / pool_k, ldy] -- the max over the neighbours of a
// set-abstraction / EdgeConv block without the [M, N] tensor ever reaching memory.
PCB_API int pcb_linear_bias_act_rows_bf16(const void *x, int64_t ldx, const void *w, int64_t ldw, int64_t M, int N, int Nw,
                                          int K, const float *bias, int Cv, int act, float slope, int pool_k, void *y,
                                          int64_t ldy, pcb_stream_t stream)
{
    GemmParams p = {};
    GemmOperands o = {};
    o.A = x, o.B = w, o.C = y, o.lda = ldx, o.ldb = ldw, o.ldc = ldy, o.Nb = Nw;
    p.M = M, p.N = N, p.K = K;
    const int rc = gemm_check(p, o);
    if (rc) return rc;
    PCB_REQUIRE(act >= 0 && act <= 2 && Cv >= 0 && Cv <= N, PCB_EINVAL);
    PCB_REQUIRE(pool_k >= 1 && (pool_k == 1 || (pool_k <= 128 && 128 % pool_k == 0 && M % pool_k == 0)), PCB_ERANGE);
    p.bias = bias, p.Cv = Cv, p.act = act, p.slope = slope, p.pool_k = pool_k, p.pooled = y, p.ldp = ldy;
    return gemm_launch<EPI_BIAS>(p, o, (cudaStream_t)stream);
}
