// Weight gradient of the shared-MLP 1x1 convolutions on point-major rows (SURVEY.md section 8 row a11,
// backward of Conv2d/Conv1d 1x1 in Partsize-identical/models/pointnet_util.py:213-215, 273-277, 343-345):
//     gw[n, k] += sum_r gy[r, n] * x[r, k],     gy [M, N] bf16, x [M, ldx] bf16, gw fp32
// M is 10^4 .. 5*10^5 while N x K is at most a few hundred squared: the contraction runs over the
// LONG dimension.  A library GEMM parallelises over the output, so this shape needed a batched GEMM
// over 2048-row chunks plus a reduction pass (and a dtype copy) per layer; here one kernel streams gy
// and x exactly once per 64x64 output tile: grid = (output tiles, row splits), every CTA walks its
// rows in 64-row chunks through a 3-stage cp.async pipeline, both operands are taken with
// ldmatrix.trans straight from the row-major tiles (the contraction index is the strided one), bf16
// mma.sync m16n8k16 with fp32 accumulators in registers, and one fp32 reduction per output element
// and CTA at the end (red.global.add.f32 into the caller's zero-initialised or accumulating buffer).
// HBM-bound: 2*M*(N + K) bytes per launch; tensor throughput is irrelevant at these widths, which is
// why this is mma.sync and not a tcgen05 pipeline (an accumulator of 64x64 would use 1/8 of TMEM).
#include <cuda_bf16.h>

#include <cstdlib>

#include "pcb_common.cuh"

namespace pcb {

constexpr int kWgThreads = 128;
constexpr int kWgTile = 64;                  // output tile (n and k) and rows per chunk
constexpr int kWgStages = 4;
constexpr int kWgTileBytes = kWgTile * kWgTile * 2;          // one operand chunk: 64 rows x 128 B

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3)
{
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float c[4], const uint32_t a[4], uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// tile element (row r, 16-byte chunk c) lives at chunk c ^ (r & 7) of its 128-byte row: the 8 row
// addresses of an ldmatrix 8x8 block then fall into 8 different bank groups
__device__ __forceinline__ uint32_t tile_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

// rows [r0, r0+64) x columns [c0, c0+64) of src [M, ld] -> swizzled smem tile; out-of-range -> zeros
__device__ __forceinline__ void load_tile(unsigned char *dst, const __nv_bfloat16 *src, int64_t M, int ncols, int ld,
                                          int64_t r0, int c0)
{
#pragma unroll
    for (int i = 0; i < (kWgTile * 8) / kWgThreads; ++i) {
        const int q = threadIdx.x + i * kWgThreads;
        const int r = q >> 3, c = q & 7;
        const int64_t gr = r0 + r;
        const int gc = c0 + c * 8;
        const bool ok = gr < M && gc < ncols;              // ncols % 8 == 0: a chunk is all in or all out
        const __nv_bfloat16 *g = ok ? src + gr * ld + gc : src;
        cp_async16(dst + tile_off(r, c), g, ok ? 16 : 0);
    }
}

__global__ void __launch_bounds__(kWgThreads, 3)
wgrad_rows_kernel(const __nv_bfloat16 *__restrict__ gy, const __nv_bfloat16 *__restrict__ x, int64_t M, int N, int K,
                  int ldgy, int ldx, int tiles_k, int64_t rows_per_split, float *__restrict__ gw, int ldw)
{
    extern __shared__ __align__(128) unsigned char smem[];
    pdl_wait();
    pdl_trigger();
    const int tile = blockIdx.x;
    const int n0 = (tile / tiles_k) * kWgTile, k0 = (tile % tiles_k) * kWgTile;
    const int64_t rbeg = (int64_t)blockIdx.y * rows_per_split;
    const int64_t rend = rbeg + rows_per_split < M ? rbeg + rows_per_split : M;
    const int nchunks = rbeg < rend ? (int)((rend - rbeg + kWgTile - 1) / kWgTile) : 0;
    const int K8 = (K + 7) & ~7;                           // x carries zero pad columns up to a multiple of 8

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32; // warp tile: 32 (n of gw) x 32 (k of gw)
    const bool warp_live = n0 + wm < N && k0 + wn < K;     // all-padding warp tiles skip the math

    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

    auto stage_a = [&](int s) { return smem + (size_t)s * 2 * kWgTileBytes; };
    auto stage_b = [&](int s) { return smem + (size_t)s * 2 * kWgTileBytes + kWgTileBytes; };
    auto issue = [&](int chunk) {
        if (chunk < nchunks) {
            const int s = chunk % kWgStages;
            const int64_t r0 = rbeg + (int64_t)chunk * kWgTile;
            // rows beyond this split's end belong to the next split: mask them through M = rend
            load_tile(stage_a(s), gy, rend, (N + 7) & ~7, ldgy, r0, n0);   // columns N..ldgy-1 of gy are zero
            load_tile(stage_b(s), x, rend, K8, ldx, r0, k0);
        }
        cp_async_commit();
    };

    for (int c = 0; c < kWgStages - 1; ++c) issue(c);
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        cp_async_wait<kWgStages - 2>();
        __syncthreads();                                   // chunk's tiles visible; stage (chunk-1)%S free again
        issue(chunk + kWgStages - 1);
        if (warp_live) {
            const uint32_t a_base = smem_u32(stage_a(chunk % kWgStages));
            const uint32_t b_base = smem_u32(stage_b(chunk % kWgStages));
#pragma unroll
            for (int kk = 0; kk < kWgTile; kk += 16) {     // 16 rows (the contraction index) per mma
                // ldmatrix.x4.trans: lanes 8j..8j+7 address the rows of 8x8 block j
                const int blk = lane >> 3, i = lane & 7;
                uint32_t a[2][4], b[4][2];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    // blocks: (rows kk..+7, n +0), (rows kk..+7, n +8), (rows kk+8.., n +0), (rows kk+8.., n +8)
                    const int r = kk + (blk >> 1) * 8 + i;
                    const int ch = (wm + mt * 16) / 8 + (blk & 1);
                    ldmatrix_x4_trans(a_base + tile_off(r, ch), a[mt][0], a[mt][1], a[mt][2], a[mt][3]);
                }
#pragma unroll
                for (int np = 0; np < 2; ++np) {
                    // blocks: (rows kk..+7, k +0), (rows kk+8.., k +0), (rows kk..+7, k +8), (rows kk+8.., k +8)
                    const int r = kk + (blk & 1) * 8 + i;
                    const int ch = (wn + np * 16) / 8 + (blk >> 1);
                    ldmatrix_x4_trans(b_base + tile_off(r, ch), b[2 * np][0], b[2 * np][1], b[2 * np + 1][0], b[2 * np + 1][1]);
                }
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) mma_bf16_16816(acc[mt][nt], a[mt], b[nt][0], b[nt][1]);
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();                                       // pipeline buffers are free: reuse them as the fp32 tile
    if (nchunks == 0) return;

    // Epilogue.  Every CTA adds its 64x64 partial into gw, so the adds of a few hundred CTAs land on
    // the same cache lines: stage the tile in shared memory and issue them as full 16-byte
    // reductions (red.global.add.v4.f32), consecutive threads on consecutive addresses -- one L2
    // sector operation per 8 values instead of per 2-4 with the accumulator's native layout.
    float *s_tile = reinterpret_cast<float *>(smem);       // [64][64 + 4]
    constexpr int kPitch = kWgTile + 4;
    if (warp_live) {
        const int g = lane >> 2, tig = lane & 3;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    *reinterpret_cast<float2 *>(s_tile + (wm + mt * 16 + g + h * 8) * kPitch + wn + nt * 8 + 2 * tig) =
                        make_float2(acc[mt][nt][2 * h], acc[mt][nt][2 * h + 1]);
    }
    __syncthreads();
    const int nvalid = N - n0 < kWgTile ? N - n0 : kWgTile;
    const int kvalid = K - k0 < kWgTile ? K - k0 : kWgTile;
    const bool vec = (ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(gw) & 15) == 0;
    if (vec) {
        const int kq = (kvalid + 3) >> 2;                  // float4 groups per row
        for (int q = threadIdx.x; q < nvalid * kq; q += kWgThreads) {
            const int n = q / kq, k4 = (q - n * kq) * 4;
            // (n < nvalid, k4 < kvalid lie in live warp tiles: a warp tile is skipped only when it starts outside gw)
            float4 v = *reinterpret_cast<const float4 *>(s_tile + n * kPitch + k4);
            float *dst = gw + (size_t)(n0 + n) * ldw + k0 + k4;
            if (k4 + 4 <= kvalid) {
                atomicAdd(reinterpret_cast<float4 *>(dst), v);
            } else {
                const float e[4] = {v.x, v.y, v.z, v.w};
                for (int j = 0; j < kvalid - k4; ++j) atomicAdd(dst + j, e[j]);
            }
        }
    } else {
        for (int q = threadIdx.x; q < nvalid * kvalid; q += kWgThreads) {
            const int n = q / kvalid, k = q - n * kvalid;
            atomicAdd(gw + (size_t)(n0 + n) * ldw + k0 + k, s_tile[n * kPitch + k]);
        }
    }
}

}  // namespace pcb

using namespace pcb;

PCB_API int pcb_wgrad_rows_bf16(const void *gy, const void *x, int64_t M, int N, int K, int ldgy, int ldx, float *gw,
                                int ldw, pcb_stream_t stream)
{
    PCB_REQUIRE(gy && x && gw, PCB_EINVAL);
    PCB_REQUIRE(M > 0 && N > 0 && K > 0 && ldw >= K, PCB_EINVAL);
    PCB_REQUIRE(ldgy % 8 == 0 && ldgy >= ((N + 7) & ~7) && ldx % 8 == 0 && ldx >= ((K + 7) & ~7), PCB_ERANGE);
    PCB_REQUIRE((reinterpret_cast<uintptr_t>(gy) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, PCB_EALIGN);
    const int tiles_n = (N + kWgTile - 1) / kWgTile, tiles_k = (K + kWgTile - 1) / kWgTile;
    const int64_t tiles = (int64_t)tiles_n * tiles_k;
    PCB_REQUIRE(tiles < (1ll << 31), PCB_ERANGE);
    // about 2 CTAs per SM in total (every CTA ends with a reduction into gw); a split is a whole number of 64-row chunks
    const int64_t chunks = ceil_div(M, kWgTile);
    static int ctas_per_sm = 0;                           // tuning hook: PCB_WGRAD_CTAS_PER_SM (default 2)
    if (!ctas_per_sm) {
        const char *e = getenv("PCB_WGRAD_CTAS_PER_SM");
        ctas_per_sm = e && atoi(e) > 0 ? atoi(e) : 2;
    }
    int64_t splits = ceil_div((int64_t)ctas_per_sm * PCB_NUM_SMS, tiles);
    if (splits > chunks) splits = chunks;
    if (splits > 65535) splits = 65535;
    const int64_t rows_per_split = ceil_div(chunks, splits) * kWgTile;
    splits = ceil_div(M, rows_per_split);
    const size_t smem = (size_t)kWgStages * 2 * kWgTileBytes;
    static bool attr_set[kMaxDevices] = {};
    if (cudaError_t e = smem_optin_once(wgrad_rows_kernel, (int)smem, attr_set)) return (int)e;
    dim3 grid((unsigned)tiles, (unsigned)splits);
    return (int)launch_pdl(wgrad_rows_kernel, grid, dim3(kWgThreads), smem, (cudaStream_t)stream, (const __nv_bfloat16 *)gy,
                           (const __nv_bfloat16 *)x, M, N, K, ldgy, ldx, tiles_k, rows_per_split, gw, ldw);
}
