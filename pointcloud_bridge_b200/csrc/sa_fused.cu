// Fused set-abstraction / EdgeConv block for inference:
//   gather neighbours -> [dxyz | feat] rows -> (1x1 conv + folded BN + activation) x L on the
//   5th-generation tensor cores (tcgen05.mma, bf16 in, fp32 accumulate in TMEM) -> max over the
//   K neighbours -> [B*S, C_out].
// Replaces, for eval-mode networks, index_points + cat + permute + Conv2d/BN/ReLU x3 + torch.max of
//   Partsize-identical/models/pointnet_util.py:137-147, 203-217, 258-279,
//   Highway_bridge/models/pointnet2_utils.py:140-154, 341-356 and the EdgeConv of
//   Highway_bridge/models/DGCNN.py:72-109, 134-148.
// The grouped tensor [B,S,K,3+D] and every intermediate activation stay on chip.
//
// One CTA = 128 threads = one 128-row M tile (rows = (group, neighbour) pairs; a tile holds
// floor(128/K) whole groups).  Thread t owns row t everywhere:
//   * stage 0: thread t gathers its neighbour's row from global memory, converts to bf16 and writes it
//     into the A operand buffer in the canonical K-major / no-swizzle UMMA layout
//     [k/8][row][8 elements] (16-byte units, rows 16 B apart, 8-row core matrices 128 B apart);
//   * one elected thread issues K/16 tcgen05.mma (M=128, N=C_layer) against the layer's weights,
//     which sit in shared memory in the same layout ([k/8][n][8], packed on the host), and
//     commits to an mbarrier;
//   * epilogue: tcgen05.ld 32x32b (lane t of TMEM = row t), + bias, activation, -> bf16 -> next
//     layer's A buffer (same layout, so thread t again writes 16-byte units of row t);
//   * after the last layer the tile is staged as bf16 [row][C] and reduced over K per group.
// Weights of all layers are copied into shared memory once per CTA (TMA bulk copy); CTAs are
// persistent over tiles.  Envelope: C_in_pad, C_l multiples of 16, C_l <= 256, total shared
// memory <= 200 KB (wider layers use the unfused path).
#include <cuda_bf16.h>

#include "pcb_common.cuh"
#include "umma.cuh"

namespace pcb {

constexpr int kSaThreads = 128;
constexpr int kSaMaxLayers = 3;

struct SaParams {
    const float *xyz;        // [B,N,3]
    const float *points;     // [B,N,D] point-major fp32 or nullptr
    const float *new_xyz;    // [B,S,3] (mode 0) / unused (mode 1)
    const int64_t *idx;      // [B,S,K]
    int B, N, S, K, D;
    int mode;                // 0: [dxyz|feat] / [feat|dxyz] grouping; 1: EdgeConv [x_nbr - x_ctr | x_ctr] from points
    int xyz_first;
    int nlayers;
    int kdim[kSaMaxLayers + 1];   // kdim[0] = padded input width, kdim[l] = padded width after layer l
    int cout;                // true (unpadded) output channels of the last layer
    const void *wblob;       // packed bf16 weights, layer after layer: [kdim[l]/8][kdim[l+1]][8]
    const float *bias;       // concatenated fp32 biases (padded widths)
    float slope;             // activation: max(x, slope * x)  (0 = ReLU, 0.2 = LeakyReLU(0.2))
    void *out;               // [B*S, cout] bf16 (out_bf16) or fp32
    int out_bf16;
    int rows_per_tile;       // floor(128 / K) * K
    int64_t total_rows;      // B*S*K
    int ntiles;
};

__global__ void __launch_bounds__(kSaThreads)
sa_fused_kernel(const SaParams p, int tmem_cols, int a_bytes, int w_bytes)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *sW = smem_raw;                         // packed weights of all layers
    unsigned char *sA0 = sW + w_bytes;                    // activation ping
    unsigned char *sA1 = sA0 + a_bytes;                   // activation pong / pooling stage
    float *sBias = reinterpret_cast<float *>(sA1 + a_bytes);
    __shared__ __align__(8) uint64_t s_bar_w, s_bar_mma;
    __shared__ uint32_t s_tmem;

    const int tid = threadIdx.x, warp = tid >> 5;
    int nbias = 0;
    for (int l = 1; l <= p.nlayers; ++l) nbias += p.kdim[l];

    if (tid == 0) {
        mbar_init(&s_bar_w, 1);
        mbar_init(&s_bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem), (uint32_t)tmem_cols);
    for (int i = tid; i < nbias; i += kSaThreads) sBias[i] = __ldg(p.bias + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = s_tmem;
    if (tid == 0) {                                        // weights: one TMA bulk copy per CTA
        mbar_expect_tx(&s_bar_w, (uint32_t)w_bytes);
        bulk_g2s(sW, p.wblob, (uint32_t)w_bytes, &s_bar_w);
    }
    mbar_wait(&s_bar_w, 0);

    uint32_t mma_phase = 0;
    const int K = p.K;
    const int groups_per_tile = p.rows_per_tile / K;
    const int C0 = (p.mode == 0) ? 3 + p.D : 2 * p.D;

    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        // ---------------- stage 0: gather row `tid` into sA0 ----------------
        {
            const int64_t gr = (int64_t)tile * p.rows_per_tile + tid;
            const bool valid = tid < p.rows_per_tile && gr < p.total_rows;
            int64_t b = 0, pt = 0;
            long long nb = 0;
            float cx = 0.f, cy = 0.f, cz = 0.f;
            if (valid) {
                const int64_t grp = gr / K;                // b*S + s
                b = grp / p.S;
                pt = grp - b * p.S;
                nb = p.idx[gr];
                nb = nb < 0 ? 0 : (nb > p.N - 1 ? p.N - 1 : nb);
                if (p.mode == 0) {
                    const float *c = p.new_xyz + grp * 3;
                    cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
                }
            }
            const float *nrow = p.points ? p.points + ((size_t)b * p.N + nb) * p.D : nullptr;
            const float *crow = p.points ? p.points + ((size_t)b * p.N + pt) * p.D : nullptr;   // mode 1: centre = point s
            const float *nx = p.xyz ? p.xyz + ((size_t)b * p.N + nb) * 3 : nullptr;
            const int nchunks = p.kdim[0] / 8;
            for (int kc = 0; kc < nchunks; ++kc) {
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = kc * 8 + j;
                    float x = 0.f;
                    if (valid && c < C0) {
                        if (p.mode == 0) {
                            const int cxyz = p.xyz_first ? c : c - p.D;
                            if (cxyz >= 0 && cxyz < 3) {
                                x = __ldg(nx + cxyz) - (cxyz == 0 ? cx : (cxyz == 1 ? cy : cz));
                            } else {
                                x = __ldg(nrow + (p.xyz_first ? c - 3 : c));
                            }
                        } else {
                            x = c < p.D ? __ldg(nrow + c) - __ldg(crow + c) : __ldg(crow + c - p.D);
                        }
                    }
                    v[j] = x;
                }
                *reinterpret_cast<uint4 *>(sA0 + (size_t)kc * (128 * 16) + tid * 16) = pack8(v);
            }
        }
        proxy_fence();
        tc_fence_before();
        __syncthreads();

        // ---------------- layers ----------------
        unsigned char *sIn = sA0, *sOut = sA1;
        int woff = 0, boff = 0;
        for (int l = 0; l < p.nlayers; ++l) {
            const int Kd = p.kdim[l], Nd = p.kdim[l + 1];
            if (tid == 0) {
                tc_fence_after();
                const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(Nd >> 3) << 17) | ((128u >> 4) << 24);
                const uint32_t a_base = smem_u32(sIn), b_base = smem_u32(sW + woff);
                const uint32_t lbo_a = 128 * 16, lbo_b = (uint32_t)Nd * 16;
                for (int kk = 0; kk < Kd / 16; ++kk) {
                    const uint64_t da = umma_desc(a_base + (uint32_t)kk * 2 * lbo_a, lbo_a, 128);
                    const uint64_t db = umma_desc(b_base + (uint32_t)kk * 2 * lbo_b, lbo_b, 128);
                    umma_bf16(tmem_base, da, db, idesc, kk > 0 ? 1u : 0u);
                }
                umma_commit(&s_bar_mma);
            }
            mbar_wait(&s_bar_mma, mma_phase);
            mma_phase ^= 1;
            tc_fence_after();

            const bool last = l == p.nlayers - 1;
            const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
            const int pitch = Nd + 8;                         // bf16 elements, pooling stage
            for (int c0 = 0; c0 < Nd; c0 += 16) {
                float v[16];
                tmem_ld16(trow + (uint32_t)c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float x = v[i] + sBias[boff + c0 + i];
                    v[i] = fmaxf(x, p.slope * x);
                }
                if (!last) {
                    *reinterpret_cast<uint4 *>(sOut + (size_t)(c0 / 8) * (128 * 16) + tid * 16) = pack8(v);
                    *reinterpret_cast<uint4 *>(sOut + (size_t)(c0 / 8 + 1) * (128 * 16) + tid * 16) = pack8(v + 8);
                } else {
                    __nv_bfloat16 *st = reinterpret_cast<__nv_bfloat16 *>(sOut) + (size_t)tid * pitch + c0;
                    *reinterpret_cast<uint4 *>(st) = pack8(v);
                    *reinterpret_cast<uint4 *>(st + 8) = pack8(v + 8);
                }
            }
            woff += (Kd / 8) * Nd * 16;
            boff += Nd;
            proxy_fence();
            tc_fence_before();
            __syncthreads();
            unsigned char *t = sIn;
            sIn = sOut;
            sOut = t;
        }

        // ---------------- max over the K neighbours of every group (stage is now in sIn) ----------------
        {
            const int Nd = p.kdim[p.nlayers];
            const int pitch = Nd + 8;
            const __nv_bfloat16 *st = reinterpret_cast<const __nv_bfloat16 *>(sIn);
            const int64_t g0 = (int64_t)tile * groups_per_tile;
            const int64_t total_groups = (int64_t)p.B * p.S;
            for (int t = tid; t < groups_per_tile * p.cout; t += kSaThreads) {
                const int g = t / p.cout, c = t - g * p.cout;
                if (g0 + g >= total_groups) break;
                float m = __bfloat162float(st[(size_t)(g * K) * pitch + c]);
                for (int k = 1; k < K; ++k) m = fmaxf(m, __bfloat162float(st[(size_t)(g * K + k) * pitch + c]));
                if (p.out_bf16)
                    reinterpret_cast<__nv_bfloat16 *>(p.out)[(g0 + g) * p.cout + c] = __float2bfloat16(m);
                else
                    reinterpret_cast<float *>(p.out)[(g0 + g) * p.cout + c] = m;
            }
        }
        __syncthreads();                                   // stage buffer is reused by the next tile
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

}  // namespace pcb

using namespace pcb;

PCB_API int pcb_sa_fused_bf16(const float *xyz, const float *points, const float *new_xyz, const int64_t *idx, int B,
                              int N, int S, int K, int D, int mode, int xyz_first, int nlayers, const int *kdim,
                              int cout, const void *wblob, const float *bias, float slope, void *out, int out_bf16,
                              pcb_stream_t stream)
{
    PCB_REQUIRE(idx && kdim && wblob && bias && out, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && S > 0 && K > 0 && D >= 0, PCB_EINVAL);
    PCB_REQUIRE(nlayers >= 1 && nlayers <= kSaMaxLayers && K <= 128, PCB_ERANGE);
    PCB_REQUIRE((mode == 0 && xyz && new_xyz && (points || D == 0)) || (mode == 1 && points && D > 0), PCB_EINVAL);
    PCB_REQUIRE((reinterpret_cast<uintptr_t>(wblob) & 15) == 0, PCB_EALIGN);
    SaParams p;
    p.xyz = xyz, p.points = points, p.new_xyz = new_xyz, p.idx = idx;
    p.B = B, p.N = N, p.S = S, p.K = K, p.D = D, p.mode = mode, p.xyz_first = xyz_first, p.nlayers = nlayers;
    int maxw = 0, w_bytes = 0, nbias = 0;
    for (int l = 0; l <= nlayers; ++l) {
        PCB_REQUIRE(kdim[l] > 0 && kdim[l] % 16 == 0 && kdim[l] <= 256, PCB_ERANGE);
        p.kdim[l] = kdim[l];
        if (kdim[l] > maxw) maxw = kdim[l];
        if (l > 0) {
            w_bytes += kdim[l - 1] * kdim[l] * 2;
            nbias += kdim[l];
        }
    }
    const int c0 = mode == 0 ? 3 + D : 2 * D;
    PCB_REQUIRE(c0 <= kdim[0] && cout > 0 && cout <= kdim[nlayers], PCB_ERANGE);
    p.cout = cout, p.wblob = wblob, p.bias = bias, p.slope = slope, p.out = out, p.out_bf16 = out_bf16;
    p.rows_per_tile = (128 / K) * K;
    p.total_rows = (int64_t)B * S * K;
    p.ntiles = (int)ceil_div(p.total_rows, p.rows_per_tile);
    // activation buffers: UMMA layout needs 128 * width * 2 B, the pooling stage 128 * (width + 8) * 2 B
    const int a_bytes = ((128 * (maxw + 8) * 2) + 127) & ~127;
    w_bytes = (w_bytes + 127) & ~127;
    const size_t smem = (size_t)w_bytes + 2 * (size_t)a_bytes + (size_t)nbias * 4 + 128;
    PCB_REQUIRE(smem <= 200 * 1024, PCB_ERANGE);
    int tmem_cols = 32;
    while (tmem_cols < maxw) tmem_cols <<= 1;
    cudaError_t e = cudaFuncSetAttribute(sa_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int ctas_per_sm = (int)((220 * 1024) / (smem + 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm * tmem_cols > 512) ctas_per_sm = 512 / tmem_cols;
    if (ctas_per_sm > 4) ctas_per_sm = 4;
    int grid = PCB_NUM_SMS * ctas_per_sm;
    if (grid > p.ntiles) grid = p.ntiles;
    sa_fused_kernel<<<grid, kSaThreads, smem, (cudaStream_t)stream>>>(p, tmem_cols, a_bytes, w_bytes);
    PCB_RETURN_LAUNCH_STATUS();
}
