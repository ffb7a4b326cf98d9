// Ball query -- replaces query_ball_point (Partsize-identical/models/pointnet_util.py:91-112,
// Highway_bridge/models/pointnet2_utils.py:97-112), which materialises a [B,S,N] distance
// matrix and an int64 [B,S,N] index tensor and fully sorts every row.
//
// Here: the cloud is staged once per CTA into shared memory by the TMA engine (one 1-D bulk
// copy per chunk, falling back to plain loads when the 16-byte alignment rule is not met), a
// pass adds |p|^2, and then one warp serves one query: lanes test 32 consecutive points per
// step, __ballot_sync + popc compacts the in-ball lanes in ascending index order, and the
// scan stops as soon as nsample indices are found.  No distance is ever written to memory.
//
// Distance arithmetic is square_distance(new_xyz, xyz) bit for bit (pcb_common.cuh:sqdist3);
// a point is in the ball iff !(d > radius2), radius2 = fp32(double(radius)^2).
#include "pcb_common.cuh"

namespace pcb {

constexpr int kBqWarps = 16;                 // queries per CTA
constexpr int kBqChunk = 8192;               // points staged per pass (8192 * 16 B = 128 KB)

__global__ void __launch_bounds__(kBqWarps * 32, 1)
ball_query_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz, int N, int S,
                  float r2, int nsample, int64_t *__restrict__ out, int chunk, int use_bulk)
{
    extern __shared__ __align__(16) float smem[];
    float *s_p = smem;                        // chunk * 3 (AoS, as in global memory)
    float *s_n = smem + (size_t)chunk * 3;    // chunk norms
    __shared__ __align__(8) uint64_t s_bar;

    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.x * kBqWarps + warp;
    const bool active = s < S;
    const float *cloud = xyz + (size_t)b * N * 3;

    float qx = 0.f, qy = 0.f, qz = 0.f, qn = 0.f;
    if (active) {
        const float *q = new_xyz + ((size_t)b * S + s) * 3;
        qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
        qn = norm3(qx, qy, qz);
    }
    int64_t *row = out + ((size_t)b * S + (active ? s : 0)) * nsample;
    int cnt = active ? 0 : nsample;           // inactive warps count as finished
    int first = N;

    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    unsigned phase = 0;
    for (int c0 = 0; c0 < N; c0 += chunk) {
        const int len = min(chunk, N - c0);
        // ---- stage xyz[c0 : c0+len] ----
        if (use_bulk) {
            if (threadIdx.x == 0) {
                uint32_t bytes = (uint32_t)len * 12u;
                mbar_expect_tx(&s_bar, bytes);
                bulk_g2s(s_p, cloud + (size_t)c0 * 3, bytes, &s_bar);
            }
            mbar_wait(&s_bar, phase);
            phase ^= 1;
        } else {
            for (int i = threadIdx.x; i < len * 3; i += blockDim.x) s_p[i] = __ldg(cloud + (size_t)c0 * 3 + i);
            __syncthreads();
        }
        for (int i = threadIdx.x; i < len; i += blockDim.x)
            s_n[i] = norm3(s_p[3 * i], s_p[3 * i + 1], s_p[3 * i + 2]);
        __syncthreads();

        // ---- scan ----
        for (int base = 0; base < len && cnt < nsample; base += 32) {
            int i = base + lane;
            bool in = false;
            if (i < len) {
                float d = sqdist3(qx, qy, qz, qn, s_p[3 * i], s_p[3 * i + 1], s_p[3 * i + 2], s_n[i]);
                in = !(d > r2);
            }
            unsigned m = __ballot_sync(PCB_FULL_MASK, in);
            if (m) {
                if (first == N) first = c0 + base + (__ffs(m) - 1);
                int pos = cnt + __popc(m & ((1u << lane) - 1u));
                if (in && pos < nsample) row[pos] = (int64_t)(c0 + i);
                cnt += __popc(m);
            }
        }
        // Block-uniform early exit; the barrier also protects the refill of the staging buffer.
        if (c0 + chunk < N && __syncthreads_and(cnt >= nsample)) break;
    }
    if (active && cnt < nsample) {
        // pad with the row's first index (N when the ball is empty)
        for (int k = cnt + lane; k < nsample; k += 32) row[k] = (int64_t)first;
    }
}

// ------------------------------------------------------------------------------------------
// Several radii around the same centroids in ONE scan (multi-scale grouping:
// PointNetSetAbstractionMsg pointnet_util.py:241-284, MultiScaleSetAbstraction pointnet2_utils.py:326-360
// call query_ball_point once per radius on identical (xyz, new_xyz)).  The distance of a pair is
// computed once and compared against every radius; each scale keeps its own count / first hit and
// stops taking points when full, the scan ends when all scales are full.  Candidates sit in shared
// memory as float4 (x, y, z, |p|^2): one LDS.128 per pair instead of four LDS.32.
// ------------------------------------------------------------------------------------------
constexpr int kBqMaxScales = 4;
#ifndef PCB_BQ_CHUNK
#define PCB_BQ_CHUNK 4096
#endif
#ifndef PCB_BQ_MINB
#define PCB_BQ_MINB 2                        // 62 registers: two CTAs (32 warps) per SM -- 140 -> 110 us at B = 16, 4096 -> 1024,
#endif                                       // two radii; chunks of 2048 / 1024 points measured 114 / 122 us (round 2)
constexpr int kBqMultiChunk = PCB_BQ_CHUNK;  // points per pass: 12 B staging + 16 B float4 each (4096: 112 KB)

struct BqScales {
    float r2[kBqMaxScales];
    int ns[kBqMaxScales];
    int64_t *out[kBqMaxScales];
    int nscales;
};

// (Serving 4 queries per warp from one candidate fetch was tried and was slower -- 0.31 ms against 0.175 ms
// for the four launches of the MSG step: fewer, longer dependent chains per SM.  One query per warp with
// 128 candidates per step is kept.)
__global__ void __launch_bounds__(kBqWarps * 32, PCB_BQ_MINB)
ball_query_multi_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz, int N, int S,
                        const BqScales p, int chunk, int use_bulk)
{
    extern __shared__ __align__(16) float smem[];
    float4 *s_p4 = reinterpret_cast<float4 *>(smem);                 // chunk float4
    float *s_raw = smem + (size_t)chunk * 4;                         // chunk * 3 (AoS as in global memory)
    __shared__ __align__(8) uint64_t s_bar;

    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = blockIdx.x * kBqWarps + warp;
    const bool active = s < S;
    const float *cloud = xyz + (size_t)b * N * 3;

    float qx = 0.f, qy = 0.f, qz = 0.f, qn = 0.f;
    if (active) {
        const float *q = new_xyz + ((size_t)b * S + s) * 3;
        qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
        qn = norm3(qx, qy, qz);
    }
    int cnt[kBqMaxScales], first[kBqMaxScales];
#pragma unroll
    for (int k = 0; k < kBqMaxScales; ++k) {
        cnt[k] = (active && k < p.nscales) ? 0 : 0x7fffffff;         // absent scales / idle warps count as full
        first[k] = N;
    }
    auto all_full = [&]() {
        bool f = true;
#pragma unroll
        for (int k = 0; k < kBqMaxScales; ++k) f = f && (k >= p.nscales || cnt[k] >= p.ns[k]);
        return f;
    };

    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    unsigned phase = 0;
    for (int c0 = 0; c0 < N; c0 += chunk) {
        const int len = min(chunk, N - c0);
        if (use_bulk) {
            if (threadIdx.x == 0) {
                const uint32_t bytes = (uint32_t)len * 12u;
                mbar_expect_tx(&s_bar, bytes);
                bulk_g2s(s_raw, cloud + (size_t)c0 * 3, bytes, &s_bar);
            }
            mbar_wait(&s_bar, phase);
            phase ^= 1;
        } else {
            for (int i = threadIdx.x; i < len * 3; i += blockDim.x) s_raw[i] = __ldg(cloud + (size_t)c0 * 3 + i);
            __syncthreads();
        }
        for (int i = threadIdx.x; i < len; i += blockDim.x) {
            const float x = s_raw[3 * i], y = s_raw[3 * i + 1], z = s_raw[3 * i + 2];
            s_p4[i] = make_float4(x, y, z, norm3(x, y, z));
        }
        __syncthreads();

        // 128 candidates per step: four independent LDS.128 + distance chains per lane, loop and fullness
        // bookkeeping once per 128 pairs; the four ballots of a scale are taken in index order
        for (int base = 0; base < len && !all_full(); base += 128) {
            float d[4];
            bool valid[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = base + u * 32 + lane;
                valid[u] = i < len;
                const float4 c = valid[u] ? s_p4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                d[u] = sqdist3(qx, qy, qz, qn, c.x, c.y, c.z, c.w);
            }
#pragma unroll
            for (int k = 0; k < kBqMaxScales; ++k) {
                if (k < p.nscales && cnt[k] < p.ns[k]) {             // warp-uniform
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const bool in = valid[u] && !(d[u] > p.r2[k]);
                        const unsigned m = __ballot_sync(PCB_FULL_MASK, in);
                        if (m) {
                            if (first[k] == N) first[k] = c0 + base + u * 32 + (__ffs(m) - 1);
                            const int pos = cnt[k] + __popc(m & ((1u << lane) - 1u));
                            if (in && pos < p.ns[k])
                                p.out[k][((size_t)b * S + s) * p.ns[k] + pos] = (int64_t)(c0 + base + u * 32 + lane);
                            cnt[k] += __popc(m);
                        }
                    }
                }
            }
        }
        // block-uniform early exit; the barrier also protects the refill of the staging buffers
        if (c0 + chunk < N && __syncthreads_and(all_full())) break;
    }
    if (active) {
#pragma unroll
        for (int k = 0; k < kBqMaxScales; ++k)
            if (k < p.nscales && cnt[k] < p.ns[k]) {                 // pad with the row's first index (N: empty ball)
                int64_t *row = p.out[k] + ((size_t)b * S + s) * p.ns[k];
                for (int j = cnt[k] + lane; j < p.ns[k]; j += 32) row[j] = (int64_t)first[k];
            }
    }
}

}  // namespace pcb

PCB_API int pcb_ball_query_multi_f32(const float *xyz, const float *new_xyz, int B, int N, int S, int nscales,
                                     const float *radius2, const int *nsample, int64_t *const *out_idx,
                                     pcb_stream_t stream)
{
    using namespace pcb;
    PCB_REQUIRE(xyz && new_xyz && radius2 && nsample && out_idx, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && S > 0 && nscales >= 1, PCB_EINVAL);
    PCB_REQUIRE(B <= 65535 && nscales <= kBqMaxScales, PCB_ERANGE);
    BqScales p;
    p.nscales = nscales;
    for (int k = 0; k < kBqMaxScales; ++k) {
        p.r2[k] = k < nscales ? radius2[k] : 0.f;
        p.ns[k] = k < nscales ? nsample[k] : 0;
        p.out[k] = k < nscales ? out_idx[k] : nullptr;
        PCB_REQUIRE(k >= nscales || (p.ns[k] > 0 && p.out[k]), PCB_EINVAL);
    }
    const int chunk = N < kBqMultiChunk ? ((N + 3) & ~3) : kBqMultiChunk;
    const int use_bulk = ((reinterpret_cast<uintptr_t>(xyz) & 15) == 0) && (N % 4 == 0);
    const size_t smem = (size_t)chunk * 28;
    static bool attr_set[kMaxDevices] = {};
    if (cudaError_t e = smem_optin_once(ball_query_multi_kernel, kBqMultiChunk * 28, attr_set)) return (int)e;
    dim3 grid((unsigned)ceil_div(S, kBqWarps), (unsigned)B);
    ball_query_multi_kernel<<<grid, kBqWarps * 32, smem, (cudaStream_t)stream>>>(xyz, new_xyz, N, S, p, chunk, use_bulk);
    PCB_RETURN_LAUNCH_STATUS();
}

PCB_API int pcb_ball_query_f32(const float *xyz, const float *new_xyz, int B, int N, int S,
                               float radius2, int nsample, int64_t *out_idx, pcb_stream_t stream)
{
    using namespace pcb;
    PCB_REQUIRE(xyz && new_xyz && out_idx, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && S > 0 && nsample > 0, PCB_EINVAL);
    PCB_REQUIRE(B <= 65535, PCB_ERANGE);
    int chunk = N < kBqChunk ? ((N + 3) & ~3) : kBqChunk;
    // bulk copies need 16-byte aligned source rows and sizes: cloud stride N*12 and every
    // chunk length*12 must be multiples of 16
    int use_bulk = ((reinterpret_cast<uintptr_t>(xyz) & 15) == 0) && (N % 4 == 0);
    size_t smem = (size_t)chunk * 16;
    static bool attr_set[kMaxDevices] = {};
    if (cudaError_t e = smem_optin_once(ball_query_kernel, kBqChunk * 16, attr_set)) return (int)e;
    dim3 grid((unsigned)ceil_div(S, kBqWarps), (unsigned)B);
    ball_query_kernel<<<grid, kBqWarps * 32, smem, (cudaStream_t)stream>>>(xyz, new_xyz, N, S, radius2, nsample,
                                                                          out_idx, chunk, use_bulk);
    PCB_RETURN_LAUNCH_STATUS();
}
