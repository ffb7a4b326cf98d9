// Nearest-neighbour searches in coordinate space (3-D points):
//   pcb_three_nn_f32   -- the k=3 (k=4) nearest of xyz2 for every xyz1 point + inverse-distance
//                         weights; replaces square_distance + full sort of
//                         Partsize-identical/models/pointnet_util.py:325-332 and
//                         Highway_bridge/models/pointnet2_utils.py:183-191, 253-262.
//   pcb_knn_f32 (D=3)  -- DGCNN.knn on coordinates, Highway_bridge/models/DGCNN.py:49-70.
//   pcb_knn_cdist_f32  -- torch.cdist + topk(largest=False),
//                         Highway_bridge/models/attention_modules.py:584-586, 736-738.
// None of them materialises the [B,N,M] distance matrix.  Candidates are staged in shared
// memory as float4 (x, y, z, |p|^2); distances use the reference's exact fp32 formula
// (pcb_common.cuh) and selection orders by (distance, index).
#include "pcb_common.cuh"

namespace pcb {

// ------------------------------------------------------------------------------------------
// staging helper: cloud[c0 : c0+len] -> s_pts[0:len] as (x, y, z, |p|^2)
// ------------------------------------------------------------------------------------------
// cf != 0: the cloud is channels-first ([3,N], as DGCNN passes x), else points-major ([N,3]).
__device__ __forceinline__ float3 load_point(const float *cloud, int i, int N, int cf)
{
    if (cf) return make_float3(__ldg(cloud + i), __ldg(cloud + N + i), __ldg(cloud + 2 * (size_t)N + i));
    const float *p = cloud + (size_t)i * 3;
    return make_float3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}
__device__ __forceinline__ void stage_points(float4 *s_pts, const float *cloud, int c0, int len,
                                             int N = 0, int cf = 0)
{
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
        float3 v = load_point(cloud, c0 + i, N, cf);
        s_pts[i] = make_float4(v.x, v.y, v.z, norm3(v.x, v.y, v.z));
    }
}

// ------------------------------------------------------------------------------------------
// three_nn: one thread per query, top-K in registers, candidates broadcast from smem.
// ------------------------------------------------------------------------------------------
constexpr int kNnThreads = 256;
constexpr int kNnChunk = 4096;               // candidates per staging pass (64 KB)

template <int K>
__global__ void __launch_bounds__(kNnThreads)
three_nn_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2, int N, int S,
                float *__restrict__ out_dist, int64_t *__restrict__ out_idx,
                float *__restrict__ out_w, int chunk)
{
    extern __shared__ __align__(16) float4 s_pts[];
    const int b = blockIdx.y;
    const int n = blockIdx.x * kNnThreads + threadIdx.x;
    const bool active = n < N;
    const float *cloud2 = xyz2 + (size_t)b * S * 3;

    float qx = 0.f, qy = 0.f, qz = 0.f, qn = 0.f;
    if (active) {
        const float *q = xyz1 + ((size_t)b * N + n) * 3;
        qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
        qn = norm3(qx, qy, qz);
    }
    float bd[K];
    int bi[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        bd[j] = __int_as_float(0x7f800000);      // +inf
        bi[j] = 0;
    }
    for (int c0 = 0; c0 < S; c0 += chunk) {
        const int len = min(chunk, S - c0);
        if (c0 > 0) __syncthreads();
        stage_points(s_pts, cloud2, c0, len);
        __syncthreads();
        if (active) {
#pragma unroll 4
            for (int s = 0; s < len; ++s) {
                float4 p = s_pts[s];
                float d = sqdist3(qx, qy, qz, qn, p.x, p.y, p.z, p.w);   // square_distance(xyz1, xyz2)
                if (d < bd[K - 1]) {
                    bd[K - 1] = d;
                    bi[K - 1] = c0 + s;
#pragma unroll
                    for (int j = K - 1; j > 0; --j) {
                        if (bd[j] < bd[j - 1]) {                         // strict: earlier index stays first
                            float td = bd[j]; bd[j] = bd[j - 1]; bd[j - 1] = td;
                            int ti = bi[j]; bi[j] = bi[j - 1]; bi[j - 1] = ti;
                        }
                    }
                }
            }
        }
    }
    if (!active) return;
    const size_t o = ((size_t)b * N + n) * K;
    float rec[K];
    float norm = 0.f;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        out_dist[o + j] = bd[j];
        out_idx[o + j] = (int64_t)bi[j];
        // dist_recip = 1.0 / (dists + 1e-8); norm = sum(dist_recip)   (pointnet_util.py:330-331)
        rec[j] = __fdiv_rn(1.0f, __fadd_rn(bd[j], 1e-8f));
        norm = (j == 0) ? rec[0] : __fadd_rn(norm, rec[j]);
    }
    if (out_w) {
#pragma unroll
        for (int j = 0; j < K; ++j) out_w[o + j] = __fdiv_rn(rec[j], norm);
    }
}

template <int K>
static int launch_three_nn(const float *xyz1, const float *xyz2, int B, int N, int S, float *od,
                           int64_t *oi, float *ow, cudaStream_t st)
{
    int chunk = S < kNnChunk ? S : kNnChunk;
    size_t smem = (size_t)chunk * sizeof(float4);
    cudaError_t e = cudaFuncSetAttribute(three_nn_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kNnChunk * (int)sizeof(float4));
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)ceil_div(N, kNnThreads), (unsigned)B);
    three_nn_kernel<K><<<grid, kNnThreads, smem, st>>>(xyz1, xyz2, N, S, od, oi, ow, chunk);
    PCB_RETURN_LAUNCH_STATUS();
}

// ------------------------------------------------------------------------------------------
// kNN with k up to 64: one warp per query, sorted top list distributed over the lanes
// (NPL entries per lane), threshold filter + ballot, warp-wide shuffle insertion.
// ------------------------------------------------------------------------------------------
enum { MODE_PD = 0, MODE_CDIST = 1 };

constexpr int kKnnWarps = 16;
constexpr int kKnnChunk = 8192;              // 128 KB of float4

template <int NPL>
struct WarpList {
    float d[NPL];
    int i[NPL];
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
            d[r] = __int_as_float(0x7f800000);
            i[r] = 0x7fffffff;
        }
    }
    // value at list position pos (warp-uniform pos)
    __device__ __forceinline__ float dist_at(int pos) const
    {
        float v = 0.f;
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
            float t = __shfl_sync(PCB_FULL_MASK, d[r], pos & 31);
            if ((pos >> 5) == r) v = t;
        }
        return v;
    }
    // insert (v, vi) keeping (distance, index) ascending order; the last entry falls off
    __device__ __forceinline__ void insert(float v, int vi, int lane)
    {
        float carry_d = 0.f;
        int carry_i = 0;
        bool carry_gt = false;
#pragma unroll
        for (int r = 0; r < NPL; ++r) {
            const bool gt = (d[r] > v) || (d[r] == v && i[r] > vi);
            float up_d = __shfl_up_sync(PCB_FULL_MASK, d[r], 1);
            int up_i = __shfl_up_sync(PCB_FULL_MASK, i[r], 1);
            bool up_gt = __shfl_up_sync(PCB_FULL_MASK, (int)gt, 1) != 0;
            // what lane 31 of this register row hands to lane 0 of the next one
            const float last_d = __shfl_sync(PCB_FULL_MASK, d[r], 31);
            const int last_i = __shfl_sync(PCB_FULL_MASK, i[r], 31);
            const bool last_gt = __shfl_sync(PCB_FULL_MASK, (int)gt, 31) != 0;
            if (lane == 0) {
                up_d = carry_d;
                up_i = carry_i;
                up_gt = (r == 0) ? false : carry_gt;
            }
            if (gt) {
                d[r] = up_gt ? up_d : v;
                i[r] = up_gt ? up_i : vi;
            }
            carry_d = last_d;
            carry_i = last_i;
            carry_gt = last_gt;
        }
    }
};

template <int MODE, int NPL>
__global__ void __launch_bounds__(kKnnWarps * 32)
knn_xyz_kernel(const float *__restrict__ xyz, int N, int k, int qpw,
               int64_t *__restrict__ out_idx, float *__restrict__ out_dist, int chunk, int cf)
{
    extern __shared__ __align__(16) float4 s_pts[];
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *cloud = xyz + (size_t)b * N * 3;
    const int nchunks = (N + chunk - 1) / chunk;
    const int q_begin = blockIdx.x * (kKnnWarps * qpw);

    for (int pass = 0; pass < qpw; ++pass) {
        const int q = q_begin + pass * kKnnWarps + warp;
        const bool active = q < N;                       // warp-uniform
        float qx = 0.f, qy = 0.f, qz = 0.f, qn = 0.f;
        if (active) {
            float3 v = load_point(cloud, q, N, cf);
            qx = v.x, qy = v.y, qz = v.z;
            qn = norm3(qx, qy, qz);
        }
        float ax = qx, ay = qy, az = qz;
        if (MODE == MODE_CDIST) {                        // x1_ = cat(-2 * x1, |x1|^2, 1)
            ax = __fmul_rn(qx, -2.0f);
            ay = __fmul_rn(qy, -2.0f);
            az = __fmul_rn(qz, -2.0f);
        }
        WarpList<NPL> list;
        list.init();
        float thr = __int_as_float(0x7f800000);
        for (int c = 0; c < nchunks; ++c) {
            const int c0 = c * chunk;
            const int len = min(chunk, N - c0);
            if (nchunks > 1 || pass == 0) {
                if (c > 0 || pass > 0) __syncthreads();
                stage_points(s_pts, cloud, c0, len, N, cf);
                __syncthreads();
            }
            if (!active) continue;
            for (int base = 0; base < len; base += 32) {
                const int j = base + lane;
                float d = __int_as_float(0x7f800000);
                if (j < len) {
                    float4 p = s_pts[j];
                    if (MODE == MODE_PD) {
                        d = sqdist3(qx, qy, qz, qn, p.x, p.y, p.z, p.w);
                    } else {
                        float t = cdist3_pre(ax, ay, az, qn, p.x, p.y, p.z, p.w);
                        t = t < 0.0f ? 0.0f : t;             // clamp_min_(0)
                        d = __fsqrt_rn(t);
                    }
                }
                unsigned m = __ballot_sync(PCB_FULL_MASK, d < thr);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float v = __shfl_sync(PCB_FULL_MASK, d, src);
                    if (v < thr) {                           // thr may have dropped within this batch
                        list.insert(v, c0 + base + src, lane);
                        thr = list.dist_at(k - 1);
                    }
                }
            }
        }
        if (active) {
#pragma unroll
            for (int r = 0; r < NPL; ++r) {
                const int pos = r * 32 + lane;
                if (pos < k) {
                    out_idx[((size_t)b * N + q) * k + pos] = (int64_t)list.i[r];
                    if (out_dist) out_dist[((size_t)b * N + q) * k + pos] = list.d[r];
                }
            }
        }
    }
}

template <int MODE>
static int launch_knn_xyz(const float *xyz, int B, int N, int k, int cf, int64_t *oi, float *od,
                          cudaStream_t st)
{
    int chunk = N < kKnnChunk ? N : kKnnChunk;
    size_t smem = (size_t)chunk * sizeof(float4);
    // enough CTAs for >= 2 waves of 148 SMs, but amortise the staging over several queries per warp
    int qpw = 8;
    while (qpw > 1 && (int64_t)B * ceil_div(N, kKnnWarps * qpw) < 2 * PCB_NUM_SMS) qpw >>= 1;
    dim3 grid((unsigned)ceil_div(N, kKnnWarps * qpw), (unsigned)B);
    cudaError_t e;
    if (k <= 32) {
        e = cudaFuncSetAttribute(knn_xyz_kernel<MODE, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kKnnChunk * (int)sizeof(float4));
        if (e != cudaSuccess) return (int)e;
        knn_xyz_kernel<MODE, 1><<<grid, kKnnWarps * 32, smem, st>>>(xyz, N, k, qpw, oi, od, chunk, cf);
    } else {
        e = cudaFuncSetAttribute(knn_xyz_kernel<MODE, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 kKnnChunk * (int)sizeof(float4));
        if (e != cudaSuccess) return (int)e;
        knn_xyz_kernel<MODE, 2><<<grid, kKnnWarps * 32, smem, st>>>(xyz, N, k, qpw, oi, od, chunk, cf);
    }
    PCB_RETURN_LAUNCH_STATUS();
}

int knn_xyz_pd(const float *xyz, int B, int N, int k, int cf, int64_t *oi, float *od, cudaStream_t st)
{
    return launch_knn_xyz<MODE_PD>(xyz, B, N, k, cf, oi, od, st);
}

}  // namespace pcb

PCB_API int pcb_three_nn_f32(const float *xyz1, const float *xyz2, int B, int N, int S, int k,
                             float *out_dist, int64_t *out_idx, float *out_weight, pcb_stream_t stream)
{
    using namespace pcb;
    PCB_REQUIRE(xyz1 && xyz2 && out_dist && out_idx, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && S > 0 && k > 0, PCB_EINVAL);
    PCB_REQUIRE(k <= 8 && k <= S && B <= 65535, PCB_ERANGE);
    cudaStream_t st = (cudaStream_t)stream;
    switch (k) {
        case 1: return launch_three_nn<1>(xyz1, xyz2, B, N, S, out_dist, out_idx, out_weight, st);
        case 2: return launch_three_nn<2>(xyz1, xyz2, B, N, S, out_dist, out_idx, out_weight, st);
        case 3: return launch_three_nn<3>(xyz1, xyz2, B, N, S, out_dist, out_idx, out_weight, st);
        case 4: return launch_three_nn<4>(xyz1, xyz2, B, N, S, out_dist, out_idx, out_weight, st);
        case 5: return launch_three_nn<5>(xyz1, xyz2, B, N, S, out_dist, out_idx, out_weight, st);
        case 6: return launch_three_nn<6>(xyz1, xyz2, B, N, S, out_dist, out_idx, out_weight, st);
        case 7: return launch_three_nn<7>(xyz1, xyz2, B, N, S, out_dist, out_idx, out_weight, st);
        default: return launch_three_nn<8>(xyz1, xyz2, B, N, S, out_dist, out_idx, out_weight, st);
    }
}

PCB_API int pcb_knn_cdist_f32(const float *xyz, int B, int N, int k, int64_t *out_idx, float *out_dist,
                              pcb_stream_t stream)
{
    using namespace pcb;
    PCB_REQUIRE(xyz && out_idx, PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && k > 0, PCB_EINVAL);
    PCB_REQUIRE(k <= 64 && k <= N && B <= 65535, PCB_ERANGE);
    return launch_knn_xyz<MODE_CDIST>(xyz, B, N, k, 0, out_idx, out_dist, (cudaStream_t)stream);
}
