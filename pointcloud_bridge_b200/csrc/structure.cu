// BridgeStructureEncoding input rows in one kernel -- replaces, for every point of a cloud,
//   Highway_bridge/models/attention_modules.py:552-574  compute_absolute_position_encoding (sin / cos of the grid-snapped
//                                                       coordinates at F frequencies -> 6F values)
//   :590-597  neighbours - centre                       (rel_pos [B,N,k,3], k nearest neighbours from the cdist-kNN kernel)
//   :622-687  get_structure_features                    (13 statistics of rel_pos: eigenvalue shape features of the local
//                                                       covariance, distance max / mean / std to the neighbourhood centre,
//                                                       direction consistency, z std / range, mean offset, spread)
//   :603-613  expand + cat                              (every neighbour row = [abs enc | rel_pos | structure features])
// which the reference (and rounds 1-2 of this repo) evaluate as ~40 ATen kernels over [B,N,k,*] temporaries, among them a
// batched cuSOLVER eigensolver that synchronises the host.  Here: one CTA per 64 points; phase 1, one thread per point:
// gather the k neighbours, all statistics in registers (covariance accumulated in float64 and rounded once to fp32 like
// the matrix the reference hands to eigh; eigenvalues by the trigonometric closed form in float64, as eig3.cu), results
// into shared memory; phase 2, all threads: the k x (6F + 16) row block of the CTA's points written with coalesced
// 16-byte stores, fp32 or bf16 (the training MLP's operand type), zero padded to the row pitch.
// Nothing here is differentiable in the reference either (coordinates are inputs), so there is no backward.
#include <cuda_bf16.h>

#include "pcb_common.cuh"

namespace pcb {

constexpr int kStPoints = 64;              // points per CTA
constexpr int kStThreads = 256;
constexpr int kStMaxK = 32;
constexpr int kStMaxF = 8;

struct StructArgs {
    const float *xyz;                      // [B, N, 3]
    const int64_t *idx;                    // [B, N, k]
    void *rows;                            // [B * N * k, pitch] fp32 / bf16, may be NULL
    float *feat;                           // [B * N, 13] fp32, may be NULL
    float freqs[kStMaxF];
    float grid_size;
    int N, k, F, pitch, bf16;
    int64_t total;                         // B * N
};

// ascending eigenvalues of the symmetric 3x3 matrix given by its lower triangle (float64 closed form)
__device__ __forceinline__ void eig3_sym(double a00, double a10, double a11, double a20, double a21, double a22, double &e0,
                                         double &e1, double &e2)
{
    const double p1 = a10 * a10 + a20 * a20 + a21 * a21;
    if (p1 == 0.0) {
        e0 = fmin(a00, fmin(a11, a22));
        e2 = fmax(a00, fmax(a11, a22));
        e1 = a00 + a11 + a22 - e0 - e2;
        return;
    }
    const double q = (a00 + a11 + a22) / 3.0;
    const double b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
    const double p2 = b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * p1;
    const double p = sqrt(p2 / 6.0);
    const double ip = 1.0 / p;
    const double c00 = b00 * ip, c11 = b11 * ip, c22 = b22 * ip, c10 = a10 * ip, c20 = a20 * ip, c21 = a21 * ip;
    const double det = c00 * (c11 * c22 - c21 * c21) - c10 * (c10 * c22 - c21 * c20) + c20 * (c10 * c21 - c11 * c20);
    double r = 0.5 * det;
    r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
    const double phi = acos(r) / 3.0;
    e2 = q + 2.0 * p * cos(phi);
    e0 = q + 2.0 * p * cos(phi + 2.0943951023931954923);   // + 2 pi / 3
    e1 = 3.0 * q - e0 - e2;
}

__global__ void __launch_bounds__(kStThreads)
structure_rows_kernel(const StructArgs a)
{
    extern __shared__ float st_smem[];
    const int k = a.k, F = a.F, nf = 6 * F + 13;            // per-point values: abs enc, then the 13 statistics
    float *s_feat = st_smem;                                // [kStPoints][nf]
    float *s_rel = st_smem + kStPoints * nf;                // [kStPoints][3 k + 1] (odd stride: no bank conflicts)
    const int rstride = 3 * k + 1;
    pdl_wait();
    pdl_trigger();
    const int64_t p0 = (int64_t)blockIdx.x * kStPoints;
    const int tid = threadIdx.x;

    // ---- phase 1: one thread per point
    if (tid < kStPoints && p0 + tid < a.total) {
        const int64_t pt = p0 + tid;
        const int64_t b = pt / a.N;
        const float *cloud = a.xyz + b * a.N * 3;
        const float cx = a.xyz[pt * 3], cy = a.xyz[pt * 3 + 1], cz = a.xyz[pt * 3 + 2];
        float *f = s_feat + tid * nf;
        float *rel = s_rel + (size_t)tid * rstride;
        // absolute position encoding of the grid-snapped coordinates
        const float gx = floorf(cx / a.grid_size) * a.grid_size, gy = floorf(cy / a.grid_size) * a.grid_size,
                    gz = floorf(cz / a.grid_size) * a.grid_size;
        for (int i = 0; i < F; ++i) {
            const float w = a.freqs[i];
            f[6 * i + 0] = sinf(gx * w), f[6 * i + 1] = sinf(gy * w), f[6 * i + 2] = sinf(gz * w);
            f[6 * i + 3] = cosf(gx * w), f[6 * i + 4] = cosf(gy * w), f[6 * i + 5] = cosf(gz * w);
        }
        // neighbour offsets; sums for the centre, the covariance (float64) and the mean unit vector
        const int64_t *id = a.idx + pt * k;
        double c00 = 0, c10 = 0, c11 = 0, c20 = 0, c21 = 0, c22 = 0;
        float sx = 0.f, sy = 0.f, sz = 0.f, ux = 0.f, uy = 0.f, uz = 0.f, zmax = -INFINITY, zmin = INFINITY;
        for (int j = 0; j < k; ++j) {
            const int64_t q = id[j];
            const float dx = cloud[q * 3] - cx, dy = cloud[q * 3 + 1] - cy, dz = cloud[q * 3 + 2] - cz;
            rel[3 * j] = dx, rel[3 * j + 1] = dy, rel[3 * j + 2] = dz;
            c00 += (double)dx * dx, c10 += (double)dy * dx, c11 += (double)dy * dy;
            c20 += (double)dz * dx, c21 += (double)dz * dy, c22 += (double)dz * dz;
            sx += dx, sy += dy, sz += dz;
            const float inv = 1.f / (sqrtf(dx * dx + dy * dy + dz * dz) + 1e-8f);
            ux += dx * inv, uy += dy * inv, uz += dz * inv;
            zmax = fmaxf(zmax, dz), zmin = fminf(zmin, dz);
        }
        const float kf = (float)k, km1 = (float)(k - 1);
        // shape features from the eigenvalues of cov = rel^T rel / (k - 1), cov held in fp32 as the reference's is
        double e0, e1, e2;
        eig3_sym((double)((float)c00 / km1), (double)((float)c10 / km1), (double)((float)c11 / km1),
                 (double)((float)c20 / km1), (double)((float)c21 / km1), (double)((float)c22 / km1), e0, e1, e2);
        const float v0 = (float)e0, v1 = (float)e1, v2 = (float)e2, den = v0 + 1e-8f;
        float *s = f + 6 * F;
        s[0] = (v0 - v1) / den, s[1] = (v1 - v2) / den, s[2] = v2 / den;
        // distances to the neighbourhood centre: max, mean, unbiased std; per-axis unbiased std of the offsets
        const float mx = sx / kf, my = sy / kf, mz = sz / kf;
        float dmax = 0.f, dsum = 0.f, vx = 0.f, vy = 0.f, vz = 0.f;
        for (int j = 0; j < k; ++j) {
            const float ex = rel[3 * j] - mx, ey = rel[3 * j + 1] - my, ez = rel[3 * j + 2] - mz;
            const float d = sqrtf(ex * ex + ey * ey + ez * ez);
            dmax = j == 0 ? d : fmaxf(dmax, d);
            dsum += d;
            vx += ex * ex, vy += ey * ey, vz += ez * ez;
        }
        const float dmean = dsum / kf;
        float dvar = 0.f;
        for (int j = 0; j < k; ++j) {
            const float ex = rel[3 * j] - mx, ey = rel[3 * j + 1] - my, ez = rel[3 * j + 2] - mz;
            const float d = sqrtf(ex * ex + ey * ey + ez * ez) - dmean;
            dvar += d * d;
        }
        s[3] = dmax, s[4] = dmean, s[5] = sqrtf(dvar / km1);
        // mean over all k x k cosines between neighbour directions == |mean unit vector|^2
        const float wx = ux / kf, wy = uy / kf, wz = uz / kf;
        s[6] = wx * wx + wy * wy + wz * wz;
        const float stdx = sqrtf(vx / km1), stdy = sqrtf(vy / km1), stdz = sqrtf(vz / km1);
        s[7] = stdz, s[8] = zmax - zmin;
        s[9] = mx, s[10] = my, s[11] = mz;
        s[12] = sqrtf(stdx * stdx + stdy * stdy + stdz * stdz);
        if (a.feat) {
            float *o = a.feat + pt * 13;
#pragma unroll
            for (int i = 0; i < 13; ++i) o[i] = s[i];
        }
    }
    __syncthreads();
    if (!a.rows) return;

    // ---- phase 2: rows [abs enc (6F) | rel (3) | statistics (13) | 0 ...] of the CTA's points, 8 channels per store
    const int npts = a.total - p0 < kStPoints ? (int)(a.total - p0) : kStPoints;
    const int chunks = a.pitch >> 3, nabs = 6 * F, C = nabs + 16;
    const int64_t items = (int64_t)npts * k * chunks;
    for (int64_t e = tid; e < items; e += kStThreads) {
        const int ch = (int)(e % chunks);
        const int64_t r = e / chunks;                       // row inside the CTA: point * k + neighbour
        const int p = (int)(r / k), j = (int)(r - (int64_t)p * k);
        const float *f = s_feat + p * nf, *rel = s_rel + (size_t)p * rstride + 3 * j;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = ch * 8 + i;
            v[i] = c < nabs ? f[c] : (c < nabs + 3 ? rel[c - nabs] : (c < C ? f[c - 3] : 0.f));
        }
        const int64_t row = p0 * k + r;
        if (a.bf16) {
            unsigned w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                w[i] = *reinterpret_cast<unsigned *>(&t);
            }
            *reinterpret_cast<uint4 *>((__nv_bfloat16 *)a.rows + row * a.pitch + ch * 8) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            float4 *o = reinterpret_cast<float4 *>((float *)a.rows + row * a.pitch + ch * 8);
            o[0] = make_float4(v[0], v[1], v[2], v[3]);
            o[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
}

}  // namespace pcb

// xyz [B,N,3], idx [B,N,k] (k nearest neighbours of every point, int64, inside the cloud) ->
//   rows [B*N*k, pitch] (fp32 when out_bf16 == 0, else bf16; pitch >= 6F + 16, multiple of 8, pad columns zero):
//         [sin/cos position encoding (6F) | neighbour - centre (3) | 13 structure statistics]   (either may be NULL)
//   feat [B*N, 13] fp32: the statistics alone (get_structure_features)
PCB_API int pcb_structure_rows_f32(const float *xyz, const int64_t *idx, int B, int N, int k, const float *freqs, int F,
                                   float grid_size, int out_bf16, int pitch, void *rows, float *feat, pcb_stream_t stream)
{
    using namespace pcb;
    PCB_REQUIRE(xyz && idx && (rows || feat) && (F == 0 || freqs), PCB_EINVAL);
    PCB_REQUIRE(B > 0 && N > 0 && k >= 2 && k <= kStMaxK && F >= 0 && F <= kStMaxF && grid_size > 0.f, PCB_ERANGE);
    PCB_REQUIRE(!rows || (pitch >= 6 * F + 16 && pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(rows) & 15) == 0), PCB_ERANGE);
    StructArgs a = {};
    a.xyz = xyz, a.idx = idx, a.rows = rows, a.feat = feat, a.grid_size = grid_size;
    a.N = N, a.k = k, a.F = F, a.pitch = pitch, a.bf16 = out_bf16, a.total = (int64_t)B * N;
    for (int i = 0; i < F; ++i) a.freqs[i] = freqs[i];
    const size_t smem = (size_t)kStPoints * ((6 * F + 13) + 3 * k + 1) * sizeof(float);
    static bool attr_set[kMaxDevices] = {};
    if (smem > 48 * 1024)
        if (cudaError_t e = smem_optin_once(structure_rows_kernel, (int)(kStPoints * (6 * kStMaxF + 13 + 3 * kStMaxK + 1) * 4), attr_set))
            return (int)e;
    const int64_t grid = ceil_div(a.total, kStPoints);
    PCB_REQUIRE(grid < (1ll << 31), PCB_ERANGE);
    return (int)launch_pdl(structure_rows_kernel, dim3((unsigned)grid), dim3(kStThreads), smem, (cudaStream_t)stream, a);
}
