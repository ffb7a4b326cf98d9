// Eigenvalues of batched symmetric 3x3 matrices -- torch.linalg.eigvalsh(cov) of the local covariance in
// BridgeStructureEncoding.get_structure_features (Highway_bridge/models/attention_modules.py:628-640,
// SURVEY.md section 8f rank 3).  cuSOLVER's batched eigensolver behind torch.linalg.eigvalsh checks its
// `info` output on the host (a stream synchronisation), which keeps the BriStruNet training step out of a
// CUDA graph; this is the sync-free replacement used in training mode: one thread per matrix, the
// trigonometric closed form evaluated in float64 (absolute error ~1e-16 * |A|, far below the fp32 solver's
// own 1e-7 * |A|), eigenvalues ascending as eigvalsh returns them.  Only the lower triangle is read, as
// LAPACK's UPLO='L' default does.
#include "pcb_common.cuh"

namespace pcb {

__global__ void __launch_bounds__(256)
eigvalsh3_kernel(const float *__restrict__ a, int64_t M, float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= M) return;
    const float *m = a + i * 9;
    const double a00 = m[0], a11 = m[4], a22 = m[8], a10 = m[3], a20 = m[6], a21 = m[7];
    double e0, e1, e2;                                    // ascending
    const double p1 = a10 * a10 + a20 * a20 + a21 * a21;
    if (p1 == 0.0) {                                      // diagonal
        e0 = fmin(a00, fmin(a11, a22));
        e2 = fmax(a00, fmax(a11, a22));
        e1 = a00 + a11 + a22 - e0 - e2;
    } else {
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
    out[i * 3 + 0] = (float)e0;
    out[i * 3 + 1] = (float)e1;
    out[i * 3 + 2] = (float)e2;
}

}  // namespace pcb

PCB_API int pcb_eigvalsh3_f32(const float *a, int64_t M, float *out, pcb_stream_t stream)
{
    using namespace pcb;
    PCB_REQUIRE(a && out, PCB_EINVAL);
    PCB_REQUIRE(M > 0 && ceil_div(M, 256) < (1ll << 31), PCB_ERANGE);
    eigvalsh3_kernel<<<(unsigned)ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(a, M, out);
    PCB_RETURN_LAUNCH_STATUS();
}
