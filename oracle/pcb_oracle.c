/*
 * pcb_oracle.c -- CPU ORACLE for the sampling-and-grouping hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (pointcloud_bridge_b200/) never imports, links or calls anything in oracle/.
 *
 * It restates, in plain C, the arithmetic that the reference's pure-PyTorch functions
 * execute on CPU (torch 2.11 / ATen / MKL).  The reference has no tests or golden vectors
 * for this path, so the oracle is pinned against outputs of the
 * reference itself, run in the authoring container by tests/golden/make_golden.py and
 * committed under tests/golden/ (see tests/test_oracle_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference root).  All floating point is IEEE fp32, round-to-nearest, compiled with
 * -ffp-contract=off; fused multiply-adds appear only where written as fmaf().
 *
 * Build: make -C oracle   (gcc -O2 -mavx2 -mfma -fopenmp -ffp-contract=off)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------------------------
 * Row norms.
 * sum(v**2, -1) as ATen's CPU sum kernel evaluates it for a contiguous last dim:
 * each square is rounded on its own (pow(2) is a separate elementwise kernel), then the
 * inner-dim reduction runs 4 interleaved vector accumulators of VEC lanes over the
 * leading floor(D/VEC) vectors, folds them ((A0+A1)+A2)+A3, adds the scalar tail
 * left-to-right starting from 0, then the VEC lane partials left-to-right.
 * (aten/src/ATen/native/cpu/SumKernel.cpp: vectorized_inner_sum / row_sum; no cascade
 * level is reached below 16*4 vectors, i.e. D < 64*VEC.)
 * For D = 3 this is (x0^2 + x1^2) + x2^2.
 * Used by: Partsize-identical/models/pointnet_util.py:41-42,84;
 *          Highway_bridge/models/DGCNN.py:64; torch.cdist's _euclidean_dist.
 * VEC = 8 matches torch 2.11 CPU in the authoring container (pinned by golden vectors).
 * --------------------------------------------------------------------------------- */
#define ORC_VEC 8
#define ORC_ILP 4

static float orc_sumsq(const float *v, int64_t D, int64_t stride)
{
    float acc[ORC_ILP][ORC_VEC];
    memset(acc, 0, sizeof(acc));
    int64_t nvec = D / ORC_VEC;          /* full vectors */
    int64_t nilp = nvec / ORC_ILP;       /* rounds of 4 vectors */
    for (int64_t r = 0; r < nilp; ++r)
        for (int a = 0; a < ORC_ILP; ++a)
            for (int l = 0; l < ORC_VEC; ++l) {
                float x = v[((r * ORC_ILP + a) * ORC_VEC + l) * stride];
                float sq = x * x;
                acc[a][l] = acc[a][l] + sq;
            }
    for (int64_t i = nilp * ORC_ILP; i < nvec; ++i)   /* leftover vectors go to acc 0 */
        for (int l = 0; l < ORC_VEC; ++l) {
            float x = v[(i * ORC_VEC + l) * stride];
            float sq = x * x;
            acc[0][l] = acc[0][l] + sq;
        }
    for (int a = 1; a < ORC_ILP; ++a)
        for (int l = 0; l < ORC_VEC; ++l)
            acc[0][l] = acc[0][l] + acc[a][l];
    float fin = 0.0f;
    for (int64_t k = nvec * ORC_VEC; k < D; ++k) {    /* scalar tail */
        float x = v[k * stride];
        float sq = x * x;
        fin = fin + sq;
    }
    for (int l = 0; l < ORC_VEC; ++l)
        fin = fin + acc[0][l];
    return fin;
}

ORC_API void orc_row_sumsq(const float *x, int64_t rows, int64_t D, float *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r)
        out[r] = orc_sumsq(x + r * D, D, 1);
}

/* dot product as MKL sgemm evaluates it for these shapes: one sequential FMA chain,
 * k ascending, starting from 0 (pinned by golden vectors). */
static inline float orc_dot_chain(const float *a, const float *b, int64_t D)
{
    float acc = 0.0f;
    for (int64_t k = 0; k < D; ++k)
        acc = fmaf(a[k], b[k], acc);
    return acc;
}

/* ---------------------------------------------------------------------------------
 * square_distance(src, dst)            pointnet_util.py:22-43, pointnet2_utils.py:7-14
 *   dist  = -2 * matmul(src, dst^T)
 *   dist += sum(src**2,-1)[:, :, None]
 *   dist += sum(dst**2,-1)[:, None, :]
 * src [B,N,C], dst [B,M,C] -> out [B,N,M]
 * --------------------------------------------------------------------------------- */
ORC_API void orc_square_distance(const float *src, const float *dst, int64_t B, int64_t N,
                                 int64_t M, int64_t C, float *out)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t b = 0; b < B; ++b)
        for (int64_t n = 0; n < N; ++n) {
            const float *s = src + (b * N + n) * C;
            float sn = orc_sumsq(s, C, 1);
            float *o = out + (b * N + n) * M;
            for (int64_t m = 0; m < M; ++m) {
                const float *d = dst + (b * M + m) * C;
                float dn = orc_sumsq(d, C, 1);
                float t = -2.0f * orc_dot_chain(s, d, C);
                t = t + sn;
                t = t + dn;
                o[m] = t;
            }
        }
}

/* ---------------------------------------------------------------------------------
 * farthest_point_sample(xyz, npoint)   pointnet_util.py:66-88, pointnet2_utils.py:63-80
 * The start index per cloud is an input here: the reference draws it with
 * torch.randint on the CPU default generator (pointnet_util.py:79); the caller of the
 * oracle makes that same call.
 *   distance = 1e10; loop: emit far; d = sum((xyz - c)**2, -1); distance = min-by-<;
 *   far = first index of max(distance)
 * --------------------------------------------------------------------------------- */
ORC_API void orc_fps(const float *xyz, int64_t B, int64_t N, const int64_t *start,
                     int64_t npoint, int64_t *out)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < B; ++b) {
        const float *p = xyz + b * N * 3;
        float *dist = (float *)malloc(sizeof(float) * (size_t)N);
        for (int64_t n = 0; n < N; ++n) dist[n] = 1e10f;
        int64_t far = start[b];
        for (int64_t i = 0; i < npoint; ++i) {
            out[b * npoint + i] = far;
            float cx = p[far * 3 + 0], cy = p[far * 3 + 1], cz = p[far * 3 + 2];
            float best = -INFINITY;
            int64_t besti = 0;
            for (int64_t n = 0; n < N; ++n) {
                float dx = p[n * 3 + 0] - cx;
                float dy = p[n * 3 + 1] - cy;
                float dz = p[n * 3 + 2] - cz;
                float sx = dx * dx, sy = dy * dy, sz = dz * dz;
                float d = sx + sy;
                d = d + sz;
                if (d < dist[n]) dist[n] = d;
                if (dist[n] > best) { best = dist[n]; besti = n; }   /* first max wins */
            }
            far = besti;
        }
        free(dist);
    }
}

/* ---------------------------------------------------------------------------------
 * query_ball_point(radius, nsample, xyz, new_xyz)
 *                                      pointnet_util.py:91-112, pointnet2_utils.py:97-112
 * r2 is fp32(double(radius)**2), computed by the caller as Python does.
 * Row = ascending indices n with not (d > r2), first nsample; short rows padded with the
 * row's first index; empty row -> all N.
 * d = square_distance(new_xyz, xyz)[b, s, n].
 * --------------------------------------------------------------------------------- */
ORC_API void orc_ball_query(const float *xyz, const float *new_xyz, int64_t B, int64_t N,
                            int64_t S, float r2, int64_t nsample, int64_t *out)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t b = 0; b < B; ++b)
        for (int64_t s = 0; s < S; ++s) {
            const float *q = new_xyz + (b * S + s) * 3;
            float qn = orc_sumsq(q, 3, 1);
            int64_t *row = out + (b * S + s) * nsample;
            int64_t cnt = 0;
            for (int64_t n = 0; n < N && cnt < nsample; ++n) {
                const float *p = xyz + (b * N + n) * 3;
                float pn = orc_sumsq(p, 3, 1);
                float t = -2.0f * orc_dot_chain(q, p, 3);
                t = t + qn;
                t = t + pn;
                if (!(t > r2)) row[cnt++] = n;
            }
            int64_t first = cnt > 0 ? row[0] : N;
            for (int64_t k = cnt; k < nsample; ++k) row[k] = first;
        }
}

/* ---------------------------------------------------------------------------------
 * Ordered top-k of one row by (value, index) ascending.  The reference's sort/topk leave
 * the order of exactly-equal keys unspecified; the build's rule, and
 * therefore the oracle's, is lowest index first.
 * --------------------------------------------------------------------------------- */
static void orc_topk_row(const float *d, int64_t n, int64_t k, float *od, int64_t *oi)
{
    int64_t cnt = 0;
    for (int64_t j = 0; j < n; ++j) {
        float v = d[j];
        if (cnt == k && !(v < od[k - 1])) continue;       /* ties lose to earlier index */
        int64_t pos = cnt < k ? cnt : k - 1;
        while (pos > 0 && v < od[pos - 1]) {
            od[pos] = od[pos - 1];
            oi[pos] = oi[pos - 1];
            --pos;
        }
        od[pos] = v;
        oi[pos] = j;
        if (cnt < k) ++cnt;
    }
}

/* ---------------------------------------------------------------------------------
 * DGCNN.knn(x, k)                      Highway_bridge/models/DGCNN.py:49-70
 *   inner = -2 * matmul(x, x^T); xx = sum(x**2, dim=2)
 *   pd = xx + inner + xx^T ; idx = topk(-pd, k)      (self included, ascending pd)
 * x here is points-major [B,N,D] (the reference transposes [B,D,N] first, DGCNN.py:60).
 * out_dist may be NULL.
 * --------------------------------------------------------------------------------- */
ORC_API void orc_knn(const float *x, int64_t B, int64_t N, int64_t D, int64_t k,
                     int64_t *out_idx, float *out_dist)
{
    for (int64_t b = 0; b < B; ++b) {
        const float *xb = x + b * N * D;
        float *xx = (float *)malloc(sizeof(float) * (size_t)N);
        float *xt = (float *)malloc(sizeof(float) * (size_t)(N * D));   /* [D][N] */
        for (int64_t n = 0; n < N; ++n) {
            xx[n] = orc_sumsq(xb + n * D, D, 1);
            for (int64_t c = 0; c < D; ++c) xt[c * N + n] = xb[n * D + c];
        }
#pragma omp parallel
        {
            float *acc = (float *)malloc(sizeof(float) * (size_t)N);
            float *od = (float *)malloc(sizeof(float) * (size_t)k);
            int64_t *oi = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
#pragma omp for schedule(static)
            for (int64_t i = 0; i < N; ++i) {
                const float *xi = xb + i * D;
                for (int64_t j = 0; j < N; ++j) acc[j] = 0.0f;
                for (int64_t c = 0; c < D; ++c) {
                    float a = xi[c];
                    const float *row = xt + c * N;
                    for (int64_t j = 0; j < N; ++j) acc[j] = fmaf(a, row[j], acc[j]);
                }
                float xi2 = xx[i];
                for (int64_t j = 0; j < N; ++j) {
                    float t = -2.0f * acc[j];
                    t = xi2 + t;
                    t = t + xx[j];
                    acc[j] = t;
                }
                orc_topk_row(acc, N, k, od, oi);
                memcpy(out_idx + (b * N + i) * k, oi, sizeof(int64_t) * (size_t)k);
                if (out_dist) memcpy(out_dist + (b * N + i) * k, od, sizeof(float) * (size_t)k);
            }
            free(acc); free(od); free(oi);
        }
        free(xx); free(xt);
    }
}

/* ---------------------------------------------------------------------------------
 * cdist-kNN                            Highway_bridge/models/attention_modules.py:584-586,736-738
 *   dist = torch.cdist(xyz, xyz); _, idx = dist.topk(k, largest=False)
 * torch.cdist (default compute mode, > 25 rows) = _euclidean_dist:
 *   sqrt(clamp_min(matmul([-2x, |x|^2, 1], [y, 1, |y|^2]^T), 0))   (K = 5 FMA chain)
 * The matmul + clamp part is reproduced bit for bit.  The square root is IEEE
 * round-to-nearest here (and in the CUDA kernels); torch 2.11's CPU sqrt goes through
 * MKL VML and differs from IEEE in the last bit for ~0.6 % of inputs (measured), so against
 * the reference the sqrt'd values are compared to 1 ulp and indices may differ only between
 * candidates whose distances are within 1 ulp of each other (tests/parity.py).
 * out_sq (optional) receives the pre-sqrt clamped values of the selected neighbours.
 * --------------------------------------------------------------------------------- */
ORC_API void orc_knn_cdist(const float *xyz, int64_t B, int64_t N, int64_t k,
                           int64_t *out_idx, float *out_dist, float *out_sq)
{
    for (int64_t b = 0; b < B; ++b) {
        const float *p = xyz + b * N * 3;
        float *nn = (float *)malloc(sizeof(float) * (size_t)N);
        for (int64_t n = 0; n < N; ++n) nn[n] = orc_sumsq(p + n * 3, 3, 1);
#pragma omp parallel
        {
            float *sq = (float *)malloc(sizeof(float) * (size_t)N);
            float *acc = (float *)malloc(sizeof(float) * (size_t)N);
            float *od = (float *)malloc(sizeof(float) * (size_t)k);
            int64_t *oi = (int64_t *)malloc(sizeof(int64_t) * (size_t)k);
#pragma omp for schedule(static)
            for (int64_t i = 0; i < N; ++i) {
                float a0 = -2.0f * p[i * 3 + 0], a1 = -2.0f * p[i * 3 + 1], a2 = -2.0f * p[i * 3 + 2];
                float a3 = nn[i];
                for (int64_t j = 0; j < N; ++j) {
                    float t = fmaf(a0, p[j * 3 + 0], 0.0f);
                    t = fmaf(a1, p[j * 3 + 1], t);
                    t = fmaf(a2, p[j * 3 + 2], t);
                    t = fmaf(a3, 1.0f, t);
                    t = fmaf(1.0f, nn[j], t);
                    t = t < 0.0f ? 0.0f : t;
                    sq[j] = t;
                    acc[j] = sqrtf(t);
                }
                orc_topk_row(acc, N, k, od, oi);
                memcpy(out_idx + (b * N + i) * k, oi, sizeof(int64_t) * (size_t)k);
                if (out_dist) memcpy(out_dist + (b * N + i) * k, od, sizeof(float) * (size_t)k);
                if (out_sq)
                    for (int64_t j = 0; j < k; ++j) out_sq[(b * N + i) * k + j] = sq[oi[j]];
            }
            free(acc); free(od); free(oi); free(sq);
        }
        free(nn);
    }
}

/* ---------------------------------------------------------------------------------
 * k nearest of xyz2 for every xyz1 point (k = 3, or 4 in EnhancedFeaturePropagation)
 *   pointnet_util.py:325-328, pointnet2_utils.py:183-186,253-256
 *   dists = square_distance(xyz1, xyz2); dists, idx = dists.sort(-1); take first k
 * xyz1 [B,N,3], xyz2 [B,S,3] -> dist [B,N,k], idx [B,N,k]
 * --------------------------------------------------------------------------------- */
ORC_API void orc_three_nn(const float *xyz1, const float *xyz2, int64_t B, int64_t N,
                          int64_t S, int64_t k, float *out_dist, int64_t *out_idx)
{
#pragma omp parallel
    {
        float *row = (float *)malloc(sizeof(float) * (size_t)S);
#pragma omp for collapse(2) schedule(static)
        for (int64_t b = 0; b < B; ++b)
            for (int64_t n = 0; n < N; ++n) {
                const float *q = xyz1 + (b * N + n) * 3;
                float qn = orc_sumsq(q, 3, 1);
                for (int64_t s = 0; s < S; ++s) {
                    const float *p = xyz2 + (b * S + s) * 3;
                    float pn = orc_sumsq(p, 3, 1);
                    float t = -2.0f * orc_dot_chain(q, p, 3);
                    t = t + qn;
                    t = t + pn;
                    row[s] = t;
                }
                orc_topk_row(row, S, k, out_dist + (b * N + n) * k, out_idx + (b * N + n) * k);
            }
        free(row);
    }
}

/* ---------------------------------------------------------------------------------
 * Inverse-distance weights and interpolation   pointnet_util.py:330-334
 *   dist_recip = 1.0 / (dists + 1e-8); norm = sum(dist_recip, 2); weight = dist_recip / norm
 *   out = sum(index_points(points2, idx) * weight[..., None], dim=2)
 * points2 [B,S,D] (points-major), idx/dist [B,N,k] -> weight [B,N,k], out [B,N,D]
 * --------------------------------------------------------------------------------- */
ORC_API void orc_interp_weights(const float *dist, int64_t rows, int64_t k, float *weight)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r) {
        float rec[8];
        float norm = 0.0f;
        for (int64_t j = 0; j < k; ++j) {
            float t = dist[r * k + j] + 1e-8f;
            rec[j] = 1.0f / t;
            norm = (j == 0) ? rec[0] : norm + rec[j];
        }
        for (int64_t j = 0; j < k; ++j) weight[r * k + j] = rec[j] / norm;
    }
}

ORC_API void orc_three_interpolate(const float *points2, const int64_t *idx, const float *weight,
                                   int64_t B, int64_t N, int64_t S, int64_t D, int64_t k,
                                   float *out)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t b = 0; b < B; ++b)
        for (int64_t n = 0; n < N; ++n) {
            const int64_t *id = idx + (b * N + n) * k;
            const float *w = weight + (b * N + n) * k;
            float *o = out + (b * N + n) * D;
            for (int64_t c = 0; c < D; ++c) {
                float acc = 0.0f;
                for (int64_t j = 0; j < k; ++j) {
                    float t = points2[(b * S + id[j]) * D + c] * w[j];
                    acc = (j == 0) ? t : acc + t;
                }
                o[c] = acc;
            }
        }
}

/* ---------------------------------------------------------------------------------
 * index_points(points, idx)            pointnet_util.py:46-63 (raises on idx >= N)
 *                                      pointnet2_utils.py:17-39 (clamps to [0, N-1])
 * points [B,N,C], idx [B,M] -> out [B,M,C].  Returns the number of out-of-range indices
 * seen (the Partsize flavour turns a non-zero count into IndexError); with clamp != 0
 * indices are clamped instead and 0 is returned.
 * --------------------------------------------------------------------------------- */
ORC_API int64_t orc_index_points(const float *points, const int64_t *idx, int64_t B, int64_t N,
                                 int64_t C, int64_t M, int clamp, float *out)
{
    int64_t bad = 0;
    for (int64_t b = 0; b < B; ++b)
        for (int64_t m = 0; m < M; ++m) {
            int64_t i = idx[b * M + m];
            if (clamp) { if (i < 0) i = 0; if (i > N - 1) i = N - 1; }
            else if (i < -N || i >= N) { ++bad; continue; }
            else if (i < 0) i += N;                      /* python negative indexing */
            memcpy(out + (b * M + m) * C, points + (b * N + i) * C, sizeof(float) * (size_t)C);
        }
    return bad;
}

/* ---------------------------------------------------------------------------------
 * get_graph_feature(x, k, idx)         Highway_bridge/models/DGCNN.py:72-109
 * x [B,D,N] channels-first, idx [B,N,k] -> out [B,2D,N,k] = cat(nbr - ctr, ctr)
 * --------------------------------------------------------------------------------- */
ORC_API void orc_graph_feature(const float *x, const int64_t *idx, int64_t B, int64_t D,
                               int64_t N, int64_t k, float *out)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t b = 0; b < B; ++b)
        for (int64_t c = 0; c < D; ++c) {
            const float *xc = x + (b * D + c) * N;
            float *o1 = out + ((b * 2 * D + c) * N) * k;
            float *o2 = out + ((b * 2 * D + D + c) * N) * k;
            for (int64_t n = 0; n < N; ++n)
                for (int64_t j = 0; j < k; ++j) {
                    int64_t nb = idx[(b * N + n) * k + j];
                    o1[n * k + j] = xc[nb] - xc[n];
                    o2[n * k + j] = xc[n];
                }
        }
}

/* ---------------------------------------------------------------------------------
 * Grouping of sample_and_group after FPS + ball query   pointnet_util.py:137-147
 *   grouped_xyz_norm = index_points(xyz, idx) - new_xyz[:, :, None]
 *   new_points = cat([grouped_xyz_norm, index_points(points, idx)], -1)   (xyz_first = 1)
 * PointNetSetAbstractionMsg uses cat([points, xyz_norm]) (pointnet_util.py:265)  (xyz_first = 0)
 * xyz [B,N,3], points [B,N,D] or NULL, new_xyz [B,S,3], idx [B,S,K] -> out [B,S,K,3+D]
 * Indices are clamped when clamp != 0 (Highway flavour).
 * --------------------------------------------------------------------------------- */
ORC_API void orc_group_points(const float *xyz, const float *points, const float *new_xyz,
                              const int64_t *idx, int64_t B, int64_t N, int64_t S, int64_t K,
                              int64_t D, int xyz_first, int clamp, float *out)
{
    int64_t C = 3 + (points ? D : 0);
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t b = 0; b < B; ++b)
        for (int64_t s = 0; s < S; ++s)
            for (int64_t j = 0; j < K; ++j) {
                int64_t i = idx[(b * S + s) * K + j];
                if (clamp) { if (i < 0) i = 0; if (i > N - 1) i = N - 1; }
                float *o = out + ((b * S + s) * K + j) * C;
                float *ox = (xyz_first || !points) ? o : o + D;
                float *of = xyz_first ? o + 3 : o;
                for (int c = 0; c < 3; ++c)
                    ox[c] = xyz[(b * N + i) * 3 + c] - new_xyz[(b * S + s) * 3 + c];
                if (points)
                    memcpy(of, points + (b * N + i) * D, sizeof(float) * (size_t)D);
            }
}

ORC_API int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* Distances of explicit (row, idx) pairs under the oracle's arithmetic, for the tie-aware
 * comparison in tests/parity.py.  mode 0: DGCNN pairwise distance (orc_knn); mode 1: cdist
 * (orc_knn_cdist, D must be 3); mode 2: cdist before the square root.  x [B,N,D] points-major, idx [B,N,k] -> out [B,N,k]. */
ORC_API void orc_pair_dist(const float *x, const int64_t *idx, int64_t B, int64_t N, int64_t D,
                           int64_t k, int mode, float *out)
{
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t b = 0; b < B; ++b)
        for (int64_t i = 0; i < N; ++i) {
            const float *xi = x + (b * N + i) * D;
            float ni = orc_sumsq(xi, D, 1);
            for (int64_t j = 0; j < k; ++j) {
                const float *xj = x + (b * N + idx[(b * N + i) * k + j]) * D;
                float nj = orc_sumsq(xj, D, 1);
                float t;
                if (mode == 0) {
                    t = -2.0f * orc_dot_chain(xi, xj, D);
                    t = ni + t;
                    t = t + nj;
                } else {
                    t = fmaf(-2.0f * xi[0], xj[0], 0.0f);
                    t = fmaf(-2.0f * xi[1], xj[1], t);
                    t = fmaf(-2.0f * xi[2], xj[2], t);
                    t = fmaf(ni, 1.0f, t);
                    t = fmaf(1.0f, nj, t);
                    t = t < 0.0f ? 0.0f : t;
                    if (mode == 1) t = sqrtf(t);
                }
                out[(b * N + i) * k + j] = t;
            }
        }
}
