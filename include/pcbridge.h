/*
 * pcbridge.h -- C ABI of libpcbridge.so: sm_100a CUDA kernels for the sampling-and-grouping
 * hot path of the bridge point-cloud segmentation networks.
 *
 * The reference (UT-Team-Chun/Pointcloud-bridge) has no FFI layer: the hot path is a set of
 * module-level Python functions executed by ATen.  Each entry point below replaces one of
 * those functions (cited as reference file:line, relative to the reference root) and is what
 * a binding of that function would call.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name says host; buffers are dense,
 *     row-major, in the layout written next to the argument;
 *   - coordinates/features are fp32, indices are int64 (the reference returns LongTensors);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and the call
 *     returns without synchronising; the library keeps no state between calls;
 *   - return value: 0 on success, a positive cudaError_t if the launch failed, or a
 *     negative PCB_E* code for arguments outside the supported envelope.  No CPU fallback
 *     exists: without a CUDA device every call fails.
 *   - callable from any host thread.
 */
#ifndef PCBRIDGE_H_
#define PCBRIDGE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCB_VERSION 100

#define PCB_EINVAL (-1)      /* null pointer / non-positive size */
#define PCB_ERANGE (-2)      /* size outside the supported envelope (see each function) */
#define PCB_EALIGN (-3)      /* pointer not aligned as required */

typedef void *pcb_stream_t;

int pcb_version(void);
/* static string for a return code of any function below */
const char *pcb_error_string(int code);

/* ---- a1  farthest_point_sample            Partsize-identical/models/pointnet_util.py:66-88
 *                                           Highway_bridge/models/pointnet2_utils.py:63-80
 * xyz [B,N,3]; start [B] = the torch.randint draw of pointnet_util.py:79 (made by the caller);
 * out_idx [B,npoint].  One persistent CTA per cloud.  N <= 49152. */
int pcb_fps_f32(const float *xyz, int B, int N, const int64_t *start, int npoint,
                int64_t *out_idx, pcb_stream_t stream);

/* ---- a2  square_distance                  pointnet_util.py:22-43; pointnet2_utils.py:7-14
 * src [B,N,C], dst [B,M,C] -> out [B,N,M].  C <= 512. */
int pcb_square_distance_f32(const float *src, const float *dst, int B, int N, int M, int C,
                            float *out, pcb_stream_t stream);

/* ---- a3  query_ball_point                 pointnet_util.py:91-112; pointnet2_utils.py:97-112
 * xyz [B,N,3], new_xyz [B,S,3]; radius2 = (float)((double)radius * radius) as the Python
 * comparison sees it; out_idx [B,S,nsample].  Row = first nsample in-ball indices in
 * ascending order, padded with the first one; empty ball -> N. */
int pcb_ball_query_f32(const float *xyz, const float *new_xyz, int B, int N, int S,
                       float radius2, int nsample, int64_t *out_idx, pcb_stream_t stream);
/* the same for `nscales` (<= 4) radii around the same centroids in one scan (multi-scale grouping,
 * pointnet_util.py:241-284; pointnet2_utils.py:326-360): radius2 / nsample / out_idx are HOST arrays of
 * length nscales, out_idx[k] a device pointer to [B,S,nsample[k]].  Results equal nscales separate calls. */
int pcb_ball_query_multi_f32(const float *xyz, const float *new_xyz, int B, int N, int S, int nscales,
                             const float *radius2, const int *nsample, int64_t *const *out_idx, pcb_stream_t stream);

/* ---- a4  index_points                     pointnet_util.py:46-63 (errors on idx >= N)
 *                                           pointnet2_utils.py:17-39 (clamps to [0,N-1])
 * points [B,N,C], idx [B,M] -> out [B,M,C].  clamp != 0: Highway flavour.  clamp == 0:
 * out-of-range entries (after Python negative-index wrap) write zeros and, when
 * err_count != NULL, atomically increment *err_count (device int32) so the host side can
 * raise IndexError as the reference does. */
int pcb_gather_f32(const float *points, const int64_t *idx, int B, int N, int C, int64_t M,
                   int clamp, float *out, int *err_count, pcb_stream_t stream);
/* backward of pcb_gather_f32 w.r.t. points: grad_points [B,N,C] += scatter(grad_out [B,M,C]).
 * grad_points must be zero-initialised (or hold a value to accumulate into). */
int pcb_gather_bwd_f32(const float *grad_out, const int64_t *idx, int B, int N, int C, int64_t M,
                       int clamp, float *grad_points, pcb_stream_t stream);

/* ---- a5  grouping step of sample_and_group       pointnet_util.py:137-147, 260-267
 *                                                  pointnet2_utils.py:50-58, 342-349
 * xyz [B,N,3], points [B,N,D] or NULL (D = 0), new_xyz [B,S,3], idx [B,S,K]
 * -> out [B,S,K,3+D] = cat(xyz[idx] - new_xyz, points[idx]) when xyz_first != 0
 *                      cat(points[idx], xyz[idx] - new_xyz) otherwise (MSG order).
 * points may be channels-first ([B,D,N]) when points_cf != 0, saving the caller a transpose.
 * pitch = row pitch of out in elements: 0 or 3+D for the dense layout; a larger value (e.g. 3+D
 * rounded up to 8) appends zero columns so that the consuming GEMM sees 16-byte aligned rows. */
int pcb_group_points_f32(const float *xyz, const float *points, const float *new_xyz,
                         const int64_t *idx, int B, int N, int S, int K, int D, int xyz_first,
                         int points_cf, int clamp, int pitch, float *out, pcb_stream_t stream);
/* backward w.r.t. points: grad_points ([B,N,D] or [B,D,N]) += scatter(grad_out[..., feature part]) */
int pcb_group_points_bwd_f32(const float *grad_out, const int64_t *idx, int B, int N, int S, int K,
                             int D, int xyz_first, int points_cf, int clamp, int pitch, float *grad_points,
                             pcb_stream_t stream);
/* bf16 variants: the grouped tensor (and its gradient) in bf16, as the autocast GEMM consumes it;
 * xyz and grad_points stay fp32; `points` is fp32, or bf16 (points_bf16 != 0: the previous layer's
 * autocast output, taken as is -- needs pitch % 8 == 0) */
int pcb_group_points_bf16(const float *xyz, const void *points, int points_bf16, const float *new_xyz,
                          const int64_t *idx, int B, int N, int S, int K, int D, int xyz_first,
                          int points_cf, int clamp, int pitch, void *out, pcb_stream_t stream);
int pcb_group_points_bwd_bf16(const void *grad_out, const int64_t *idx, int B, int N, int S, int K,
                              int D, int xyz_first, int points_cf, int clamp, int pitch, float *grad_points,
                              pcb_stream_t stream);

/* ---- a7  k nearest of xyz2 for each xyz1 point + inverse-distance weights
 *          pointnet_util.py:325-332; pointnet2_utils.py:183-191 (k=3), 253-262 (k=4)
 * xyz1 [B,N,3], xyz2 [B,S,3] -> out_dist [B,N,k] (squared, ascending, ties by index),
 * out_idx [B,N,k], out_weight [B,N,k] (may be NULL).  1 <= k <= 8, k <= S. */
int pcb_three_nn_f32(const float *xyz1, const float *xyz2, int B, int N, int S, int k,
                     float *out_dist, int64_t *out_idx, float *out_weight, pcb_stream_t stream);
/* ---- a7  interpolation     pointnet_util.py:333-334; pointnet2_utils.py:196, 264-268
 * channels_first == 0: points2 [B,S,D] -> out [B,N,D];  != 0: points2 [B,D,S] -> out [B,D,N]. */
int pcb_interpolate_f32(const float *points2, const int64_t *idx, const float *weight, int B, int N,
                        int S, int D, int k, int channels_first, float *out, pcb_stream_t stream);
int pcb_interpolate_bwd_f32(const float *grad_out, const int64_t *idx, const float *weight, int B,
                            int N, int S, int D, int k, int channels_first, float *grad_points2,
                            pcb_stream_t stream);

/* a7 fused for training under bf16 autocast: the input rows of a feature-propagation MLP in one pass,
 *          pointnet_util.py:325-340 (interpolate, torch.cat([points1, interpolated]), cast)
 * points1 [B,N,D1] (fp32 or bf16; NULL when D1 == 0), points2 [B,S,D2] (fp32 or bf16), idx/weight [B,N,k]
 * -> out [B,N,pitch] bf16 = [points1 | sum_j weight_j * points2[idx_j] | zeros]; D1, D2, pitch even.
 * backward w.r.t. points2 (zero-initialised fp32 [B,S,D2]); the gradient of points1 is grad_out[..., :D1]. */
int pcb_fp_concat_bf16(const void *points1, int p1_bf16, const void *points2, int p2_bf16, const int64_t *idx,
                       const float *weight, int B, int N, int S, int D1, int D2, int k, int pitch, void *out,
                       pcb_stream_t stream);
int pcb_fp_concat_bwd_bf16(const void *grad_out, const int64_t *idx, const float *weight, int B, int N, int S,
                           int D1, int D2, int k, int pitch, float *grad_points2, pcb_stream_t stream);

/* ---- a8  DGCNN.knn                        Highway_bridge/models/DGCNN.py:49-70
 * x [B,D,N] (channels_first != 0, as DGCNN passes it) or [B,N,D]; out_idx [B,N,k] ordered by
 * (pairwise distance, index), self included; out_dist [B,N,k] may be NULL.  1 <= k <= 64,
 * k <= N, D <= 512. */
int pcb_knn_f32(const float *x, int B, int N, int D, int k, int channels_first, int64_t *out_idx,
                float *out_dist, pcb_stream_t stream);

/* ---- a10 cdist + topk(largest=False)      Highway_bridge/models/attention_modules.py:584-586,
 *                                           736-738
 * xyz [B,N,3]; out_idx [B,N,k]; out_dist [B,N,k] = Euclidean (sqrt'd) distances, may be NULL. */
int pcb_knn_cdist_f32(const float *xyz, int B, int N, int k, int64_t *out_idx, float *out_dist,
                      pcb_stream_t stream);

/* ---- a9  DGCNN.get_graph_feature          Highway_bridge/models/DGCNN.py:72-109
 * x [B,D,N], idx [B,N,k] -> out [B,2D,N,k] = cat(x[idx] - x, x) on the channel axis. */
int pcb_graph_feature_f32(const float *x, const int64_t *idx, int B, int D, int N, int k,
                          float *out, pcb_stream_t stream);
/* backward w.r.t. x: grad_x [B,D,N] (zero-initialised by the caller) */
int pcb_graph_feature_bwd_f32(const float *grad_out, const int64_t *idx, int B, int D, int N, int k,
                              float *grad_x, pcb_stream_t stream);

/* ---- a6/a11 (elementwise half)  BatchNorm(train) + ReLU (+ max over nsample) on point-major rows
 *          pointnet_util.py:213-217, 273-279, 343-345; pointnet2_utils.py:150-154, 353-356
 * Activations y [M,C] are fp32 (dtype 0) or bf16 (dtype 1); C % 4 == 0; statistics fp32.
 * The 1x1-conv bias is NOT added to y: BN(y + b) == BN(y) for batch statistics, b only enters
 * the running mean -- so no bias-add / bias-gradient pass exists.
 * Each direction is ONE persistent cooperative kernel (column sums -> grid barrier -> fold ->
 * grid barrier -> elementwise pass; the second read of y comes from L2).
 * `work` is a caller-provided fp32 scratch of pcb_bn_work_floats(C) = 3*C*(1+296) floats: the first
 * 3*C floats receive results, the rest holds per-CTA partial sums (deterministic two-stage reduction).
 *   pcb_bn_fwd_rows : mean/invstd [C] of y (batch statistics, biased variance, eps); running_mean /
 *                     running_var update with `momentum`, unbiased variance and the conv `bias` added
 *                     to the mean (all three may be NULL);
 *                     out[r] = max_{k<pool_k} act((y[r*pool_k+k]-mean)*invstd*gamma+beta), out is
 *                     [M/pool_k, C]; argmax [M/pool_k, C] uint8 = winning k, first on ties (may be
 *                     NULL when pool_k == 1); M % pool_k == 0, pool_k <= 255
 *   pcb_bn_bwd_rows : gy = gamma*invstd*(dy - sum_dy/M - yhat*sum_dy_yhat/M) with dy = gz masked by the
 *                     ReLU / argmax of the forward pass, gz is [M/pool_k, C];
 *                     work[0:C] = sum dy (grad beta), work[C:2C] = sum dy*yhat (grad gamma),
 *                     work[2C:3C] = gradient of the folded conv bias
 * C is the row pitch, Cv <= C the number of real channels: columns Cv..C-1 of y are zero padding that keeps
 * rows 16-byte aligned (196 -> 200 channels in the MSG network); parameter arrays have Cv entries, mean /
 * invstd / work are sized for C, and the padded columns of out / gy are written as zeros.
 * out_pitch / gz_pitch (elements, 0 = C): `out` resp. `gz` may be a column slice of a wider row-major buffer -- the
 * scales of a multi-scale module write straight into their concatenated output and read their slice of its gradient. */
int64_t pcb_bn_work_floats(int C);
/* SMs the cooperative BN grids are sized for from now on (0 = all): lets a step runner keep a few SMs free for a
 * long-running kernel of another stream (FPS of the next batch) that would otherwise stall every cooperative launch. */
int pcb_bn_set_coop_sms(int sms);
int pcb_bn_fwd_rows(const void *y, int dtype, int64_t M, int C, int Cv, int pool_k, const float *bias, const float *gamma,
                    const float *beta, float eps, float momentum, float *running_mean, float *running_var, int relu,
                    float *mean, float *invstd, void *out, int64_t out_pitch, unsigned char *argmax, float *work,
                    pcb_stream_t stream);
int pcb_bn_bwd_rows(const void *gz, int64_t gz_pitch, const void *y, const unsigned char *argmax, int dtype, int64_t M,
                    int C, int Cv, int pool_k, const float *mean, const float *invstd, const float *gamma, const float *beta,
                    int relu, float *work, void *gy, pcb_stream_t stream);

/* ---- a11 (training, backward of the 1x1 convolutions on rows)
 *          pointnet_util.py:213-215, 273-277, 343-345 (Conv2d / Conv1d 1x1 weight gradient)
 * gw[n, k] += sum_r gy[r, n] * x[r, k]:  gy [M, ldgy] and x [M, ldx] bf16 row-major, ldgy % 8 == ldx % 8 == 0,
 * zero pad columns beyond N resp. K, gw [N, ldw] fp32 -- ACCUMULATED into (zero it for a plain
 * gradient, or point it at a gradient bucket).  One pass over gy and x per 64x64 output tile, fp32
 * accumulation, atomics only for the per-CTA partial results. */
int pcb_wgrad_rows_bf16(const void *gy, const void *x, int64_t M, int N, int K, int ldgy, int ldx, float *gw,
                        int ldw, pcb_stream_t stream);

/* ---- section 8f rank 2: whole-scene tiling and vote scatter-back
 *          ScannetDatasetWholeScene.__getitem__  Highway_bridge/utils/BridgeDataLoader.py:214-277
 *          add_vote / argmax                     Partsize-identical/test_sem_seg.py:58-65, 162
 * points [P, point_stride] fp32 rows (x, y, z, r, g, b, ...).  Windows: index (iy, ix), closed membership
 * intervals lo_x/hi_x [grid_x], lo_y/hi_y [grid_y] (device arrays of doubles, padding included, computed by
 * the host in float64 as the reference does); x0, y0, stride locate the candidate windows of a point and
 * `reach` = ceil(block_size / stride) + 1 bounds how far below floor((x - x0) / stride) they can start.
 *   count : counts[iy*grid_x + ix] += 1 per member point (counts zeroed by the caller)
 *   fill  : keys[offsets[w] + k] = (w << 47) | (hash16(seed, point, w) << 31) | point index, k handed out through
 *           cursor[w] (zeroed); offsets = exclusive prefix sum of counts.  The caller sorts `keys` (ascending): every
 *           window's members are then in a pseudo-random order that does not depend on the order of the atomics;
 *           members = keys & 0x7fffffff.  grid_x * grid_y < 65536.
 *   blocks: the padded window has tot = ceil(n / block_points) * block_points entries (n = blk_cnt[b]); entry
 *           q = blk_first[b] + j of it takes slot s = (1000003 q + seed % 999983) mod tot of the member segment at
 *           blk_off[b], slots >= n wrapping to (s - n) mod n (the reference's np.random.choice padding + shuffle,
 *           BridgeDataLoader.py:239-242, from a counter-based hash) -> data [nblocks, block_points, 9] = (x - cx,
 *           y - cy, z, r, g, b, x/ext_x, y/ext_y, z/ext_z) computed in double and rounded to fp32,
 *           point_idx [nblocks, block_points]
 *   vote  : pool[point_idx[t] * num_classes + pred[t]] += 1;  vote_argmax: first maximum per point */
int pcb_scene_window_count_f32(const float *points, int64_t P, int point_stride, int grid_x, int grid_y,
                               const double *lo_x, const double *hi_x, const double *lo_y, const double *hi_y, double x0,
                               double y0, double stride, int reach, int *counts, pcb_stream_t stream);
int pcb_scene_window_fill_f32(const float *points, int64_t P, int point_stride, int grid_x, int grid_y,
                              const double *lo_x, const double *hi_x, const double *lo_y, const double *hi_y, double x0,
                              double y0, double stride, int reach, const int64_t *offsets, int *cursor, unsigned seed,
                              int64_t *keys, pcb_stream_t stream);
int pcb_scene_blocks_f32(const float *points, int point_stride, const int *members, const int64_t *blk_off,
                         const int *blk_cnt, const int64_t *blk_first, const double *blk_center, int64_t nblocks,
                         int block_points, double ext_x, double ext_y, double ext_z, unsigned seed, float *data,
                         int64_t *point_idx, pcb_stream_t stream);
int pcb_scene_vote(const int64_t *point_idx, const unsigned char *pred, int64_t total, int64_t P, int num_classes,
                   int *pool, pcb_stream_t stream);
int pcb_scene_vote_argmax(const int *pool, int64_t P, int num_classes, unsigned char *labels, pcb_stream_t stream);

/* ---- section 8f rank 3: eigenvalues of batched symmetric 3x3 matrices (BriStruNet structure statistics)
 *          torch.linalg.eigvalsh(cov), Highway_bridge/models/attention_modules.py:628-640
 * a [M,3,3] fp32 (lower triangle read), out [M,3] fp32 ascending; float64 closed form, no host synchronisation
 * (cuSOLVER's batched solver checks `info` on the host, which prevents CUDA-graph capture of the train step). */
int pcb_eigvalsh3_f32(const float *a, int64_t M, float *out, pcb_stream_t stream);

/* ---- section 8f rank 4 (training runner): mean NLL of the segmentation head on logits rows
 *          pointnet2_sem_seg.py:46-47, 56 (F.log_softmax + F.nll_loss, mean, no class weights)
 * logits [M, pitch] fp32 (dtype 0) or bf16 (1), `classes` (<= 32) real columns, `bias` [classes] fp32 added to them
 * (NULL: none -- the classifier GEMM runs bias-free); labels [M] int64.
 * fwd: partial[b] (b < pcb_nll_rows_blocks(M)) = sum of (logsumexp(x) - x[label]) over the rows of CTA b;
 *      loss = sum(partial) / M (reduced by the caller in a fixed order).
 * bwd: grad_logits[r, c] = (softmax(x[r])[c] - [c == label]) * grad_loss[0] / M, pad columns 0;
 *      grad_bias[c] += column sums of grad_logits (may be NULL). */
int pcb_nll_rows_blocks(int64_t M);
int pcb_nll_rows_fwd(const void *logits, int dtype, const float *bias, const int64_t *labels, int64_t M, int classes,
                     int pitch, float *partial, pcb_stream_t stream);
int pcb_nll_rows_bwd(const void *logits, int dtype, const float *bias, const int64_t *labels, int64_t M, int classes,
                     int pitch, const float *grad_loss, void *grad_logits, float *grad_bias, pcb_stream_t stream);

/* ---- section 8f rank 4 (training runner): Adam over one flat fp32 parameter buffer
 *          Highway_bridge/train_MulSca_BriStruNet_CB.py:158-190 (torch.optim.Adam, L2 weight decay)
 * p, g, m, v [n] fp32; `lr` [1] fp32 and `step` [1] int64 (already incremented, >= 1) are DEVICE scalars so a
 * captured CUDA graph follows a scheduler.  shadow_index [n] int32 (may be NULL): destination element of
 * parameter i in `shadow_bf16`, the bf16 copies of the GEMM weights ([N8, K8] zero padded), or -1;
 * shadow_index_t [n] (may be NULL): where the same element goes in the TRANSPOSED copy ([K8, N8], the operand of the
 * data-gradient GEMM), read only where shadow_index[i] >= 0.  skip (may be NULL): [n] bytes, non-zero = the element
 * belongs to a parameter that never receives a gradient and is left untouched, as torch.optim.Adam does for grad = None. */
int pcb_adam_flat_f32(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, const float *lr,
                      float beta1, float beta2, float eps, float weight_decay, const int64_t *step,
                      const int *shadow_index, const int *shadow_index_t, void *shadow_bf16, const unsigned char *skip,
                      pcb_stream_t stream);

/* ---- section 8f rank 3: input rows of BridgeStructureEncoding in one kernel
 *          Highway_bridge/models/attention_modules.py:552-574 (absolute position encoding), :590-597 (neighbours - centre),
 *          :622-687 (get_structure_features: 13 statistics per point), :603-613 (expand + cat)
 * xyz [B,N,3] fp32, idx [B,N,k] int64 (the k nearest neighbours of every point, pcb_knn_cdist_f32), 2 <= k <= 32;
 * freqs: HOST array of F <= 8 frequencies (the module's `freqs` buffer), grid_size > 0.
 * rows (may be NULL) [B*N*k, pitch]: [sin/cos encoding (6F) | neighbour - centre (3) | statistics (13) | 0 pad], fp32
 * (out_bf16 == 0) or bf16; pitch >= 6F + 16, multiple of 8.  feat (may be NULL) [B*N, 13] fp32: the statistics alone.
 * Covariance accumulated in float64, eigenvalues by the float64 closed form of pcb_eigvalsh3_f32; no host sync. */
int pcb_structure_rows_f32(const float *xyz, const int64_t *idx, int B, int N, int k, const float *freqs, int F,
                           float grid_size, int out_bf16, int pitch, void *rows, float *feat, pcb_stream_t stream);

/* ---- a11 (training): shared-MLP contractions on the tcgen05 tensor cores, BatchNorm statistics in the epilogue
 *          pointnet_util.py:213-217, 275-277, 343-345; pointnet2_sem_seg.py:43-46; pointnet2_utils.py:150-154, 353-356;
 *          DGCNN.py:134-148  (1x1 Conv2d / Conv1d = [M,K] x [K,N] on point-major rows)
 * y[M, N] = x[M, K] . w[Nw, K]^T with bf16 operands (row-major, leading dimensions ldx / ldw / ldy in elements,
 * multiples of 8, 16-byte aligned bases), fp32 accumulation in TMEM, bf16 result; N, K multiples of 8; rows >= Nw of
 * w count as zero.  Persistent warp-specialised kernel (cp.async producers -> tcgen05.mma -> TMEM -> epilogue).
 *   pcb_linear_rows_bf16          plain product (forward of a conv without BatchNorm; data gradient with w = W^T)
 *   pcb_linear_bn_stats_rows_bf16 + training-mode BatchNorm statistics of the result: mean / invstd / var [N] (biased
 *                                 variance, eps) of the bias-free output, from (count, mean, M2) triples merged with
 *                                 Chan's update in a fixed order (stateless, no cancellation).  Cv <= N real channels.
 *                                 work: pcb_gemm_work_floats(M, N, K) floats of scratch; tickets: pcb_gemm_tickets()
 *                                 zeroed 32-bit words (left zeroed).  Deterministic sums.  Deferred final fold
 *                                 (gparts != NULL): the kernel stops after its first fold level and leaves the group
 *                                 triples in gparts [*groups_out][3][N] (capacity pcb_gemm_max_groups() rows; *groups_out
 *                                 is written on the host at launch time); mean / invstd / var may be NULL -- the
 *                                 pcb_bn_apply_rows call that follows merges the groups and writes them.
 *   pcb_dgrad_bn_rows_bf16        data gradient THROUGH the previous layer's BN + ReLU: gz = gy . wt^T; dy = gz * [z > 0]
 *                                 with z = BN(yprev) recomputed from yprev [M, ldyp] and mean / invstd / gamma / beta;
 *                                 writes dy [M, lddy] and sums [3][N] = (sum dy, sum dy * yhat, 0); deferred as above
 *                                 (gparts [*groups_out][3][N] = (0, sum dy * yhat, sum dy) per group, sums may be NULL)
 *   pcb_bn_apply_rows             out = [max over pool_k rows of] act(BN(y)) with given statistics (ordinary launch);
 *                                 updates running_mean / running_var [Cv] (may be NULL) from mean / var with `momentum`
 *                                 (unbiased variance; `bias` [Cv], may be NULL, is added to the running mean only).
 *                                 gparts != NULL: mean / invstd / var are outputs, merged from `groups` group triples;
 *                                 ymax != NULL (pooled): also writes y of the winning rows [M / pool_k][C]
 *   pcb_bn_pool_bwd_rows          backward of a pooled last layer (pool_k > 1) in two ordinary launches: the BatchNorm
 *                                 sums from gz / ymax / argmax alone (M / pool_k rows), then gy in ONE pass over y
 *                                 (pcb_bn_bwd_rows: cooperative, two passes); work >= 99 * C floats,
 *                                 work[0 : 3C] = (sum dy, sum dy * yhat, 0) on return
 *   pcb_bn_bwd_apply_rows         gy = gamma * invstd * (dy - sums[0] / M - yhat * sums[1] / M); gy may alias dy.
 *                                 gparts != NULL: sums [3][C] is an output, added up from `groups` group partials */
int64_t pcb_gemm_work_floats(int64_t M, int N, int K);
int pcb_gemm_tickets(void);
int pcb_gemm_max_groups(void);
int pcb_linear_rows_bf16(const void *x, int64_t ldx, const void *w, int64_t ldw, int64_t M, int N, int Nw, int K, void *y,
                         int64_t ldy, pcb_stream_t stream);
/* inference layer, BatchNorm folded into w / bias: y = act(x . w^T + bias) (act 0 identity, 1 ReLU, 2 leaky ReLU with
 * `slope`; bias fp32 [Cv] or NULL); pool_k > 1 (divides 128 and M): only the max over every pool_k consecutive rows is
 * written, y = [M / pool_k, ldy]  (pointnet_util.py:213-217, 273-279, 343-345; DGCNN.py:134-148 in evaluation mode) */
int pcb_linear_bias_act_rows_bf16(const void *x, int64_t ldx, const void *w, int64_t ldw, int64_t M, int N, int Nw, int K,
                                  const float *bias, int Cv, int act, float slope, int pool_k, void *y, int64_t ldy,
                                  pcb_stream_t stream);
int pcb_linear_bn_stats_rows_bf16(const void *x, int64_t ldx, const void *w, int64_t ldw, int64_t M, int N, int Nw, int K,
                                  void *y, int64_t ldy, int Cv, float eps, float *mean, float *invstd, float *var,
                                  float *work, unsigned *tickets, float *gparts, int *groups_out, pcb_stream_t stream);
int pcb_dgrad_bn_rows_bf16(const void *gy, int64_t ldg, const void *wt, int64_t ldwt, int64_t M, int N, int Nw, int K,
                           const void *yprev, int64_t ldyp, const float *mean, const float *invstd, const float *gamma,
                           const float *beta, int Cv, int relu, void *dy, int64_t lddy, float *sums, float *work,
                           unsigned *tickets, float *gparts, int *groups_out, pcb_stream_t stream);
int pcb_bn_apply_rows(const void *y, int dtype, int64_t M, int C, int Cv, int pool_k, float *mean, float *invstd,
                      const float *gamma, const float *beta, int relu, void *out, int64_t out_pitch,
                      unsigned char *argmax, float *var, const float *bias, float momentum, float *running_mean,
                      float *running_var, const float *gparts, int groups, float eps, void *ymax, pcb_stream_t stream);
int pcb_bn_pool_bwd_rows(const void *gz, int64_t gz_pitch, const void *ymax, const void *y, const unsigned char *argmax,
                         int dtype, int64_t M, int C, int Cv, int pool_k, const float *mean, const float *invstd,
                         const float *gamma, const float *beta, int relu, float *work, void *gy, pcb_stream_t stream);
int pcb_bn_bwd_apply_rows(const void *dy, const void *y, int dtype, int64_t M, int C, int Cv, const float *mean,
                          const float *invstd, const float *gamma, float *sums, void *gy, const float *gparts, int groups,
                          pcb_stream_t stream);

/* ---- a11 / section 8f rank 1: fused set-abstraction / EdgeConv block for inference
 *          pointnet_util.py:137-147, 203-217, 258-279; pointnet2_utils.py:140-154, 341-356;
 *          DGCNN.py:72-109, 134-148
 * gather -> [dxyz|feat] rows (mode 0; [feat|dxyz] when xyz_first == 0) or EdgeConv rows
 * [x_nbr - x_ctr | x_ctr] (mode 1, points [B,N,D], centre = point s, S == N) -> nlayers x
 * (1x1 conv with folded BatchNorm + max(x, slope*x)) on tcgen05 tensor cores (bf16 operands,
 * fp32 accumulation in TMEM) -> max over the K neighbours -> out [B*S, cout] (bf16 or fp32).
 * kdim (HOST array, nlayers+1 entries): padded widths, multiples of 16, <= 256; wblob: bf16
 * weights packed per layer as [kdim[l]/8][kdim[l+1]][8] (zero padded), 16-byte aligned, padded
 * to a multiple of 128 bytes; bias: fp32, concatenated padded widths.  K <= 128.
 * Returns PCB_ERANGE when the weights + activation tiles do not fit in shared memory. */
int pcb_sa_fused_bf16(const float *xyz, const float *points, const float *new_xyz, const int64_t *idx, int B,
                      int N, int S, int K, int D, int mode, int xyz_first, int nlayers, const int *kdim, int cout,
                      const void *wblob, const float *bias, float slope, void *out, int out_bf16,
                      pcb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PCBRIDGE_H_ */
