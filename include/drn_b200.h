/* drn_b200.h -- C ABI of libdrn_sm100.so: the B200 (sm_100a) kernels behind DRN's dense-regression hot path.
 *
 * Caller = the Python drop-in `model/` package through ctypes (drn_b200/lib.py).  Conventions
 * (SURVEY.md section 8b):
 *   - every entry point returns 0 on success, a positive cudaError_t, or a negative DRN_E* code; the
 *     text of the last error of the calling thread is drn_last_error().  Nothing throws / exits.
 *   - all buffers are device memory owned by the caller (torch caching allocator); the library
 *     allocates nothing persistent and never synchronises the device.
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work on it.
 *   - activations are channels-last: fp32 [B, T, C] row-major, or "split planes": two bf16 tensors
 *     hi = bf16(x), lo = bf16(x - hi) stored as [2][B][T][C] (plane stride given in elements).  The
 *     tensor-core kernels contract hi/lo planes with 3 BF16 MMAs per product (fp32 TMEM accumulate),
 *     which reproduces fp32 results to ~2^-16 (the 1e-3 parity contract cannot be met by one
 *     BF16/TF32 pass -- SURVEY.md section 7 hard part 1).
 *
 * Each function cites the reference code (paths relative to the DRN repository) it replaces.
 */
#ifndef DRN_B200_H
#define DRN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRN_VERSION 100

#define DRN_EINVAL (-1)   /* bad argument / unsupported shape */
#define DRN_EARCH (-2)    /* device is not compute capability 10.x */
#define DRN_EDRIVER (-3)  /* driver entry point (cuTensorMapEncodeTiled) unavailable */

int drn_version(void);
const char* drn_last_error(void);
/* 0 iff the current device is a CC 10.x part (B200); DRN_EARCH otherwise. */
int drn_device_check(void);
int drn_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * Split-plane tensor view: bf16 [2 planes][B][T][P][C]; element (plane,b,t,p,c) lives at
 *   ptr + plane*plane_stride + ((b*T + t)*P + p)*C + c.
 * P is a "parity" split of the time axis used by the stride-2 temporal convs (P=2: time index
 * 2*t+p); P=1 otherwise.  C must be a multiple of 8, ptr 16-byte aligned.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  void* ptr;
  int64_t plane_stride; /* elements between the hi and the lo plane */
  int32_t B, T, P, C;
} drn_planes_t;

#define DRN_MAX_TAPS 4

/* ------------------------------------------------------------------------------------------------
 * drn_gemm: the one tensor-core contraction behind every dense layer of the path, forward and backward.
 *
 *   form DRN_GEMM_ROWS (temporal conv / linear forward and data-gradient):
 *     D[(b,t), n] = sum_taps sum_k A[b, t + shift_tap, par_tap, a_c0 + k] * W[wtap_tap][..]
 *       b_mn = 0: W element = Bm[b=wtap][t=n][c=b_c0+k]   (weights stored [tap][N][K]: conv/linear forward)
 *       b_mn = 1: W element = Bm[b=wtap][t=k][c=n]        (weights stored [tap][K][N]: data gradient)
 *     replaces nn.Linear / nn.Conv1d forward+dgrad at model/main_model.py:59 (prop_fc),
 *     model/basic_blocks.py:9-31 (backbone + FPN convs), model/fcos.py:31-69 (head towers, mix_fc, iou_scores).
 *     Rows outside [0,T) read as zero (= Conv1d zero padding) through TMA out-of-bounds fill.
 *   form DRN_GEMM_WGRAD (weight gradient):
 *     D[m, n] = sum_{b,t} A[b, t, 0, m] * Bm[b, t + shift_0, par_0, n]      (A = dY planes, Bm = X planes)
 *
 * Epilogue (per output element v = acc): v += bias[n]; out2[row,n] = v (optional);
 *   v *= rowscale[b*rowscale_ld + n] (optional); out[row, out_col0+n] (=, +=, atomic +=) v (optional);
 *   outp planes [row, outp_col0+n] = split(v) (optional).
 *   ROWS: row = b*out_T + t*out_t_mul + out_t_add.  WGRAD: row = m.
 * ---------------------------------------------------------------------------------------------- */
#define DRN_GEMM_ROWS 0
#define DRN_GEMM_WGRAD 2

#define DRN_OUT_STORE 0
#define DRN_OUT_ADD 1
#define DRN_OUT_ATOMIC 2

typedef struct {
  int32_t form;
  int32_t b_mn;          /* ROWS only: weight operand layout, see above */
  drn_planes_t a, b;
  int32_t B, T;          /* ROWS: output rows = B samples x T time slots (in units of A's t axis) */
  int32_t N, K;          /* ROWS: N output channels, K contraction length per tap (multiple of 64).
                            WGRAD: M = a.C-range [m0..), N = columns of Bm used, K ignored */
  int32_t M;             /* WGRAD only: rows of D (channels of A used, starting at a_c0) */
  int32_t ntaps;
  int32_t tap_shift[DRN_MAX_TAPS], tap_par[DRN_MAX_TAPS], tap_w[DRN_MAX_TAPS];
  int32_t a_c0, b_c0;    /* channel offsets into A / Bm */
  int32_t nprod;         /* 1: hi*hi only (bf16 speed mode); 3: hi*hi+hi*lo+lo*hi (parity mode); 4: + lo*lo */
  int32_t split_k;       /* WGRAD: number of K splits (>1 requires out_mode ATOMIC on a zeroed buffer) */
  /* epilogue */
  float* out; int64_t out_ld; int32_t out_col0; int32_t out_mode;
  int64_t out_tap_stride; /* WGRAD: tap i is written at out + tap_w[i]*out_tap_stride */
  int64_t out_split_stride; /* WGRAD, split_k > 1: != 0 -> K-split s STORES its partial sum at out + s*out_split_stride
                               (deterministic, no atomics, no zero-fill; the caller sums the slices, see drn_unpack_conv_wgrads) */
  int32_t out_T, out_t_mul, out_t_add;
  const float* bias;
  const float* rowscale; int32_t rowscale_ld;
  float* out2; int64_t out2_ld;
  void* outp; int64_t outp_ld; int32_t outp_col0; int64_t outp_plane_stride;
  /* engine: 0 = tcgen05 tensor cores (product path), 1 = fp32 CUDA-core checker kernel (tests only) */
  int32_t engine;
  int32_t dbg_lbo, dbg_sbo, dbg_kadv; /* 0 = defaults; descriptor overrides used by the bring-up sweep test */
  /* Optional, ROWS form through drn_gemm_group only (split_k == 1): BatchNorm statistics of the conv output fused into the
   * epilogue.  Every 32-row block of the output writes the partial column sums of v (after the bias) over its valid rows:
   *   stats[(blk * 2 + 0) * N + n] = sum v,   stats[(blk * 2 + 1) * N + n] = sum v^2,   blk < drn_gemm_stats_rows(g)
   * (plain stores, deterministic; blocks made of padding rows write zeros).  drn_bn_stats_multi reduces them
   * (drn_bn_job_t.partials) instead of re-reading the whole tensor. */
  float* stats;
} drn_gemm_t;

int drn_gemm(const drn_gemm_t* g, void* stream);
/* ONE launch of the persistent CTA-pair kernel over the 256 x 256 tiles of n <= 6 independent problems (any mix of forms):
 * the three pyramid levels of a shared head / FPN conv, or the data- and weight-gradients of one layer.  Small problems
 * launched one by one leave most of the 148 SMs idle and each pay pipeline fill and drain. */
int drn_gemm_group(int n, const drn_gemm_t* descs, void* stream);
/* The same launches with a caller-owned workspace (drn_gemm_workspace_bytes() bytes, 32-byte aligned, ZERO-FILLED ONCE by the
 * caller, never shared by launches that may run concurrently): switches the persistent kernel from whole tiles round-robin
 * to a stream-K schedule -- the k-iterations of all tiles are cut into equal contiguous ranges, one per SM pair; a range that
 * starts inside a tile leaves its fp32 partial accumulator in the workspace and the pair that owns the tile folds the
 * partials in a fixed order before the epilogue (deterministic).  Removes the tile-wave quantisation of the small layers
 * (224 tower tiles on 74 pairs = 3.03 waves -> 72.6 of 73 iterations per pair).  workspace == NULL: static schedule. */
size_t drn_gemm_workspace_bytes(void);
int drn_gemm_ws(const drn_gemm_t* g, void* workspace, size_t workspace_bytes, void* stream);
int drn_gemm_group_ws(int n, const drn_gemm_t* descs, void* workspace, size_t workspace_bytes, void* stream);
/* Profiling aid: a one-thread kernel that stores the GPU's %globaltimer (ns) into *slot (device memory).  Enqueued between
 * the kernels of a step -- also inside a CUDA-graph capture -- it yields their in-situ durations (scripts/insitu_timeline.py). */
int drn_timestamp(uint64_t* slot, void* stream);
/* Cap (0 = none) on the SM pairs the persistent CTA-pair kernel occupies in the launches that follow (process-wide; one host
 * thread per process).  The data-parallel schedule confines the prop_fc weight gradient to 70 of the 74 pairs so that the
 * NCCL all-reduce of the gradients already complete runs beside it. */
void drn_set_pair_clusters(int n);
/* number of 32-row blocks (rows of the `stats` buffer, each 2*N floats) a drn_gemm_group launch writes for this problem */
int drn_gemm_stats_rows(const drn_gemm_t* g);

/* ------------------------------------------------------------------------------------------------
 * HBM-bound kernels (drn_b200/csrc/elementwise.cu).  "planes" arguments are the hi plane pointer of a
 * split-plane tensor plus the element stride to its lo plane.
 * ---------------------------------------------------------------------------------------------- */
/* fp32 rows [rows][C] (leading dim src_ld) -> planes at column dst_col0 of a [rows][dst_ld] planes tensor.
 * Used for the C3D clip features (input of prop_fc, model/main_model.py:59) and small host-side vectors. */
int drn_split_planes(const float* src, int64_t rows, int C, int64_t src_ld, void* dst, int64_t dst_ld, int dst_col0,
                     int64_t dst_plane_stride, void* stream);
/* Level-0 query gate (model/backbone.py:28-30) on the fp32 prop_fc output x [B][T][C]: planes[b,t,dst_col0+c] = q[b][c] * x[b,t,c]. */
int drn_gate_planes(const float* x, const float* q, int B, int T, int C, void* dst, int64_t dst_ld, int dst_col0,
                    int64_t dst_plane_stride, void* stream);
/* Table-driven weight packing, ONE launch for all layers.  Item: nn.Conv1d / nn.Linear weight `src` [O][C][k] fp32 ->
 * planes [k][Ototal][C] at row offset o0 (tap-major operand of drn_gemm).  drn_unpack_conv_wgrads goes the other way for
 * weight gradients: workspace `src` [nslices][k][Ototal][C] fp32 (partial sums of the K-splits / pyramid levels, see
 * drn_gemm_t.out_split_stride) -> parameter gradient `grad` [O][C][k] = sum over the slices. */
typedef struct {
  const float* src;
  void* planes;   /* pack: destination hi plane */
  float* grad;    /* unpack: destination */
  int32_t O, C, k, Ototal, o0;
  int64_t plane_stride;
  int32_t nslices;      /* unpack: `src` holds nslices partial sums [k][Ototal][C], slice_stride elements apart, to be added */
  int64_t slice_stride;
} drn_pack_item_t;
int drn_pack_conv_weights(int n, const drn_pack_item_t* items, void* stream);
int drn_unpack_conv_wgrads(int n, const drn_pack_item_t* items, void* stream);
/* position feature, model/main_model.py:53-55: planes[row, dst_col0 + c] = Wp[c] . (s, e, e-s) + bp[c]; pos_in[row][3] saved for backward. */
int drn_pos_feature(const double* pse, const float* Wp, const float* bp, int64_t rows, int Cp, void* dst, int64_t dst_ld,
                    int dst_col0, int64_t plane_stride, float* pos_in, void* stream);
int drn_pos_bwd(const float* dx, int64_t dx_ld, int col0, const float* pos_in, int64_t rows, int Cp, float* dWp, float* dbp, void* stream);

/* Train-mode BatchNorm1d + ReLU (model/basic_blocks.py:22-30, model/fcos.py:31-38) on a conv output y [rows][C] fp32.
 * A BatchNorm "part" names the parameters of one nn.BatchNorm1d covering channels [c0, c0+n) of y (the fused cls|bbox tower
 * output carries two modules side by side).
 *   drn_bn_stats      training: per-channel sum / sum of squares (fp64 atomics); the last CTA to finish writes
 *                     coef[0..4][c] = scale, shift, mean, invstd, unbiased variance, updates running_mean/var (momentum, unbiased variance) and
 *                     num_batches_tracked, and re-zeroes `sums` and `counter` (both must be zero before the first use).
 *                     eval (training = 0): coef from the running statistics.
 *   drn_bn_relu_apply a = relu(y*scale+shift) [+ nearest-x2 upsample of `up` (FPN top-down add, model/FPN.py:63-68)]
 *                     -> planes out_a; optional planes out_qa = gate[b][c] * a (query gating, model/backbone.py:28-30)
 *   drn_bn_bwd_reduce sums of g and g*xhat (g = da masked by the ReLU); last CTA writes bcoef[0..1][c] = mean(g), mean(g*xhat),
 *                     accumulates dgamma / dbeta of each part (when non-null) and re-zeroes sums / counter.
 *   drn_bn_bwd_apply  dy planes = scale * (g - mean(g) - xhat * mean(g*xhat)). */
typedef struct {
  int32_t c0, n;
  const float* gamma; const float* beta;
  float* running_mean; float* running_var; int64_t* num_batches_tracked;
  float* dgamma; float* dbeta; /* backward only; may be null (frozen parameters) */
} drn_bn_part_t;
int drn_bn_stats(const float* y, int64_t rows, int C, int nparts, const drn_bn_part_t* parts, float momentum, float eps,
                 int training, float* coef, double* sums, unsigned* counter, void* stream);
int drn_bn_relu_apply(const float* y, int B, int T, int C, const float* coef, const void* up, int64_t up_plane_stride,
                      const float* gate, void* out_a, int64_t a_plane_stride, void* out_qa, int64_t qa_plane_stride, void* stream);
int drn_bn_bwd_reduce(const float* da, const float* y, int64_t rows, int C, float* coef, int nparts, const drn_bn_part_t* parts,
                      double* sums, unsigned* counter, float* bcoef, void* stream);
int drn_bn_bwd_apply(const float* da, const float* y, int64_t rows, int C, const float* coef, const float* bcoef, void* dy,
                     int64_t dy_plane_stride, void* stream);
/* The same four kernels over up to 3 independent BatchNorm applications per launch (the three pyramid levels of a shared head
 * or FPN block): fewer launches and full SM occupancy for the small levels.  Fields as in the single forms above. */
typedef struct {
  const float* y;            /* conv output [B*T][C] fp32 */
  const float* y2;           /* optional: second K-split slice of the contraction; drn_bn_stats_multi (training) folds it into y */
  int32_t B, T, C;
  int32_t nparts; drn_bn_part_t parts[2];
  float* coef; double* sums; unsigned* counter; float* bcoef;
  const void* up; int64_t up_plane_stride; const float* gate;                  /* apply: FPN upsample-add source, query gate */
  void* out_a; int64_t a_plane_stride; void* out_qa; int64_t qa_plane_stride;  /* apply outputs (planes) */
  const float* da; void* dy; int64_t dy_plane_stride;                          /* backward */
  /* drn_bn_stats_multi, training: partial column sums written by the contraction epilogue (drn_gemm_t.stats),
   * [partial_rows][2][C]; when set (for ALL jobs of a launch) the statistics are reduced from them and y is not read */
  const float* partials; int64_t partial_rows;
} drn_bn_job_t;
/* training: 0 = eval (coef from the running statistics); 1 = batch statistics, each job's finaliser also updates its running
 * statistics (jobs must own DISTINCT modules); 2 = batch statistics only -- the caller follows with drn_bn_running_update,
 * which absorbs the jobs' statistics into the (shared) running buffers job after job, the order of the reference's level loop
 * (model/fcos.py:93-102).  coef is [5][C]: scale, shift, mean, invstd, unbiased variance. */
int drn_bn_stats_multi(int n, const drn_bn_job_t* jobs, float momentum, float eps, int training, void* stream);
int drn_bn_running_update(int n, const drn_bn_job_t* jobs, float momentum, void* stream);
int drn_bn_relu_apply_multi(int n, const drn_bn_job_t* jobs, void* stream);
int drn_bn_bwd_reduce_multi(int n, const drn_bn_job_t* jobs, void* stream);
int drn_bn_bwd_apply_multi(int n, const drn_bn_job_t* jobs, void* stream);
/* backward of the FPN nearest-x2 upsample: dst[b][j] += src[b][2j] + src[b][2j+1]. */
int drn_pair_sum_add(float* dst, const float* src, int64_t rows_half, int C, void* stream);
/* backward of the query gate x = q[b][c] * a[b][t][c]: dq[b][c] += sum_t g*a; optionally dp planes = q*g and dbias[c] += sum q*g
 * (level 0: a = prop_fc output, dp = gradient w.r.t. it, dbias = d prop_fc.bias). */
int drn_gate_reduce(const float* g, int64_t g_ld, const void* a, int64_t a_ld, int64_t a_plane_stride, int a_is_planes, int B,
                    int T, int C, float* dq, const float* q, void* dp, int64_t dp_plane_stride, float* dbias, void* stream);
int drn_colsum(const float* x, int64_t rows, int C, int64_t ld, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * FCOS head projections and losses (drn_b200/csrc/head.cu).
 * ---------------------------------------------------------------------------------------------- */
/* cls_logits / bbox_pred / iou_scores.3 (model/fcos.py:41-69): out[b][t][o] = bias[o] + sum_r sum_c W[o][c][r] x[b][t+r-(k-1)/2][c0+c]. */
int drn_skinny_conv_fwd(const void* x, int64_t x_plane_stride, int x_ld, int c0, int Cw, int B, int T, int nout, int k,
                        const float* W, const float* bias, float* out, void* stream);
int drn_skinny_conv_bwd(const float* d, const void* x, int64_t x_plane_stride, int x_ld, int c0, int Cw, int B, int T, int nout,
                        int k, const float* W, float* dx, int dx_ld, int dx_accumulate, float* dW, void* stream);
/* The three skinny projections of the shared head on ALL pyramid levels in one launch (model/fcos.py:93-102):
 * cls_logits (k3, F -> 1) on tower channels [0,F), bbox_pred (k3, F -> 2) on tower channels [F,2F) -- the tower tensor is the
 * fused [cls_tower | bbox_tower] output -- and iou_scores.3 (k1, F/2 -> 1) on the iou_scores hidden activations (skipped where
 * iou_hidden[l] is null).  Outputs in the reference's flatten order (level, sample, t).  Backward: gradient w.r.t. the tower
 * tensor (fp32 [B*T_l][2F], stored) and weight gradients (accumulated); bias gradients come from drn_fcos_loss_bwd's pgrad. */
typedef struct {
  int32_t nlevels, B, F;
  int32_t T[3];
  const void* tower[3]; int64_t tower_plane_stride[3];            /* planes [B][T_l][2F] */
  const void* iou_hidden[3]; int64_t iou_hidden_plane_stride[3];  /* planes [B][T_l][F/2] */
  float* d_tower[3];                                              /* backward only */
} drn_head_levels_t;
int drn_head_proj_fwd(const drn_head_levels_t* h, const float* Wc, const float* bc, const float* Wb, const float* bb, const float* Wi,
                      const float* bi, float* cls_raw, float* box_raw, float* iou_raw, void* stream);
int drn_head_proj_bwd(const drn_head_levels_t* h, const float* dcls, const float* dbox, const float* Wc, const float* Wb, float* dWc,
                      float* dWb, void* stream);
/* FCOSLossComputation.__call__ (model/loss.py:134-239) on raw head outputs in the reference's flatten order
 * (level-major, then sample, then t).  losses = {loss_cls, loss_reg, loss_iou, n_pos, n_iou}; acc = 8 doubles of scratch kept for backward. */
int drn_fcos_loss_fwd(int nlevels, int B, const int* T, const float* strides, const float* cls_raw, const float* box_raw,
                      const float* iou_raw, const float* scales, const float* gt, float gamma, float alpha, int iou_branch_on,
                      float* bbox_out, double* acc, float* losses, void* stream);
int drn_fcos_loss_bwd(int nlevels, int B, const int* T, const float* strides, const float* cls_raw, const float* box_raw,
                      const float* iou_raw, const float* scales, const float* gt, float gamma, float alpha, int iou_branch_on,
                      const double* acc, const float* upstream, float* dcls, float* dbox, float* diou, float* pgrad, void* stream);

/* Eval-mode candidate selection, FCOSPostProcessor.forward_for_single_feature_map (model/inference.py:49-136) for every
 * (sample, level) in one launch: sigmoid, threshold `thr` on the class score, top_n by score (class score, or class x
 * sigmoid(IoU score) when use_iou), decode (loc -/+ reg)/32 clamped to [0,1], sqrt score.  bbox = exp'd regression
 * [B*P][2] (drn_fcos_loss_fwd's bbox_out).  Outputs [B][nlevels][top_n] (x2 for out_det), out_count [B][nlevels]; slots
 * beyond the count are untouched.  The concatenation over levels and the empty-result fallback (inference.py:167-215) are
 * host-side list assembly on these arrays. */
int drn_postprocess(int nlevels, int B, const int* T, const float* strides, const float* cls_raw, const float* bbox,
                    const float* iou_raw, float thr, int top_n, int use_iou, float* out_det, float* out_score, float* out_loc,
                    int* out_count, void* stream);

/* Proposal feature pooling + padding, the step right before the path (dataset.py:105-155 CharadesSTA.get_data, 180-206
 * collate_data): feats = the batch's per-video window features concatenated [sum n_win][D] fp32, video b owning rows
 * [win_off[b], win_off[b+1]); proposal (b, p) = (p_start float frame, p_end = min(int(end), num_frames)).  For p < nprops[b]:
 * out_feats[b][p] = element-wise max over the feature windows the proposal covers (index arithmetic of dataset.py:126-145 with
 * window = ft_window_size, interval = int(window * (1 - ft_overlap)), clamped to the last window present), out_pse[b][p] =
 * (p_start, p_end) / num_frames[b] in float64; rows p >= nprops[b] are zero.  Bit-exact with the reference (integer + max). */
int drn_pool_proposals(const float* feats, const int64_t* win_off, const double* p_start, const int32_t* p_end,
                       const int32_t* nprops, const int32_t* num_frames, int B, int P, int D, int window, int interval,
                       float* out_feats, double* out_pse, void* stream);

/* Language-guided pooling, model/LGP.py:29-51 (dead code in the reference; standalone op `model/LGP.py` of this repo), reference
 * layout: x [B][C][t] channels-first, t even.  Forward: z = query W^T (drn_sgemm_batch, store mode: exact fp32) -> drn_lgp_bn: BatchNorm1d of the query
 * tiled over t (statistics over the batch; running variance with n = B*t) -> qn, xhat [B][C], invstd [C] -> drn_lgp_pool_fwd:
 * pair scores, softmax over each pair (att [B][t/2][2]), out [B][C][t/2].  Backward: drn_lgp_pool_bwd (dx, dqn) ->
 * drn_lgp_bn_bwd (dz [B][C]; dgamma / dbeta accumulated) -> d query = dz W, dW += dz^T query (drn_sgemm_batch). */
int drn_lgp_bn(const float* z, int B, int C, int t, const float* gamma, const float* beta, float* running_mean, float* running_var,
               int64_t* num_batches_tracked, float momentum, float eps, int training, float* qn, float* xhat, float* invstd,
               void* stream);
int drn_lgp_pool_fwd(const float* x, const float* qn, int B, int C, int t, float* att, float* out, void* stream);
int drn_lgp_pool_bwd(const float* x, const float* qn, const float* att, const float* dout, int B, int C, int t, float* dx, float* dqn,
                     void* stream);
int drn_lgp_bn_bwd(const float* dqn, const float* xhat, const float* invstd, const float* gamma, int B, int C, int training, float* dz,
                   float* dgamma, float* dbeta, void* stream);

/* Fused clip_grad_norm_ + Adam step, what main.py:239-244 does after every backward (opt-in: `drn_b200.optim.FusedClipAdam`).
 * items: device array of drn_adam_item_t; (chunk_item[i], chunk_off[i]) = tensor and element offset of chunk i (chunks of
 * drn_clip_adam_chunk() elements).  total_norm over every item with a gradient; clip_coef = min(1, max_norm / (norm + 1e-6));
 * items with update = 1 take a torch.optim.Adam step (no amsgrad / weight decay) on the clipped gradient; steps[i] is the per-tensor
 * step counter (device, zero before the first call; advanced only when the tensor has a gradient, as torch does).
 * scratch: 3 doubles (sum of squares, total norm, clip coefficient) readable after the call.  No host synchronisation. */
typedef struct {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
  int64_t numel;
  int32_t update;
} drn_adam_item_t;
int drn_clip_adam(int nchunks, const void* items, const int32_t* chunk_item, const int64_t* chunk_off, double* scratch,
                  int32_t* steps, float max_norm, float lr, float beta1, float beta2, float eps, void* stream);
int drn_clip_adam_chunk(void);

/* ------------------------------------------------------------------------------------------------
 * Query encoder (drn_b200/csrc/query.cu): model/language_module.py:27-62 (QueryEncoder.forward +
 * extract_textual_command) with model/ops.py:16-25,74-85, forward and backward (recurrence, attention and backward
 * contractions in exact fp32 FMA arithmetic; input projection and forward linears in split-BF16 products).
 * ---------------------------------------------------------------------------------------------- */
/* Small fp32 contraction on CUDA cores: C[m][n] (= | +=) sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] (+ bias[n]), optional ReLU.
 * accumulate = 1 adds into C.  Serves nn.Linear forward (x W^T), data gradient (dy W) and weight gradient (dy^T x) of the
 * query encoder's projections and of the gates qInput0-2 (model/main_model.py:36-40,49-50) -- M is the batch (<= a few hundred). */
int drn_sgemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C, int64_t ldc, int M,
              int N, int K, const float* bias, int relu, int accumulate, void* stream);

/* Deterministic (fixed reduction order, no atomics) small-batch nn.Linear forward out[b][n] = act(bias[n] + sum_k x[b][k] W[n][k]):
 * query_encoder.qInput / qInput0-2 (model/language_module.py:57-58,30-31) and the gates qInput0-2 (model/main_model.py:49-50).
 * Weight-streaming on the warp-level tensor cores with the dense path's split-BF16 products (hi*hi + hi*lo + lo*hi, fp32
 * accumulate: ~1e-5 relative; use drn_sgemm_batch with store = 1 where exact fp32 FMA arithmetic is required). */
int drn_linear_fwd(const float* x, int64_t ldx, const float* W, int64_t ldw, const float* bias, float* out, int64_t ldo, int B,
                   int N, int K, int relu, void* stream);

/* Up to 12 small contractions / 4 small Linear forwards in ONE launch (the gate and query-encoder backward is a chain of ~20 of
 * them).  A drn_sgemm_batch job accumulates into C (which must then hold valid data) or overwrites it (`store`); a column sum is
 * the job B = a constant 1 with strides 0.  drn_linear_fwd_batch jobs are independent deterministic forwards. */
typedef struct {
  const float* A; int64_t sam, sak;
  const float* B; int64_t sbk, sbn;
  float* C; int64_t ldc;
  int32_t M, N, K;
  const float* bias;
  int32_t store;   /* 0: C += (atomics, K-split when small); 1: C = (plain stores, no split) */
} drn_sgemm_job_t;
int drn_sgemm_batch(int n, const drn_sgemm_job_t* jobs, void* stream);
typedef struct {
  const float* x; int64_t ldx;
  const float* W; int64_t ldw;
  const float* bias;
  float* out; int64_t ldo;
  int32_t B, N, K, relu;
} drn_linear_job_t;
int drn_linear_fwd_batch(int n, const drn_linear_job_t* jobs, void* stream);

typedef struct {
  int32_t B, L;            /* batch, padded query length = number of token columns processed (<= 64) */
  int32_t H, E;            /* LSTM hidden size per direction (512), embedding width (300) */
  int32_t tok_ld;          /* row stride of `tokens` (>= L) */
  const int64_t* tokens;   /* [B][tok_ld] device; 0 = padding */
  const int64_t* lengths;  /* [B] device; 1 <= length <= L (any order: no sorting requirement) */
  /* parameters, torch layouts (state_dict names under query_encoder.) */
  const float* emb;                          /* embedding.weight [V][E] */
  const float* w_ih[2]; const float* w_hh[2]; /* biLSTM.weight_{ih,hh}_l0{,_reverse} [4H][E], [4H][H] */
  const float* b_ih[2]; const float* b_hh[2]; /* biLSTM.bias_* [4H] */
  const float* w1; const float* b1;           /* qInput [H][4H], [H] */
  const float* w2[3]; const float* b2[3];     /* qInput0..2 [2H][H], [2H] */
  const float* wa; const float* ba;           /* cmd_inter2logits [1][2H], [1] */
  /* forward output: the three command vectors [B][2H] */
  float* cmd[3];
  /* backward input d cmd [B][2H] and gradient destinations (ACCUMULATED into; null = parameter frozen) */
  const float* dcmd[3];
  float* g_emb; float* g_w_ih[2]; float* g_w_hh[2]; float* g_b_ih[2]; float* g_b_hh[2];
  float* g_w1; float* g_b1; float* g_w2[3]; float* g_b2[3]; float* g_wa; float* g_ba;
  /* scratch that carries the forward state to the backward call: drn_qe_workspace_bytes(B, L, H, E) bytes */
  void* workspace; size_t workspace_bytes;
} drn_qe_t;

size_t drn_qe_workspace_bytes(int B, int L, int H, int E);
/* Staging of the caller's query tensors (tokens [B][tok_ld], first `ncols` columns used; lengths [B]; both int64, device) into
 * the static [B][L] / [B] buffers drn_qe_t points at, WITH VALIDATION: the kernels index the embedding table (and its gradient)
 * by token id and the LSTM output by length-1.  A token id outside [0, vocab) is staged as 0 (padding), a length outside
 * [1, min(L, ncols)] is clamped; *err gets the sticky bits 1 (token) / 2 (length) and *poison = NaN for this batch, else 0 --
 * the caller adds *poison to the losses, so a bad batch is loud without a device synchronisation (the reference raises:
 * nn.Embedding IndexError / pack_padded_sequence, model/language_module.py:41-42). */
int drn_qe_stage(const int64_t* tokens, int64_t tok_ld, int ncols, const int64_t* lengths, int B, int L, int vocab,
                 int64_t* tokens_out, int64_t* lengths_out, int32_t* err, float* poison, void* stream);
/* number of kernels one drn_qe_forward (backward = 0) / drn_qe_backward (backward = 1) call launches for this shape */
int drn_qe_launch_count(int B, int L, int H, int backward);
int drn_qe_forward(const drn_qe_t* q, void* stream);
int drn_qe_backward(const drn_qe_t* q, void* stream);
/* The same work in two calls, split where the latency-bound recurrence starts, so that the caller can fork an independent
 * HBM-bound branch there (weight packing beside the forward LSTM, weight-gradient unpacking beside the BPTT):
 * part 1 = everything before the recurrence, part 2 = the recurrence and everything after it.  part 1 then part 2 on one
 * stream == drn_qe_forward / drn_qe_backward. */
int drn_qe_forward_part(const drn_qe_t* q, int part, void* stream);
int drn_qe_backward_part(const drn_qe_t* q, int part, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Batched temporal NMS + recall@k (drn_b200/csrc/metric.cu): the evaluation step right after the hot path.  Replaces the
 * pure-Python per-query loops of utils/evaluate_utils.py:132-190 (compute_IoU_recall_top_n_ours), 192-215 (nms_temporal) and
 * 232-236 (calculate_IoU) as main.py:450-454 runs them.  One warp per query, IEEE double arithmetic in the reference's
 * operation order: picks are bit-exact.
 *   det [Q][G*K][2], score [Q][G*K] fp32: candidates of query q = G groups (pyramid levels) of K slots, group g holding
 *   count[q*G+g] valid entries at its front (the layout drn_postprocess writes; a plain list is G = 1, K = list capacity);
 *   gt [Q][2] f64; overlap = NMS threshold (the reference passes iou - 0.05); topk [ntopk] int32 (device).
 * Candidate order = (group, slot); among equal scores the later candidate is picked first (stable ascending sort, best last:
 * evaluate_utils.py:200).  nms == 0: no suppression, picks = all candidates by descending score, among equal scores the EARLIER
 * first (the reference's stable descending sort, evaluate_utils.py:97,165); `overlap` is ignored.  Zero-length candidates are dropped (the reference raises ZeroDivisionError on them);
 * empty_fallback != 0: a query without candidates is given the detection (0, 1), score 1 (model/inference.py:192-197).
 * Outputs (each optional): picks [Q][G*K] candidate indices in pick order, -1 padded; npicks [Q]; hits [Q][ntopk] 0/1;
 * correct [ntopk] += number of hit queries (caller zeroes).  G*K <= 256, ntopk <= 8. */
int drn_nms_recall(const float* det, const float* score, const int32_t* count, const double* gt, int Q, int G, int K,
                   int nms, double overlap, double iou_thr, const int32_t* topk, int ntopk, int empty_fallback, int32_t* picks,
                   int32_t* npicks, int32_t* hits, int32_t* correct, void* stream);

/* Schedule of the persistent contraction kernel for launches that are given a workspace (drn_gemm_ws / drn_gemm_group_ws):
 * 0 = static tile round-robin only; 1 = hybrid (default; environment DRN_SCHEDULE=static|hybrid|streamk): full waves static, the
 * k-iterations of the last, partial wave cut into equal ranges over all SM pairs and folded through the workspace in pair order
 * (deterministic); 2 = full stream-K.  Groups whose problems have tiles of different lengths use the host-balanced static
 * schedule instead. */
void drn_gemm_set_schedule(int mode);
/* The launch planner on its own (host arithmetic only, no GPU): `nprob` problems of tiles[k] tiles with nk[k] k-iterations each,
 * sorted by decreasing nk as the launcher does, on `pairs` SM pairs, with or without a workspace, under schedule `mode` (< 0: the
 * process setting).  Returns the SM pairs that would be launched; *kind = 0 round-robin, 1 host-balanced tile lists
 * (counts[pair], lists[pair * 16 + i]), 2 hybrid, 3 stream-K; *quota, *static_tiles = k-range length per pair and number of
 * whole tiles walked statically.  tests/test_host_cpu.py checks the planner through it. */
int drn_gemm_schedule_probe(int nprob, const int* tiles, const int* nk, int pairs, int has_ws, int mode, int* kind, int* quota,
                            int* static_tiles, unsigned char* counts, unsigned short* lists);

/* Diagnostic: the next `launches` launches of the persistent contraction kernel (eager, or captured into a CUDA graph -- the
 * slot is part of the captured launch) write %globaltimer stamps to buf[launch][160 CTAs][8]: 0 entry, 1 prologue done, 2 first
 * operands landed, 3 last MMA issued, 4 / 5 first / last accumulator complete, 6 last accumulator drained, 7 exit.  buf = null
 * switches it off.  drn_gemm_trace_info returns the CTA count, tile count and schedule kind of a traced launch
 * (scripts/gemm_trace.py). */
void drn_gemm_trace(uint64_t* buf, int launches);
int drn_gemm_trace_info(int launch, int* ctas, int* tiles, int* lpt);

/* ------------------------------------------------------------------------------------------------
 * Gradient exchange over NVLink peer memory (drn_b200/csrc/p2p.cu): the ONE collective of the data-parallel path
 * (all-reduce of the fp32 gradients, main.py:99 nn.DataParallel semantics), written as a kernel that
 * reads and writes the peers' flat gradient buffers directly instead of calling NCCL.
 *   drn_ipc_export: CUDA IPC handle (64 bytes) of the cudaMalloc allocation that contains `ptr`, and ptr's byte offset in it.
 *   drn_ipc_open:   maps a peer process's allocation (peer access enabled lazily) and returns base + offset.
 *   drn_ipc_close:  unmaps what drn_ipc_open returned (pass the same offset).
 *   drn_p2p_allreduce_avg: buf[r] = rank r's flat fp32 buffer (own pointer at buf[rank], peers' mapped pointers elsewhere),
 *     flags[r] = rank r's zero-initialised uint32[DRN_P2P_FLAG_WORDS] flag block, mapped the same way.  Elements
 *     [offset, offset + n) (both multiples of 4) of EVERY rank's buffer are replaced by their mean over the ranks: rank r
 *     sums slice r of the range from all ranks in rank order (bit-identical result on every rank, deterministic), scales by
 *     1/world and stores the result into all `world` buffers.  Two flag exchanges (st.release.sys / ld.acquire.sys on the
 *     peers' flag blocks, an epoch counter in flags[rank][0]) order it: nobody reads before every rank has launched the call
 *     (its gradients are complete: stream order), nobody returns before every rank's stores have landed.  Every rank must
 *     issue the same sequence of calls.  `ctas` = CTAs of 512 threads (0 = default).  A peer that never arrives traps the
 *     kernel after ~30 s instead of hanging the GPU. */
#define DRN_P2P_MAX_RANKS 8
#define DRN_P2P_FLAG_WORDS 64
typedef struct {
  int32_t world, rank;
  float* buf[DRN_P2P_MAX_RANKS];
  uint32_t* flags[DRN_P2P_MAX_RANKS];
} drn_p2p_t;
int drn_ipc_export(const void* ptr, unsigned char* handle64, int64_t* offset);
int drn_ipc_open(const unsigned char* handle64, int64_t offset, void** out);
int drn_ipc_close(void* ptr, int64_t offset);
int drn_p2p_allreduce_avg(const drn_p2p_t* comm, int64_t offset, int64_t n, int ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DRN_B200_H */
