/* morgana_b200 -- C ABI of the B200-native (sm_100a) kernels for morgana's per-batch frame-rate feature path.
 *
 * The reference (ZackHodari/morgana) is pure Python/PyTorch and has no FFI: the "interface" each entry point
 * replaces is a Python callable, cited as <file>:<lines> relative to the reference root.  INTEGRATION.md shows the
 * ctypes binding and the monkey-patch a maintainer of the reference would add.
 *
 * Conventions (every entry point):
 *   - plain C types only: raw DEVICE pointers, sizes, strides (in ELEMENTS unless a name ends in _bytes) and a
 *     CUDA stream handle; no torch types.
 *   - returns MG_OK (0) or a negative MG_ERR_* code; never throws; mg_last_error() gives a thread-local message.
 *   - never allocates, never synchronises the stream, never touches host memory of the caller: all buffers,
 *     including workspaces, are owned by the caller.  Launches go on the stream that is passed in.
 *   - inputs are read-only; only the documented outputs are written (mg_ema_update_f32 updates `shadow` in place,
 *     as the reference does: utils.py:436-448).
 */
#ifndef MORGANA_B200_H_
#define MORGANA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MG_ABI_VERSION 1

#define MG_OK 0
#define MG_ERR_INVALID_ARG (-1)
#define MG_ERR_CUDA (-2)
#define MG_ERR_UNSUPPORTED (-3)

typedef void* mg_stream_t; /* cudaStream_t */

int mg_abi_version(void);
const char* mg_last_error(void);
/* Number of SMs of the current device (grid sizing), or a negative error code. */
int mg_sm_count(void);

/* ---------------------------------------------------------------------------------------------------------------
 * K1  duration scan -- replaces `repeated_lens = sum(repeats, 1); max(repeated_lens).item()` and the per-utterance
 *     np.repeat index construction of utils.upsample_to_repetitions (morgana/utils.py:198-199, 214-220).
 *
 * dur        (B, P) durations, int64 (dur_is_i32 = 0) or int32 (= 1); row stride dur_stride_b, item stride 1.
 * ends       (B, P) int32 out: inclusive running sum per utterance; item p covers frames [ends[p-1], ends[p]).
 * n_frames   (B,)   int64 out: sum_p dur[b, p] (what the reference calls repeated_lens).
 * summary    int64[4] out: [0] max_b n_frames (the T of the output), [1] number of negative durations (the reference
 *            raises ValueError for any, utils.py:220), [2] sum_b n_frames, [3] number of utterances whose total
 *            exceeds INT32_MAX (unsupported).  The caller reads it back (32 bytes) to size the output.  May be NULL when
 *            the caller already knows the padded length (nothing is accumulated or cleared then).
 */
int mg_dur_scan(const void* dur, int dur_is_i32, int64_t dur_stride_b, int B, int P,
                int32_t* ends, int64_t* n_frames, int64_t* summary, mg_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K2  fused normalise + duration-driven expansion -- replaces utils.upsample_to_repetitions (utils.py:175-228) and,
 *     with norm_mode != 0, its composition with data.normalise_mvn / normalise_minmax (data.py:533-534, 579-583).
 *
 *     out[b, t, :] = norm(x[b, p(b,t), :])  for t <  n_frames[b]   (p(b,t): the item whose interval holds t)
 *     out[b, t, :] = 0                      for t >= n_frames[b]   (never norm(0): utils.py:206-207, 214)
 *
 * x          (B, P, D) fp32; strides x_stride_b / x_stride_p in elements, innermost contiguous.
 * ends       (B, P) int32 from mg_dur_scan.
 * norm_mode  MG_NORM_NONE | MG_NORM_MVN (p0 = mean, p1 = std_dev: (x - p0) / (p1 + 1e-8))
 *                         | MG_NORM_MINMAX (p0 = mmin, p1 = mmax: (x - p0) / scale, scale = p1 - p0, 1 where |scale| <= 1e-8)
 * p0, p1     (D,) fp32 when param_stride_b = 0, else (B, D) with that row stride (speaker-dependent, data.py:460-501).
 * out        (B, T, D) fp32, contiguous.  T may exceed max n_frames (extra rows are zero) but not be smaller.
 * path       MG_PATH_AUTO, or force MG_PATH_BULK (smem-staged rows + cp.async.bulk stores; needs D % 4 == 0 and
 *            16-byte aligned `out`) / MG_PATH_DIRECT (register-staged vector stores; any D).
 */
#define MG_NORM_NONE 0
#define MG_NORM_MVN 1
#define MG_NORM_MINMAX 2
#define MG_PATH_AUTO 0
#define MG_PATH_BULK 1
#define MG_PATH_DIRECT 2
int mg_upsample_norm_f32(const float* x, int64_t x_stride_b, int64_t x_stride_p, const int32_t* ends,
                         const float* p0, const float* p1, int64_t param_stride_b, int norm_mode,
                         float* out, int B, int P, int D, int64_t T, int path, mg_stream_t stream);

/* The same fused op (morgana/utils.py:175-228 composed with data.py:533-534 / 579-583) with a bfloat16 output (additive; not
 * in the reference): the exact fp32 result rounded to nearest-even,
 * written once at half the bytes -- the activation format of the tensor-core layers (K7).  D % 8 == 0; out (B, T, D) bf16. */
int mg_upsample_norm_f32_bf16out(const float* x, int64_t x_stride_b, int64_t x_stride_p, const int32_t* ends,
                                 const float* p0, const float* p1, int64_t param_stride_b, int norm_mode,
                                 void* out, int B, int P, int D, int64_t T, mg_stream_t stream);

/* Packed (ragged) input -- "next" row 3: the items of the batch arrive as they are on the wire, utterance after utterance
 * in one flat array, with no phone padding to read or to produce (the padded layout is what FilesDataset.collate_fn builds
 * on the host, morgana/data.py:184-193).
 *
 * item_ends  (B,) int32: inclusive scan of the item counts (mg_dur_scan over the counts viewed as one (1, B) row).
 * mg_dur_scan_packed: dur (sum_b n_items_b,) int64 / int32; ends (sum_b n_items_b,) int32 out, the scan restarting at every
 *            utterance; n_frames / summary as for mg_dur_scan.
 * mg_upsample_packed_norm_f32: x (sum_b n_items_b, D) fp32 with row stride x_stride_p; max_items >= every item count;
 *            everything else as for mg_upsample_norm_f32 (same kernels, same results as padding first).
 * total_items = rows of dur / x / ends: item counts that overrun it are clamped on the device (never read past the arrays). */
int mg_dur_scan_packed(const void* dur, int dur_is_i32, const int32_t* item_ends, int B, int64_t total_items, int32_t* ends,
                       int64_t* n_frames, int64_t* summary, mg_stream_t stream);
int mg_upsample_packed_norm_f32(const float* x, int64_t x_stride_p, const int32_t* item_ends, int64_t total_items, const int32_t* ends,
                                const float* p0, const float* p1, int64_t param_stride_b, int norm_mode, float* out,
                                int B, int max_items, int D, int64_t T, mg_stream_t stream);

/* Dtype-agnostic expansion (the advanced-index gather of morgana/utils.py:226 preserves any dtype): rows of
 * row_bytes bytes are copied.
 * Strides in BYTES.  out is (B, T, row_bytes) contiguous. */
int mg_upsample_bytes(const void* x, int64_t x_stride_b_bytes, int64_t x_stride_p_bytes, const int32_t* ends,
                      void* out, int B, int P, int64_t row_bytes, int64_t T, int path, mg_stream_t stream);

/* Backward of K2 w.r.t. x -- replaces autograd's IndexBackward0 of utils.py:226 (an atomic index_put_) with a
 * deterministic per-item segment sum:  grad_x[b, p, :] = (sum_{t in item p} grad_out[b, t, :]) / denom(norm).
 * grad_out (B, T, D) contiguous; grad_x (B, P, D) contiguous; p0/p1/param_stride_b/norm_mode as in K2. */
int mg_upsample_norm_bwd_f32(const float* grad_out, const int32_t* ends, const float* p0, const float* p1,
                             int64_t param_stride_b, int norm_mode, float* grad_x,
                             int B, int P, int D, int64_t T, mg_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K0  on-device collate -- replaces the zero-padding of FilesDataset.collate_fn (morgana/data.py:159-224, :184-193) and
 *     the host->device copy of the PADDED batch (data.py:655-663): the host ships only valid rows, packed back to back.
 *
 * packed     (sum_b len_b, row_bytes) bytes: utterance 0's rows, then utterance 1's, ...
 * ends       (B,) int32 inclusive scan of the lengths (mg_dur_scan over the lengths viewed as one (1, B) row).
 * out        (B, T, row_bytes): rows t < len_b copied, rows t >= len_b zero.  Utterances longer than T are truncated.
 * total_rows rows held by `packed`: lengths that sum past it are truncated there (nothing is read beyond the buffer).
 */
int mg_pad_collate(const void* packed, const int32_t* ends, void* out, int B, int64_t row_bytes, int64_t T,
                   int64_t total_rows, mg_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Sibling segment operations ("next" row 4): the same scan drives three more dtype-agnostic row movers.  Strides in BYTES.
 *
 * mg_pack_rows -- utils.batched_masked_select (morgana/utils.py:147-166): rows t < len_b of every utterance, back to back.
 *     ends: (B,) int32 inclusive scan of the lengths; out: (ends[B-1], row_bytes).  The inverse of mg_pad_collate.
 * mg_segment_ends -- utils.get_segment_ends (morgana/utils.py:287-330): out[b, s] = x[b, cumsum(lens)[b, s] - 1], zero for
 *     empty segments.  seg_ends: (B, S) int32 inclusive scan of the segment lengths (mg_dur_scan); out: (B, S, row_bytes).
 * mg_split_to_segments -- utils.split_to_segments (morgana/utils.py:231-284): out[b, s, j] = x[b, begin_s + j] for
 *     j < len_s, else zero; out: (B, S, L, row_bytes) with L = the longest segment (summary of the caller's choice).
 */
int mg_pack_rows(const void* x, int64_t x_stride_b_bytes, int64_t x_stride_t_bytes, const int32_t* ends, void* out,
                 int B, int64_t T, int64_t row_bytes, mg_stream_t stream);
int mg_segment_ends(const void* x, int64_t x_stride_b_bytes, int64_t x_stride_t_bytes, const int32_t* seg_ends, void* out,
                    int B, int S, int64_t T, int64_t row_bytes, mg_stream_t stream);
int mg_split_to_segments(const void* x, int64_t x_stride_b_bytes, int64_t x_stride_t_bytes, const int32_t* seg_ends, void* out,
                         int B, int S, int64_t L, int64_t T, int64_t row_bytes, mg_stream_t stream);
/* Backward of mg_segment_ends (L == 0; grad_out (B, S, row_bytes)) and of mg_split_to_segments (L > 0; grad_out
 * (B, S, L, row_bytes)): what autograd's index_put_ does for the reference's advanced indexing (utils.py:281, 328).  Every
 * frame row receives at most one output row, so this is a gather into the contiguous grad_x (B, T, row_bytes); rows no output
 * row came from are zero.  (mg_pack_rows' backward is mg_pad_collate.) */
int mg_segments_bwd(const void* grad_out, const int32_t* seg_ends, void* grad_x, int B, int S, int64_t L, int64_t T,
                    int64_t row_bytes, mg_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K3  standalone normalise / denormalise -- replaces data.normalise_mvn, denormalise_mvn, normalise_minmax,
 *     denormalise_minmax on torch tensors (morgana/data.py:533-538, 579-590).
 *
 * x, out     (rows, D) fp32 contiguous (out may alias x).
 * rows_per_param  0: p0/p1 are (D,) shared by all rows; > 0: row r uses parameter row r / rows_per_param of a
 *            contiguous (n, D) table (speaker-dependent: rows_per_param = T).
 * inverse    0: normalise; 1: denormalise (x * s + p0 as a multiply then an add, data.py:537-538, 586-590).
 */
int mg_normalise_f32(const float* x, const float* p0, const float* p1, int norm_mode, int inverse, float* out,
                     int64_t rows, int D, int64_t rows_per_param, mg_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K4 / K5  seq_len-masked one-pass reductions -- replace losses.sequence_loss -> mse / bce (morgana/losses.py:9-56)
 *     and the accumulate() arithmetic of metrics.Mean / RMSE / MAE / Error / Accuracy / F0Distortion / LF0Distortion /
 *     Distortion / MelCepDistortion (morgana/metrics.py:383-394, 492-495, 520-522, 547-549, 574-576, 597-609, 630-634,
 *     657-665, 690-694), without materialising utils.sequence_mask (utils.py:115-144).
 *
 * Up to MG_MAX_TERMS terms are reduced in ONE launch; term i sees rows t < n_b = clamp(seq_len[b], 0, T) of utterance b
 * (seq_len == NULL: every row).  For each term the kernel produces
 *     sum   = sum_b S_b,  S_b = sum_{t < n_b} rowfn(a[b,t,:], b[b,t,:]) * m[b,t]
 *     count = sum_b (sum_{t < n_b} m[b,t])           (frames, not frames x D: SURVEY.md Q2; m == NULL -> 1)
 *             -- or T*B*D when seq_len == NULL and the kind is not mask-carrying, as Mean.accumulate does (numel)
 *     loss  = (1 / (B * D)) * sum_b S_b / n_b        (losses.py:37-42; 0/0 -> nan as in the reference)
 * and, when `grad` is non-NULL (loss kinds only), d loss / d a scaled by grad_scale * (*grad_scale_dev, if non-NULL),
 * written over all T rows (zero in the padding).
 *
 * Reduction order is fixed by the launch geometry (per-thread -> warp shuffle -> CTA -> per-CTA slot -> last CTA sums
 * the slots in index order), so results are bit-reproducible run to run; no floating-point atomics.
 */
#define MG_MAX_TERMS 12

#define MG_RED_SQDIFF 0      /* (a - b)^2             mse loss; RMSE / MelCepDistortion (caller offsets the pointers) */
#define MG_RED_ABSDIFF 1     /* |a - b|               L1 loss; MAE */
#define MG_RED_BCE 2         /* -(b log a + (1-b) log(1-a)), logs clamped at -100 (ATen); a = probability, b = label */
#define MG_RED_SUM 3         /* a                     Mean */
#define MG_RED_ROOT_SQDIFF 4 /* sqrt(sum_d (a-b)^2)   Distortion: one value per frame */
#define MG_RED_SQDIFF_EXP 5  /* (exp a - exp b)^2     LF0Distortion (with m = voiced) */
#define MG_RED_XOR 6         /* a ^ b on uint8/bool   Error (exact integer sum) */
#define MG_RED_AND 7         /* a & b on uint8/bool   Accuracy */
#define MG_RED_EQ 8          /* (a == b) ? 1 : 0      V/UV accuracy as models/RNN_SPSS.py:127 builds it, fused */
#define MG_RED_SQ 9          /* a * a                 Variance / StandardDeviation (morgana/metrics.py:427-442), with MG_RED_SUM */
#define MG_RED_CE 10         /* logsumexp_d(a) - a[b] cross-entropy loss, losses.ce (morgana/losses.py:59-61): a = (B, T, D) logits,
                                b = (B, T) int64 class indices (b_sb / b_st in elements); one value per frame, the loss's
                                feature axis has size 1.  Gradient: (softmax(a) - onehot(b)) * scale / (n_b * B) */

/* Fuse the `output_features['vuv'] > 0.5` of models/RNN_SPSS.py:122 into the reduction instead of a separate pass: */
#define MG_FLAG_M_GT_HALF 1 /* the per-frame weight is (m > 0.5) ? 1 : 0 */
#define MG_FLAG_A_GT_HALF 2 /* MG_RED_EQ compares (a > 0.5) with (b != 0) */
#define MG_FLAG_IN_TOTAL 4  /* add grad_scale * loss of this term into the first record's weighted_loss_f32 */

#define MG_DT_F32 0
#define MG_DT_U8 1 /* uint8 or bool */

struct mg_term_result;

typedef struct mg_term {
  const void* a; /* (B, T, D) view: element (b, t, d) at a + b*a_sb + t*a_st + d */
  const void* b; /* second operand or NULL (MG_RED_SUM, MG_RED_SQ) */
  const void* m; /* optional per-frame weight (B, T) -- F0Distortion's is_voiced; NULL = 1 */
  float* grad;   /* optional (B, T, D) gradient w.r.t. a (loss kinds), strides g_sb / g_st */
  const float* grad_scale_dev; /* optional device scalar multiplied into the gradient (upstream grad_output) */
  struct mg_term_result* result; /* device record this term writes (or adds to, see `accumulate`) */
  int64_t a_sb, a_st, b_sb, b_st, m_sb, m_st, g_sb, g_st;
  int32_t D;
  int32_t kind;    /* MG_RED_* */
  int32_t ab_dtype; /* MG_DT_F32 (all float kinds) or MG_DT_U8 (XOR / AND / EQ / SUM) */
  int32_t m_dtype;  /* MG_DT_F32 or MG_DT_U8 */
  int32_t b_is_u8;  /* MG_RED_EQ only: b is uint8/bool while a is f32 (probability > 0.5 is NOT applied here) */
  int32_t accumulate; /* 0: result = this batch; 1: result.sum / count / isum += this batch (streaming-metric state:
                         `self.sum += ...; self.count += ...` of metrics.py:389-394 without leaving the device) */
  float grad_scale; /* host-side factor of the gradient (e.g. 0.25 for loss / 4) */
  int32_t flags;    /* MG_FLAG_* */
} mg_term;

/* Result record per term (device memory, 48 bytes, 16-byte aligned, written by the last CTA). */
typedef struct mg_term_result {
  double sum;
  double count;
  double loss;
  int64_t isum;  /* exact integer sum for MG_DT_U8 operands */
  float sum_f32; /* the same three values rounded once to fp32, for zero-copy views from the host framework */
  float count_f32;
  float loss_f32;
  float weighted_loss_f32; /* record of term 0 only: sum of grad_scale * loss over the MG_FLAG_IN_TOTAL terms, i.e. the
                              model's total loss `(mse + mse + mse + bce) / 4` (models/RNN_SPSS.py:131-139) in one launch */
} mg_term_result;

/* Bytes of workspace mg_masked_reduce needs for this geometry.  The workspace must be zero-filled ONCE when it is
 * allocated; the kernel leaves it clean for the next launch -- of ANY geometry the workspace is large enough for (the
 * ticket area at its start has a fixed size, so changing B or n_terms between launches is safe).  One workspace per
 * concurrently-used stream. */
int64_t mg_masked_reduce_workspace_bytes(int n_terms, int B, int64_t T);

int mg_masked_reduce(const mg_term* terms /* host array */, int n_terms, const int64_t* seq_len /* (B,) or NULL */,
                     int B, int64_t T, void* workspace, int64_t workspace_bytes, mg_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K4b  whole-row masked objective -- the fused form of K4/K5 for an acoustic model's loss() (reference
 *      models/RNN_SPSS.py:120-139: 4 metric accumulations + 3 x losses.mse + losses.bce over column groups of ONE
 *      (B, T, D) prediction / target pair).  One pass: every valid row is read once, whole, by a CTA whose thread c owns
 *      column c; the per-column program `cols[c]` says which loss slot and which metric slot the column feeds, so each
 *      byte of pred / target is read once and the gradient is written once, coalesced.
 *      Algorithmic HBM bytes: 8*D*sum_b n_b (+ 4*D*B*T with a gradient).
 *
 * pred, target   (B, T, D) fp32 views, unit inner stride, strides in elements.
 * grad           NULL, or (B, T, D) fp32: d(total loss)/d(pred) = l'(p, y) * cols[c].loss_weight * (*grad_scale_dev or 1)
 *                / (n_b * B) for valid rows, 0 elsewhere and for columns without a loss.
 * cols           DEVICE array of D column programs.
 * slots          HOST array of n_slots (<= MG_MAX_TERMS) slot descriptors; slot ids in `cols` index it.
 * workspace      as for mg_masked_reduce (mg_masked_reduce_workspace_bytes(n_slots, B, T)).
 */
#define MG_COL_NONE (-1)
typedef struct mg_column {
  int8_t loss_kind;   /* MG_COL_NONE, MG_RED_SQDIFF, MG_RED_ABSDIFF or MG_RED_BCE (pred = probability, target = label) */
  int8_t loss_slot;
  int8_t metric_kind; /* MG_COL_NONE, MG_RED_SQDIFF, MG_RED_ABSDIFF, MG_RED_SQDIFF_EXP, MG_RED_EQ ((pred > 0.5) == (target != 0))
                         or MG_RED_ROOT_SQDIFF (this column LEADS a group of `width` columns: sqrt of the group's squared error) */
  int8_t metric_slot;
  int16_t mask_col;   /* MG_COL_NONE, or the metric is weighted per frame by (pred[row, mask_col] > 0.5): F0Distortion's
                         is_voiced as models/RNN_SPSS.py:122 derives it */
  int16_t width;      /* MG_RED_ROOT_SQDIFF group width (>= 1) */
  float loss_weight;  /* term weight / term width, e.g. 0.25 / 180 for the mcep stream of `loss / 4` */
} mg_column;

typedef struct mg_slot {
  struct mg_term_result* result; /* device record written (or added to) by the last CTA */
  int32_t D;          /* number of columns feeding the slot (loss: mean over D; metric without seq_len: numel = B*T*D) */
  int32_t per_frame;  /* 1: the slot holds one value per frame (ROOT_SQDIFF, weighted metrics) */
  int32_t weighted;   /* 1: count = sum of the per-frame weights (mask_col metrics) */
  int32_t accumulate; /* 1: streaming-metric state, result += batch */
  int32_t in_total;   /* 1: weight * loss is added to slots[0].result->weighted_loss_f32 */
  float weight;
} mg_slot;

/* host-only: the stage -> CTA partition of the persistent row-stream form of mg_masked_objective_f32 for given utterance lengths
 * (HOST array, or NULL for full-length utterances) on a device with `sms` SMs, computed with the same functions the device runs:
 * out[0] = grid, out[1] = first stage behind the fixed head starts, out[2] = number of 8-row stages, out[3 .. 3 + grid] = range
 * boundaries.  Returns 1 when the shape is not one the stream form takes.  Used by the CPU tests (partition invariants). */
int mg_objective_stream_plan(const int64_t* seq_len_host, int B, int64_t T, int D, int has_grad, int n_slots, int sms,
                             int64_t* out, int64_t out_len);
int mg_masked_objective_f32(const float* pred, int64_t p_sb, int64_t p_st, const float* target, int64_t t_sb, int64_t t_st,
                            float* grad, int64_t g_sb, int64_t g_st, const float* grad_scale_dev,
                            const mg_column* cols, int D, const mg_slot* slots, int n_slots,
                            const int64_t* seq_len, int B, int64_t T, void* workspace, int64_t workspace_bytes,
                            mg_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K6  multi-tensor EMA -- replaces utils.ExponentialMovingAverage._update_param / update_params (utils.py:443-456):
 *     shadow[i] = shadow[i] - one_minus_decay * (shadow[i] - param[i]), two roundings then the subtract (bit-exact
 *     with the reference's two ATen kernels; not the FMA / decay*s + (1-decay)*x form).
 * shadow, param   HOST arrays of n_tensors DEVICE pointers; numel the element counts.  One launch per 64 tensors.
 */
int mg_ema_update_f32(float* const* shadow, const float* const* param, const int64_t* numel, int n_tensors,
                      float one_minus_decay, mg_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K7  dense layer on the tcgen05 tensor cores -- replaces nn.Linear (+ nn.Sigmoid) of the example models
 *     (README.rst:65-73; models/RNN_SPSS.py:33,38,41; models/f0_test_model.py:29,41,44):
 *     y[M, N] = act(x[M, K] @ w[N, K]^T + bias[N]),  bf16 operands, fp32 accumulation in TMEM.
 *
 * x          (M, K) bf16, row stride ldx (elements, multiple of 8); K multiple of 8.
 * w          (N, K) bf16, row stride ldw (multiple of 8); N <= 2048 (the bias is staged in shared memory).
 * y          (M, N) fp32 (y_is_bf16 = 0) or bf16 (= 1), row stride ldy.
 * For N > 256 and M >= 256 * #SMs the kernel runs as 2-CTA clusters (tcgen05.mma.cta_group::2); same results.
 * act        MG_ACT_NONE | MG_ACT_SIGMOID.
 */
#define MG_ACT_NONE 0
#define MG_ACT_SIGMOID 1
int mg_linear_bf16(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, void* y, int64_t ldy,
                   int y_is_bf16, int M, int N, int K, int act, mg_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K7 backward -- what autograd runs for the same layers in the training step (README.rst:86-99, experiment_builder.py:470-479
 *     loss.backward() over the nn.Linear / nn.Sigmoid modules of README.rst:65-73, models/RNN_SPSS.py:33,38,41):
 *
 * mg_act_grad_bf16      g = grad_y * (1 - y) * y (ATen's sigmoid_backward; y = NULL: g = grad_y) written as bf16 rows of
 *                       ld_out = round_up(N, 8) columns (padding zero) -- the operand of both backward GEMMs -- and, when
 *                       bias_grad != NULL, bias_grad[n] = sum_m g[m, n] from the fp32 values of the same pass (fp64 partial
 *                       sums, fixed order).  grad_y / y are (M, N) fp32 or bf16 with row strides ldg / ldy (elements).
 *                       workspace: mg_act_grad_workspace_bytes(M, N) bytes (only read when bias_grad != NULL).
 * mg_linear_wgrad_bf16  grad_w[N, K] = g[M, N]^T @ x[M, K]: bf16 operands exactly as the forward pass holds them (row-major,
 *                       frames outermost, row strides multiples of 8), fp32 accumulation in tensor memory; the frames are
 *                       split over the SMs and the slices summed in slice order (deterministic).  grad_w fp32, row stride ldw.
 *                       workspace: mg_linear_wgrad_workspace_bytes(M, N, K) bytes, 16-byte aligned, no initialisation needed.
 */
int64_t mg_act_grad_workspace_bytes(int64_t M, int N);
int mg_act_grad_bf16(const void* grad_y, int grad_is_bf16, int64_t ldg, const void* y, int y_is_bf16, int64_t ldy,
                     void* out, int64_t ld_out, float* bias_grad, int64_t M, int N, void* workspace, int64_t workspace_bytes,
                     mg_stream_t stream);
int64_t mg_linear_wgrad_workspace_bytes(int64_t M, int N, int K);
/* host-only: the launch plan of mg_linear_wgrad_bf16 (tile_k, n_tiles, k_tiles, splits, blocks_per_split, n_fblocks, tile_rows,
 * pair) for the shapes of nn.Linear's weight gradient (README.rst:65-73); used by the CPU tests to check the plan's invariants */
int mg_linear_wgrad_plan(int64_t M, int N, int K, int64_t* out8);
int mg_linear_wgrad_bf16(const void* g, int64_t ldg, const void* x, int64_t ldx, float* grad_w, int64_t ldw, int64_t M, int N,
                         int K, void* workspace, int64_t workspace_bytes, mg_stream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * K8  batched MLPG ("next" row 1 of the scope table) -- replaces viz.synthesis.MLPG (morgana/viz/synthesis.py:79-180)
 *     for windows that reach at most one frame to either side (the reference's defaults [1], [-0.5, 0, 0.5], [1, -2, 1],
 *     synthesis.py:122-127, or any other 1 .. 4 such windows): for every utterance and static dimension, the pentadiagonal
 *     system (sum_k W_k^T diag(1/var_k) W_k) c = sum_k W_k^T (mean_k / var_k) is built and solved in fp64 over n + 2 * padding
 *     edge-replicated frames; c without the padding is the trajectory.
 *
 * means      (B, T, n_windows * feat_dim) fp32, layout [window 0 | window 1 | ...] (static | delta | delta-delta for the
 *            defaults); strides in elements, inner contiguous.
 * windows    HOST array of n_windows x 3 doubles, window k's coefficients at frame offsets (-1, 0, +1) (0 where the window
 *            does not reach), or NULL for the reference's three default windows (n_windows is then ignored).
 * variances  same layout; per frame (v_sb, v_st as for means), per utterance (v_st = 0) or global (v_sb = v_st = 0).
 * seq_len    (B,) int64 or NULL (all T frames).  out (B, T, feat_dim) fp32; frames past seq_len are zero.
 * workspace  mg_mlpg_workspace_bytes(B, T, feat_dim, padding) bytes of device memory (no initialisation needed).
 */
int64_t mg_mlpg_workspace_bytes(int B, int64_t T, int feat_dim, int padding);
int mg_mlpg_f32(const float* means, int64_t m_sb, int64_t m_st, const float* variances, int64_t v_sb, int64_t v_st,
                const int64_t* seq_len, float* out, int64_t o_sb, int64_t o_st, int B, int64_t T, int feat_dim, int padding,
                const double* windows, int n_windows, void* workspace, int64_t workspace_bytes, mg_stream_t stream);

/* losses.KLD_standard_normal (morgana/losses.py:64-67): loss = mean over rows of -0.5 * sum_d (1 + log_variance - mean^2 -
 * exp(log_variance)), mean / log_variance contiguous (rows, latent_dim) fp32.  loss: one float on the device (or NULL).
 * grad_mean / grad_log_variance: both NULL, or outputs of the operands' shape = d loss / d operand times *grad_scale_dev
 * (NULL: 1).  workspace: mg_kld_workspace_bytes(rows * latent_dim) bytes, 8-byte aligned; fixed summation order. */
int64_t mg_kld_workspace_bytes(int64_t n);
int mg_kld_standard_normal_f32(const float* mean, const float* log_variance, int64_t rows, int latent_dim, float* loss,
                               float* grad_mean, float* grad_log_variance, const float* grad_scale_dev, void* workspace,
                               int64_t workspace_bytes, mg_stream_t stream);

/* utils.both_voiced_mask (morgana/utils.py:169-172): out[i] = 1 when every feature is non-zero at i (~torch.eq(x, 0.), so NaN
 * counts as voiced), else 0.  features: HOST array of n_features (<= 8) DEVICE pointers to contiguous fp32 tensors of n elements. */
int mg_both_nonzero_u8(const float* const* features, int n_features, int64_t n, unsigned char* out, mg_stream_t stream);

/* fp32 -> bf16 row conversion with K padding: feeds K7 from the fp32 frame-rate features and weights of the example models'
 * nn.Linear layers (README.rst:65-73, models/RNN_SPSS.py:33-41); pads K to ld_out with 0. */
int mg_cast_pad_bf16(const float* x, int64_t ldx, void* out, int64_t ld_out, int64_t rows, int K, mg_stream_t stream);
/* fp32 weight (N, K), row stride ldw -> its transpose (K, ld_out >= N) in bf16, columns N.. zero: the B operand of the input
 * gradient of nn.Linear, grad_x = g @ W (autograd's mm in `loss.backward()`, experiment_builder.py:470), run as g @ (W^T)^T
 * through mg_linear_bf16.  One small launch per layer and step, instead of a transpose + pad + copy of the bf16 weight. */
int mg_cast_transpose_bf16(const float* w, int64_t ldw, void* out, int64_t ld_out, int N, int K, mg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MORGANA_B200_H_ */
