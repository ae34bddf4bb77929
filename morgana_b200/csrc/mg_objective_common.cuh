// Device helpers shared by the two whole-row objective kernels (mg_objective.cu: chunk-per-CTA, any layout;
// mg_objective_stream.cu: persistent row-stream for contiguous (B, T, D <= 256) tensors).  Column programs as in
// include/morgana_b200.h (mg_column / mg_slot); arithmetic in ATen's rounding order (explicit _rn intrinsics).
#pragma once

#include "mg_common.cuh"
#include "mg_finish.cuh"

namespace mgobj {

__device__ __forceinline__ float bce_value(float p, float y) {
  // ATen binary_cross_entropy: (y - 1) * max(log1p(-p), -100) - y * max(log(p), -100)
  const float log_p = fmaxf(logf(p), -100.f);
  const float log_1mp = fmaxf(log1pf(-p), -100.f);
  return __fsub_rn(__fmul_rn(__fsub_rn(y, 1.f), log_1mp), __fmul_rn(y, log_p));
}

// ---- rare columns (BCE, exp, equality, per-frame root, voiced weighting) ---------------------------------------------
template <bool GRAD>
__device__ __forceinline__ void general_one(const mg_column& col, float pv, float yv, float mask_v, bool has_mask,
                                            const float* p, const float* y, float* g, float w_row, double& loss_acc,
                                            double& metric_acc, double& count_acc) {
  const float d = __fsub_rn(pv, yv);
  const float sq = __fmul_rn(d, d);
  // ---- loss term of this column (+ its gradient) ----
  if (col.loss_kind == MG_RED_SQDIFF) {
    loss_acc += static_cast<double>(sq);
    if (GRAD) __stcs(g, __fmul_rn(__fmul_rn(2.f, d), w_row));
  } else if (col.loss_kind == MG_RED_BCE) {
    loss_acc += static_cast<double>(bce_value(pv, yv));
    if (GRAD) {
      // ATen binary_cross_entropy_backward: (p - y) / max((1 - p) * p, 1e-12)
      const float slope = __fdiv_rn(d, fmaxf(__fmul_rn(__fsub_rn(1.f, pv), pv), 1e-12f));
      __stcs(g, __fmul_rn(slope, w_row));
    }
  } else if (col.loss_kind == MG_RED_ABSDIFF) {
    loss_acc += static_cast<double>(fabsf(d));
    if (GRAD) __stcs(g, __fmul_rn(d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f), w_row));
  } else if (GRAD) {
    __stcs(g, 0.f);
  }
  // ---- metric of this column ----
  if (col.metric_kind != MG_COL_NONE) {
    float v;
    if (col.metric_kind == MG_RED_SQDIFF) v = sq;
    else if (col.metric_kind == MG_RED_ABSDIFF) v = fabsf(d);
    else if (col.metric_kind == MG_RED_SQDIFF_EXP) {
      const float e = __fsub_rn(expf(yv), expf(pv));   // reference metrics.py:631-632 then :607
      v = __fmul_rn(e, e);
    } else if (col.metric_kind == MG_RED_EQ) {
      v = ((pv > 0.5f) == (yv != 0.f)) ? 1.f : 0.f;     // models/RNN_SPSS.py:122, 127
    } else {   // MG_RED_ROOT_SQDIFF: this column leads a group of `width` columns (metrics.py:657-662)
      float acc = sq;
      for (int k = 1; k < col.width; ++k) {
        const float dk = __fsub_rn(__ldg(y + k), __ldg(p + k));
        acc = __fadd_rn(acc, __fmul_rn(dk, dk));
      }
      v = sqrtf(acc);
    }
    if (has_mask) {
      const float voiced = mask_v > 0.5f ? 1.f : 0.f;
      v = __fmul_rn(v, voiced);
      count_acc += static_cast<double>(voiced);
    }
    metric_acc += static_cast<double>(v);
  }
}

// One column program as per-thread scalars.
struct LaneProgram {
  int loss_slot, metric_slot;
  bool use_loss, use_metric;   // simple column with a loss / metric slot (special columns never accumulate here)
  bool loss_sq, metric_sq;
  float w_row;                 // gradient factor loss_weight * scale / (n_b * B); 0 for columns without a simple loss
};

__device__ __forceinline__ bool column_is_simple(const mg_column& col) {
  const bool simple_loss = col.loss_kind == MG_COL_NONE || col.loss_kind == MG_RED_SQDIFF || col.loss_kind == MG_RED_ABSDIFF;
  const bool simple_metric = col.mask_col == MG_COL_NONE && (col.metric_kind == MG_COL_NONE ||
                             col.metric_kind == MG_RED_SQDIFF || col.metric_kind == MG_RED_ABSDIFF);
  return simple_loss && simple_metric;
}

__device__ __forceinline__ LaneProgram make_lane(const mg_column& col, double inv_rows) {
  LaneProgram lp;
  const bool simple = column_is_simple(col);
  lp.loss_slot = col.loss_slot;
  lp.metric_slot = col.metric_slot;
  lp.use_loss = simple && col.loss_kind != MG_COL_NONE;
  lp.use_metric = simple && col.metric_kind != MG_COL_NONE;
  lp.loss_sq = col.loss_kind == MG_RED_SQDIFF;
  lp.metric_sq = col.metric_kind == MG_RED_SQDIFF;
  lp.w_row = lp.use_loss ? static_cast<float>(static_cast<double>(col.loss_weight) * inv_rows) : 0.f;
  return lp;
}

__device__ __forceinline__ LaneProgram idle_lane() {
  LaneProgram lp;
  lp.loss_slot = lp.metric_slot = 0;
  lp.use_loss = lp.use_metric = lp.loss_sq = lp.metric_sq = false;
  lp.w_row = 0.f;
  return lp;
}

__device__ __forceinline__ float simple_slope(bool sq, float d) {
  return sq ? __fmul_rn(2.f, d) : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
}

}  // namespace mgobj

// Arguments of the persistent row-stream form (host side); returns MG_OK, an error, or MG_STREAM_NOT_APPLICABLE when the
// tensors do not fit its preconditions (the caller then runs the chunk-per-CTA kernel).
#define MG_STREAM_NOT_APPLICABLE 1
struct MgObjectiveArgs {
  const float* pred; const float* target; float* grad; const float* grad_scale_dev;
  const mg_column* cols; const mg_slot* slots; const int64_t* seq_len;
  int64_t p_sb, p_st, t_sb, t_st, g_sb, g_st, T;
  int D, B, n_slots;
  void* workspace; int64_t workspace_bytes;
};
int mg_objective_stream_launch(const MgObjectiveArgs& args, cudaStream_t stream);
