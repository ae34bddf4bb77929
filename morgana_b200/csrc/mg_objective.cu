// K4b -- whole-row masked objective: every loss term and streaming metric of an acoustic model in one pass.
//
// Replaces the body of LSTMAcousticModel.loss (reference models/RNN_SPSS.py:120-139): metrics.accumulate x4
// (morgana/metrics.py:383-394, 597-609, 630-634, 657-665, 690-694) + losses.mse x3 + losses.bce (morgana/losses.py:29-56),
// all of which slice the same (B, T, D) prediction / target pair by column group.
//
// Mapping.  CTA (chunk, b) owns rows [chunk*R, (chunk+1)*R) of utterance b and only touches those below n_b.
// Thread c owns COLUMN c for all of the CTA's rows, so its program (loss kind, metric kind, slots, weight) is a handful
// of per-thread constants, a row is read by consecutive threads (coalesced 4-byte loads, 8 rows in flight per thread)
// and every byte of pred / target / grad crosses HBM exactly once.  Only the warp that straddles a group boundary
// diverges.  Rare cross-column needs (the voiced mask of the F0 metric, the feature-axis sum of Distortion) are extra
// loads that hit L1.
//
// Determinism: per-thread fp64 accumulators -> shared memory -> one thread per slot sums the threads of its slot in
// column order -> per-CTA slot in the workspace -> last CTA (integer ticket) combines in index order (mg_finish.cuh).
#include <string.h>

#include "mg_common.cuh"
#include "mg_finish.cuh"

namespace {

constexpr int kObjUnroll = 8;       // rows in flight per thread
constexpr int64_t kObjTargetElems = 24576;   // elements of one operand per CTA (~96 KB)

struct ObjectiveParams {
  MgFinishSlot slots[MG_MAX_TERMS];
  const float* pred;
  const float* target;
  float* grad;
  const float* grad_scale_dev;
  const mg_column* cols;
  const int64_t* seq_len;
  double2* partials;
  unsigned int* ticket;
  int64_t p_sb, p_st, t_sb, t_st, g_sb, g_st, T;
  int D, B, n_slots, rows_per_cta;
};

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

// MAXT = 256: D <= 256 (the acoustic layouts: 187 / 199 columns); MAXT = 1024: wider tensors.
template <bool GRAD, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT == 256 ? 4 : 1)
masked_objective_kernel(const __grid_constant__ ObjectiveParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_loss = reinterpret_cast<double*>(smem_raw);         // [blockDim.x] per-column loss partial
  double* s_metric = s_loss + blockDim.x;                        // [blockDim.x] per-column metric partial
  double* s_count = s_metric + blockDim.x;                       // [blockDim.x] per-column weight count
  __shared__ double s_red[96];
  __shared__ unsigned s_special[32];                             // per warp: lanes whose column needs general_one()
  __shared__ bool s_is_last;

  const int c = threadIdx.x, b = blockIdx.y, chunk = blockIdx.x;
  const int warp = c >> 5, lane = c & 31, n_warps = blockDim.x >> 5;
  const int D = prm.D;
  const int64_t T = prm.T;
  const int64_t n_b = mg_valid_frames(prm.seq_len, b, T);
  const int64_t r0 = static_cast<int64_t>(chunk) * prm.rows_per_cta;
  const int64_t r1 = min(r0 + prm.rows_per_cta, T);
  const int64_t valid_end = min(r1, n_b);
  const int64_t n_rows = max(static_cast<int64_t>(0), valid_end - r0);
  const int64_t p_st = prm.p_st, t_st = prm.t_st, g_st = prm.g_st;
  if (n_rows == 0 && !GRAD) {   // CTA-uniform: nothing to read, nothing to write
    if (!mg_take_ticket(prm.ticket, &s_is_last)) return;
    mg_finish(prm.slots, prm.n_slots, prm.seq_len, prm.B, T, prm.partials, prm.ticket, s_red);
    return;
  }

  // Column classes.  "Simple": squared / absolute error into a loss slot and / or a metric slot (the mcep / lf0 / bap
  // streams) -- a few predicated ops per element, handled by the column's own thread in the streaming loop below.
  // "Special" (BCE, exp, equality, per-frame root, voiced weighting; 3 of 187 columns in the acoustic layout) would make
  // its warp a straggler, so those columns are processed afterwards by the WHOLE CTA, one row per thread.
  mg_column col;
  col.loss_kind = col.metric_kind = MG_COL_NONE;
  col.loss_slot = col.metric_slot = 0;
  col.mask_col = MG_COL_NONE;
  col.width = 1;
  col.loss_weight = 0.f;
  if (c < D) col = prm.cols[c];
  const bool simple_loss = col.loss_kind == MG_COL_NONE || col.loss_kind == MG_RED_SQDIFF || col.loss_kind == MG_RED_ABSDIFF;
  const bool simple_metric = col.mask_col == MG_COL_NONE && (col.metric_kind == MG_COL_NONE ||
                             col.metric_kind == MG_RED_SQDIFF || col.metric_kind == MG_RED_ABSDIFF);
  const bool simple = simple_loss && simple_metric;
  const unsigned special_lanes = __ballot_sync(MG_FULL_MASK, c < D && !simple);
  if (lane == 0) s_special[warp] = special_lanes;

  double scale = 1.;
  if (GRAD && prm.grad_scale_dev != nullptr) scale = static_cast<double>(__ldg(prm.grad_scale_dev));
  const double inv_rows = scale / (static_cast<double>(n_b) * prm.B);   // shared factor of the gradient in this utterance

  double loss_acc = 0., metric_acc = 0.;
  if (c < D && simple) {
    const bool has_loss = col.loss_kind != MG_COL_NONE, has_metric = col.metric_kind != MG_COL_NONE;
    const bool loss_sq = col.loss_kind == MG_RED_SQDIFF, metric_sq = col.metric_kind == MG_RED_SQDIFF;
    const float w_row = static_cast<float>(static_cast<double>(col.loss_weight) * inv_rows);
    const float* p = prm.pred + b * prm.p_sb + r0 * p_st + c;
    const float* y = prm.target + b * prm.t_sb + r0 * t_st + c;
    float* g = GRAD ? prm.grad + b * prm.g_sb + r0 * g_st + c : nullptr;
    for (int64_t r = 0; r < n_rows; r += kObjUnroll) {
      const int live = static_cast<int>(min(static_cast<int64_t>(kObjUnroll), n_rows - r));   // warp-uniform
      float pv[kObjUnroll], yv[kObjUnroll];
#pragma unroll
      for (int u = 0; u < kObjUnroll; ++u) {   // all loads of up to 8 rows first
        pv[u] = u < live ? __ldcs(p + u * p_st) : 0.f;
        yv[u] = u < live ? __ldcs(y + u * t_st) : 0.f;
      }
      float l_part = 0.f, m_part = 0.f;   // rows summed in fp32 in row order, folded into fp64 once per 8
#pragma unroll
      for (int u = 0; u < kObjUnroll; ++u) {
        const float d = __fsub_rn(pv[u], yv[u]);   // 0 for the dead rows of a partial group: adds nothing
        const float sq = __fmul_rn(d, d), ab = fabsf(d);
        l_part = __fadd_rn(l_part, loss_sq ? sq : ab);
        m_part = __fadd_rn(m_part, metric_sq ? sq : ab);
        if (GRAD && u < live) {
          const float slope = loss_sq ? __fmul_rn(2.f, d) : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
          __stcs(g + u * g_st, has_loss ? __fmul_rn(slope, w_row) : 0.f);
        }
      }
      if (has_loss) loss_acc += static_cast<double>(l_part);
      if (has_metric) metric_acc += static_cast<double>(m_part);
      p += kObjUnroll * p_st;
      y += kObjUnroll * t_st;
      if (GRAD) g += kObjUnroll * g_st;
    }
  }
  if (GRAD && c < D) {   // padding rows of this chunk: the gradient is defined (zero) over the whole (B, T, D) tensor
    float* gz = prm.grad + b * prm.g_sb + c;
    for (int64_t r = max(r0, n_b); r < r1; ++r) __stcs(gz + r * g_st, 0.f);
  }
  s_loss[c] = loss_acc;
  s_metric[c] = metric_acc;
  s_count[c] = 0.;
  __syncthreads();

  // ---- special columns: the whole CTA shares each one, a row per thread; CTA-wide sum in a fixed order ----------------
  if (n_rows > 0) {
    for (int w = 0; w < n_warps; ++w) {
      unsigned todo = s_special[w];
      while (todo) {   // CTA-uniform loop
        const int k = w * 32 + __ffs(todo) - 1;
        todo &= todo - 1;
        const mg_column sc = prm.cols[k];
        const float w_row = static_cast<float>(static_cast<double>(sc.loss_weight) * inv_rows);
        const bool has_mask = sc.mask_col != MG_COL_NONE;
        double l = 0., m = 0., n = 0.;
        for (int64_t r = c; r < n_rows; r += blockDim.x) {
          const float* p = prm.pred + b * prm.p_sb + (r0 + r) * p_st + k;
          const float* y = prm.target + b * prm.t_sb + (r0 + r) * t_st + k;
          float* g = GRAD ? prm.grad + b * prm.g_sb + (r0 + r) * g_st + k : nullptr;
          const float mask_v = has_mask ? __ldg(prm.pred + b * prm.p_sb + (r0 + r) * p_st + sc.mask_col) : 1.f;
          general_one<GRAD>(sc, __ldg(p), __ldg(y), mask_v, has_mask, p, y, g, w_row, l, m, n);
        }
        l = mg_warp_sum(l);
        m = mg_warp_sum(m);
        n = mg_warp_sum(n);
        if (lane == 0) { s_red[warp] = l; s_red[32 + warp] = m; s_red[64 + warp] = n; }
        __syncthreads();
        if (c == 0) {
          double ls = 0., ms = 0., ns = 0.;
          for (int i = 0; i < n_warps; ++i) { ls += s_red[i]; ms += s_red[32 + i]; ns += s_red[64 + i]; }
          s_loss[k] = ls;
          s_metric[k] = ms;
          s_count[k] = ns;
        }
        __syncthreads();
      }
    }

    // ---- per-CTA, per-slot partials: warp s sums the columns of slot s, lanes striding the columns, fixed order ------
    for (int slot = warp; slot < prm.n_slots; slot += n_warps) {
      double s = 0., n = 0.;
      for (int k = lane; k < D; k += 32) {
        const mg_column kc = prm.cols[k];
        if (kc.loss_kind != MG_COL_NONE && kc.loss_slot == slot) s += s_loss[k];
        if (kc.metric_kind != MG_COL_NONE && kc.metric_slot == slot) { s += s_metric[k]; n += s_count[k]; }
      }
      s = mg_warp_sum(s);
      n = mg_warp_sum(n);
      if (lane == 0) prm.partials[(static_cast<int64_t>(slot) * prm.B + b) * kMgMaxChunks + chunk] = make_double2(s, n);
    }
  }

  if (!mg_take_ticket(prm.ticket, &s_is_last)) return;
  mg_finish(prm.slots, prm.n_slots, prm.seq_len, prm.B, T, prm.partials, prm.ticket, s_red);
}

}  // namespace

extern "C" int mg_masked_objective_f32(const float* pred, int64_t p_sb, int64_t p_st, const float* target, int64_t t_sb,
                                       int64_t t_st, float* grad, int64_t g_sb, int64_t g_st, const float* grad_scale_dev,
                                       const mg_column* cols, int D, const mg_slot* slots, int n_slots,
                                       const int64_t* seq_len, int B, int64_t T, void* workspace, int64_t workspace_bytes,
                                       mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(D >= 1 && D <= 1024, "mg_masked_objective_f32: D=%d outside [1, 1024]", D);
  MG_REQUIRE(n_slots >= 1 && n_slots <= MG_MAX_TERMS, "mg_masked_objective_f32: n_slots=%d outside [1, %d]", n_slots, MG_MAX_TERMS);
  MG_REQUIRE(B >= 1 && B <= 65535, "mg_masked_objective_f32: B=%d outside [1, 65535]", B);
  MG_REQUIRE(T >= 0 && T < (int64_t(1) << 31), "mg_masked_objective_f32: bad T");
  MG_REQUIRE(cols != nullptr && slots != nullptr && workspace != nullptr, "mg_masked_objective_f32: NULL buffer");
  MG_REQUIRE((pred != nullptr && target != nullptr) || T == 0, "mg_masked_objective_f32: NULL operand");
  MG_REQUIRE(workspace_bytes >= mg_masked_reduce_workspace_bytes(n_slots, B, T), "mg_masked_objective_f32: workspace too small");
  MG_REQUIRE(mg_aligned(workspace, 16), "mg_masked_objective_f32: workspace must be 16-byte aligned");

  ObjectiveParams prm;
  memset(&prm, 0, sizeof(prm));
  // Rows per CTA: ~96 KB per operand, but at least 4 CTAs per SM in the grid and at most kMgMaxChunks chunks.
  int64_t rows = kObjTargetElems / D;
  if (rows < kObjUnroll) rows = kObjUnroll;
  if (rows > T) rows = T;
  const int64_t sms = mg_cached_sm_count();
  while (rows > 2 * kObjUnroll && static_cast<int64_t>(B) * ((T + rows - 1) / rows) < 4 * sms) rows = (rows + 1) / 2;
  const int64_t min_rows = (T + kMgMaxChunks - 1) / kMgMaxChunks;
  if (rows < min_rows) rows = min_rows;
  rows = ((rows + kObjUnroll - 1) / kObjUnroll) * kObjUnroll;   // whole groups of 8 rows except at an utterance's end
  const int n_chunks = T > 0 ? static_cast<int>((T + rows - 1) / rows) : 1;

  for (int i = 0; i < n_slots; ++i) {
    MG_REQUIRE(slots[i].result != nullptr && mg_aligned(slots[i].result, 16), "mg_masked_objective_f32: slot %d needs a 16-byte aligned result record", i);
    MG_REQUIRE(slots[i].D >= 1, "mg_masked_objective_f32: slot %d has D=%d", i, slots[i].D);
    MgFinishSlot& sl = prm.slots[i];
    sl.result = slots[i].result;
    sl.D = slots[i].D;
    sl.rows_per_cta = static_cast<int>(rows);
    sl.n_chunks = n_chunks;
    sl.per_frame = slots[i].per_frame;
    sl.weighted = slots[i].weighted;
    sl.accumulate = slots[i].accumulate;
    sl.in_total = slots[i].in_total;
    sl.weight = slots[i].weight;
  }
  prm.pred = pred; prm.target = target; prm.grad = grad; prm.grad_scale_dev = grad_scale_dev;
  prm.cols = cols; prm.seq_len = seq_len;
  prm.ticket = static_cast<unsigned int*>(workspace);
  prm.partials = reinterpret_cast<double2*>(static_cast<unsigned char*>(workspace) + 256);
  prm.p_sb = p_sb; prm.p_st = p_st; prm.t_sb = t_sb; prm.t_st = t_st; prm.g_sb = g_sb; prm.g_st = g_st; prm.T = T;
  prm.D = D; prm.B = B; prm.n_slots = n_slots; prm.rows_per_cta = static_cast<int>(rows);

  const int threads = ((D + 31) / 32) * 32;
  const size_t smem = static_cast<size_t>(threads) * 3 * sizeof(double);
  dim3 grid(static_cast<unsigned>(n_chunks), static_cast<unsigned>(B));
  if (threads <= 256) {
    if (grad != nullptr) masked_objective_kernel<true, 256><<<grid, threads, smem, stream>>>(prm);
    else masked_objective_kernel<false, 256><<<grid, threads, smem, stream>>>(prm);
  } else {
    if (grad != nullptr) masked_objective_kernel<true, 1024><<<grid, threads, smem, stream>>>(prm);
    else masked_objective_kernel<false, 1024><<<grid, threads, smem, stream>>>(prm);
  }
  MG_LAUNCH_OK();
  return MG_OK;
}
