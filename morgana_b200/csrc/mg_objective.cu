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

template <bool GRAD>
__global__ void __launch_bounds__(1024)
masked_objective_kernel(const __grid_constant__ ObjectiveParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_loss = reinterpret_cast<double*>(smem_raw);         // [blockDim.x]
  double* s_metric = s_loss + blockDim.x;                        // [blockDim.x]
  double* s_count = s_metric + blockDim.x;                       // [blockDim.x]
  __shared__ double s_red[96];
  __shared__ bool s_is_last;

  const int c = threadIdx.x, b = blockIdx.y, chunk = blockIdx.x;
  const int D = prm.D;
  const int64_t T = prm.T;
  const int64_t n_b = mg_valid_frames(prm.seq_len, b, T);
  const int64_t r0 = static_cast<int64_t>(chunk) * prm.rows_per_cta;
  const int64_t r1 = min(r0 + prm.rows_per_cta, T);
  const int64_t valid_end = min(r1, n_b);

  int loss_kind = MG_COL_NONE, metric_kind = MG_COL_NONE, mask_col = MG_COL_NONE, width = 1;
  float loss_weight = 0.f;
  if (c < D) {
    const mg_column col = prm.cols[c];
    loss_kind = col.loss_kind;
    metric_kind = col.metric_kind;
    mask_col = col.mask_col;
    width = col.width;
    loss_weight = col.loss_weight;
  }
  double loss_acc = 0., metric_acc = 0., count_acc = 0.;

  if (c < D && (r0 < valid_end || GRAD)) {
    const float* p_base = prm.pred + b * prm.p_sb + c;
    const float* y_base = prm.target + b * prm.t_sb + c;
    float* g_base = GRAD ? prm.grad + b * prm.g_sb + c : nullptr;
    float w_row = 0.f;
    if (GRAD) {
      double scale = static_cast<double>(loss_weight);
      if (prm.grad_scale_dev != nullptr) scale *= static_cast<double>(__ldg(prm.grad_scale_dev));
      w_row = static_cast<float>(scale / (static_cast<double>(n_b) * prm.B));
    }

    for (int64_t r = r0; r < valid_end; r += kObjUnroll) {
      float p[kObjUnroll], y[kObjUnroll];
#pragma unroll
      for (int u = 0; u < kObjUnroll; ++u) {
        if (r + u < valid_end) {
          p[u] = __ldcs(p_base + (r + u) * prm.p_st);
          y[u] = __ldcs(y_base + (r + u) * prm.t_st);
        }
      }
#pragma unroll
      for (int u = 0; u < kObjUnroll; ++u) {
        if (r + u >= valid_end) break;
        const float d = __fsub_rn(p[u], y[u]);
        const float sq = __fmul_rn(d, d);
        // ---- loss term of this column (+ its gradient) ----
        if (loss_kind == MG_RED_SQDIFF) {
          loss_acc += static_cast<double>(sq);
          if (GRAD) __stcs(g_base + (r + u) * prm.g_st, __fmul_rn(__fmul_rn(2.f, d), w_row));
        } else if (loss_kind == MG_RED_BCE) {
          loss_acc += static_cast<double>(bce_value(p[u], y[u]));
          if (GRAD) {
            // ATen binary_cross_entropy_backward: (p - y) / max((1 - p) * p, 1e-12)
            const float g = __fdiv_rn(d, fmaxf(__fmul_rn(__fsub_rn(1.f, p[u]), p[u]), 1e-12f));
            __stcs(g_base + (r + u) * prm.g_st, __fmul_rn(g, w_row));
          }
        } else if (loss_kind == MG_RED_ABSDIFF) {
          loss_acc += static_cast<double>(fabsf(d));
          if (GRAD) __stcs(g_base + (r + u) * prm.g_st, __fmul_rn(d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f), w_row));
        } else if (GRAD) {
          __stcs(g_base + (r + u) * prm.g_st, 0.f);
        }
        // ---- metric of this column ----
        if (metric_kind != MG_COL_NONE) {
          float v;
          if (metric_kind == MG_RED_SQDIFF) v = sq;
          else if (metric_kind == MG_RED_ABSDIFF) v = fabsf(d);
          else if (metric_kind == MG_RED_SQDIFF_EXP) {
            const float e = __fsub_rn(expf(y[u]), expf(p[u]));   // metrics.py:631-632 then :607
            v = __fmul_rn(e, e);
          } else if (metric_kind == MG_RED_EQ) {
            v = ((p[u] > 0.5f) == (y[u] != 0.f)) ? 1.f : 0.f;     // RNN_SPSS.py:122, 127
          } else {   // MG_RED_ROOT_SQDIFF: this column leads a group of `width` columns (metrics.py:657-662)
            float acc = sq;
            for (int k = 1; k < width; ++k) {
              const float dk = __fsub_rn(__ldg(y_base + (r + u) * prm.t_st + k), __ldg(p_base + (r + u) * prm.p_st + k));
              acc = __fadd_rn(acc, __fmul_rn(dk, dk));
            }
            v = sqrtf(acc);
          }
          if (mask_col != MG_COL_NONE) {
            const float voiced = __ldg(prm.pred + b * prm.p_sb + (r + u) * prm.p_st + mask_col) > 0.5f ? 1.f : 0.f;
            v = __fmul_rn(v, voiced);
            count_acc += static_cast<double>(voiced);
          }
          metric_acc += static_cast<double>(v);
        }
      }
    }
    if (GRAD) {   // padding rows of this chunk: the gradient is defined (zero) over the whole (B, T, D) tensor
      for (int64_t r = max(r0, n_b); r < r1; ++r) __stcs(g_base + r * prm.g_st, 0.f);
    }
  }

  // ---- per-CTA, per-slot partials in a fixed order -------------------------------------------------------------
  if (r0 < valid_end) {   // CTA-uniform
    s_loss[c] = loss_acc;
    s_metric[c] = metric_acc;
    s_count[c] = count_acc;
    __syncthreads();
    if (c < prm.n_slots) {
      double s = 0., n = 0.;
      for (int k = 0; k < D; ++k) {
        const mg_column col = prm.cols[k];
        if (col.loss_kind != MG_COL_NONE && col.loss_slot == c) s += s_loss[k];
        if (col.metric_kind != MG_COL_NONE && col.metric_slot == c) { s += s_metric[k]; n += s_count[k]; }
      }
      prm.partials[(static_cast<int64_t>(c) * prm.B + b) * kMgMaxChunks + chunk] = make_double2(s, n);
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
  if (rows < 1) rows = 1;
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

  int threads = ((D + 31) / 32) * 32;
  if (threads < 32 * ((n_slots + 31) / 32)) threads = 32 * ((n_slots + 31) / 32);
  const size_t smem = static_cast<size_t>(threads) * 3 * sizeof(double);
  dim3 grid(static_cast<unsigned>(n_chunks), static_cast<unsigned>(B));
  if (grad != nullptr) masked_objective_kernel<true><<<grid, threads, smem, stream>>>(prm);
  else masked_objective_kernel<false><<<grid, threads, smem, stream>>>(prm);
  MG_LAUNCH_OK();
  return MG_OK;
}
