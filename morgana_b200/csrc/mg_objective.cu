// K4b -- whole-row masked objective: every loss term and streaming metric of an acoustic model in one pass.
//
// Replaces the body of LSTMAcousticModel.loss (reference models/RNN_SPSS.py:120-139): metrics.accumulate x4
// (morgana/metrics.py:383-394, 597-609, 630-634, 657-665, 690-694) + losses.mse x3 + losses.bce (morgana/losses.py:29-56),
// all of which slice the same (B, T, D) prediction / target pair by column group.
//
// Mapping.  CTA (chunk, b) owns rows [chunk*R, (chunk+1)*R) of utterance b and only touches those below n_b -- one
// contiguous span of n_rows*D floats in pred / target / grad.
//
//   "Simple" columns (squared / absolute error into a loss slot and / or a metric slot: the mcep / lf0 / bap streams,
//   184 of 187 columns) are streamed through shared memory by the TMA engine: the span is cut into stages of 8 rows
//   (8*D floats, a multiple of 16 bytes for any D); one elected thread keeps a ring of stages in flight with
//   cp.async.bulk global->shared (SASS UBLKCP) completing on mbarriers, so the bytes in flight live in shared memory,
//   not in registers (2 stages x 5 co-resident CTAs per SM at D = 187).  Thread t reads positions t, t + D, t + 2D, ... of every stage: always
//   the same column, so its program is a few per-thread constants, shared-memory reads are conflict-free, and each byte
//   of pred / target crosses HBM once.  The gradient is written straight from registers, coalesced.
//   (Fallback when pred / target / grad disagree on 16-byte alignment or rows are strided: thread = column, 8 rows in flight.)
//
//   "Special" columns (BCE, exp, equality, per-frame root, voiced weighting; 3 of 187) carry ~100 instructions per
//   element and would make their warp a 16x straggler.  While a stage sits in shared memory one warp copies the special
//   columns' operands (and their mask columns) into a small side buffer; after the stream the WHOLE CTA evaluates them,
//   one row per thread, with a single CTA-wide reduction.
//
// Determinism: per-thread fp64 accumulators -> per-slot CTA sum (entries in index order, fixed shuffle tree) -> special
// columns added in column order -> per-CTA slot in the workspace -> the last CTA of each utterance folds its chunks, the last
// of those combines the utterances (integer tickets only; mg_finish.cuh).
#include <stdlib.h>
#include <string.h>

#include "mg_objective_common.cuh"

namespace {

using namespace mgobj;

constexpr int kObjUnroll = 8;       // rows in flight per thread (column-per-thread fallback)
constexpr int kStageRows = 8;       // rows per shared-memory stage (8*D floats: a multiple of 16 bytes for every D)
constexpr int kMaxStages = 8;       // ring depth of the bulk-load pipeline
constexpr int kMaxSpecial = 4;       // special columns served from the shared-memory side buffer
constexpr int kSpecialInfo = 1024;   // ballot positions (32 words x 32 lanes)
constexpr int kMaxCapture = 8;       // columns copied to the side buffer (specials, their mask columns, root groups)
constexpr int kLanes = 2;           // column programs per thread: the streamed column + one edge element (head / tail)
constexpr int64_t kObjTargetElems = 24576;   // elements of one operand per CTA (~96 KB): 131 rows at D = 187.  200-row chunks
                                             // are 3 % faster in isolation (0.148 vs 0.153 ms) but leave more dirty lines in L2 for
                                             // the kernel that follows: in the benchmark step K2 then takes 0.149 instead of 0.146 ms

struct ObjectiveParams {
  MgFinishSlot slots[MG_MAX_TERMS];
  const float* pred;
  const float* target;
  float* grad;
  const float* grad_scale_dev;
  const mg_column* cols;
  const int64_t* seq_len;
  MgWorkspace ws;
  int64_t p_sb, p_st, t_sb, t_st, g_sb, g_st, T;
  int D, B, n_slots, rows_per_cta;
  int side_offset;  // byte offset of the side buffer in dynamic shared memory
  int n_stages;     // ring depth of the staged stream (0: staged path not applicable, use the column-per-thread loop)
};

// MAXT = 256: D <= 256 (the acoustic layouts: 187 / 199 columns), five CTAs per SM; MAXT = 1024: wider tensors.
template <bool GRAD, int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT == 256 ? 4 : 1)
masked_objective_kernel(const __grid_constant__ ObjectiveParams prm) {
  // dynamic: [ring of (pred stage | target stage)] [side buffer: kMaxCapture x rows_per_cta x (pred, target)]
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[kMaxStages];
  __shared__ double s_red[3 * kMaxSpecial * 32];
  __shared__ double s_slot_sum[MG_MAX_TERMS], s_slot_cnt[MG_MAX_TERMS];
  __shared__ unsigned s_special[32];          // per 32 columns: which need general_one()
  __shared__ int s_cap_col[kMaxCapture];      // columns copied to the side buffer during the stream
  __shared__ int s_sp_col[kMaxSpecial];       // special columns served from the side buffer
  constexpr int kInfo = MAXT <= 256 ? 256 : kSpecialInfo;   // ballot positions in use: D rounded up to a warp
  __shared__ short s_sp_mask[kInfo], s_sp_width[kInfo];   // per special column, by ballot position
  __shared__ int s_n_cap, s_n_sp, s_n_special;
  __shared__ bool s_is_last;

  const int tid = threadIdx.x, b = blockIdx.y, chunk = blockIdx.x;
  const int warp = tid >> 5, lane = tid & 31, n_warps = blockDim.x >> 5;
  const int D = prm.D;
  const int64_t T = prm.T;
  const int64_t n_b = mg_valid_frames(prm.seq_len, b, T);
  const int64_t r0 = static_cast<int64_t>(chunk) * prm.rows_per_cta;
  const int64_t r1 = min(r0 + prm.rows_per_cta, T);
  const int64_t valid_end = min(r1, n_b);
  const int64_t n_rows = max(static_cast<int64_t>(0), valid_end - r0);
  const int64_t p_st = prm.p_st, t_st = prm.t_st, g_st = prm.g_st;
  const int rows_cap = prm.rows_per_cta;
  float* side = reinterpret_cast<float*>(smem_raw + prm.side_offset);   // [cap][rows_cap][2]

  double scale = 1.;
  if (GRAD && prm.grad_scale_dev != nullptr) scale = static_cast<double>(__ldg(prm.grad_scale_dev));
  const double inv_rows = scale / (static_cast<double>(n_b) * prm.B);   // shared factor of the gradient in this utterance

  const float* p_chunk = prm.pred + b * prm.p_sb + r0 * p_st;
  const float* y_chunk = prm.target + b * prm.t_sb + r0 * t_st;
  float* g_chunk = GRAD ? prm.grad + b * prm.g_sb + r0 * g_st : nullptr;

  // ---- geometry of the staged stream, and its first loads: issued before anything else so that the column-program
  // bookkeeping below runs under their latency -------------------------------------------------------------------------
  const uintptr_t mis = reinterpret_cast<uintptr_t>(p_chunk) & 15;
  const bool staged = n_rows > 0 && prm.n_stages > 0 && p_st == D && t_st == D && (!GRAD || g_st == D) &&
                      (reinterpret_cast<uintptr_t>(y_chunk) & 15) == mis;   // CTA-uniform
  const int64_t total = n_rows * D;                                   // floats in this CTA's span
  const int64_t head = min(total, static_cast<int64_t>(((16 - mis) & 15) >> 2));
  const int64_t body = ((total - head) >> 2) << 2;                    // floats that travel through shared memory
  const int stage_elems = kStageRows * D;
  const int n_iter = static_cast<int>((body + stage_elems - 1) / stage_elems);
  const int ring = prm.n_stages;
  float* s_stage = reinterpret_cast<float*>(smem_raw);
  const float* p_body = p_chunk + head;
  const float* y_body = y_chunk + head;
  auto issue = [&](int it) {   // one elected thread: both operands of stage `it` into ring slot it % ring
    const int slot = it % ring;
    const int64_t off = static_cast<int64_t>(it) * stage_elems;
    const uint32_t bytes = static_cast<uint32_t>(min(static_cast<int64_t>(stage_elems), body - off)) * 4u;
    float* dst = s_stage + static_cast<size_t>(slot) * 2 * stage_elems;
    mg_mbar_expect_tx(&s_bar[slot], 2 * bytes);
    mg_bulk_load(dst, p_body + off, bytes, &s_bar[slot]);
    mg_bulk_load(dst + stage_elems, y_body + off, bytes, &s_bar[slot]);
  };
  if (tid == 0 && staged) {
    for (int i = 0; i < ring; ++i) mg_mbar_init(&s_bar[i], 1);
    mg_mbar_fence_init();
    for (int it = 0; it < min(ring, n_iter); ++it) issue(it);
  }

  // ---- which columns are special, and which columns must be captured for them (themselves + their mask columns).
  // Every thread reads its own column's program once (one parallel round of loads) and leaves what thread 0 needs in
  // shared memory, so the list is built without a chain of dependent global loads. --------------------------------------
  for (int c0 = warp * 32; c0 < D; c0 += n_warps * 32) {
    const int k = c0 + lane;
    mg_column col;
    col.loss_kind = col.metric_kind = MG_COL_NONE;
    col.mask_col = MG_COL_NONE;
    col.width = 1;
    if (k < D) col = prm.cols[k];
    const bool special = k < D && !column_is_simple(col);
    const unsigned bits = __ballot_sync(MG_FULL_MASK, special);
    if (lane == 0) s_special[c0 >> 5] = bits;
    if (special) {   // at most a handful of columns: park (mask column, root-group width) by position in the ballot word
      const int slot = (c0 >> 5) * 32 + __popc(bits & ((1u << lane) - 1));
      if (slot < kInfo) { s_sp_mask[slot] = col.mask_col; s_sp_width[slot] = col.metric_kind == MG_RED_ROOT_SQDIFF ? col.width : 1; }
    }
  }
  __syncthreads();
  if (tid == 0) {
    int n_cap = 0, n_sp = 0, n_special = 0;
    auto capture = [&](int col) {   // index in the capture list, or -1 when the list is full
      for (int i = 0; i < n_cap; ++i) if (s_cap_col[i] == col) return i;
      if (n_cap == kMaxCapture) return -1;
      s_cap_col[n_cap] = col;
      return n_cap++;
    };
    for (int w = 0; w < (D + 31) / 32; ++w) {
      unsigned todo = s_special[w];
      int rank_in_word = 0;
      while (todo) {
        const int k = w * 32 + __ffs(todo) - 1;
        todo &= todo - 1;
        const int info = w * 32 + rank_in_word++;
        ++n_special;
        if (n_sp == kMaxSpecial || info >= kInfo) continue;
        const int mask_col = s_sp_mask[info], width = s_sp_width[info];
        const int saved = n_cap;
        bool ok = capture(k) >= 0 && (mask_col == MG_COL_NONE || capture(mask_col) >= 0);
        for (int j = 1; j < width && ok; ++j) ok = capture(k + j) >= 0;
        if (ok) s_sp_col[n_sp++] = k; else n_cap = saved;
      }
    }
    s_n_cap = n_cap;
    s_n_sp = n_sp;
    s_n_special = n_special;
  }

  LaneProgram lanes[kLanes];
  double l_acc[kLanes], m_acc[kLanes];
#pragma unroll
  for (int k = 0; k < kLanes; ++k) { lanes[k] = idle_lane(); l_acc[k] = 0.; m_acc[k] = 0.; }

  __syncthreads();   // capture list and barriers are ready
  auto capture_index = [&](int col) {
    int idx = -1;
    for (int i = 0; i < s_n_cap; ++i) if (s_cap_col[i] == col) idx = i;
    return idx;
  };

  // ---- simple columns: the stream --------------------------------------------------------------------------------------
  if (staged) {
    // Thread t owns position t of every row of every stage, i.e. column (head + t) % D of the tensor.
    const bool active = tid < D;
    const int my_col = static_cast<int>((head + tid) % D);
    if (active) lanes[0] = make_lane(prm.cols[my_col], inv_rows);
    const LaneProgram lp = lanes[0];
    float* g_body = GRAD ? g_chunk + head : nullptr;
    const int n_cap = s_n_cap;
    for (int it = 0; it < n_iter; ++it) {
      const int slot = it % ring;
      mg_mbar_wait(&s_bar[slot], static_cast<uint32_t>((it / ring) & 1));
      const int64_t off = static_cast<int64_t>(it) * stage_elems;
      const int elems = static_cast<int>(min(static_cast<int64_t>(stage_elems), body - off));
      const float* stage_p = s_stage + static_cast<size_t>(slot) * 2 * stage_elems;
      if (active) {
        const float* sp = stage_p + tid;
        const float* sy = sp + stage_elems;
        float* g = GRAD ? g_body + off + tid : nullptr;
        float l_part = 0.f, m_part = 0.f;   // <= 8 rows summed in fp32 in row order, then one fp64 add
        if (elems == stage_elems && lp.loss_sq && (lp.metric_sq || !lp.use_metric)) {
          // Full stage, squared error into both slots (the mcep / lf0 / bap columns): the hot loop.  32-bit shared
          // addresses, one product for loss and metric, gradient = d * (2 * w_row) (same single rounding as (2d) * w).
          uint32_t pa = mg_smem_addr(sp), ya = mg_smem_addr(sy);
          const uint32_t step = static_cast<uint32_t>(D) * 4u;
          const float w2 = __fmul_rn(2.f, lp.w_row);
#pragma unroll
          for (int u = 0; u < kStageRows; ++u) {
            float pv, yv;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(pv) : "r"(pa));
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(yv) : "r"(ya));
            const float d = __fsub_rn(pv, yv);
            l_part = __fadd_rn(l_part, __fmul_rn(d, d));
            if (GRAD) { __stcs(g, __fmul_rn(d, w2)); g += D; }
            pa += step;
            ya += step;
          }
          m_part = l_part;
        } else {   // partial last stage, or a column with an absolute-error program
          for (int i = tid; i < elems; i += D) {
            const float d = __fsub_rn(*sp, *sy);
            const float sq = __fmul_rn(d, d), ab = fabsf(d);
            l_part = __fadd_rn(l_part, lp.loss_sq ? sq : ab);
            m_part = __fadd_rn(m_part, lp.metric_sq ? sq : ab);
            if (GRAD) { __stcs(g, __fmul_rn(simple_slope(lp.loss_sq, d), lp.w_row)); g += D; }
            sp += D;
            sy += D;
          }
        }
        if (lp.use_loss) l_acc[0] += static_cast<double>(l_part);
        if (lp.use_metric) m_acc[0] += static_cast<double>(m_part);
      }
      // The last warp also copies this stage's elements of the captured columns (<= 8 columns x 8 rows) to the side
      // buffer: operands of the special phase, picked up while they are in shared memory anyway.
      if (warp == n_warps - 1) {
        for (int idx = lane; idx < n_cap * kStageRows; idx += 32) {
          const int ci = idx / kStageRows, u = idx % kStageRows;
          const int pos = static_cast<int>((s_cap_col[ci] - head + D) % D);       // position of that column in a stage row
          const int i = u * D + pos;
          if (i < elems) {
            const int row = it * kStageRows + u + ((head + pos) >= D ? 1 : 0);
            float* dst = side + (static_cast<size_t>(ci) * rows_cap + row) * 2;
            dst[0] = stage_p[i];
            dst[1] = stage_p[stage_elems + i];
          }
        }
      }
      __syncthreads();   // every thread is done with this slot: it may be refilled
      if (tid == 0 && it + ring < n_iter) issue(it + ring);
    }
    // head / tail floats around the aligned body (at most 3 + 3): one per thread, through the second lane
    const int64_t tail_begin = head + body;
    const int64_t n_edge = head + (total - tail_begin);
    if (tid < n_edge) {
      const int64_t e = tid < head ? tid : tail_begin + (tid - head);
      const int ecol = static_cast<int>(e % D);
      const mg_column col = prm.cols[ecol];
      const float pv = __ldcs(p_chunk + e), yv = __ldcs(y_chunk + e);
      const int ecap = capture_index(ecol);
      if (ecap >= 0) {
        float* dst = side + (static_cast<size_t>(ecap) * rows_cap + e / D) * 2;
        dst[0] = pv;
        dst[1] = yv;
      }
      if (column_is_simple(col)) {   // special columns are evaluated after the stream
        lanes[1] = make_lane(col, inv_rows);
        const float d = __fsub_rn(pv, yv);
        const float sq = __fmul_rn(d, d), ab = fabsf(d);
        if (lanes[1].use_loss) l_acc[1] += static_cast<double>(lanes[1].loss_sq ? sq : ab);
        if (lanes[1].use_metric) m_acc[1] += static_cast<double>(lanes[1].metric_sq ? sq : ab);
        if (GRAD) __stcs(g_chunk + e, __fmul_rn(simple_slope(lanes[1].loss_sq, d), lanes[1].w_row));
      }
    }
  } else if (n_rows > 0) {
    // Fallback (strided rows / operands that disagree on alignment): thread = column, 8 rows in flight.
    if (tid < D) {
      const mg_column col = prm.cols[tid];
      const int cap = capture_index(tid);
      float* my_side = cap >= 0 ? side + static_cast<size_t>(cap) * rows_cap * 2 : nullptr;
      if (column_is_simple(col) || cap >= 0) {
        lanes[0] = make_lane(col, inv_rows);
        const LaneProgram lp = lanes[0];
        const float* p = p_chunk + tid;
        const float* y = y_chunk + tid;
        float* g = GRAD ? g_chunk + tid : nullptr;
        for (int64_t r = 0; r < n_rows; r += kObjUnroll) {
          const int live = static_cast<int>(min(static_cast<int64_t>(kObjUnroll), n_rows - r));
          float pv[kObjUnroll], yv[kObjUnroll];
#pragma unroll
          for (int u = 0; u < kObjUnroll; ++u) {
            pv[u] = u < live ? __ldcs(p + u * p_st) : 0.f;
            yv[u] = u < live ? __ldcs(y + u * t_st) : 0.f;
          }
          float l_part = 0.f, m_part = 0.f;
#pragma unroll
          for (int u = 0; u < kObjUnroll; ++u) {
            if (cap >= 0 && u < live) { my_side[2 * (r + u)] = pv[u]; my_side[2 * (r + u) + 1] = yv[u]; }
            const float d = __fsub_rn(pv[u], yv[u]);
            const float sq = __fmul_rn(d, d), ab = fabsf(d);
            l_part = __fadd_rn(l_part, lp.loss_sq ? sq : ab);
            m_part = __fadd_rn(m_part, lp.metric_sq ? sq : ab);
            if (GRAD && u < live) __stcs(g + u * g_st, __fmul_rn(simple_slope(lp.loss_sq, d), lp.w_row));
          }
          if (lp.use_loss) l_acc[0] += static_cast<double>(l_part);
          if (lp.use_metric) m_acc[0] += static_cast<double>(m_part);
          p += kObjUnroll * p_st;
          y += kObjUnroll * t_st;
          if (GRAD) g += kObjUnroll * g_st;
        }
      }
    }
  }

  // ---- padding rows of this chunk: the gradient is defined (zero) over the whole (B, T, D) tensor -------------------
  if (GRAD) {
    const int64_t z0 = max(r0, n_b), z1 = r1;
    if (z1 > z0) {
      float* gz = prm.grad + b * prm.g_sb;
      if (g_st == D) {
        float* z = gz + z0 * g_st;
        const int64_t n = (z1 - z0) * D;
        const int64_t zh = min(n, static_cast<int64_t>(((16 - (reinterpret_cast<uintptr_t>(z) & 15)) & 15) >> 2));
        for (int64_t i = tid; i < zh; i += blockDim.x) z[i] = 0.f;
        const int64_t nz = (n - zh) >> 2;
        float4* z4 = reinterpret_cast<float4*>(z + zh);
        for (int64_t i = tid; i < nz; i += blockDim.x) __stcs(z4 + i, make_float4(0.f, 0.f, 0.f, 0.f));
        for (int64_t i = zh + (nz << 2) + tid; i < n; i += blockDim.x) z[i] = 0.f;
      } else {
        for (int c = tid; c < D; c += blockDim.x)
          for (int64_t r = z0; r < z1; ++r) gz[r * g_st + c] = 0.f;
      }
    }
  }
  __syncthreads();   // stream done: side buffer complete, stage ring free, zero gradients of special columns ordered first

  if (n_rows > 0) {
    // ---- per-slot CTA sums of the streamed columns: every thread parks its (value, slot) pairs in shared memory (the
    // stage ring is free by now), then warp s sums slot s over the entries in index order with a fixed shuffle tree --------
    {
      const int n_entries = kLanes * blockDim.x;
      double* s_lv = reinterpret_cast<double*>(smem_raw);          // [n_entries] loss partials
      double* s_mv = s_lv + n_entries;                             // [n_entries] metric partials
      signed char* s_ls = reinterpret_cast<signed char*>(s_mv + n_entries);   // [n_entries] loss slot or -1
      signed char* s_ms = s_ls + n_entries;                        // [n_entries] metric slot or -1
#pragma unroll
      for (int k = 0; k < kLanes; ++k) {
        const int e = k * blockDim.x + tid;
        s_lv[e] = l_acc[k];
        s_mv[e] = m_acc[k];
        s_ls[e] = lanes[k].use_loss ? static_cast<signed char>(lanes[k].loss_slot) : static_cast<signed char>(-1);
        s_ms[e] = lanes[k].use_metric ? static_cast<signed char>(lanes[k].metric_slot) : static_cast<signed char>(-1);
      }
      __syncthreads();
      for (int slot = warp; slot < prm.n_slots; slot += n_warps) {
        double v = 0.;
        for (int e = lane; e < n_entries; e += 32) {
          if (s_ls[e] == slot) v += s_lv[e];
          if (s_ms[e] == slot) v += s_mv[e];
        }
        v = mg_warp_sum(v);
        if (lane == 0) { s_slot_sum[slot] = v; s_slot_cnt[slot] = 0.; }
      }
    }

    // ---- special columns (BCE, exp, equality, per-frame root, voiced weighting; 3 of 187 here): ~100 instructions per
    // element, so the WHOLE CTA shares them, one row per thread, operands from the side buffer filled by the stream ------
    const int n_sp = s_n_sp;
    {
      double l[kMaxSpecial], m[kMaxSpecial], n[kMaxSpecial];
#pragma unroll
      for (int q = 0; q < kMaxSpecial; ++q) {
        l[q] = m[q] = n[q] = 0.;
        if (q < n_sp) {   // CTA-uniform
          const int k = s_sp_col[q];
          const mg_column sc = prm.cols[k];
          const float w_row = static_cast<float>(static_cast<double>(sc.loss_weight) * inv_rows);
          const bool has_mask = sc.mask_col != MG_COL_NONE;
          const float* own = side + static_cast<size_t>(capture_index(k)) * rows_cap * 2;
          const float* msk = has_mask ? side + static_cast<size_t>(capture_index(sc.mask_col)) * rows_cap * 2 : nullptr;
          for (int64_t r = tid; r < n_rows; r += blockDim.x) {
            const float pv = own[2 * r], yv = own[2 * r + 1];
            const float mask_v = has_mask ? msk[2 * r] : 1.f;
            mg_column one = sc;
            one.width = 1;
            if (sc.metric_kind == MG_RED_ROOT_SQDIFF && sc.width > 1) {
              // feature-axis sum of the group (metrics.py:661) from the captured neighbours, then the root (:662)
              const float d0 = __fsub_rn(yv, pv);
              float acc = __fmul_rn(d0, d0);
              for (int j = 1; j < sc.width; ++j) {
                const float* oth = side + static_cast<size_t>(capture_index(k + j)) * rows_cap * 2;
                const float dj = __fsub_rn(oth[2 * r + 1], oth[2 * r]);
                acc = __fadd_rn(acc, __fmul_rn(dj, dj));
              }
              float root = sqrtf(acc);
              if (has_mask) {
                const float voiced = mask_v > 0.5f ? 1.f : 0.f;
                root = __fmul_rn(root, voiced);
                n[q] += static_cast<double>(voiced);
              }
              m[q] += static_cast<double>(root);
              one.metric_kind = MG_COL_NONE;   // the loss part of the column (if any) still goes through general_one
            }
            general_one<GRAD>(one, pv, yv, mask_v, has_mask, nullptr, nullptr, GRAD ? g_chunk + r * g_st + k : nullptr, w_row,
                              l[q], m[q], n[q]);
          }
          l[q] = mg_warp_sum(l[q]);
          m[q] = mg_warp_sum(m[q]);
          n[q] = mg_warp_sum(n[q]);
          if (lane == 0) { s_red[(3 * q) * 32 + warp] = l[q]; s_red[(3 * q + 1) * 32 + warp] = m[q]; s_red[(3 * q + 2) * 32 + warp] = n[q]; }
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      for (int q = 0; q < n_sp; ++q) {
        double ls = 0., ms = 0., ns = 0.;
        for (int i = 0; i < n_warps; ++i) { ls += s_red[(3 * q) * 32 + i]; ms += s_red[(3 * q + 1) * 32 + i]; ns += s_red[(3 * q + 2) * 32 + i]; }
        const mg_column sc = prm.cols[s_sp_col[q]];
        if (sc.loss_kind != MG_COL_NONE) s_slot_sum[sc.loss_slot] += ls;
        if (sc.metric_kind != MG_COL_NONE) { s_slot_sum[sc.metric_slot] += ms; s_slot_cnt[sc.metric_slot] += ns; }
      }
    }
    __syncthreads();

    // Special columns that did not fit the side buffer (more than kMaxSpecial / kMaxCapture; not the acoustic layouts):
    // one at a time from global memory (L2-hot).
    if (s_n_special > n_sp) {
      for (int w = 0; w < (D + 31) / 32; ++w) {
        unsigned todo = s_special[w];
        while (todo) {   // CTA-uniform loop
          const int k = w * 32 + __ffs(todo) - 1;
          todo &= todo - 1;
          bool served = false;
          for (int q = 0; q < n_sp; ++q) served = served || s_sp_col[q] == k;
          if (served) continue;
          const mg_column sc = prm.cols[k];
          const float w_row = static_cast<float>(static_cast<double>(sc.loss_weight) * inv_rows);
          const bool has_mask = sc.mask_col != MG_COL_NONE;
          double l = 0., m = 0., n = 0.;
          for (int64_t r = tid; r < n_rows; r += blockDim.x) {
            const float* p = p_chunk + r * p_st + k;
            const float* y = y_chunk + r * t_st + k;
            const float mask_v = has_mask ? __ldg(p_chunk + r * p_st + sc.mask_col) : 1.f;
            general_one<GRAD>(sc, __ldg(p), __ldg(y), mask_v, has_mask, p, y, GRAD ? g_chunk + r * g_st + k : nullptr, w_row,
                              l, m, n);
          }
          l = mg_warp_sum(l);
          m = mg_warp_sum(m);
          n = mg_warp_sum(n);
          if (lane == 0) { s_red[warp] = l; s_red[32 + warp] = m; s_red[64 + warp] = n; }
          __syncthreads();
          if (tid == 0) {
            double ls = 0., ms = 0., ns = 0.;
            for (int i = 0; i < n_warps; ++i) { ls += s_red[i]; ms += s_red[32 + i]; ns += s_red[64 + i]; }
            if (sc.loss_kind != MG_COL_NONE) s_slot_sum[sc.loss_slot] += ls;
            if (sc.metric_kind != MG_COL_NONE) { s_slot_sum[sc.metric_slot] += ms; s_slot_cnt[sc.metric_slot] += ns; }
          }
          __syncthreads();
        }
      }
    }
    if (tid < prm.n_slots)
      prm.ws.partials[(static_cast<int64_t>(tid) * prm.B + b) * kMgMaxChunks + chunk] = make_double2(s_slot_sum[tid], s_slot_cnt[tid]);
  }

  mg_finish(prm.slots, prm.n_slots, prm.seq_len, prm.B, T, prm.ws, b, gridDim.x, s_red, &s_is_last);
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

  {   // contiguous, aligned, D <= 256: the persistent row-stream form (mg_objective_stream.cu)
    MgObjectiveArgs args;
    args.pred = pred; args.target = target; args.grad = grad; args.grad_scale_dev = grad_scale_dev;
    args.cols = cols; args.slots = slots; args.seq_len = seq_len;
    args.p_sb = p_sb; args.p_st = p_st; args.t_sb = t_sb; args.t_st = t_st; args.g_sb = g_sb; args.g_st = g_st; args.T = T;
    args.D = D; args.B = B; args.n_slots = n_slots; args.workspace = workspace; args.workspace_bytes = workspace_bytes;
    for (int i = 0; i < n_slots; ++i) {
      MG_REQUIRE(slots[i].result != nullptr && mg_aligned(slots[i].result, 16), "mg_masked_objective_f32: slot %d needs a 16-byte aligned result record", i);
      MG_REQUIRE(slots[i].D >= 1, "mg_masked_objective_f32: slot %d has D=%d", i, slots[i].D);
    }
    const int status = mg_objective_stream_launch(args, stream);
    if (status != MG_STREAM_NOT_APPLICABLE) return status;
  }

  ObjectiveParams prm;
  memset(&prm, 0, sizeof(prm));
  // Rows per CTA: ~96 KB per operand, but at least 4 CTAs per SM in the grid and at most kMgMaxChunks chunks.
  int64_t rows = kObjTargetElems / D;
  { const char* e = getenv("MG_OBJ_ROWS"); if (e && atoi(e) > 0) rows = atoi(e); }
  if (rows < kObjUnroll) rows = kObjUnroll;
  if (rows > T) rows = T;
  const int64_t sms = mg_cached_sm_count();
  while (rows > 2 * kObjUnroll && static_cast<int64_t>(B) * ((T + rows - 1) / rows) < 4 * sms) rows = (rows + 1) / 2;
  const int64_t min_rows = (T + kMgMaxChunks - 1) / kMgMaxChunks;
  if (rows < min_rows) rows = min_rows;
  if (rows > 512 && min_rows <= 512) rows = 512;                // bounds the side buffer (8 columns x rows x 8 bytes)
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
  prm.ws = mg_carve_workspace(workspace, n_slots, B);
  prm.p_sb = p_sb; prm.p_st = p_st; prm.t_sb = t_sb; prm.t_st = t_st; prm.g_sb = g_sb; prm.g_st = g_st; prm.T = T;
  prm.D = D; prm.B = B; prm.n_slots = n_slots; prm.rows_per_cta = static_cast<int>(rows);

  int threads = ((D + 31) / 32) * 32;
  if (threads < 64) threads = 64;
  // Dynamic shared memory: ring of (pred, target) stages of 8 rows + the side buffer.  Measured on B200 (config 2,
  // D = 187): 74 KB (5 stages, 3 CTAs / SM) 0.179 ms, 54 KB 0.163, 44 KB 0.163, 34 KB (2 stages, 5 CTAs / SM) 0.160 ms --
  // the per-CTA serial phases (special columns, slot sums, ticket) want more co-resident CTAs, not a deeper ring.
  const size_t side_bytes = static_cast<size_t>(kMaxCapture) * rows * 2 * sizeof(float);
  const size_t stage_pair_bytes = static_cast<size_t>(2) * kStageRows * D * sizeof(float);
  // Default: a 2-stage ring + the side buffer (measured on B200 at D = 187: rows per CTA 131 / 160 / 200 / 262 / 400 ->
  // 0.153 / 0.152 / 0.148 / 0.154 / 0.168 ms with 2 stages; a third stage at the same rows is slower).  MG_OBJ_ROWS overrides.
  static int budget_kb = -1;
  if (budget_kb < 0) { const char* e = getenv("MG_OBJ_SMEM_KB"); budget_kb = e ? atoi(e) : 0; }
  const size_t budget = budget_kb > 0 ? static_cast<size_t>(budget_kb) * 1024 : 2 * stage_pair_bytes + side_bytes + 64;
  int ring = side_bytes < budget ? static_cast<int>((budget - side_bytes) / stage_pair_bytes) : 0;
  if (ring > kMaxStages) ring = kMaxStages;
  static int force_fallback = -1;
  if (force_fallback < 0) force_fallback = getenv("MG_OBJECTIVE_NO_STAGING") != nullptr;
  if (ring < 2) ring = (2 * stage_pair_bytes + side_bytes <= 200 * 1024) ? 2 : 0;
  if (force_fallback) ring = 0;
  prm.n_stages = ring;
  size_t ring_bytes = static_cast<size_t>(ring) * stage_pair_bytes;
  const size_t reduce_bytes = static_cast<size_t>(kLanes) * threads * (2 * sizeof(double) + 2);
  if (ring_bytes < reduce_bytes) ring_bytes = reduce_bytes;     // the slot-sum scratch reuses the ring
  ring_bytes = (ring_bytes + 127) / 128 * 128;
  prm.side_offset = static_cast<int>(ring_bytes);
  const size_t smem = ring_bytes + side_bytes;
  MG_REQUIRE(smem <= 200 * 1024, "mg_masked_objective_f32: D=%d needs %zu bytes of shared memory", D, smem);
  dim3 grid(static_cast<unsigned>(n_chunks), static_cast<unsigned>(B));
  static bool attr_done[64] = {};   // per device
  int device = 0;
  MG_CUDA_OK(cudaGetDevice(&device));
  if (smem > 32 * 1024 && !attr_done[device & 63]) {   // static + dynamic shared memory above 48 KB needs the opt-in (the kernel has ~5 KB static)
    attr_done[device & 63] = true;
    MG_CUDA_OK(cudaFuncSetAttribute(masked_objective_kernel<true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MG_CUDA_OK(cudaFuncSetAttribute(masked_objective_kernel<false, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MG_CUDA_OK(cudaFuncSetAttribute(masked_objective_kernel<true, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MG_CUDA_OK(cudaFuncSetAttribute(masked_objective_kernel<false, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
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
