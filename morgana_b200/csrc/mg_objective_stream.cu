// K4b, persistent row-stream form -- the whole-row masked objective (models/RNN_SPSS.py:120-139: 3 x losses.mse + losses.bce
// with the gradient, and the four streaming metrics; morgana/losses.py:29-56, morgana/metrics.py:383-394, 597-694) for
// contiguous (B, T, D <= 224) float32 tensors whose bases are 16-byte aligned.  Same results contract as mg_objective.cu.
//
// Why a second form.  The chunk-per-CTA kernel spends ~20 us of a 150 us launch in per-CTA work that does not scale with
// the bytes (column-program set-up, barrier init, special-column phase, slot sums, two ticket levels per 131-row chunk) and
// keeps only one 12 KB stage per CTA in flight while it computes (profiles/r1b_summary.md, r1d_summary.md).  Here:
//
//   * The (B*T, D) row space is cut into STAGES of 8 rows (8*D floats = 32*D bytes: 16-byte aligned for every D, and a
//     whole number of rows, so thread t always sees column t).  One persistent CTA per half SM owns a contiguous range of
//     stages chosen on the device so that every CTA carries the same cost (6 per valid row -- two reads, one gradient write and
//     the arithmetic --, 1 per padding row -- the zero gradient; 3 : 1, the ratio of the bytes, left the padding-heavy ranges
//     early: measured).
//   * Warp roles.  A producer lane keeps a ring of stages in flight with cp.async.bulk global->shared (TMA engine, SASS
//     UBLKCP) on "full" mbarriers and refills a slot as soon as the consumer warps have arrived on its "empty" mbarrier -- no
//     CTA-wide barrier inside the stream, the bytes in flight live in shared memory (ring x 12 KB per CTA).  Padding-only
//     stages cost it one bulk store from a zero tile and are never loaded.  ceil(D/32) consumer warps own one column per
//     thread: 16 shared-memory loads in flight, squared error into an fp32 partial per stage (row order), one fp64 add per
//     stage, gradient straight from registers.  Runs of fully valid stages of one utterance are a tight loop without any
//     per-stage classification.
//   * "Special" columns (BCE, exp, equality, per-frame root, voiced weighting: 3 of 187, ~100 instructions per element) would
//     make the warp that owns them a straggler.  Instead consumer warp w adopts special column w: as each stage passes, eight
//     of its lanes park that column's operands (and the mask / group columns they depend on, same row of the stage) in
//     registers; after four stages the warp holds 32 rows, lane = row, and evaluates them on one code path.  (A first version
//     with a dedicated warp evaluating 8 rows x 3 columns per stage ran at 2.2 us per stage -- a dependent chain of ~500
//     instructions in one warp -- and was slower than the kernel it replaces; profiles/r2b_summary.md.)
//   * Head start.  CTA c first takes the fixed stages [4c, 4c + 4): the producer lane issues their loads as soon as the barriers
//     exist, before anything is known about the batch, so the DRAM latency of the first bytes overlaps the prologue (column
//     programs, lengths, cost prefix, range search: ~5 us of every launch before this).  The cost-balanced partition then covers
//     the stages behind those head starts; it is computed by the producer warp alone while the consumers already stream.
//     (Handing the last sixth of the work out dynamically -- guided self-scheduling over an atomic counter, a record per job so
//     the sums stay deterministic -- was built and measured: every job switch costs each warp ~600 instructions (flush of the
//     special-column batch, shuffle folds, range search) = ~2 us, more than the ~8 us of imbalance it removes; dropped.)
//   * Accumulators are carried across stages AND utterances in registers: per thread (sum, sum / n_b) of its loss slot and the
//     sum of its metric slot, folded into the weighted form only when the utterance changes.  At the end the consumer warps
//     fold their lanes with a fixed shuffle tree, warp 0 adds the warps in order and writes the CTA's record [slot][cta].
//     One ticket; the last CTA adds the records in CTA order.
//
// Determinism: the stage -> CTA assignment is a pure function of (B, T, seq_len, grid), every sum has a fixed order, no
// floating-point atomics.
#include <stdlib.h>
#include <string.h>

#include "mg_objective_common.cuh"

namespace {

using namespace mgobj;

constexpr int kRows = 8;            // rows per stage
constexpr int kMaxRing = 16;
constexpr int kHead = 4;            // stages of a CTA's head start (<= ring)
constexpr int kMaxB = 1024;         // utterance lengths and the cost prefix live in shared memory
constexpr int kSpPerWarp = 1;       // special columns a consumer warp evaluates in batches of 32 rows
constexpr int kMaxSp = 16;          // batched special columns per CTA (any further one is evaluated by its own thread)
constexpr int kMaxWarps = 8;        // 7 consumer warps (D <= 224) + producer: 256 threads, 128 registers at two CTAs per SM
constexpr int kMaxCtas = 1024;      // bounds the per-CTA records in the workspace
constexpr int kFinishLoads = 10;    // records a lane of the last CTA keeps in flight: one round for 2 x 148 CTAs

struct StreamParams {
  MgFinishSlot slots[MG_MAX_TERMS];
  const float* pred;
  const float* target;
  float* grad;
  const float* grad_scale_dev;
  const mg_column* cols;
  const int64_t* seq_len;
  unsigned int* ticket;
  double2* records;      // [n_slots][grid][2]: (sum, sum / n_b), (weighted count, -)
  int64_t T;
  int D, B, n_slots, ring, n_consumer_warps;
  int cost_valid, cost_pad;   // relative cost of a valid / padding row in the stage -> CTA partition
  float load_first_frac;   // tuning: share of the bulk loads that carry the L2 evict-first hint
  int debug;   // tuning experiments only (MG_OBJ_DEBUG): 1 consumers skip the arithmetic, 2 no special columns, 4 no zero stores,
               // 8 per-CTA %globaltimer stamps -> `stamps`, 16 flip the L2 hint of the bulk loads, 32 plain gradient stores
  unsigned long long* stamps;
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mg_smem_addr(bar)) : "memory");
}

__device__ __forceinline__ void consumer_barrier(int n_threads) {   // named barrier 1: the consumer warps only
  asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory");
}

// Walks a range of stages; every role keeps its own copy and sees the same sequence.
struct StageCursor {
  int64_t stage, stage_end;   // current stage and end of the range
  int64_t total_rows;
  int64_t T;
  int b;                      // utterance of the stage's first row
  int64_t t0;                 // that row's index inside the utterance
  __device__ __forceinline__ void init(int64_t s_lo, int64_t s_hi, int64_t T_, int64_t total_rows_) {
    stage = s_lo; stage_end = s_hi; T = T_; total_rows = total_rows_;
    const unsigned r = static_cast<unsigned>(s_lo * kRows);      // B * T < 2^31 (host check): one 32-bit division
    b = static_cast<int>(r / static_cast<unsigned>(T_));
    t0 = static_cast<int64_t>(r) - static_cast<int64_t>(b) * T_;
  }
  __device__ __forceinline__ bool done() const { return stage >= stage_end; }
  __device__ __forceinline__ void next() {
    ++stage;
    t0 += kRows;
    while (t0 >= T) { t0 -= T; ++b; }
  }
  __device__ __forceinline__ void advance(int64_t n) {
    stage += n;
    t0 += n * kRows;
    while (t0 >= T) { t0 -= T; ++b; }
  }
  // consecutive FULL stages (8 valid rows of utterance b) from here, inside the range
  __device__ __forceinline__ int64_t full_run(const int* s_nb) const {
    const int64_t n_b = s_nb[b];
    if (t0 + kRows > n_b) return 0;
    const int64_t run = (n_b - t0) / kRows, cap = stage_end - stage;
    return run < cap ? run : cap;
  }
  // consecutive PAD stages (8 padding rows of utterance b) from here, inside the range
  __device__ __forceinline__ int64_t pad_run(const int* s_nb) const {
    if (t0 < s_nb[b] || t0 + kRows > T) return 0;
    const int64_t run = (T - t0) / kRows, cap = stage_end - stage;
    return run < cap ? run : cap;
  }
  // rows of this stage that exist (8 except for the last stage of the tensor)
  __device__ __forceinline__ int rows() const {
    const int64_t left = total_rows - stage * kRows;
    return left < kRows ? static_cast<int>(left) : kRows;
  }
};

enum StageKind { STAGE_FULL = 0, STAGE_MIXED = 1, STAGE_PAD = 2, STAGE_TAIL = 3 };

// FULL: 8 valid rows of one utterance.  PAD: 8 padding rows (no load outside the head start; zero gradient).  TAIL: the
// tensor's last, short stage (its byte count need not be a multiple of 16: read from global memory).  MIXED: anything else
// that has 8 rows.
__device__ __forceinline__ int classify(const StageCursor& c, const int* s_nb, int B) {
  if (c.rows() < kRows) return STAGE_TAIL;
  const int64_t n_b = s_nb[c.b];
  if (c.t0 + kRows <= c.T) {
    if (c.t0 + kRows <= n_b) return STAGE_FULL;
    if (c.t0 >= n_b) return STAGE_PAD;
    return STAGE_MIXED;
  }
  // the stage runs into the next utterance(s): padding only if every row is padding
  int b = c.b;
  int64_t t = c.t0;
#pragma unroll 1
  for (int u = 0; u < kRows; ++u) {
    if (t < s_nb[b]) return STAGE_MIXED;
    if (++t >= c.T) { t = 0; ++b; }
  }
  return STAGE_PAD;
}

// The cost-balanced partition covers the rows behind the head starts, [first_row, B * T).  Inside an utterance the rows left
// of it count in row order: valid rows (cost_valid each), then padding rows (cost_pad each).
struct CostModel {
  int64_t T, first_row;
  unsigned cost_valid, cost_pad;
  __host__ __device__ __forceinline__ unsigned cut(int b) const {       // rows of utterance b that belong to the head starts
    const int64_t c = first_row - static_cast<int64_t>(b) * T;
    return static_cast<unsigned>(c < 0 ? 0 : (c > T ? T : c));
  }
  __host__ __device__ __forceinline__ unsigned cost(int b, unsigned n) const {
    const unsigned c = cut(b), t = static_cast<unsigned>(T);
    const unsigned valid = n > c ? n - c : 0u, from = n > c ? n : c;
    return cost_valid * valid + cost_pad * (t - from);
  }
};

// Cost position -> stage, rounded up to a stage.  Monotone in x, and the same function at both ends of every CTA's range, so the
// ranges tile the stages exactly.
__host__ __device__ __forceinline__ int64_t cost_to_stage(unsigned x, const CostModel& cm, const int* s_pref, const int* s_nb, int B,
                                                 int64_t first_stage, int64_t n_stages) {
  if (x == 0) return first_stage;
  if (x >= static_cast<unsigned>(s_pref[B])) return n_stages;
  int lo = 0, hi = B;                      // last b with pref[b] <= x
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (static_cast<unsigned>(s_pref[mid]) <= x) lo = mid; else hi = mid; }
  const unsigned y = x - static_cast<unsigned>(s_pref[lo]), n = static_cast<unsigned>(s_nb[lo]), c = cm.cut(lo);
  const unsigned valid = n > c ? n - c : 0u, from = n > c ? n : c;
  unsigned r_in = y < cm.cost_valid * valid ? c + y / cm.cost_valid
                                            : (cm.cost_pad > 0 ? from + (y - cm.cost_valid * valid) / cm.cost_pad : static_cast<unsigned>(cm.T));
  if (r_in > static_cast<unsigned>(cm.T)) r_in = static_cast<unsigned>(cm.T);
  int64_t stage = (static_cast<int64_t>(lo) * cm.T + r_in + kRows - 1) / kRows;
  if (stage < first_stage) stage = first_stage;
  return stage > n_stages ? n_stages : stage;
}

// Everything outside the stage loop runs as cold code at ~25 cycles per instruction (instruction-cache misses served by an L2 that
// is saturated with the data stream: profiles/r2e_summary.md), and lambdas inlined at several call sites made the kernel 11 k
// instructions long.  The rare routines are therefore real functions with ONE copy each.
struct Contribution { double l, m, n; };   // loss value, metric value, metric weight of one element

__device__ __noinline__ double div_f64(double a, double b) { return a / b; }

// One element of a special column from values held in registers.  `root_acc`: squared error summed over the column's
// ROOT_SQDIFF group (only read when the column leads a group of more than one column).
template <bool GRAD>
__device__ __noinline__ Contribution special_value(mg_column sc, float pv, float yv, float mask_v, float root_acc, float* g, float w_row) {
  const bool has_mask = sc.mask_col != MG_COL_NONE;
  Contribution c;
  c.l = c.m = c.n = 0.;
  if (sc.metric_kind == MG_RED_ROOT_SQDIFF && sc.width > 1) {
    float root = sqrtf(root_acc);                     // feature-axis sum of the group (metrics.py:661), then the root (:662)
    if (has_mask) {
      const float voiced = mask_v > 0.5f ? 1.f : 0.f;
      root = __fmul_rn(root, voiced);
      c.n += static_cast<double>(voiced);
    }
    c.m += static_cast<double>(root);
    sc.metric_kind = MG_COL_NONE;   // the loss part of the column (if any) still goes through general_one
  }
  sc.width = 1;
  general_one<GRAD>(sc, pv, yv, mask_v, has_mask, nullptr, nullptr, g, w_row, c.l, c.m, c.n);
  return c;
}

// The lanes of a warp that feed slot s are summed with a fixed shuffle tree (slots in lane order of their first column) and
// added to the warp's row `mine` [slot][3]: `a` to [s][0], `b2` to [s][second].
__device__ __noinline__ void fold_lanes(double* mine, int slot_id, double a, double b2, int second, bool with_second) {
  const int lane = threadIdx.x & 31;
  unsigned todo = __ballot_sync(MG_FULL_MASK, slot_id >= 0);
  while (todo) {
    const int leader = __ffs(todo) - 1;
    const int s = __shfl_sync(MG_FULL_MASK, slot_id, leader);
    const bool in = slot_id == s;
    todo &= ~__ballot_sync(MG_FULL_MASK, in);
    const double va = mg_warp_sum(in ? a : 0.);
    const double vb = with_second ? mg_warp_sum(in ? b2 : 0.) : 0.;
    if (lane == 0) { mine[s * 3] += va; mine[s * 3 + second] += vb; }
  }
}

__device__ __forceinline__ float root_group_acc(const mg_column& sc, int k, const float* row_p, const float* row_y) {
  const float d0 = __fsub_rn(row_y[k], row_p[k]);
  float acc = __fmul_rn(d0, d0);
#pragma unroll 1
  for (int j = 1; j < sc.width; ++j) {
    const float dj = __fsub_rn(row_y[k + j], row_p[k + j]);
    acc = __fadd_rn(acc, __fmul_rn(dj, dj));
  }
  return acc;
}

template <bool GRAD>
__global__ void __launch_bounds__(kMaxWarps * 32, 2)
objective_stream_kernel(const __grid_constant__ StreamParams prm) {
  extern __shared__ __align__(128) unsigned char smem_raw[];   // [ring x (pred stage | target stage)] [zero tile]
  __shared__ __align__(8) uint64_t s_full[kMaxRing], s_empty[kMaxRing], s_range_bar;
  __shared__ mg_column s_cols[256];
  __shared__ int s_nb[kMaxB];
  __shared__ int s_pref[kMaxB + 1];          // exclusive prefix of the per-utterance cost
  __shared__ int s_has_empty;
  __shared__ int64_t s_range[2];
  __shared__ long long s_valid_total;
  __shared__ double s_flush[kMaxWarps - 1][MG_MAX_TERMS][3];   // per-warp sums
  __shared__ double s_slot[MG_MAX_TERMS][3];
  __shared__ bool s_is_last;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int D = prm.D, B = prm.B, ring = prm.ring, n_cw = prm.n_consumer_warps;
  const int producer_warp = n_cw;
  const int64_t T = prm.T;
  const int64_t total_rows = static_cast<int64_t>(B) * T;
  const int64_t n_stages = (total_rows + kRows - 1) / kRows;
  const int stage_elems = kRows * D;
  const uint32_t stage_bytes = static_cast<uint32_t>(stage_elems) * 4u;
  float* s_ring = reinterpret_cast<float*>(smem_raw);
  float* s_zero = s_ring + static_cast<size_t>(ring) * 2 * stage_elems;
  // head start of this CTA: stages [head_lo, head_hi); the partition covers [first_stage, n_stages)
  const int n_head = ring < kHead ? ring : kHead;
  const int64_t first_stage = min(static_cast<int64_t>(gridDim.x) * n_head, n_stages);
  const int64_t head_lo = min(static_cast<int64_t>(blockIdx.x) * n_head, n_stages), head_hi = min(head_lo + n_head, n_stages);

  const bool stamp = (prm.debug & 8) != 0;
  if (stamp && tid == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    prm.stamps[blockIdx.x * 16 + 0] = global_ns();
    prm.stamps[blockIdx.x * 16 + 14] = smid;
  }
  // producer state (lane 0 of the producer warp)
  int p_slot = 0;
  uint32_t p_phase = 1;       // parity to wait for on the slot's "empty" barrier; the first pass over the ring does not wait
  bool p_first_pass = true;
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, %1;" : "=l"(policy) : "f"(prm.load_first_frac));
  // L2 policies, measured at config 2 (forward + gradient): loads normal + gradient stores evict-first 0.118 ms, both evict-first
  // 0.121, loads evict-first + plain stores 0.129, both plain 0.1185.  Evict-first loads keep the kernel's own code and tables in
  // L2 from launch to launch (prologue 3.1 -> 1.9 us, last CTA 4.3 -> 1.9 us) but cost the stream 4 % as soon as there are stores
  // to write back, so only the forward-only form uses them (0.0767 -> 0.0746 ms); the small re-used data asks for evict-last.
  auto load_stage = [&](int64_t stage) {
    if (!p_first_pass) mg_mbar_wait(&s_empty[p_slot], p_phase);
    const int64_t off = stage * stage_elems;
    float* dst = s_ring + static_cast<size_t>(p_slot) * 2 * stage_elems;
    mg_mbar_expect_tx(&s_full[p_slot], 2 * stage_bytes);
    if (GRAD ? !(prm.debug & 16) : (prm.debug & 16) != 0) {
      mg_bulk_load(dst, prm.pred + off, stage_bytes, &s_full[p_slot]);
      mg_bulk_load(dst + stage_elems, prm.target + off, stage_bytes, &s_full[p_slot]);
    } else {
      mg_bulk_load_hint(dst, prm.pred + off, stage_bytes, &s_full[p_slot], policy);
      mg_bulk_load_hint(dst + stage_elems, prm.target + off, stage_bytes, &s_full[p_slot], policy);
    }
    if (++p_slot == ring) { p_slot = 0; p_phase ^= 1u; p_first_pass = false; }
  };
  // ---- prologue: the producer lane goes straight to the barriers and the head start's loads; the other warps fetch the column
  // programs and the utterance lengths meanwhile ---------------------------------------------------------------------------------
  // Programmatic dependent launch: the barriers and the zero tile are set up while the kernel before this one drains; nothing
  // above the wait touches global memory.  All CTAs of this grid are resident from the start, so the next kernel in the stream
  // may launch at once: its CTAs take the SMs of the CTAs that finish early and wait there for this grid to complete.
  if (tid == producer_warp * 32) {
    for (int i = 0; i < ring; ++i) { mg_mbar_init(&s_full[i], 1); mg_mbar_init(&s_empty[i], static_cast<uint32_t>(n_cw)); }
    mg_mbar_init(&s_range_bar, 1);
    mg_mbar_fence_init();
  } else if (GRAD && warp != producer_warp) {
    const int t2 = warp < producer_warp ? tid : tid - 32, nt = static_cast<int>(blockDim.x) - 32;
#pragma unroll 1
    for (int i = t2; i < stage_elems; i += nt) s_zero[i] = 0.f;
    mg_fence_proxy_async_smem();   // the zero tile is read by the bulk-copy engine
  }
  mg_pdl_wait();
  mg_pdl_launch_dependents();
  double scale = 1.;
  if (GRAD && prm.grad_scale_dev != nullptr) scale = static_cast<double>(__ldg(prm.grad_scale_dev));   // in flight under the loads below
  if (warp == producer_warp) {
    if (lane == 0) {
      for (int64_t st = head_lo; st < head_hi; ++st)
        if ((st + 1) * kRows <= total_rows) load_stage(st);      // every full-size stage, whatever it holds (nothing is known yet)
      if (stamp) prm.stamps[blockIdx.x * 16 + 12] = global_ns();
    }
  } else {
    const int t2 = warp < producer_warp ? tid : tid - 32, nt = static_cast<int>(blockDim.x) - 32;
    const uint64_t keep = mg_policy_evict_last();
#pragma unroll 1
    for (int b = t2; b < B; b += nt) {
      int64_t n = T;
      if (prm.seq_len != nullptr) { n = mg_ld_keep_s64(prm.seq_len + b, keep); n = n < 0 ? 0 : (n > T ? T : n); }   // as mg_valid_frames
      s_nb[b] = static_cast<int>(n);
    }
    static_assert(sizeof(mg_column) == 12, "mg_column is read as three 32-bit words");
#pragma unroll 1
    for (int k = t2; k < 3 * D; k += nt) reinterpret_cast<uint32_t*>(s_cols)[k] = mg_ld_keep_u32(reinterpret_cast<const uint32_t*>(prm.cols) + k, keep);
  }
  __syncthreads();
  if (stamp && tid == 0) prm.stamps[blockIdx.x * 16 + 8] = global_ns();
  const double scale_over_b = scale / static_cast<double>(B);

  if (warp == producer_warp) {
    // ---- producer warp: cost prefix and this CTA's range (the consumers are already on the head start), then the loads ---------
    CostModel cm;
    cm.T = T; cm.first_row = first_stage * kRows;
    cm.cost_valid = static_cast<unsigned>(prm.cost_valid); cm.cost_pad = GRAD ? static_cast<unsigned>(prm.cost_pad) : 0u;
    {
      unsigned carry = 0;
      int empty = 0;
      long long valid = 0;
#pragma unroll 1
      for (int base = 0; base < B; base += 32) {
        const int b = base + lane;
        const unsigned n = b < B ? static_cast<unsigned>(s_nb[b]) : 0u;
        const unsigned cost = b < B ? cm.cost(b, n) : 0u;
        unsigned incl = cost;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned v = __shfl_up_sync(MG_FULL_MASK, incl, o);
          if (lane >= o) incl += v;
        }
        if (b < B) s_pref[b] = static_cast<int>(carry + incl - cost);
        carry += __shfl_sync(MG_FULL_MASK, incl, 31);
        empty |= (b < B && n == 0) ? 1 : 0;
        valid += n;
      }
      empty = __any_sync(MG_FULL_MASK, empty);
      valid = mg_warp_sum(valid);
      if (lane == 0) { s_pref[B] = static_cast<int>(carry); s_has_empty = empty; s_valid_total = valid; }
      __syncwarp();
      if (stamp && lane == 0) prm.stamps[blockIdx.x * 16 + 13] = global_ns();
    }
    if (lane < 2) {
      // one end of the range per lane
      const unsigned total_cost = static_cast<unsigned>(s_pref[B]);
      const unsigned c = blockIdx.x + lane;
      int64_t stage;
      if (c >= gridDim.x) stage = n_stages;
      else {
        // floor(total_cost * c / grid) up to the rounding of one double division: any monotone function of c works
        const unsigned x = static_cast<unsigned>(static_cast<double>(total_cost) * static_cast<double>(c) / static_cast<double>(gridDim.x));
        stage = cost_to_stage(x, cm, s_pref, s_nb, B, first_stage, n_stages);
      }
      s_range[lane] = stage;
    }
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(&s_range_bar);                   // release: the consumers read s_range after their wait
      const int64_t s_lo = s_range[0], s_hi = s_range[1];
      if (stamp) { prm.stamps[blockIdx.x * 16 + 1] = global_ns(); prm.stamps[blockIdx.x * 16 + 5] = static_cast<unsigned long long>(s_hi - s_lo); }
      StageCursor cur;
      cur.init(s_lo, s_hi, T, total_rows);
      while (!cur.done()) {
        int64_t run = cur.full_run(s_nb);
        if (run > 0) {                                   // a run of fully valid stages of one utterance
          for (int64_t i = 0; i < run; ++i) load_stage(cur.stage + i);
          cur.advance(run);
          continue;
        }
        run = cur.pad_run(s_nb);
        if (run > 0) {                                   // a run of padding-only stages: zero gradient, nothing to load
          if (GRAD && !(prm.debug & 4)) {
            for (int64_t i = 0; i < run; ++i) {
              mg_bulk_store_hint(prm.grad + (cur.stage + i) * stage_elems, mg_smem_addr(s_zero), stage_bytes, policy);
              mg_bulk_commit();
            }
          }
          cur.advance(run);
          continue;
        }
        const int kind = classify(cur, s_nb, B);         // boundary stages
        if (kind == STAGE_PAD) {
          if (GRAD && !(prm.debug & 4)) { mg_bulk_store_hint(prm.grad + cur.stage * stage_elems, mg_smem_addr(s_zero), stage_bytes, policy); mg_bulk_commit(); }
        } else if (kind != STAGE_TAIL) {
          load_stage(cur.stage);
        }
        cur.next();
      }
      if (GRAD) mg_bulk_wait_read<0>();   // the zero tile has been read by every bulk store (their writes complete asynchronously)
    }
  } else if (warp < n_cw) {
    // ---- consumers: thread t owns column t; warp w also evaluates special column w (in column order) in batches of 32 rows ----
    const bool active = tid < D;
    const mg_column col = active ? s_cols[tid] : s_cols[0];
    const bool simple = active && column_is_simple(col);
    // special columns in column order: the first min(n_cw, kMaxSp) of them are evaluated in batches, warp w takes the w-th; any
    // further one is evaluated by the thread that owns the column
    int my_rank = -1, my_special_col = -1;
    {
      int n_before = 0;
#pragma unroll 1
      for (int c0 = 0; c0 < D; c0 += 32) {
        const int k = c0 + lane;
        const bool special = k < D && !column_is_simple(s_cols[k]);
        const unsigned bits = __ballot_sync(MG_FULL_MASK, special);
        if (c0 == warp * 32 && special) my_rank = n_before + __popc(bits & ((1u << lane) - 1));
        const int want = warp - n_before;          // the warp-th special column sits in this chunk?
        if (want >= 0 && want < __popc(bits)) my_special_col = c0 + __fns(bits, 0, want + 1);
        n_before += __popc(bits);
      }
    }
    const int sp_cap = kSpPerWarp * n_cw < kMaxSp ? kSpPerWarp * n_cw : kMaxSp;
    const bool served = active && !simple && my_rank >= 0 && my_rank < sp_cap;   // evaluated in batches by warp `my_rank`
    const bool own_special = active && !simple && !served;   // any further special column: its own thread evaluates it
    const bool use_loss = (simple || own_special) && col.loss_kind != MG_COL_NONE;
    const bool use_metric = (simple || own_special) && col.metric_kind != MG_COL_NONE;
    const bool loss_sq = col.loss_kind == MG_RED_SQDIFF, metric_sq = col.metric_kind == MG_RED_SQDIFF;
    const bool hot = simple && use_loss && loss_sq && (metric_sq || !use_metric);
    int b_cur = -1;
    double n_cur = 1.;
    float w_row = 0.f, w2 = 0.f;
    double cur_l = 0., cur_m = 0., cur_n = 0., sum_l = 0., wsum_l = 0., sum_m = 0., sum_n = 0.;
    auto enter = [&](int b) {   // the utterance changes: fold the finished one into the weighted sums
      if (b == b_cur) return;
      sum_l += cur_l; wsum_l += div_f64(cur_l, n_cur); sum_m += cur_m; sum_n += cur_n;
      cur_l = cur_m = cur_n = 0.;
      b_cur = b;
      n_cur = b >= 0 ? static_cast<double>(s_nb[b]) : 1.;
      w_row = use_loss ? static_cast<float>(static_cast<double>(col.loss_weight) * div_f64(scale_over_b, n_cur)) : 0.f;
      w2 = __fmul_rn(2.f, w_row);
    };
    // rows of a stage one at a time, from shared or global memory (mixed / tail stages and non-squared programs)
    auto slow_rows = [&](const float* sp, const float* sy, float* g, const StageCursor& cur, int n_rows) {
      int b = cur.b;
      int64_t t = cur.t0;
#pragma unroll 1
      for (int u = 0; u < n_rows; ++u) {
        const bool valid = t < s_nb[b];
        if (valid && own_special) {
          enter(b);
          const float* row_p = sp + u * D - tid;
          const float* row_y = sy + u * D - tid;
          const float root = (col.metric_kind == MG_RED_ROOT_SQDIFF && col.width > 1) ? root_group_acc(col, tid, row_p, row_y) : 0.f;
          const Contribution c = special_value<GRAD>(col, row_p[tid], row_y[tid], col.mask_col != MG_COL_NONE ? row_p[col.mask_col] : 1.f, root,
                                                     GRAD ? g + u * D : nullptr, w_row);
          cur_l += c.l; cur_m += c.m; cur_n += c.n;
        } else if (valid && simple) {
          enter(b);
          const float d = __fsub_rn(sp[u * D], sy[u * D]);
          const float sq = __fmul_rn(d, d), ab = fabsf(d);
          if (use_loss) cur_l += static_cast<double>(loss_sq ? sq : ab);
          if (use_metric) cur_m += static_cast<double>(metric_sq ? sq : ab);
          if (GRAD) __stcs(g + u * D, __fmul_rn(simple_slope(loss_sq, d), w_row));
        } else if (GRAD && !valid) {
          __stcs(g + u * D, 0.f);      // padding rows, every column (special ones included)
        }
        if (++t >= T) { t = 0; ++b; }
      }
    };

    // -- special column of this warp: operands of 32 rows (four stages) are parked in registers, lane = row, then the
    // column is evaluated by the whole warp on one code path (~100 instructions per element: BCE, exp, root ...) ------------
    int n_my = 0;
    int my_k[kSpPerWarp];
    mg_column my_sc[kSpPerWarp];
#pragma unroll
    for (int j = 0; j < kSpPerWarp; ++j) {
      my_k[j] = 0;
      my_sc[j] = s_cols[0];
      if (j == 0 && my_special_col >= 0 && warp < sp_cap && !(prm.debug & 2)) { my_k[j] = my_special_col; my_sc[j] = s_cols[my_special_col]; n_my = j + 1; }
    }
    float b_pv[kSpPerWarp], b_yv[kSpPerWarp], b_mv[kSpPerWarp], b_root[kSpPerWarp];
    int b_utt = -1;               // utterance of this lane's parked row, -1: nothing parked
    int64_t b_off = 0;            // float offset of that row in the tensors
    double sp_sum_l[kSpPerWarp], sp_w_l[kSpPerWarp], sp_sum_m[kSpPerWarp], sp_sum_n[kSpPerWarp];
#pragma unroll
    for (int j = 0; j < kSpPerWarp; ++j) { sp_sum_l[j] = sp_w_l[j] = sp_sum_m[j] = sp_sum_n[j] = 0.; b_pv[j] = b_yv[j] = b_mv[j] = b_root[j] = 0.f; }
    int batch_pos = 0;
    auto park = [&](const float* stage_p, const float* stage_y, const StageCursor& cur, int64_t off, int kind) {
      // lanes [8 * batch_pos, 8 * batch_pos + 8) take the stage's rows
      const int u = lane & 7;
      if ((lane >> 3) == batch_pos) {
        int b = cur.b;
        int64_t t = cur.t0 + u;
        if (kind != STAGE_FULL) while (t >= T && b < B - 1) { t -= T; ++b; }
        const bool valid = kind == STAGE_FULL || (u < cur.rows() && t < T && t < s_nb[b]);
        b_utt = valid ? b : -1;
        b_off = off + static_cast<int64_t>(u) * D;
        if (valid) {
          const float* row_p = stage_p + u * D;
          const float* row_y = stage_y + u * D;
#pragma unroll
          for (int j = 0; j < kSpPerWarp; ++j) {
            if (j < n_my) {
              const mg_column& sc = my_sc[j];
              b_pv[j] = row_p[my_k[j]];
              b_yv[j] = row_y[my_k[j]];
              b_mv[j] = sc.mask_col != MG_COL_NONE ? row_p[sc.mask_col] : 1.f;
              b_root[j] = (sc.metric_kind == MG_RED_ROOT_SQDIFF && sc.width > 1) ? root_group_acc(sc, my_k[j], row_p, row_y) : 0.f;
            }
          }
        }
      }
      ++batch_pos;
    };
    auto evaluate = [&]() {
      if (b_utt >= 0) {
        const double n_b = static_cast<double>(s_nb[b_utt]);
#pragma unroll
        for (int j = 0; j < kSpPerWarp; ++j) {
          if (j < n_my) {   // warp-uniform
            const mg_column& sc = my_sc[j];
            const float w = static_cast<float>(static_cast<double>(sc.loss_weight) * div_f64(scale_over_b, n_b));
            const Contribution c = special_value<GRAD>(sc, b_pv[j], b_yv[j], b_mv[j], b_root[j], GRAD ? prm.grad + b_off + my_k[j] : nullptr, w);
            sp_sum_l[j] += c.l; sp_w_l[j] += div_f64(c.l, n_b); sp_sum_m[j] += c.m; sp_sum_n[j] += c.n;
          }
        }
      }
      b_utt = -1;
      batch_pos = 0;
    };

    // one fully valid stage of utterance cur.b in ring slot `slot` (the hot path)
    const bool generic_lane = active && !hot && !served;    // absolute-error programs, metric-only columns, a 7th special column ...
    auto full_stage = [&](const float* stage_p, int64_t off, const StageCursor& cur) {
      if (hot && !(prm.debug & 1)) {
        const float* sp = stage_p + tid;
        const float* sy = sp + stage_elems;
        float pv[kRows], yv[kRows];
#pragma unroll
        for (int u = 0; u < kRows; ++u) { pv[u] = sp[u * D]; yv[u] = sy[u * D]; }    // 16 loads in flight, then the arithmetic
        float part = 0.f;     // 8 rows in fp32, row order, then one fp64 add
        float* g = GRAD ? prm.grad + off + tid : nullptr;
#pragma unroll
        for (int u = 0; u < kRows; ++u) {
          const float d = __fsub_rn(pv[u], yv[u]);
          part = __fadd_rn(part, __fmul_rn(d, d));
          if (GRAD) { if (prm.debug & 32) g[u * D] = __fmul_rn(d, w2); else __stcs(g + u * D, __fmul_rn(d, w2)); }
        }
        cur_l += static_cast<double>(part);
        if (use_metric) cur_m += static_cast<double>(part);
      } else if (generic_lane && !(prm.debug & 1)) {
        slow_rows(stage_p + tid, stage_p + stage_elems + tid, GRAD ? prm.grad + off + tid : nullptr, cur, kRows);
      }
    };

    int slot = 0;
    uint32_t phase = 0;
    int n_loaded = 0;    // stages that went through the ring (debug stamps)
    if (stamp && tid == 0) prm.stamps[blockIdx.x * 16 + 9] = global_ns();
    // pass 0: the head start (every full-size stage was loaded, padding-only ones included: their zero gradient is written here);
    // pass 1: this CTA's range of the partition (padding-only stages are the producer's)
    for (int pass = 0; pass < 2; ++pass) {
      StageCursor cur;
      if (pass == 0) {
        cur.init(head_lo, head_hi, T, total_rows);
      } else {
        mg_mbar_wait(&s_range_bar, 0);
        cur.init(s_range[0], s_range[1], T, total_rows);
      }
      while (!cur.done()) {
        int64_t run = cur.full_run(s_nb);
        if (run > 0) {
          n_loaded += static_cast<int>(run);
          enter(cur.b);
          for (int64_t i = 0; i < run; ++i) {
            const int64_t off = cur.stage * stage_elems;
            mg_mbar_wait(&s_full[slot], phase);
            if (stamp && tid == 0 && prm.stamps[blockIdx.x * 16 + 2] == 0) prm.stamps[blockIdx.x * 16 + 2] = global_ns();
            const float* stage_p = s_ring + static_cast<size_t>(slot) * 2 * stage_elems;
            full_stage(stage_p, off, cur);
            if (n_my > 0) park(stage_p, stage_p + stage_elems, cur, off, STAGE_FULL);
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[slot]);   // the stage's operands are in registers: the slot may be refilled
            if (++slot == ring) { slot = 0; phase ^= 1u; }
            if (batch_pos == 4) evaluate();
            cur.stage += 1;
            cur.t0 += kRows;
          }
          cur.advance(0);
          continue;
        }
        if (pass == 1) {
          run = cur.pad_run(s_nb);
          if (run > 0) { cur.advance(run); continue; }
        }
        const int kind = classify(cur, s_nb, B);
        if (kind == STAGE_PAD && pass == 1) { cur.next(); continue; }
        const int64_t off = cur.stage * stage_elems;
        float* g = GRAD ? prm.grad + off + tid : nullptr;
        if (kind == STAGE_TAIL) {
          if (active) slow_rows(prm.pred + off + tid, prm.target + off + tid, g, cur, cur.rows());
          if (n_my > 0) { park(prm.pred + off, prm.target + off, cur, off, kind); if (batch_pos == 4) evaluate(); }
          cur.next();
          continue;
        }
        ++n_loaded;
        mg_mbar_wait(&s_full[slot], phase);
        if (stamp && tid == 0 && prm.stamps[blockIdx.x * 16 + 2] == 0) prm.stamps[blockIdx.x * 16 + 2] = global_ns();
        const float* stage_p = s_ring + static_cast<size_t>(slot) * 2 * stage_elems;
        if (kind == STAGE_PAD) {
          // a padding-only stage of the head start: it was loaded before anything was known; its gradient rows are zero
          if (GRAD && active) {
#pragma unroll 1
            for (int u = 0; u < kRows; ++u) __stcs(g + u * D, 0.f);
          }
        } else if (active && !(prm.debug & 1)) {
          if ((hot || served) && T >= kRows) {
            // a boundary stage is at most: valid rows of utterance b | padding | valid rows of utterance b + 1 | padding
            const int rows_a = static_cast<int>(min(static_cast<int64_t>(kRows), T - cur.t0));
            const int valid_a = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(rows_a), s_nb[cur.b] - cur.t0)));
            const int valid_b = rows_a < kRows ? min(kRows - rows_a, s_nb[cur.b + 1]) : 0;
            const float* sp = stage_p + tid;
            const float* sy = sp + stage_elems;
            auto valid_rows = [&](int u0, int u1, int b) {
              if (u1 <= u0 || !hot) return;
              enter(b);
              float part = 0.f;
#pragma unroll 1
              for (int u = u0; u < u1; ++u) {
                const float d = __fsub_rn(sp[u * D], sy[u * D]);
                part = __fadd_rn(part, __fmul_rn(d, d));
                if (GRAD) __stcs(g + u * D, __fmul_rn(d, w2));
              }
              cur_l += static_cast<double>(part);
              if (use_metric) cur_m += static_cast<double>(part);
            };
            valid_rows(0, valid_a, cur.b);
            valid_rows(rows_a, rows_a + valid_b, cur.b + 1);
            if (GRAD) {
#pragma unroll 1
              for (int u = valid_a; u < rows_a; ++u) __stcs(g + u * D, 0.f);
#pragma unroll 1
              for (int u = rows_a + valid_b; u < kRows; ++u) __stcs(g + u * D, 0.f);
            }
          } else {
            slow_rows(stage_p + tid, stage_p + stage_elems + tid, g, cur, kRows);
          }
        }
        if (n_my > 0 && kind != STAGE_PAD) park(stage_p, stage_p + stage_elems, cur, off, kind);
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_empty[slot]);
        if (++slot == ring) { slot = 0; phase ^= 1u; }
        if (batch_pos == 4) evaluate();
        cur.next();
      }
    }
    if (stamp && tid == 0) { prm.stamps[blockIdx.x * 16 + 3] = global_ns(); prm.stamps[blockIdx.x * 16 + 7] = static_cast<unsigned long long>(n_loaded); }

    // ---- the CTA's record: the lanes of a warp that feed slot s are summed with a fixed shuffle tree (slots in lane order of
    // their first column), the warps are added in warp order by warp 0 ---------------------------------------------------------
    if (n_my > 0) evaluate();
    enter(-2);   // fold the last utterance (b = -2 never matches)
    double* mine = &s_flush[warp][0][0];
#pragma unroll 1
    for (int i = lane; i < 3 * prm.n_slots; i += 32) mine[i] = 0.;
    __syncwarp();
    const bool warp_counts = __any_sync(MG_FULL_MASK, own_special && use_metric);   // only per-thread special columns count frames here
    fold_lanes(mine, use_loss ? col.loss_slot : -1, sum_l, wsum_l, 1, true);
    fold_lanes(mine, use_metric ? col.metric_slot : -1, sum_m, sum_n, 2, warp_counts);
#pragma unroll
    for (int j = 0; j < kSpPerWarp; ++j) {
      if (j < n_my) {   // warp-uniform
        fold_lanes(mine, my_sc[j].loss_kind != MG_COL_NONE ? my_sc[j].loss_slot : -1, sp_sum_l[j], sp_w_l[j], 1, true);
        fold_lanes(mine, my_sc[j].metric_kind != MG_COL_NONE ? my_sc[j].metric_slot : -1, sp_sum_m[j], sp_sum_n[j], 2, true);
      }
    }
    consumer_barrier(n_cw * 32);
    if (warp == 0 && lane < prm.n_slots) {
      double a = 0., w = 0., n = 0.;
#pragma unroll 1
      for (int cw = 0; cw < n_cw; ++cw) { a += s_flush[cw][lane][0]; w += s_flush[cw][lane][1]; n += s_flush[cw][lane][2]; }
      double2* out = prm.records + (static_cast<int64_t>(lane) * gridDim.x + blockIdx.x) * 2;
      const uint64_t keep = mg_policy_evict_last();
      mg_st_keep_f64x2(out, make_double2(a, w), keep);
      mg_st_keep_f64x2(out + 1, make_double2(n, 0.), keep);
    }
  }

  // ---- ticket: the last CTA adds the records in CTA order and writes the result records.  The records were written by warp 0
  // before the barrier; the fence of the ticket thread is cumulative over what the barrier ordered before it. -----------------
  __syncthreads();
  if (tid == producer_warp * 32) {
    __threadfence();
    if (stamp) prm.stamps[blockIdx.x * 16 + 4] = global_ns();
    const bool last = atomicAdd(prm.ticket, 1u) == gridDim.x - 1;
    if (last) __threadfence();     // acquire: the other threads read after the barrier below, through L2 (__ldcg)
    s_is_last = last;
  }
  __syncthreads();
  if (!s_is_last) return;
  if (stamp && tid == 0) prm.stamps[blockIdx.x * 16 + 10] = global_ns();
  {
    // records [slot][cta]: a warp per slot, lane i takes CTAs i, i + 32, ... in CTA order (consecutive lanes on consecutive
    // records, kFinishLoads of them in flight), the lanes are summed with a fixed shuffle tree and lane 0 writes the slot's result
    const unsigned n_rec = gridDim.x;
    const int n_warps = blockDim.x >> 5;
    const uint64_t keep = mg_policy_evict_last();
    for (int sl_i = warp; sl_i < prm.n_slots; sl_i += n_warps) {
      const MgFinishSlot& sl = prm.slots[sl_i];
      const bool weighted = sl.weighted != 0;
      double old_s = 0., old_c = 0.;
      if (lane == 0 && sl.accumulate) { old_s = sl.result->sum; old_c = sl.result->count; }   // streaming-metric state, in flight under the loads
      const double2* rec = prm.records + static_cast<int64_t>(sl_i) * n_rec * 2;
      double s = 0., w = 0., n = 0.;
      for (unsigned j0 = lane; j0 < n_rec; j0 += kFinishLoads * 32) {
        double2 v0[kFinishLoads];
        double v1[kFinishLoads];
#pragma unroll
        for (int u = 0; u < kFinishLoads; ++u) {
          const unsigned j = j0 + u * 32;
          v0[u] = make_double2(0., 0.);
          v1[u] = 0.;
          if (j < n_rec) {
            v0[u] = mg_ld_keep_cg_f64x2(rec + 2 * static_cast<int64_t>(j), keep);
            if (weighted) v1[u] = mg_ld_keep_cg_f64x2(rec + 2 * static_cast<int64_t>(j) + 1, keep).x;
          }
        }
#pragma unroll
        for (int u = 0; u < kFinishLoads; ++u) { s += v0[u].x; w += v0[u].y; n += v1[u]; }
      }
      if (stamp && tid == 0) prm.stamps[blockIdx.x * 16 + 11] = global_ns();
      s = mg_warp_sum(s); w = mg_warp_sum(w);
      if (weighted) n = mg_warp_sum(n);
      if (stamp && tid == 0) prm.stamps[blockIdx.x * 16 + 15] = global_ns();
      if (lane == 0) {
        double c;
        double l = w / (static_cast<double>(B) * static_cast<double>(sl.D));   // torch.mean over (B, D), losses.py:42
        if (s_has_empty) l = __longlong_as_double(0x7ff8000000000000LL);   // an empty utterance contributes 0 / 0 (losses.py:39)
        if (weighted) c = n;
        else if (prm.seq_len != nullptr) c = static_cast<double>(s_valid_total);                          // frames (metrics.py:393-394)
        else c = static_cast<double>(B) * static_cast<double>(T) * (sl.per_frame ? 1. : static_cast<double>(sl.D));   // numel (:390)
        s += old_s;
        c += old_c;
        mg_term_result res;
        res.sum = s;
        res.count = c;
        res.loss = l;
        res.isum = static_cast<int64_t>(s);
        res.sum_f32 = static_cast<float>(s);
        res.count_f32 = static_cast<float>(c);
        res.loss_f32 = static_cast<float>(l);
        res.weighted_loss_f32 = 0.f;
        *sl.result = res;
        s_slot[sl_i][0] = sl.in_total ? static_cast<double>(sl.weight) * l : 0.;
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    double weighted_total = 0.;
    for (int t = 0; t < prm.n_slots; ++t) weighted_total += s_slot[t][0];
    prm.slots[0].result->weighted_loss_f32 = static_cast<float>(weighted_total);
    *prm.ticket = 0u;   // leave the workspace clean for the next launch
    if (stamp) prm.stamps[blockIdx.x * 16 + 6] = global_ns();
  }
}

int env_int(const char* name, int fallback) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : fallback;
}

// Tuning knobs, read from the environment once per process (a dozen getenv calls per launch cost ~1.5 us of host time).
struct Tuning {
  int cost_valid_grad, cost_valid_fwd, cost_pad, ctas_per_sm, ring, debug, load_first_pct;
  Tuning()
      : cost_valid_grad(env_int("MG_OBJ_COST_VALID", 6)), cost_valid_fwd(env_int("MG_OBJ_COST_VALID", 2)), cost_pad(env_int("MG_OBJ_COST_PAD", 1)),
        ctas_per_sm(env_int("MG_OBJ_CTAS_PER_SM", 2)), ring(env_int("MG_OBJ_RING", 0)), debug(env_int("MG_OBJ_DEBUG", 0)),
        load_first_pct(env_int("MG_OBJ_LOAD_FIRST_PCT", 100)) {}
};

// Launch geometry: consumer warps, threads, ring depth, shared memory and grid of the stream kernel for one problem size.
struct StreamGeometry { int n_cw, threads, ring; size_t smem; int64_t grid, n_stages; };

StreamGeometry stream_geometry(int B, int64_t T, int D, bool has_grad, int n_slots, int sms, int ctas_per_sm, int ring_override) {
  StreamGeometry g;
  g.n_cw = (D + 31) / 32;
  int n_warps = g.n_cw + 1;                      // consumers + producer ...
  if (n_warps < n_slots) n_warps = n_slots < kMaxWarps ? n_slots : kMaxWarps;   // ... + idle warps so that the last CTA has a warp per slot
  g.threads = n_warps * 32;
  const size_t stage_pair = static_cast<size_t>(2) * kRows * D * sizeof(float);
  const size_t zero_bytes = has_grad ? static_cast<size_t>(kRows) * D * sizeof(float) : 0;
  // shared memory per CTA: 227 KB per SM minus ~13 KB of static arrays and 1 KB of reserve per CTA
  const size_t budget = (static_cast<size_t>(227) * 1024) / ctas_per_sm - 15 * 1024;
  int ring = ring_override;
  // measured at config 2 (D = 187, two CTAs per SM): ring 2 / 3 / 4 / 5 / 6 / 8 -> 0.140 / 0.124 / 0.123 / 0.132 / 0.136 / 0.138 ms:
  // four 12 KB stages per CTA cover the latency; a deeper ring only takes shared memory away from L1
  if (ring <= 0) {
    ring = budget > zero_bytes ? static_cast<int>((budget - zero_bytes) / stage_pair) : 0;
    if (ring > 4) ring = 4;
  }
  if (ring > kMaxRing) ring = kMaxRing;
  g.ring = ring;
  g.smem = static_cast<size_t>(ring) * stage_pair + zero_bytes;
  g.n_stages = (static_cast<int64_t>(B) * T + kRows - 1) / kRows;
  int64_t grid = static_cast<int64_t>(sms) * ctas_per_sm;
  if (grid > g.n_stages / 4) grid = g.n_stages / 4 > 0 ? g.n_stages / 4 : 1;     // tiny batches: at least four stages per CTA
  if (grid > kMaxCtas) grid = kMaxCtas;
  g.grid = grid;
  return g;
}

}  // namespace

int mg_objective_stream_launch(const MgObjectiveArgs& a, cudaStream_t stream) {
  static const int enabled = env_int("MG_OBJECTIVE_STREAM", 1);
  if (!enabled) return MG_STREAM_NOT_APPLICABLE;
  const int D = a.D, B = a.B;
  const int64_t T = a.T;
  static const Tuning tune;
  const int cost_valid = a.grad != nullptr ? tune.cost_valid_grad : tune.cost_valid_fwd;
  const int cost_pad = tune.cost_pad;
  // preconditions: contiguous (B, T, D) tensors on 16-byte aligned bases, D <= 224, B <= 1024, costs in 32 bits
  if (D > (kMaxWarps - 1) * 32 || B > kMaxB || T < 1 || cost_valid < 1 || cost_valid > 16 || cost_pad < 0 || cost_pad > 16) return MG_STREAM_NOT_APPLICABLE;
  if (static_cast<int64_t>(B) * T * (cost_valid > cost_pad ? cost_valid : cost_pad) >= (int64_t(1) << 31) - (int64_t(1) << 24)) return MG_STREAM_NOT_APPLICABLE;
  if (a.p_st != D || a.t_st != D || a.p_sb != T * D || a.t_sb != T * D) return MG_STREAM_NOT_APPLICABLE;
  if (!mg_aligned(a.pred, 16) || !mg_aligned(a.target, 16)) return MG_STREAM_NOT_APPLICABLE;
  if (a.grad != nullptr && (a.g_st != D || a.g_sb != T * D || !mg_aligned(a.grad, 16))) return MG_STREAM_NOT_APPLICABLE;
  if (static_cast<int64_t>(B) * T < 4 * kRows) return MG_STREAM_NOT_APPLICABLE;

  const StreamGeometry geo = stream_geometry(B, T, D, a.grad != nullptr, a.n_slots, mg_cached_sm_count(), tune.ctas_per_sm, tune.ring);
  if (geo.ring < 2) return MG_STREAM_NOT_APPLICABLE;
  const int n_cw = geo.n_cw, threads = geo.threads, ring = geo.ring;
  const size_t smem = geo.smem;
  const int64_t grid = geo.grid;
  const int64_t need = kMgTicketBytes + grid * a.n_slots * static_cast<int64_t>(2 * sizeof(double2));
  if (a.workspace_bytes < need) return MG_STREAM_NOT_APPLICABLE;

  // special columns: the stream serves at most kMaxSp of them, each with its dependencies inside the row
  // (checked on the device-side table by the caller's contract: mask_col / group columns are < D)
  StreamParams prm;
  memset(&prm, 0, sizeof(prm));
  for (int i = 0; i < a.n_slots; ++i) {
    MgFinishSlot& sl = prm.slots[i];
    sl.result = a.slots[i].result;
    sl.D = a.slots[i].D;
    sl.per_frame = a.slots[i].per_frame;
    sl.weighted = a.slots[i].weighted;
    sl.accumulate = a.slots[i].accumulate;
    sl.in_total = a.slots[i].in_total;
    sl.weight = a.slots[i].weight;
  }
  prm.pred = a.pred; prm.target = a.target; prm.grad = a.grad; prm.grad_scale_dev = a.grad_scale_dev;
  prm.cols = a.cols; prm.seq_len = a.seq_len;
  unsigned char* base = static_cast<unsigned char*>(a.workspace);
  prm.ticket = reinterpret_cast<unsigned int*>(base);
  prm.records = reinterpret_cast<double2*>(base + kMgTicketBytes);
  prm.T = T; prm.D = D; prm.B = B; prm.n_slots = a.n_slots; prm.ring = ring; prm.n_consumer_warps = n_cw;
  prm.debug = tune.debug;
  prm.stamps = reinterpret_cast<unsigned long long*>(base + kMgTicketBytes + 512 * 1024);
  if ((prm.debug & 8) && a.workspace_bytes < kMgTicketBytes + 512 * 1024 + grid * 128) prm.debug &= ~8;
  prm.load_first_frac = static_cast<float>(tune.load_first_pct) / 100.f;
  prm.cost_valid = cost_valid;
  prm.cost_pad = cost_pad;

  if (smem > 200 * 1024) return MG_STREAM_NOT_APPLICABLE;
  static bool attr_done[64] = {};   // the attribute is per device (several GPUs in one process are allowed)
  int device = 0;
  MG_CUDA_OK(cudaGetDevice(&device));
  if (!attr_done[device & 63]) {
    MG_CUDA_OK(cudaFuncSetAttribute(objective_stream_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    MG_CUDA_OK(cudaFuncSetAttribute(objective_stream_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_done[device & 63] = true;
  }
  if (a.grad != nullptr) MG_CUDA_OK(mg_launch_pdl(objective_stream_kernel<true>, dim3(static_cast<unsigned>(grid)), dim3(threads), smem, stream, prm));
  else MG_CUDA_OK(mg_launch_pdl(objective_stream_kernel<false>, dim3(static_cast<unsigned>(grid)), dim3(threads), smem, stream, prm));
  MG_LAUNCH_OK();
  return MG_OK;
}

// Host-only: the stage -> CTA partition of the stream kernel for given utterance lengths, computed with the SAME functions the
// device runs (CostModel, cost_to_stage, the double division of the range ends).  out[0] = grid, out[1] = first stage behind the head
// starts, out[2] = number of stages, out[3 .. 3 + grid] = the range boundaries (CTA c owns [out[3 + c], out[4 + c])).  Returns
// MG_STREAM_NOT_APPLICABLE for shapes the stream kernel does not take.  Used by the CPU tests to check the partition's invariants.
extern "C" int mg_objective_stream_plan(const int64_t* seq_len_host, int B, int64_t T, int D, int has_grad, int n_slots, int sms,
                                        int64_t* out, int64_t out_len) {
  MG_REQUIRE(B >= 1 && T >= 1 && D >= 1 && n_slots >= 1 && n_slots <= MG_MAX_TERMS && sms >= 1 && out != nullptr, "mg_objective_stream_plan: bad argument");
  const int cost_valid = has_grad ? 6 : 2, cost_pad = has_grad ? 1 : 0;
  if (D > (kMaxWarps - 1) * 32 || B > kMaxB) return MG_STREAM_NOT_APPLICABLE;
  if (static_cast<int64_t>(B) * T * cost_valid >= (int64_t(1) << 31) - (int64_t(1) << 24)) return MG_STREAM_NOT_APPLICABLE;
  if (static_cast<int64_t>(B) * T < 4 * kRows) return MG_STREAM_NOT_APPLICABLE;
  const StreamGeometry geo = stream_geometry(B, T, D, has_grad != 0, n_slots, sms, 2, 0);
  if (geo.ring < 2) return MG_STREAM_NOT_APPLICABLE;
  MG_REQUIRE(out_len >= 4 + geo.grid, "mg_objective_stream_plan: out holds %lld values, %lld needed", static_cast<long long>(out_len), static_cast<long long>(4 + geo.grid));
  const int n_head = geo.ring < kHead ? geo.ring : kHead;
  const int64_t first_stage = geo.grid * n_head < geo.n_stages ? geo.grid * n_head : geo.n_stages;
  CostModel cm;
  cm.T = T; cm.first_row = first_stage * kRows; cm.cost_valid = static_cast<unsigned>(cost_valid); cm.cost_pad = static_cast<unsigned>(cost_pad);
  int* nb = static_cast<int*>(malloc(sizeof(int) * (2 * static_cast<size_t>(B) + 1)));
  MG_REQUIRE(nb != nullptr, "mg_objective_stream_plan: out of memory");
  int* pref = nb + B;
  unsigned carry = 0;
  for (int b = 0; b < B; ++b) {
    int64_t n = seq_len_host != nullptr ? seq_len_host[b] : T;
    n = n < 0 ? 0 : (n > T ? T : n);
    nb[b] = static_cast<int>(n);
    pref[b] = static_cast<int>(carry);
    carry += cm.cost(b, static_cast<unsigned>(n));
  }
  pref[B] = static_cast<int>(carry);
  out[0] = geo.grid; out[1] = first_stage; out[2] = geo.n_stages;
  for (int64_t c = 0; c <= geo.grid; ++c) {
    int64_t stage;
    if (c >= geo.grid) stage = geo.n_stages;
    else {
      const unsigned x = static_cast<unsigned>(static_cast<double>(carry) * static_cast<double>(c) / static_cast<double>(geo.grid));
      stage = cost_to_stage(x, cm, pref, nb, B, first_stage, geo.n_stages);
    }
    out[3 + c] = stage;
  }
  free(nb);
  return MG_OK;
}
