// Shared tail of the masked reductions.  Two integer tickets, no floating-point atomics, fixed summation order:
//   1. every CTA of utterance b bumps ticket[b]; the LAST one folds that utterance's per-CTA slots (chunk order, fixed
//      shuffle tree) into one (sum, count) per slot -- this runs while other utterances are still streaming;
//   2. those B "utterance finishers" bump the global ticket; the last one sums the B per-utterance values per slot in
//      index order and writes the result records.
// The serial tail after the last byte is therefore ~B loads per slot, not B x chunks.
#pragma once

#include "mg_common.cuh"

constexpr int kMgMaxChunks = 64;   // CTAs (row chunks) per utterance per slot: bounds the workspace
constexpr int kMgMaxSlots = MG_MAX_TERMS;

struct MgFinishSlot {
  mg_term_result* result;
  int D;             // columns feeding the slot: loss normalisation 1 / (B * D); numel count without seq_len
  int rows_per_cta;  // chunking of this slot's partials
  int n_chunks;
  int per_frame;     // one value per frame (ROOT_SQDIFF, weighted metrics): numel count = B * T
  int weighted;      // count = sum of the per-frame weights accumulated by the CTAs
  int accumulate;    // result.sum / count += this batch
  int in_total;      // contributes weight * loss to results[0].weighted_loss_f32
  float weight;
};

// Workspace carved by the host: [global ticket | per-utterance tickets | per-utterance totals | per-CTA partials].
// The kernels leave every TICKET at zero, so a workspace is zeroed once and reused; totals and partials are overwritten
// before they are read and may hold anything.  The ticket area therefore has a FIXED size (the largest supported batch):
// with a batch-dependent size, the tickets of a larger batch would land on stale totals / partials of a smaller one.
constexpr int64_t kMgMaxBatch = 65536;
constexpr int64_t kMgTicketBytes = 256 + kMgMaxBatch * 4;
struct MgWorkspace {
  unsigned int* ticket;       // 1
  unsigned int* utt_ticket;   // [B]
  double2* utt_total;         // [slot][B]
  double2* partials;          // [slot][B][kMgMaxChunks]
};

static inline int64_t mg_workspace_bytes(int n_slots, int B) {
  return kMgTicketBytes + static_cast<int64_t>(n_slots) * B * (1 + kMgMaxChunks) * static_cast<int64_t>(sizeof(double2));
}

static inline MgWorkspace mg_carve_workspace(void* workspace, int n_slots, int B) {
  MgWorkspace ws;
  unsigned char* base = static_cast<unsigned char*>(workspace);
  ws.ticket = reinterpret_cast<unsigned int*>(base);
  ws.utt_ticket = reinterpret_cast<unsigned int*>(base + 256);
  base += kMgTicketBytes;
  ws.utt_total = reinterpret_cast<double2*>(base);
  ws.partials = ws.utt_total + static_cast<int64_t>(n_slots) * B;
  return ws;
}

__device__ __forceinline__ int64_t mg_valid_frames(const int64_t* seq_len, int b, int64_t T) {
  if (seq_len == nullptr) return T;
  const int64_t n = __ldg(seq_len + b);
  return n < 0 ? 0 : (n > T ? T : n);   // mask = arange(T) < seq_len  (reference utils.py:140-142)
}

// Returns true in every thread of exactly one CTA among `total` arrivals on `ticket`: the last one.
// `writers_in_warp0`: the data the winner will read was written by threads of warp 0 only, so only they need the
// release fence (a fence in every thread also waits for that thread's unrelated streaming stores to drain).
__device__ __forceinline__ bool mg_take_ticket(unsigned int* ticket, unsigned total, bool* s_flag, bool writers_in_warp0 = false) {
  if (!writers_in_warp0 || threadIdx.x < 32) __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) *s_flag = atomicAdd(ticket, 1u) == total - 1;
  __syncthreads();
  const bool last = *s_flag;
  if (last) __threadfence();
  return last;
}

// Step 1b, run by the last CTA of utterance b: chunk partials (+ optional per-slot extras already summed for the whole
// utterance, in shared memory) -> one (sum, count) per slot in ws.utt_total.
__device__ __forceinline__ void mg_fold_utterance(const MgFinishSlot* slots, int n_slots, const int64_t* seq_len, int B,
                                                  int64_t T, const MgWorkspace& ws, int b, const double* extra_sum,
                                                  const double* extra_cnt) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  const int64_t n_b = mg_valid_frames(seq_len, b, T);
  for (int t = warp; t < n_slots; t += n_warps) {
    const int64_t R = slots[t].rows_per_cta;
    const int used = static_cast<int>(min(static_cast<int64_t>(slots[t].n_chunks), (n_b + R - 1) / R));
    const double2* part = ws.partials + (static_cast<int64_t>(t) * B + b) * kMgMaxChunks;
    double s = 0., c = 0.;
    for (int k = lane; k < used; k += 32) {   // <= 2 independent loads per lane
      const double2 v = __ldcg(part + k);
      s += v.x;
      c += v.y;
    }
    s = mg_warp_sum(s);
    c = mg_warp_sum(c);
    if (lane == 0) {
      if (extra_sum != nullptr) { s += extra_sum[t]; c += extra_cnt[t]; }
      ws.utt_total[static_cast<int64_t>(t) * B + b] = make_double2(s, c);
    }
  }
  if (threadIdx.x == 0) ws.utt_ticket[b] = 0u;   // clean for the next launch
}

// Step 2: the B utterance finishers bump the global ticket; the last one writes the result records.
__device__ __forceinline__ void mg_finish_global(const MgFinishSlot* slots, int n_slots, const int64_t* seq_len, int B,
                                                 int64_t T, const MgWorkspace& ws, double* s_red, bool* s_flag) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  if (!mg_take_ticket(ws.ticket, static_cast<unsigned>(B), s_flag)) return;
  double* s_loss = s_red;   // [n_slots] weight * loss per slot, for the weighted total
  for (int t = warp; t < n_slots; t += n_warps) {
    const MgFinishSlot& sl = slots[t];
    const double2* tot = ws.utt_total + static_cast<int64_t>(t) * B;
    double sum_acc = 0., cnt_acc = 0., loss_acc = 0.;
    for (int b0 = lane; b0 < B; b0 += 32 * 8) {
      double2 v[8];
      int64_t nb[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {   // 8 independent loads in flight per lane
        const int bb = b0 + 32 * j;
        v[j] = bb < B ? __ldcg(tot + bb) : make_double2(0., 0.);
        nb[j] = bb < B ? mg_valid_frames(seq_len, bb, T) : 1;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (b0 + 32 * j >= B) break;
        sum_acc += v[j].x;
        loss_acc += v[j].x / static_cast<double>(nb[j]);   // reference losses.py:39 (0/0 -> nan for an empty utterance)
        if (sl.weighted) cnt_acc += v[j].y;
        else if (seq_len != nullptr) cnt_acc += static_cast<double>(nb[j]);                          // frames (metrics.py:393-394)
        else cnt_acc += static_cast<double>(T) * (sl.per_frame ? 1. : static_cast<double>(sl.D));   // numel (metrics.py:390)
      }
    }
    double s = mg_warp_sum(sum_acc), c = mg_warp_sum(cnt_acc), l = mg_warp_sum(loss_acc);
    if (lane == 0) {
      l /= static_cast<double>(B) * static_cast<double>(sl.D);   // torch.mean over (B, D), losses.py:42
      if (sl.accumulate) {   // running state of a streaming metric: self.sum += ..., self.count += ...
        const mg_term_result old = *sl.result;
        s += old.sum;
        c += old.count;
      }
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
      s_loss[t] = sl.in_total ? static_cast<double>(sl.weight) * l : 0.;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double weighted_total = 0.;
    for (int t = 0; t < n_slots; ++t) weighted_total += s_loss[t];
    slots[0].result->weighted_loss_f32 = static_cast<float>(weighted_total);
    *ws.ticket = 0u;   // leave the workspace clean for the next launch
  }
}

// Whole CTA (blockDim.x a multiple of 32, <= 1024).  `b`: the CTA's utterance; `ctas_per_utt`: CTAs that share it.
// s_red: 3 * 32 doubles, s_flag: one bool of shared memory.
__device__ __forceinline__ void mg_finish(const MgFinishSlot* slots, int n_slots, const int64_t* seq_len, int B, int64_t T,
                                          const MgWorkspace& ws, int b, unsigned ctas_per_utt, double* s_red, bool* s_flag) {
  if (!mg_take_ticket(ws.utt_ticket + b, ctas_per_utt, s_flag, true)) return;
  mg_fold_utterance(slots, n_slots, seq_len, B, T, ws, b, nullptr, nullptr);
  mg_finish_global(slots, n_slots, seq_len, B, T, ws, s_red, s_flag);
}
