// Shared tail of the masked reductions: integer ticket -> the last CTA combines the per-CTA slots in index order.
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

__device__ __forceinline__ int64_t mg_valid_frames(const int64_t* seq_len, int b, int64_t T) {
  if (seq_len == nullptr) return T;
  const int64_t n = __ldg(seq_len + b);
  return n < 0 ? 0 : (n > T ? T : n);   // mask = arange(T) < seq_len  (reference utils.py:140-142)
}

// Returns true in every thread of exactly one CTA: the last one to get here.  `s_flag` is a shared-memory bool.
__device__ __forceinline__ bool mg_take_ticket(unsigned int* ticket, bool* s_flag) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    *s_flag = atomicAdd(ticket, 1u) == total - 1;
  }
  __syncthreads();
  const bool last = *s_flag;
  if (last) __threadfence();
  return last;
}

// Run by the whole last CTA (blockDim.x threads, a multiple of 32, at most 1024).  s_red: 3 * 32 doubles of shared memory.
__device__ __forceinline__ void mg_finish(const MgFinishSlot* slots, int n_slots, const int64_t* seq_len, int B, int64_t T,
                                          const double2* partials, unsigned int* ticket, double* s_red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  double* s_a = s_red;
  double* s_b = s_red + 32;
  double* s_c = s_red + 64;
  double weighted_total = 0.;   // meaningful in thread 0 only
  for (int t = 0; t < n_slots; ++t) {
    const MgFinishSlot& sl = slots[t];
    const int64_t R = sl.rows_per_cta;
    double sum_acc = 0., cnt_acc = 0., loss_acc = 0.;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
      const int64_t n_b = mg_valid_frames(seq_len, b, T);
      const int64_t used = min(static_cast<int64_t>(sl.n_chunks), (n_b + R - 1) / R);
      const double2* slot = partials + (static_cast<int64_t>(t) * B + b) * kMgMaxChunks;
      double s = 0., c = 0.;
      for (int64_t k = 0; k < used; ++k) {
        const double2 v = __ldcg(slot + k);
        s += v.x;
        c += v.y;
      }
      sum_acc += s;
      loss_acc += s / static_cast<double>(n_b);   // reference losses.py:39 (0/0 -> nan for an empty utterance)
      if (sl.weighted) cnt_acc += c;
      else if (seq_len != nullptr) cnt_acc += static_cast<double>(n_b);                       // frames (metrics.py:393-394)
      else cnt_acc += static_cast<double>(T) * (sl.per_frame ? 1. : static_cast<double>(sl.D));  // numel (metrics.py:390)
    }
    sum_acc = mg_warp_sum(sum_acc);
    cnt_acc = mg_warp_sum(cnt_acc);
    loss_acc = mg_warp_sum(loss_acc);
    __syncthreads();
    if (lane == 0) { s_a[warp] = sum_acc; s_b[warp] = cnt_acc; s_c[warp] = loss_acc; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0., c = 0., l = 0.;
      for (int i = 0; i < n_warps; ++i) { s += s_a[i]; c += s_b[i]; l += s_c[i]; }
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
      if (sl.in_total) weighted_total += static_cast<double>(sl.weight) * l;
    }
  }
  if (threadIdx.x == 0) {
    slots[0].result->weighted_loss_f32 = static_cast<float>(weighted_total);
    *ticket = 0u;   // leave the workspace clean for the next launch
  }
}
