// K1 -- per-utterance inclusive scan of durations (+ max / total / validity summary).
//
// Replaces `repeated_lens = sum(repeats, 1)`, `max(repeated_lens).item()` and the host-side np.repeat index build of
// utils.upsample_to_repetitions (reference morgana/utils.py:198-199, 214-222).  One warp per utterance; lanes take
// 32 consecutive items at a time, a shuffle scan plus a running carry gives the inclusive sums.  HBM traffic is
// 8*B*P bytes in + 4*B*P out -- negligible next to K2; the kernel exists to keep the index build on the device.
#include "mg_common.cuh"

namespace {

constexpr int kScanWarpsPerCta = 8;

template <typename DurT>
__global__ void __launch_bounds__(kScanWarpsPerCta * 32)
dur_scan_kernel(const DurT* __restrict__ dur, int64_t dur_stride_b, int B, int P, int32_t* __restrict__ ends,
                int64_t* __restrict__ n_frames, unsigned long long* __restrict__ summary,
                const int32_t* __restrict__ item_ends, int64_t total_items) {
  mg_pdl_wait();                  // nothing before this touches global memory
  mg_pdl_launch_dependents();     // one short wave: the next kernel's CTAs may take the free SMs right away (they wait for this grid)
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * kScanWarpsPerCta + (threadIdx.x >> 5);
  if (b >= B) return;
  // Padded layout: utterance b's items are row b of a (B, P) array.  Packed (ragged) layout: its items follow those of
  // utterance b - 1 in one flat array; item_ends is the inclusive scan of the item counts.
  int64_t base = static_cast<int64_t>(b) * P;
  if (item_ends) {   // counts that overrun the arrays are clamped (the host checks the total when it reads the summary back)
    base = min(max(b > 0 ? static_cast<int64_t>(__ldg(item_ends + b - 1)) : 0, static_cast<int64_t>(0)), total_items);
    const int64_t stop = min(max(static_cast<int64_t>(__ldg(item_ends + b)), base), total_items);
    P = static_cast<int>(stop - base);
  }
  const DurT* row = item_ends ? dur + base : dur + static_cast<int64_t>(b) * dur_stride_b;
  int32_t* ends_row = ends + base;

  long long carry = 0;
  int negatives = 0;
  for (int p0 = 0; p0 < P; p0 += 32) {
    const int p = p0 + lane;
    long long d = (p < P) ? static_cast<long long>(row[p]) : 0;
    if (d < 0) {  // counted and treated as 0 so the scan stays monotone; the host raises ValueError
      negatives += 1;
      d = 0;
    }
    long long s = d;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long up = __shfl_up_sync(MG_FULL_MASK, s, o);
      if (lane >= o) s += up;
    }
    s += carry;
    if (p < P) ends_row[p] = static_cast<int32_t>(s > 2147483647LL ? 2147483647LL : s);
    carry = __shfl_sync(MG_FULL_MASK, s, 31);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) negatives += __shfl_xor_sync(MG_FULL_MASK, negatives, o);

  if (lane == 0) {
    n_frames[b] = carry;
    if (summary == nullptr) return;   // the caller knows the padded length and trusts the durations: nothing to read back
    // Integer atomics: exact and order-independent.
    atomicMax(summary + 0, static_cast<unsigned long long>(carry));
    if (negatives) atomicAdd(summary + 1, static_cast<unsigned long long>(negatives));
    atomicAdd(summary + 2, static_cast<unsigned long long>(carry));
    if (carry > 2147483647LL) atomicAdd(summary + 3, 1ull);
  }
}

}  // namespace

extern "C" int mg_dur_scan(const void* dur, int dur_is_i32, int64_t dur_stride_b, int B, int P, int32_t* ends,
                           int64_t* n_frames, int64_t* summary, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(B >= 0 && P >= 0, "mg_dur_scan: negative shape (B=%d, P=%d)", B, P);
  if (summary != nullptr) MG_CUDA_OK(cudaMemsetAsync(summary, 0, 4 * sizeof(int64_t), stream));
  if (B == 0) return MG_OK;
  MG_REQUIRE(n_frames != nullptr && (P == 0 || (dur != nullptr && ends != nullptr)), "mg_dur_scan: NULL buffer");
  const int grid = (B + kScanWarpsPerCta - 1) / kScanWarpsPerCta;
  auto* summary_u = reinterpret_cast<unsigned long long*>(summary);
  if (dur_is_i32) {
    MG_CUDA_OK(mg_launch_pdl(dur_scan_kernel<int32_t>, dim3(grid), dim3(kScanWarpsPerCta * 32), 0, stream, static_cast<const int32_t*>(dur),
                             dur_stride_b, B, P, ends, n_frames, summary_u, static_cast<const int32_t*>(nullptr), static_cast<int64_t>(0)));
  } else {
    MG_CUDA_OK(mg_launch_pdl(dur_scan_kernel<long long>, dim3(grid), dim3(kScanWarpsPerCta * 32), 0, stream, static_cast<const long long*>(dur),
                             dur_stride_b, B, P, ends, n_frames, summary_u, static_cast<const int32_t*>(nullptr), static_cast<int64_t>(0)));
  }
  MG_LAUNCH_OK();
  return MG_OK;
}

extern "C" int mg_dur_scan_packed(const void* dur, int dur_is_i32, const int32_t* item_ends, int B, int64_t total_items,
                                  int32_t* ends, int64_t* n_frames, int64_t* summary, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(B >= 0 && total_items >= 0, "mg_dur_scan_packed: negative size");
  if (summary != nullptr) MG_CUDA_OK(cudaMemsetAsync(summary, 0, 4 * sizeof(int64_t), stream));
  if (B == 0) return MG_OK;
  MG_REQUIRE(n_frames != nullptr && item_ends != nullptr && (total_items == 0 || (dur != nullptr && ends != nullptr)),
             "mg_dur_scan_packed: NULL buffer");
  const int grid = (B + kScanWarpsPerCta - 1) / kScanWarpsPerCta;
  auto* summary_u = reinterpret_cast<unsigned long long*>(summary);
  if (dur_is_i32) {
    MG_CUDA_OK(mg_launch_pdl(dur_scan_kernel<int32_t>, dim3(grid), dim3(kScanWarpsPerCta * 32), 0, stream, static_cast<const int32_t*>(dur),
                             static_cast<int64_t>(0), B, 0, ends, n_frames, summary_u, item_ends, total_items));
  } else {
    MG_CUDA_OK(mg_launch_pdl(dur_scan_kernel<long long>, dim3(grid), dim3(kScanWarpsPerCta * 32), 0, stream, static_cast<const long long*>(dur),
                             static_cast<int64_t>(0), B, 0, ends, n_frames, summary_u, item_ends, total_items));
  }
  MG_LAUNCH_OK();
  return MG_OK;
}
