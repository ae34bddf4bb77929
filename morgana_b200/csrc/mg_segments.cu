// Sibling segment operations of the expansion kernel ("next" row 4 of the scope table): the same duration / length scan
// (K1) drives three more row movers, each replacing a host-side index build + ATen advanced-index gather of the reference:
//
//   mg_pack_rows          utils.batched_masked_select   (morgana/utils.py:147-166)  (B, T, D) + lengths -> (sum len, D)
//   mg_segment_ends       utils.get_segment_ends        (morgana/utils.py:287-330)  row at the last frame of each segment
//   mg_split_to_segments  utils.split_to_segments       (morgana/utils.py:231-284)  (B, T, D) -> (B, S, L_max, D), zero padded
//
// All three are pure byte movers (dtype-agnostic), one warp per output row, 16-byte vectors when the row allows.
#include <string.h>

#include "mg_common.cuh"

namespace {

constexpr int kSegWarps = 8;

__device__ __forceinline__ void copy_row(const unsigned char* src, unsigned char* dst, int64_t row_bytes, int vec, int lane) {
  if (vec == 16) {
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int64_t i = lane; i < row_bytes / 16; i += 32) d[i] = __ldg(s + i);
  } else if (vec == 4) {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
    for (int64_t i = lane; i < row_bytes / 4; i += 32) d[i] = __ldg(s + i);
  } else {
    for (int64_t i = lane; i < row_bytes; i += 32) dst[i] = src[i];
  }
}
__device__ __forceinline__ void zero_row(unsigned char* dst, int64_t row_bytes, int vec, int lane) {
  if (vec == 16) {
    uint4* d = reinterpret_cast<uint4*>(dst);
    for (int64_t i = lane; i < row_bytes / 16; i += 32) d[i] = make_uint4(0, 0, 0, 0);
  } else if (vec == 4) {
    uint32_t* d = reinterpret_cast<uint32_t*>(dst);
    for (int64_t i = lane; i < row_bytes / 4; i += 32) d[i] = 0u;
  } else {
    for (int64_t i = lane; i < row_bytes; i += 32) dst[i] = 0;
  }
}

// (B, T, row) -> packed rows: warp per (b, t); ends = inclusive scan of min(len_b, T) over utterances
__global__ void __launch_bounds__(kSegWarps * 32)
pack_rows_kernel(const unsigned char* __restrict__ x, int64_t x_sb, int64_t x_st, const int32_t* __restrict__ ends,
                 unsigned char* __restrict__ out, int B, int64_t T, int64_t row_bytes, int vec) {
  const int lane = threadIdx.x & 31;
  const int64_t w = static_cast<int64_t>(blockIdx.x) * kSegWarps + (threadIdx.x >> 5);
  if (w >= static_cast<int64_t>(B) * T) return;
  const int b = static_cast<int>(w / T);
  const int64_t t = w - static_cast<int64_t>(b) * T;
  const int64_t begin = b > 0 ? static_cast<int64_t>(__ldg(ends + b - 1)) : 0;
  const int64_t n_b = static_cast<int64_t>(__ldg(ends + b)) - begin;
  if (t >= n_b) return;
  copy_row(x + b * x_sb + t * x_st, out + (begin + t) * row_bytes, row_bytes, vec, lane);
}

// out[b, s] = x[b, ends[b, s] - 1] when segment s is non-empty, else 0
__global__ void __launch_bounds__(kSegWarps * 32)
segment_ends_kernel(const unsigned char* __restrict__ x, int64_t x_sb, int64_t x_st, const int32_t* __restrict__ seg_ends,
                    unsigned char* __restrict__ out, int B, int S, int64_t T, int64_t row_bytes, int vec) {
  const int lane = threadIdx.x & 31;
  const int64_t w = static_cast<int64_t>(blockIdx.x) * kSegWarps + (threadIdx.x >> 5);
  if (w >= static_cast<int64_t>(B) * S) return;
  const int b = static_cast<int>(w / S), s = static_cast<int>(w - static_cast<int64_t>(b) * S);
  const int64_t end = __ldg(seg_ends + w);
  const int64_t begin = s > 0 ? static_cast<int64_t>(__ldg(seg_ends + w - 1)) : 0;
  unsigned char* dst = out + w * row_bytes;
  if (end > begin && end <= T) copy_row(x + b * x_sb + (end - 1) * x_st, dst, row_bytes, vec, lane);
  else zero_row(dst, row_bytes, vec, lane);
}

// out[b, s, j] = x[b, begin_s + j] for j < len_s, else 0; warp per (b, s, j)
__global__ void __launch_bounds__(kSegWarps * 32)
split_segments_kernel(const unsigned char* __restrict__ x, int64_t x_sb, int64_t x_st, const int32_t* __restrict__ seg_ends,
                      unsigned char* __restrict__ out, int B, int S, int64_t L, int64_t T, int64_t row_bytes, int vec) {
  const int lane = threadIdx.x & 31;
  const int64_t w = static_cast<int64_t>(blockIdx.x) * kSegWarps + (threadIdx.x >> 5);
  if (w >= static_cast<int64_t>(B) * S * L) return;
  const int64_t bs = w / L, j = w - bs * L;
  const int b = static_cast<int>(bs / S), s = static_cast<int>(bs - static_cast<int64_t>(b) * S);
  const int64_t end = __ldg(seg_ends + bs);
  const int64_t begin = s > 0 ? static_cast<int64_t>(__ldg(seg_ends + bs - 1)) : 0;
  unsigned char* dst = out + w * row_bytes;
  if (begin + j < end && begin + j < T) copy_row(x + b * x_sb + (begin + j) * x_st, dst, row_bytes, vec, lane);
  else zero_row(dst, row_bytes, vec, lane);
}

// Backward of the two segment gathers: every frame row (b, t) belongs to at most one output row, so the gradient is a gather
// too -- no atomics, no accumulation order.  Warp per (b, t): binary search of the segment whose interval holds t, then
//   ENDS  (L == 0): grad_x[b, t] = grad_out[b, s]     if t is the segment's last frame, else 0
//   SPLIT (L  > 0): grad_x[b, t] = grad_out[b, s, j]  with j = t - begin_s, if j < L, else 0 (rows cut off by a short L)
// Rows past the last segment get zeros (the reference's padder row swallows nothing from them either).
__global__ void __launch_bounds__(kSegWarps * 32)
segments_bwd_kernel(const unsigned char* __restrict__ grad_out, const int32_t* __restrict__ seg_ends,
                    unsigned char* __restrict__ grad_x, int B, int S, int64_t L, int64_t T, int64_t row_bytes, int vec) {
  const int lane = threadIdx.x & 31;
  const int64_t w = static_cast<int64_t>(blockIdx.x) * kSegWarps + (threadIdx.x >> 5);
  if (w >= static_cast<int64_t>(B) * T) return;
  const int b = static_cast<int>(w / T);
  const int64_t t = w - static_cast<int64_t>(b) * T;
  const int32_t* e = seg_ends + static_cast<int64_t>(b) * S;
  int lo = 0, hi = S;                      // first s with e[s] > t
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (static_cast<int64_t>(__ldg(e + mid)) > t) hi = mid; else lo = mid + 1;
  }
  unsigned char* dst = grad_x + w * row_bytes;
  if (lo < S) {
    const int64_t end = __ldg(e + lo);
    const int64_t begin = lo > 0 ? static_cast<int64_t>(__ldg(e + lo - 1)) : 0;
    const int64_t bs = static_cast<int64_t>(b) * S + lo;
    if (L == 0) {
      if (t == end - 1) { copy_row(grad_out + bs * row_bytes, dst, row_bytes, vec, lane); return; }
    } else if (t - begin < L) {
      copy_row(grad_out + (bs * L + (t - begin)) * row_bytes, dst, row_bytes, vec, lane);
      return;
    }
  }
  zero_row(dst, row_bytes, vec, lane);
}

int pick_vec(const void* x, int64_t x_sb, int64_t x_st, const void* out, int64_t row_bytes) {
  auto ok = [&](int64_t a) { return row_bytes % a == 0 && mg_aligned(x, a) && mg_aligned(out, a) && x_sb % a == 0 && x_st % a == 0; };
  return ok(16) ? 16 : (ok(4) ? 4 : 1);
}

}  // namespace

extern "C" int mg_pack_rows(const void* x, int64_t x_stride_b_bytes, int64_t x_stride_t_bytes, const int32_t* ends, void* out,
                            int B, int64_t T, int64_t row_bytes, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(B >= 0 && T >= 0 && row_bytes >= 0, "mg_pack_rows: negative shape");
  if (B == 0 || T == 0 || row_bytes == 0) return MG_OK;
  MG_REQUIRE(x != nullptr && ends != nullptr && out != nullptr, "mg_pack_rows: NULL buffer");
  const int64_t warps = static_cast<int64_t>(B) * T;
  MG_REQUIRE(warps / kSegWarps < (int64_t(1) << 31), "mg_pack_rows: too many rows");
  pack_rows_kernel<<<static_cast<unsigned>((warps + kSegWarps - 1) / kSegWarps), kSegWarps * 32, 0, stream>>>(
      static_cast<const unsigned char*>(x), x_stride_b_bytes, x_stride_t_bytes, ends, static_cast<unsigned char*>(out), B, T, row_bytes,
      pick_vec(x, x_stride_b_bytes, x_stride_t_bytes, out, row_bytes));
  MG_LAUNCH_OK();
  return MG_OK;
}

extern "C" int mg_segment_ends(const void* x, int64_t x_stride_b_bytes, int64_t x_stride_t_bytes, const int32_t* seg_ends,
                               void* out, int B, int S, int64_t T, int64_t row_bytes, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(B >= 0 && S >= 0 && T >= 0 && row_bytes >= 0, "mg_segment_ends: negative shape");
  if (B == 0 || S == 0 || row_bytes == 0) return MG_OK;
  MG_REQUIRE(seg_ends != nullptr && out != nullptr && (T == 0 || x != nullptr), "mg_segment_ends: NULL buffer");
  const int64_t warps = static_cast<int64_t>(B) * S;
  segment_ends_kernel<<<static_cast<unsigned>((warps + kSegWarps - 1) / kSegWarps), kSegWarps * 32, 0, stream>>>(
      static_cast<const unsigned char*>(x), x_stride_b_bytes, x_stride_t_bytes, seg_ends, static_cast<unsigned char*>(out), B, S, T,
      row_bytes, pick_vec(x, x_stride_b_bytes, x_stride_t_bytes, out, row_bytes));
  MG_LAUNCH_OK();
  return MG_OK;
}

extern "C" int mg_split_to_segments(const void* x, int64_t x_stride_b_bytes, int64_t x_stride_t_bytes, const int32_t* seg_ends,
                                    void* out, int B, int S, int64_t L, int64_t T, int64_t row_bytes, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(B >= 0 && S >= 0 && L >= 0 && T >= 0 && row_bytes >= 0, "mg_split_to_segments: negative shape");
  if (B == 0 || S == 0 || L == 0 || row_bytes == 0) return MG_OK;
  MG_REQUIRE(seg_ends != nullptr && out != nullptr && (T == 0 || x != nullptr), "mg_split_to_segments: NULL buffer");
  const int64_t warps = static_cast<int64_t>(B) * S * L;
  MG_REQUIRE(warps / kSegWarps < (int64_t(1) << 31), "mg_split_to_segments: too many rows");
  split_segments_kernel<<<static_cast<unsigned>((warps + kSegWarps - 1) / kSegWarps), kSegWarps * 32, 0, stream>>>(
      static_cast<const unsigned char*>(x), x_stride_b_bytes, x_stride_t_bytes, seg_ends, static_cast<unsigned char*>(out), B, S, L, T,
      row_bytes, pick_vec(x, x_stride_b_bytes, x_stride_t_bytes, out, row_bytes));
  MG_LAUNCH_OK();
  return MG_OK;
}

extern "C" int mg_segments_bwd(const void* grad_out, const int32_t* seg_ends, void* grad_x, int B, int S, int64_t L, int64_t T,
                               int64_t row_bytes, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(B >= 0 && S >= 0 && L >= 0 && T >= 0 && row_bytes >= 0, "mg_segments_bwd: negative shape");
  if (B == 0 || T == 0 || row_bytes == 0) return MG_OK;
  MG_REQUIRE(grad_x != nullptr && (S == 0 || (seg_ends != nullptr && grad_out != nullptr)), "mg_segments_bwd: NULL buffer");
  const int64_t warps = static_cast<int64_t>(B) * T;
  MG_REQUIRE(warps / kSegWarps < (int64_t(1) << 31), "mg_segments_bwd: too many rows");
  segments_bwd_kernel<<<static_cast<unsigned>((warps + kSegWarps - 1) / kSegWarps), kSegWarps * 32, 0, stream>>>(
      static_cast<const unsigned char*>(grad_out), seg_ends, static_cast<unsigned char*>(grad_x), B, S, L, T, row_bytes,
      pick_vec(grad_out, 16, 16, grad_x, row_bytes));
  MG_LAUNCH_OK();
  return MG_OK;
}
