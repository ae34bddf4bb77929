// On-device collate: ragged (packed) rows -> zero-padded (B, T, row) batch.
//
// Replaces the zero-padding loop of FilesDataset.collate_fn (reference morgana/data.py:159-224, in particular :184-193:
// `torch.zeros(batch, max_len, dim)` then one Python-level copy per utterance per feature) followed by the host->device
// copy of the PADDED tensor (ToDeviceWrapper, data.py:655-663).  With this kernel the host ships only the valid rows
// (33 % fewer bytes over PCIe at config 2) and the padding is produced where it is consumed.
// HBM bytes: 2 * valid bytes + padding bytes.
#include "mg_common.cuh"

namespace {

constexpr int kCollateThreads = 256;

template <typename Vec>
__device__ __forceinline__ void copy_span(const unsigned char* src, unsigned char* dst, int64_t bytes) {
  const int64_t n = bytes / static_cast<int64_t>(sizeof(Vec));
  const Vec* s = reinterpret_cast<const Vec*>(src);
  Vec* d = reinterpret_cast<Vec*>(dst);
  int64_t i = threadIdx.x;
  for (; i + 3 * kCollateThreads < n; i += 4 * kCollateThreads) {
    Vec v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = __ldcs(s + i + j * kCollateThreads);
#pragma unroll
    for (int j = 0; j < 4; ++j) d[i + j * kCollateThreads] = v[j];
  }
  for (; i < n; i += kCollateThreads) d[i] = __ldcs(s + i);
  for (int64_t k = n * static_cast<int64_t>(sizeof(Vec)) + threadIdx.x; k < bytes; k += kCollateThreads) dst[k] = src[k];
}

template <typename Vec>
__device__ __forceinline__ void zero_span(unsigned char* dst, int64_t bytes) {
  const int64_t n = bytes / static_cast<int64_t>(sizeof(Vec));
  Vec* d = reinterpret_cast<Vec*>(dst);
  Vec z;
  memset(&z, 0, sizeof(Vec));
  for (int64_t i = threadIdx.x; i < n; i += kCollateThreads) d[i] = z;
  for (int64_t k = n * static_cast<int64_t>(sizeof(Vec)) + threadIdx.x; k < bytes; k += kCollateThreads) dst[k] = 0;
}

// ends: inclusive scan of the lengths (int32, from mg_dur_scan on a (1, B) view).  vec: 16 / 4 / 1 bytes.
__global__ void __launch_bounds__(kCollateThreads)
pad_collate_kernel(const unsigned char* __restrict__ packed, const int32_t* __restrict__ ends, unsigned char* __restrict__ out,
                   int64_t row_bytes, int64_t T, int64_t total_rows, int rows_per_cta, int vec) {
  const int b = blockIdx.y;
  // lengths that sum past the packed rows (unchecked on the no-sync path) are truncated, never read out of bounds
  const int64_t begin = min(b > 0 ? static_cast<int64_t>(__ldg(ends + b - 1)) : 0, total_rows);
  const int64_t n_b = min(min(static_cast<int64_t>(__ldg(ends + b)), total_rows) - begin, T);
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_cta;
  const int64_t r1 = min(r0 + rows_per_cta, T);
  const int64_t valid_end = min(r1, n_b);
  unsigned char* dst = out + (static_cast<int64_t>(b) * T + r0) * row_bytes;
  if (valid_end > r0) {
    const unsigned char* src = packed + (begin + r0) * row_bytes;
    const int64_t bytes = (valid_end - r0) * row_bytes;
    // vec was chosen on the host from the row size and base alignment; spans start on row boundaries.
    if (vec == 16) copy_span<uint4>(src, dst, bytes);
    else if (vec == 4) copy_span<uint32_t>(src, dst, bytes);
    else copy_span<unsigned char>(src, dst, bytes);
  }
  const int64_t z0 = max(r0, n_b);
  if (r1 > z0) {
    unsigned char* z = out + (static_cast<int64_t>(b) * T + z0) * row_bytes;
    const int64_t bytes = (r1 - z0) * row_bytes;
    if (vec == 16) zero_span<uint4>(z, bytes);
    else if (vec == 4) zero_span<uint32_t>(z, bytes);
    else zero_span<unsigned char>(z, bytes);
  }
}

}  // namespace

extern "C" int mg_pad_collate(const void* packed, const int32_t* ends, void* out, int B, int64_t row_bytes, int64_t T,
                              int64_t total_rows, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(B >= 0 && row_bytes >= 0 && T >= 0 && total_rows >= 0, "mg_pad_collate: negative shape");
  MG_REQUIRE(B <= 65535, "mg_pad_collate: B=%d exceeds 65535 utterances per call", B);
  if (B == 0 || T == 0 || row_bytes == 0) return MG_OK;
  MG_REQUIRE(ends != nullptr && out != nullptr && packed != nullptr, "mg_pad_collate: NULL buffer");
  int vec = 1;
  if (row_bytes % 16 == 0 && mg_aligned(packed, 16) && mg_aligned(out, 16)) vec = 16;
  else if (row_bytes % 4 == 0 && mg_aligned(packed, 4) && mg_aligned(out, 4)) vec = 4;
  int64_t rows = (256 * 1024 + row_bytes - 1) / row_bytes;   // ~256 KB of output per CTA
  if (rows < 1) rows = 1;
  const int64_t sms = mg_cached_sm_count();
  while (rows > 8 && static_cast<int64_t>(B) * ((T + rows - 1) / rows) < 8 * sms) rows = (rows + 1) / 2;
  if (rows > T) rows = T;
  dim3 grid(static_cast<unsigned>((T + rows - 1) / rows), static_cast<unsigned>(B));
  pad_collate_kernel<<<grid, kCollateThreads, 0, stream>>>(static_cast<const unsigned char*>(packed), ends,
                                                           static_cast<unsigned char*>(out), row_bytes, T,
                                                           total_rows, static_cast<int>(rows), vec);
  MG_LAUNCH_OK();
  return MG_OK;
}
