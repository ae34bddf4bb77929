// K7 -- dense layer on the tcgen05 tensor cores (placeholder: the kernel lands in a later milestone).
#include "mg_common.cuh"

#include <cuda_bf16.h>

namespace {

constexpr int kCastThreads = 256;

// fp32 (rows, K) with row stride ldx -> bf16 (rows, ld_out), columns >= K zero-filled.  8 outputs (16 bytes) per thread.
__global__ void __launch_bounds__(kCastThreads)
cast_pad_kernel(const float* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out, int64_t ld_out, int64_t rows,
                int K, int groups_per_row) {
  const int64_t total = rows * groups_per_row;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * kCastThreads + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * kCastThreads) {
    const int64_t r = idx / groups_per_row;
    const int c0 = static_cast<int>(idx - r * groups_per_row) * 8;
    const float* src = x + r * ldx + c0;
    __align__(16) __nv_bfloat16 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = __float2bfloat16_rn(c0 + k < K ? __ldg(src + k) : 0.f);
    *reinterpret_cast<uint4*>(out + r * ld_out + c0) = *reinterpret_cast<const uint4*>(v);
  }
}

}  // namespace

extern "C" int mg_cast_pad_bf16(const float* x, int64_t ldx, void* out, int64_t ld_out, int64_t rows, int K,
                                mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(rows >= 0 && K >= 0 && ld_out >= K && ldx >= K, "mg_cast_pad_bf16: bad shape");
  MG_REQUIRE(ld_out % 8 == 0 && mg_aligned(out, 16), "mg_cast_pad_bf16: output rows must be 16-byte aligned");
  if (rows == 0 || ld_out == 0) return MG_OK;
  MG_REQUIRE(x != nullptr && out != nullptr, "mg_cast_pad_bf16: NULL buffer");
  const int groups = static_cast<int>(ld_out / 8);
  const int64_t total = rows * groups;
  int64_t blocks = (total + kCastThreads - 1) / kCastThreads;
  const int64_t cap = static_cast<int64_t>(mg_cached_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  cast_pad_kernel<<<static_cast<unsigned>(blocks), kCastThreads, 0, stream>>>(
      x, ldx, static_cast<__nv_bfloat16*>(out), ld_out, rows, K, groups);
  MG_LAUNCH_OK();
  return MG_OK;
}

extern "C" int mg_linear_bf16(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, void* y,
                              int64_t ldy, int y_is_bf16, int M, int N, int K, int act, mg_stream_t stream_) {
  (void)x; (void)ldx; (void)w; (void)ldw; (void)bias; (void)y; (void)ldy; (void)y_is_bf16; (void)M; (void)N; (void)K;
  (void)act; (void)stream_;
  mg_set_error("mg_linear_bf16: not built yet");
  return MG_ERR_UNSUPPORTED;
}
