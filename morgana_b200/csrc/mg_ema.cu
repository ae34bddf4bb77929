// K6 -- multi-tensor exponential-moving-average update.
//
// Replaces utils.ExponentialMovingAverage.update_params / _update_param (reference morgana/utils.py:443-456): per
// trainable tensor `delta = shadow - x` then `shadow -= (1 - decay) * delta`, i.e. 3 ATen kernels and a temporary per
// tensor (38 tensors for LSTMAcousticModel -> launch-bound).  Here: one launch per 64 tensors, 12 bytes of HBM traffic
// per parameter (read shadow, read param, write shadow).
//
// Bit-exactness: the reference rounds three times -- fl(s - x), fl(c * that), fl(s - that) with c = fl32(1 - decay) --
// so the kernel uses explicit round-to-nearest intrinsics and never contracts into an FMA.
#include <string.h>

#include "mg_common.cuh"

namespace {

constexpr int kEmaThreads = 256;
constexpr int kEmaMaxTensors = 64;
constexpr int kEmaChunk = 8192;   // elements per CTA: 96 KB of traffic

struct EmaParams {
  float* shadow[kEmaMaxTensors];
  const float* param[kEmaMaxTensors];
  int64_t numel[kEmaMaxTensors];
  int chunk_begin[kEmaMaxTensors + 1];   // prefix sum of chunks per tensor
  int n_tensors;
  float one_minus_decay;
};

__device__ __forceinline__ float ema1(float s, float x, float c) {
  return __fsub_rn(s, __fmul_rn(c, __fsub_rn(s, x)));
}

__global__ void __launch_bounds__(kEmaThreads) ema_kernel(const __grid_constant__ EmaParams prm) {
  mg_pdl_wait();                 // programmatic dependent launch (mg_common.cuh): nothing above touches global memory
  mg_pdl_launch_dependents();
  // Which tensor does this CTA's chunk belong to?  (<= 64 entries in constant-bank parameter space.)
  int lo = 0, hi = prm.n_tensors;
  const int cta = blockIdx.x;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (prm.chunk_begin[mid] <= cta) lo = mid; else hi = mid;
  }
  const int t = lo;
  const int64_t begin = static_cast<int64_t>(cta - prm.chunk_begin[t]) * kEmaChunk;
  const int64_t n = min(static_cast<int64_t>(kEmaChunk), prm.numel[t] - begin);
  float* __restrict__ s = prm.shadow[t] + begin;
  const float* __restrict__ x = prm.param[t] + begin;
  const float c = prm.one_minus_decay;

  // Chunk starts are multiples of 8192 elements, so the chunk is 16-byte aligned iff the tensor is.
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(x)) & 15) == 0;
  if (vec_ok) {
    const int nvec = static_cast<int>(n >> 2);
    float4* s4 = reinterpret_cast<float4*>(s);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    int i = threadIdx.x;
    for (; i + 3 * kEmaThreads < nvec; i += 4 * kEmaThreads) {
      float4 sv[4], xv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sv[j] = s4[i + j * kEmaThreads];
        xv[j] = __ldcs(x4 + i + j * kEmaThreads);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sv[j].x = ema1(sv[j].x, xv[j].x, c);
        sv[j].y = ema1(sv[j].y, xv[j].y, c);
        sv[j].z = ema1(sv[j].z, xv[j].z, c);
        sv[j].w = ema1(sv[j].w, xv[j].w, c);
        s4[i + j * kEmaThreads] = sv[j];
      }
    }
    for (; i < nvec; i += kEmaThreads) {
      float4 sv = s4[i];
      const float4 xv = __ldcs(x4 + i);
      sv.x = ema1(sv.x, xv.x, c);
      sv.y = ema1(sv.y, xv.y, c);
      sv.z = ema1(sv.z, xv.z, c);
      sv.w = ema1(sv.w, xv.w, c);
      s4[i] = sv;
    }
    for (int k = (nvec << 2) + threadIdx.x; k < n; k += kEmaThreads) s[k] = ema1(s[k], x[k], c);
  } else {
    for (int k = threadIdx.x; k < n; k += kEmaThreads) s[k] = ema1(s[k], x[k], c);
  }
}

}  // namespace

extern "C" int mg_ema_update_f32(float* const* shadow, const float* const* param, const int64_t* numel, int n_tensors,
                                 float one_minus_decay, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(n_tensors >= 0, "mg_ema_update_f32: negative tensor count");
  if (n_tensors == 0) return MG_OK;
  MG_REQUIRE(shadow != nullptr && param != nullptr && numel != nullptr, "mg_ema_update_f32: NULL table");
  int next = 0;     // running index: empty tensors are skipped without using up a slot of the launch
  while (next < n_tensors) {
    EmaParams prm;
    memset(&prm, 0, sizeof(prm));
    int count = 0, chunks = 0;
    int i = next;
    for (; i < n_tensors && count < kEmaMaxTensors; ++i) {
      MG_REQUIRE(numel[i] >= 0, "mg_ema_update_f32: tensor %d has negative numel", i);
      if (numel[i] == 0) continue;
      MG_REQUIRE(shadow[i] != nullptr && param[i] != nullptr, "mg_ema_update_f32: tensor %d is NULL", i);
      MG_REQUIRE(shadow[i] != param[i], "mg_ema_update_f32: tensor %d: shadow aliases param (utils.py:452 asserts the models differ)", i);
      prm.shadow[count] = shadow[i];
      prm.param[count] = param[i];
      prm.numel[count] = numel[i];
      prm.chunk_begin[count] = chunks;
      const int64_t c = (numel[i] + kEmaChunk - 1) / kEmaChunk;
      MG_REQUIRE(chunks + c < (int64_t(1) << 30), "mg_ema_update_f32: too many elements in one call");
      chunks += static_cast<int>(c);
      ++count;
    }
    next = i;
    if (count == 0) continue;
    prm.chunk_begin[count] = chunks;
    prm.n_tensors = count;
    prm.one_minus_decay = one_minus_decay;
    MG_CUDA_OK(mg_launch_pdl(ema_kernel, dim3(static_cast<unsigned>(chunks)), dim3(kEmaThreads), 0, stream, prm));
    MG_LAUNCH_OK();
  }
  return MG_OK;
}
