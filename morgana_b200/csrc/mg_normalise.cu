// K3 -- standalone normalise / denormalise of a (rows, D) fp32 stream.
//
// Replaces data.normalise_mvn / denormalise_mvn / normalise_minmax / denormalise_minmax on torch tensors (reference
// morgana/data.py:533-538, 579-590): 2-5 elementwise ATen kernels with full-size temporaries (and, for minmax, a
// `scale` rebuild every call) become one pass.  HBM bytes: 8 * D * rows.
//
// The tensor is walked as a flat stream of 16-byte vectors regardless of D (rows of 187 or 3 floats are not 16-byte
// multiples); each thread tracks its (row, column) incrementally so there is no division in the loop.  The
// arithmetic keeps ATen's rounding sequence: normalise = IEEE sub then IEEE div; denormalise = mul then add (no FMA).
#include "mg_common.cuh"

namespace {

constexpr int kNormThreads = 256;

__device__ __forceinline__ float norm_scale(int mode, float p0, float p1) {
  if (mode == MG_NORM_MVN) return p1;     // std_dev
  float scale = __fsub_rn(p1, p0);        // data.py:580
  if (fabsf(scale) <= 1e-8f) scale = 1.0f;  // data.py:581
  return scale;
}

template <int MODE, bool INVERSE>
__device__ __forceinline__ float apply1(float x, float p0, float p1) {
  const float s = norm_scale(MODE, p0, p1);
  if (INVERSE) return __fadd_rn(__fmul_rn(x, s), p0);                       // data.py:538 / 590
  const float denom = (MODE == MG_NORM_MVN) ? __fadd_rn(s, 1e-8f) : s;      // data.py:534
  return __fdiv_rn(__fsub_rn(x, p0), denom);
}

// VEC = 4: x/out viewed as float4 (n_units = numel / 4, tail handled by the scalar instantiation); VEC = 1: scalar.
template <int VEC, int MODE, bool INVERSE>
__global__ void __launch_bounds__(kNormThreads)
normalise_kernel(const float* __restrict__ x, const float* __restrict__ p0, const float* __restrict__ p1,
                 float* __restrict__ out, int64_t first_elem, int64_t n_units, int D, int64_t rows_per_param,
                 int64_t step_rows, int step_cols) {
  mg_pdl_wait();                 // programmatic dependent launch (mg_common.cuh): nothing above touches global memory
  mg_pdl_launch_dependents();
  const int64_t tid = static_cast<int64_t>(blockIdx.x) * kNormThreads + threadIdx.x;
  const int64_t n_threads = static_cast<int64_t>(gridDim.x) * kNormThreads;
  // (row, col) of this thread's first element; afterwards advance by n_threads * VEC elements per iteration:
  // step_rows = (n_threads * VEC) / D and step_cols = (n_threads * VEC) % D are precomputed on the host.
  int64_t elem = first_elem + tid * VEC;
  int64_t row = elem / D;
  int col = static_cast<int>(elem - row * D);
  // Two units per iteration, both loads issued before any arithmetic: with one 16-byte load in flight per thread the
  // kernel held ~65 KB per SM in flight and ran at 0.79 of the copy peak.
  constexpr int kUnits = 2;
  for (int64_t u0 = tid; u0 < n_units; u0 += kUnits * n_threads) {
    float v[kUnits][VEC];
    bool live[kUnits];
#pragma unroll
    for (int j = 0; j < kUnits; ++j) {
      const int64_t u = u0 + j * n_threads;
      live[j] = u < n_units;
      if (!live[j]) continue;
      if constexpr (VEC == 4) {
        const float4 t = __ldcs(reinterpret_cast<const float4*>(x + first_elem) + u);
        v[j][0] = t.x; v[j][1] = t.y; v[j][2] = t.z; v[j][3] = t.w;
      } else {
        v[j][0] = __ldcs(x + first_elem + u);
      }
    }
#pragma unroll
    for (int j = 0; j < kUnits; ++j) {
      if (live[j]) {
        int64_t r = row;
        int c = col;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const int64_t prow = rows_per_param > 0 ? r / rows_per_param : 0;
          const float a = __ldg(p0 + prow * D + c), s = __ldg(p1 + prow * D + c);
          v[j][k] = apply1<MODE, INVERSE>(v[j][k], a, s);
          if (++c == D) { c = 0; ++r; }
        }
        const int64_t u = u0 + j * n_threads;
        if constexpr (VEC == 4) {
          __stcs(reinterpret_cast<float4*>(out + first_elem) + u, make_float4(v[j][0], v[j][1], v[j][2], v[j][3]));
        } else {
          __stcs(out + first_elem + u, v[j][0]);
        }
      }
      row += step_rows;      // (row, col) of the next unit of this thread
      col += step_cols;
      if (col >= D) { col -= D; ++row; }
    }
  }
}

template <int VEC>
int launch(const float* x, const float* p0, const float* p1, int mode, int inverse, float* out, int64_t first_elem,
           int64_t n_units, int D, int64_t rows_per_param, cudaStream_t stream) {
  if (n_units <= 0) return MG_OK;
  const int64_t max_blocks = static_cast<int64_t>(mg_cached_sm_count()) * 16;
  int64_t blocks = (n_units + kNormThreads - 1) / kNormThreads;
  // Each thread handles ~4 units when the tensor is large (enough loads in flight, few tail waves).
  if (blocks > max_blocks) blocks = max_blocks;
  const int64_t stride_elems = blocks * kNormThreads * VEC;
  const int64_t step_rows = stride_elems / D;
  const int step_cols = static_cast<int>(stride_elems % D);
  const unsigned grid = static_cast<unsigned>(blocks);
#define MG_NORM_LAUNCH(MODE, INV)                                                                             \
  MG_CUDA_OK(mg_launch_pdl(normalise_kernel<VEC, MODE, INV>, dim3(grid), dim3(kNormThreads), 0, stream, x, p0, p1, out, first_elem, \
                           n_units, D, rows_per_param, step_rows, step_cols))
  if (mode == MG_NORM_MVN) {
    if (inverse) MG_NORM_LAUNCH(MG_NORM_MVN, true); else MG_NORM_LAUNCH(MG_NORM_MVN, false);
  } else {
    if (inverse) MG_NORM_LAUNCH(MG_NORM_MINMAX, true); else MG_NORM_LAUNCH(MG_NORM_MINMAX, false);
  }
#undef MG_NORM_LAUNCH
  MG_LAUNCH_OK();
  return MG_OK;
}

}  // namespace

extern "C" int mg_normalise_f32(const float* x, const float* p0, const float* p1, int norm_mode, int inverse, float* out,
                                int64_t rows, int D, int64_t rows_per_param, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(rows >= 0 && D >= 0 && rows_per_param >= 0, "mg_normalise_f32: negative shape");
  MG_REQUIRE(norm_mode == MG_NORM_MVN || norm_mode == MG_NORM_MINMAX, "mg_normalise_f32: bad norm_mode %d", norm_mode);
  if (rows == 0 || D == 0) return MG_OK;
  MG_REQUIRE(x != nullptr && out != nullptr && p0 != nullptr && p1 != nullptr, "mg_normalise_f32: NULL buffer");
  const int64_t numel = rows * D;
  if (mg_aligned(x, 16) && mg_aligned(out, 16) && numel >= 4) {
    const int64_t n_vec = numel / 4;
    int rc = launch<4>(x, p0, p1, norm_mode, inverse, out, 0, n_vec, D, rows_per_param, stream);
    if (rc != MG_OK) return rc;
    return launch<1>(x, p0, p1, norm_mode, inverse, out, n_vec * 4, numel - n_vec * 4, D, rows_per_param, stream);
  }
  return launch<1>(x, p0, p1, norm_mode, inverse, out, 0, numel, D, rows_per_param, stream);
}


// ---- utils.both_voiced_mask (reference morgana/utils.py:169-172): out[i] = all_k (feature_k[i] != 0), one byte per element ----
namespace {

constexpr int kMaxVoicedFeatures = 8;
struct VoicedParams {
  const float* features[kMaxVoicedFeatures];
  int n_features;
  int64_t n;
  unsigned char* out;
};

__global__ void __launch_bounds__(256)
both_nonzero_kernel(const VoicedParams prm) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < prm.n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    bool voiced = true;
    for (int k = 0; k < prm.n_features; ++k) voiced = voiced && !(__ldg(prm.features[k] + i) == 0.f);   // ~torch.eq(x, 0.): NaN counts as voiced
    prm.out[i] = voiced ? 1 : 0;
  }
}

}  // namespace

extern "C" int mg_both_nonzero_u8(const float* const* features, int n_features, int64_t n, unsigned char* out, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(n_features >= 1 && n_features <= kMaxVoicedFeatures, "mg_both_nonzero_u8: %d features outside [1, %d]", n_features, kMaxVoicedFeatures);
  MG_REQUIRE(n >= 0, "mg_both_nonzero_u8: negative size");
  if (n == 0) return MG_OK;
  MG_REQUIRE(features != nullptr && out != nullptr, "mg_both_nonzero_u8: NULL buffer");
  VoicedParams prm;
  prm.n_features = n_features; prm.n = n; prm.out = out;
  for (int k = 0; k < kMaxVoicedFeatures; ++k) prm.features[k] = k < n_features ? features[k] : nullptr;
  for (int k = 0; k < n_features; ++k) MG_REQUIRE(features[k] != nullptr, "mg_both_nonzero_u8: feature %d is NULL", k);
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = static_cast<int64_t>(mg_cached_sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  both_nonzero_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(prm);
  MG_LAUNCH_OK();
  return MG_OK;
}
