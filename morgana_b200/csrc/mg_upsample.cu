// K2 -- fused normalise + duration-driven item->frame expansion, and its backward.
//
// Replaces utils.upsample_to_repetitions (reference morgana/utils.py:175-228): zeros+cat copy of the input, CPU-side
// np.repeat index build, two (B, T) int64 index uploads and an ATen advanced-index gather -- and, when norm_mode != 0,
// the data.normalise_mvn / normalise_minmax pass that precedes it (morgana/data.py:533-534, 579-583).
//
// Geometry.  The output (B, T, D) is dense and write-dominated: every item row (D*4 bytes) is read once and written
// `dur` (~15) times, and ~1/3 of the rows are zero padding.  Work is therefore split by OUTPUT rows: CTA (c, b) owns
// rows [c*R, (c+1)*R) of utterance b, which are one contiguous R*D*4-byte span of `out`.  Each warp takes items that
// overlap the span, normalises the item row ONCE, and replicates it:
//
//   BULK path   (D % 4 == 0): the normalised row is staged in shared memory and one elected lane issues `dur`
//               cp.async.bulk shared->global copies of the whole row (TMA engine, SASS UBLKCP); padding rows are bulk
//               copies from a zero tile.  No per-element store instructions; smem slots are recycled with
//               cp.async.bulk.wait_group.read.
//   DIRECT path (any D / dtype): the row lives in registers and is written with vector stores, row-major.
//
// Algorithmic HBM bytes per launch: 4*D*(B*T + sum_b n_items_b) + 4*B*P + 8*D  (SURVEY.md section 8d).
#include <stdlib.h>
#include <string.h>

#include <cuda_bf16.h>

#include "mg_common.cuh"

namespace {

constexpr int kUpWarps = 4;        // warps per CTA
constexpr int kUpThreads = kUpWarps * 32;
constexpr int kUpSlots = 2;        // shared-memory row buffers per warp (bulk path)
constexpr int kBatch = 8;          // vectors per lane held in registers at once
constexpr int kMaxEndsSmem = 8192; // items whose scan row is staged in shared memory (32 KB)

__device__ __forceinline__ float mg_denominator(int mode, float p0, float p1) {
  if (mode == MG_NORM_MVN) return __fadd_rn(p1, 1e-8f);  // data.py:534  std_dev + 1e-8 (fp32 add)
  float scale = __fsub_rn(p1, p0);                       // data.py:580  mmax - mmin
  if (fabsf(scale) <= 1e-8f) scale = 1.0f;               // data.py:581
  return scale;
}

template <int MODE>
__device__ __forceinline__ float mg_norm1(float x, float p0, float p1) {
  if (MODE == MG_NORM_NONE) return x;
  // (x - p0) / denom with an IEEE subtract and an IEEE divide: bit-identical to ATen's sub then div kernels.
  return __fdiv_rn(__fsub_rn(x, p0), mg_denominator(MODE, p0, p1));
}

template <typename Vec>
struct VecTraits;
template <>
struct VecTraits<float4> {
  static constexpr int kFloats = 4;
};
template <>
struct VecTraits<float> {
  static constexpr int kFloats = 1;
};

template <int MODE>
__device__ __forceinline__ float4 mg_norm_vec(float4 v, const float* p0, const float* p1, int i) {
  if (MODE == MG_NORM_NONE) return v;
  const float4 a = __ldg(reinterpret_cast<const float4*>(p0) + i);
  const float4 s = __ldg(reinterpret_cast<const float4*>(p1) + i);
  v.x = mg_norm1<MODE>(v.x, a.x, s.x);
  v.y = mg_norm1<MODE>(v.y, a.y, s.y);
  v.z = mg_norm1<MODE>(v.z, a.z, s.z);
  v.w = mg_norm1<MODE>(v.w, a.w, s.w);
  return v;
}
template <int MODE>
__device__ __forceinline__ float mg_norm_vec(float v, const float* p0, const float* p1, int i) {
  if (MODE == MG_NORM_NONE) return v;
  return mg_norm1<MODE>(v, __ldg(p0 + i), __ldg(p1 + i));
}
// Raw (dtype-agnostic) vectors are never normalised.
template <int MODE, typename Vec>
__device__ __forceinline__ Vec mg_norm_vec(Vec v, const float*, const float*, int) {
  return v;
}

// Stage this utterance's inclusive scan row in shared memory (or fall back to global for very long rows) and return
// the pointer to search.  Must be called by the whole CTA.
__device__ __forceinline__ const int32_t* mg_stage_ends(const int32_t* ends_row, int P, int32_t* smem_ends, bool fits) {
  if (fits)
    for (int p = threadIdx.x; p < P; p += blockDim.x) smem_ends[p] = __ldg(ends_row + p);
  __syncthreads();   // unconditional: callers also rely on it to order their own shared-memory writes
  return fits ? smem_ends : ends_row;
}

// First item whose interval ends after frame t, i.e. the item that holds t (items with zero duration are skipped).
__device__ __forceinline__ int mg_item_of_frame(const int32_t* e, int P, int64_t t) {
  int lo = 0, hi = P;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (static_cast<int64_t>(e[mid]) <= t) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ------------------------------------------------------------------------------------------------------------------
// BULK path
// ------------------------------------------------------------------------------------------------------------------
// Dynamic shared memory: [kUpWarps * kUpSlots rows][zero tile of zero_rows rows][ends row (int32 x P, if it fits)]
// OUT_BF16: the staged row is rounded to bf16 (round-to-nearest-even of the exact fp32 result) before it is replicated, so
// the frame-rate tensor that feeds the tensor-core layers is written once at half the bytes (input rows stay fp32).
template <int MODE, bool OUT_BF16>
__global__ void __launch_bounds__(kUpThreads)
upsample_bulk_kernel(const unsigned char* __restrict__ x, int64_t x_sb, int64_t x_sp, const int32_t* __restrict__ ends,
                     const float* __restrict__ p0, const float* __restrict__ p1, int64_t p_sb,
                     unsigned char* __restrict__ out, int P, int nvec /* input row bytes / 16 */, int64_t T,
                     int rows_per_cta, int zero_rows, const int32_t* __restrict__ item_ends, int64_t total_items) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t row_bytes = static_cast<uint32_t>(nvec) * (OUT_BF16 ? 8u : 16u);   // OUTPUT row
  unsigned char* slots = smem;
  unsigned char* zero_tile = smem + static_cast<size_t>(kUpWarps * kUpSlots) * row_bytes;
  int32_t* smem_ends = reinterpret_cast<int32_t*>(zero_tile + static_cast<size_t>(zero_rows) * row_bytes);

  mg_pdl_wait();   // programmatic dependent launch: this grid may be resident before the duration scan has finished
  mg_pdl_launch_dependents();   // counts once the LAST wave has started: the next grid then takes the slots this one frees
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * rows_per_cta;
  const int64_t t1 = min(t0 + rows_per_cta, T);
  const bool ends_fit = P <= kMaxEndsSmem;   // the host sized the shared-memory scan row from this (largest) item count
  // padded items: row b of (B, P); packed items: a flat array, item_ends = inclusive scan of the item counts
  int64_t item_base = static_cast<int64_t>(b) * P;
  if (item_ends) {   // clamped to the arrays, and to the item count the host sized shared memory for
    item_base = min(max(b > 0 ? static_cast<int64_t>(__ldg(item_ends + b - 1)) : 0, static_cast<int64_t>(0)), total_items);
    const int64_t stop = min(max(static_cast<int64_t>(__ldg(item_ends + b)), item_base), total_items);
    P = static_cast<int>(min(stop - item_base, static_cast<int64_t>(P)));
    x_sb = 0;
    x += item_base * x_sp;
  }
  const int32_t* ends_row = ends + item_base;
  const int64_t n_b = P > 0 ? static_cast<int64_t>(__ldg(ends_row + P - 1)) : 0;
  const int64_t valid_end = min(t1, n_b);
  unsigned char* out_b = out + static_cast<int64_t>(b) * T * row_bytes;
  const uint64_t policy = mg_policy_evict_first();

  // ---- padding rows: bulk copies from a zero tile -------------------------------------------------------------
  const int64_t pad_begin = max(t0, n_b);
  const bool has_padding = pad_begin < t1;     // CTA-uniform
  if (has_padding) {
    uint4* z = reinterpret_cast<uint4*>(zero_tile);
    const int nz = static_cast<int>(zero_rows * (row_bytes / 16));
    for (int i = threadIdx.x; i < nz; i += kUpThreads) z[i] = make_uint4(0, 0, 0, 0);
    mg_fence_proxy_async_smem();
  }
  const int32_t* e = ends_row;
  if (t0 < valid_end) e = mg_stage_ends(ends_row, P, smem_ends, ends_fit);   // contains a __syncthreads
  else if (has_padding) __syncthreads();
  if (has_padding && lane == 0) {
    int piece = 0;
    for (int64_t r = pad_begin; r < t1; r += zero_rows, ++piece) {
      if ((piece % kUpWarps) != warp) continue;
      const uint32_t rows = static_cast<uint32_t>(min(static_cast<int64_t>(zero_rows), t1 - r));
      mg_bulk_store_hint(out_b + r * row_bytes, mg_smem_addr(zero_tile), rows * row_bytes, policy);
    }
    mg_bulk_commit();
  }

  // ---- valid rows: one item per warp at a time -----------------------------------------------------------------
  if (t0 < valid_end) {
    const float* q0 = p0 + static_cast<int64_t>(b) * p_sb;
    const float* q1 = p1 + static_cast<int64_t>(b) * p_sb;
    const int first_item = mg_item_of_frame(e, P, t0);
    int use = 0;
    for (int p = first_item + warp; p < P; p += kUpWarps) {
      const int64_t start = p > 0 ? static_cast<int64_t>(e[p - 1]) : 0;
      if (start >= valid_end) break;
      const int64_t r_begin = max(start, t0);
      const int64_t r_end = min(static_cast<int64_t>(e[p]), valid_end);
      if (r_end <= r_begin) continue;   // zero-duration item

      unsigned char* slot = slots + static_cast<size_t>(warp * kUpSlots + (use % kUpSlots)) * row_bytes;
      // The slot was last read by the bulk group issued kUpSlots uses ago: let all but the newest kUpSlots-1 drain.
      if (use >= kUpSlots) {
        if (lane == 0) mg_bulk_wait_read<kUpSlots - 1>();
        __syncwarp();
      }
      ++use;

      const float4* src = reinterpret_cast<const float4*>(x + static_cast<int64_t>(b) * x_sb + static_cast<int64_t>(p) * x_sp);
      float4* dst = reinterpret_cast<float4*>(slot);
      for (int i0 = lane; i0 < nvec; i0 += 32 * kBatch) {
        float4 v[kBatch];
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
          const int i = i0 + 32 * j;
          if (i < nvec) v[j] = __ldcs(src + i);
        }
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
          const int i = i0 + 32 * j;
          if (i < nvec) {
            const float4 r = mg_norm_vec<MODE>(v[j], q0, q1, i);
            if (OUT_BF16) {
              const __nv_bfloat162 lo = __floats2bfloat162_rn(r.x, r.y), hi = __floats2bfloat162_rn(r.z, r.w);
              reinterpret_cast<uint2*>(slot)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo),
                                                             *reinterpret_cast<const uint32_t*>(&hi));
            } else {
              dst[i] = r;
            }
          }
        }
      }
      mg_fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        const uint32_t src_addr = mg_smem_addr(slot);
        for (int64_t r = r_begin; r < r_end; ++r) mg_bulk_store_hint(out_b + r * row_bytes, src_addr, row_bytes, policy);
        mg_bulk_commit();
      }
    }
  }
  // Shared memory must outlive every bulk read that sources from it.
  if (lane == 0) mg_bulk_wait_read<0>();
}

// ------------------------------------------------------------------------------------------------------------------
// DIRECT path (register-staged rows, vector stores).  Vec: float4 / float (f32, optionally normalised) or
// uint4 / uint2 / uint32_t / uint16_t / uint8_t (raw bytes of any dtype).
// ------------------------------------------------------------------------------------------------------------------
template <typename Vec>
__device__ __forceinline__ Vec mg_zero_vec() {
  Vec z;
  memset(&z, 0, sizeof(Vec));
  return z;
}

template <typename Vec, int MODE>
__global__ void __launch_bounds__(kUpThreads)
upsample_direct_kernel(const unsigned char* __restrict__ x, int64_t x_sb, int64_t x_sp, const int32_t* __restrict__ ends,
                       const float* __restrict__ p0, const float* __restrict__ p1, int64_t p_sb,
                       unsigned char* __restrict__ out, int P, int nvec /* row_bytes / sizeof(Vec) */, int64_t T,
                       int rows_per_cta, const int32_t* __restrict__ item_ends, int64_t total_items) {
  extern __shared__ __align__(128) unsigned char smem[];
  int32_t* smem_ends = reinterpret_cast<int32_t*>(smem);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * rows_per_cta;
  const int64_t t1 = min(t0 + rows_per_cta, T);
  const bool ends_fit = P <= kMaxEndsSmem;
  int64_t item_base = static_cast<int64_t>(b) * P;
  if (item_ends) {   // clamped to the arrays, and to the item count the host sized shared memory for
    item_base = min(max(b > 0 ? static_cast<int64_t>(__ldg(item_ends + b - 1)) : 0, static_cast<int64_t>(0)), total_items);
    const int64_t stop = min(max(static_cast<int64_t>(__ldg(item_ends + b)), item_base), total_items);
    P = static_cast<int>(min(stop - item_base, static_cast<int64_t>(P)));
    x_sb = 0;
    x += item_base * x_sp;
  }
  const int32_t* ends_row = ends + item_base;
  const int64_t n_b = P > 0 ? static_cast<int64_t>(__ldg(ends_row + P - 1)) : 0;
  const int64_t valid_end = min(t1, n_b);
  Vec* out_b = reinterpret_cast<Vec*>(out) + static_cast<int64_t>(b) * T * nvec;

  // padding rows: one flat, fully coalesced zero fill
  const int64_t pad_begin = max(t0, n_b);
  if (pad_begin < t1) {
    Vec* z = out_b + pad_begin * nvec;
    const int64_t n = (t1 - pad_begin) * nvec;
    const Vec zero = mg_zero_vec<Vec>();
    for (int64_t i = threadIdx.x; i < n; i += kUpThreads) z[i] = zero;
  }
  if (t0 >= valid_end) return;   // CTA-uniform

  const int32_t* e = mg_stage_ends(ends_row, P, smem_ends, ends_fit);
  const float* q0 = p0 + static_cast<int64_t>(b) * p_sb;
  const float* q1 = p1 + static_cast<int64_t>(b) * p_sb;
  const int first_item = mg_item_of_frame(e, P, t0);
  for (int p = first_item + warp; p < P; p += kUpWarps) {
    const int64_t start = p > 0 ? static_cast<int64_t>(e[p - 1]) : 0;
    if (start >= valid_end) break;
    const int64_t r_begin = max(start, t0);
    const int64_t r_end = min(static_cast<int64_t>(e[p]), valid_end);
    if (r_end <= r_begin) continue;
    const Vec* src = reinterpret_cast<const Vec*>(x + static_cast<int64_t>(b) * x_sb + static_cast<int64_t>(p) * x_sp);
    for (int i0 = lane; i0 < nvec; i0 += 32 * kBatch) {
      Vec v[kBatch];
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        const int i = i0 + 32 * j;
        if (i < nvec) v[j] = mg_norm_vec<MODE>(__ldg(src + i), q0, q1, i);
      }
      for (int64_t r = r_begin; r < r_end; ++r) {
        Vec* dst = out_b + r * nvec;
#pragma unroll
        for (int j = 0; j < kBatch; ++j) {
          const int i = i0 + 32 * j;
          if (i < nvec) dst[i] = v[j];
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Backward: deterministic per-item segment sum of grad_out rows (ascending t), divided by the normaliser's denominator.
// One warp per (utterance, item); lanes across the feature axis.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kBwdWarps = 8;

template <typename Vec, int MODE>
__global__ void __launch_bounds__(kBwdWarps * 32)
upsample_bwd_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ ends, const float* __restrict__ p0,
                    const float* __restrict__ p1, int64_t p_sb, float* __restrict__ grad_x, int64_t n_items_total,
                    int P, int nvec, int64_t T) {
  constexpr int F = VecTraits<Vec>::kFloats;
  const int lane = threadIdx.x & 31;
  const int64_t w = static_cast<int64_t>(blockIdx.x) * kBwdWarps + (threadIdx.x >> 5);
  if (w >= n_items_total) return;
  const int b = static_cast<int>(w / P), p = static_cast<int>(w % P);
  const int32_t* e = ends + static_cast<int64_t>(b) * P;
  // clamped to the rows grad_out has: with a caller-supplied max_len < n_b the forward truncated the utterance, so do we
  const int64_t start = min(p > 0 ? static_cast<int64_t>(__ldg(e + p - 1)) : 0, T);
  const int64_t stop = min(static_cast<int64_t>(__ldg(e + p)), T);
  const Vec* g = reinterpret_cast<const Vec*>(grad_out) + static_cast<int64_t>(b) * T * nvec;
  Vec* gx = reinterpret_cast<Vec*>(grad_x) + w * nvec;
  const float* q0 = p0 + static_cast<int64_t>(b) * p_sb;
  const float* q1 = p1 + static_cast<int64_t>(b) * p_sb;

  for (int i = lane; i < nvec; i += 32) {
    float acc[F];
#pragma unroll
    for (int k = 0; k < F; ++k) acc[k] = 0.f;
    int64_t t = start;
    for (; t + 4 <= stop; t += 4) {   // four loads in flight, summed in ascending t
      Vec v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = __ldcs(g + (t + j) * nvec + i);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float* f = reinterpret_cast<const float*>(&v[j]);
#pragma unroll
        for (int k = 0; k < F; ++k) acc[k] = __fadd_rn(acc[k], f[k]);
      }
    }
    for (; t < stop; ++t) {
      const Vec v = __ldcs(g + t * nvec + i);
      const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
      for (int k = 0; k < F; ++k) acc[k] = __fadd_rn(acc[k], f[k]);
    }
    if (MODE != MG_NORM_NONE) {
#pragma unroll
      for (int k = 0; k < F; ++k)
        acc[k] = __fdiv_rn(acc[k], mg_denominator(MODE, __ldg(q0 + i * F + k), __ldg(q1 + i * F + k)));
    }
    Vec o;
    float* of = reinterpret_cast<float*>(&o);
#pragma unroll
    for (int k = 0; k < F; ++k) of[k] = acc[k];
    gx[i] = o;
  }
}

// Row-streaming form for rows of up to 32 * NS float4 (D <= 640 at NS = 5): a lane owns the NS vectors lane, lane + 32, ... of
// EVERY row, so a warp reads whole contiguous rows (2400 bytes at D = 600) with two rows -- 2 * NS 16-byte loads per lane -- in
// flight, instead of walking the item's rows once per 512-byte column strip.  Same ascending-t order per element: same bits.
template <int MODE, int NS>
__global__ void __launch_bounds__(kBwdWarps * 32, 3)   // <= 85 registers: three CTAs per SM (86 registers gave two)
upsample_bwd_rows_kernel(const float* __restrict__ grad_out, const int32_t* __restrict__ ends, const float* __restrict__ p0,
                         const float* __restrict__ p1, int64_t p_sb, float* __restrict__ grad_x, int64_t n_items_total,
                         int P, int nvec, int64_t T) {
  const int lane = threadIdx.x & 31;
  const int64_t w = static_cast<int64_t>(blockIdx.x) * kBwdWarps + (threadIdx.x >> 5);
  if (w >= n_items_total) return;
  const int b = static_cast<int>(w / P), p = static_cast<int>(w % P);
  const int32_t* e = ends + static_cast<int64_t>(b) * P;
  // clamped to the rows grad_out has: with a caller-supplied max_len < n_b the forward truncated the utterance, so do we
  const int64_t start = min(p > 0 ? static_cast<int64_t>(__ldg(e + p - 1)) : 0, T);
  const int64_t stop = min(static_cast<int64_t>(__ldg(e + p)), T);
  const float4* g = reinterpret_cast<const float4*>(grad_out) + static_cast<int64_t>(b) * T * nvec;
  float4 acc[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
  auto add = [](float4& a, const float4& v) {
    a.x = __fadd_rn(a.x, v.x); a.y = __fadd_rn(a.y, v.y); a.z = __fadd_rn(a.z, v.z); a.w = __fadd_rn(a.w, v.w);
  };
  int64_t t = start;
  for (; t + 2 <= stop; t += 2) {
    float4 v0[NS], v1[NS];
    const float4* r0 = g + t * nvec;
    const float4* r1 = r0 + nvec;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int i = lane + 32 * s;
      if (i < nvec) { v0[s] = __ldcs(r0 + i); v1[s] = __ldcs(r1 + i); }
    }
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (lane + 32 * s < nvec) { add(acc[s], v0[s]); add(acc[s], v1[s]); }
  }
  if (t < stop) {
    const float4* r0 = g + t * nvec;
#pragma unroll
    for (int s = 0; s < NS; ++s)
      if (lane + 32 * s < nvec) add(acc[s], __ldcs(r0 + lane + 32 * s));
  }
  const float* q0 = p0 + static_cast<int64_t>(b) * p_sb;
  const float* q1 = p1 + static_cast<int64_t>(b) * p_sb;
  float4* gx = reinterpret_cast<float4*>(grad_x) + w * nvec;
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int i = lane + 32 * s;
    if (i >= nvec) continue;
    float4 o = acc[s];
    if (MODE != MG_NORM_NONE) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(q0) + i), c = __ldg(reinterpret_cast<const float4*>(q1) + i);
      o.x = __fdiv_rn(o.x, mg_denominator(MODE, a.x, c.x));
      o.y = __fdiv_rn(o.y, mg_denominator(MODE, a.y, c.y));
      o.z = __fdiv_rn(o.z, mg_denominator(MODE, a.z, c.z));
      o.w = __fdiv_rn(o.w, mg_denominator(MODE, a.w, c.w));
    }
    gx[i] = o;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------------------------
int mg_rows_per_cta(int64_t T, int B, int64_t row_bytes) {
  // Target ~128 KB of output per CTA (amortises the search + first-load latency) while keeping >= ~8 CTAs per SM's
  // worth of work in the grid.  Overridable for experiments.
  static int forced = -2;
  if (forced == -2) {
    const char* env = getenv("MG_UPSAMPLE_ROWS_PER_CTA");
    forced = env ? atoi(env) : -1;
  }
  if (forced > 0) return forced;
  int64_t rows = (128 * 1024 + row_bytes - 1) / row_bytes;
  if (rows < 8) rows = 8;
  const int64_t sms = mg_cached_sm_count();
  while (rows > 8 && static_cast<int64_t>(B) * ((T + rows - 1) / rows) < 8 * sms) rows = (rows + 1) / 2;
  if (rows > T) rows = T > 0 ? T : 1;
  return static_cast<int>(rows);
}

template <int MODE, bool OUT_BF16 = false>
int launch_bulk(const unsigned char* x, int64_t x_sb, int64_t x_sp, const int32_t* ends, const float* p0, const float* p1,
                int64_t p_sb, unsigned char* out, int B, int P, int64_t in_row_bytes, int64_t T, cudaStream_t stream,
                const int32_t* item_ends = nullptr, int64_t total_items = 0) {
  const int64_t row_bytes = OUT_BF16 ? in_row_bytes / 2 : in_row_bytes;   // output row
  const int rows = mg_rows_per_cta(T, B, row_bytes);
  int zero_rows = static_cast<int>((16 * 1024) / row_bytes);
  if (zero_rows < 1) zero_rows = 1;
  if (zero_rows > rows) zero_rows = rows;
  const size_t ends_bytes = (P <= kMaxEndsSmem) ? static_cast<size_t>(P) * 4 : 0;
  const size_t smem = static_cast<size_t>(kUpWarps * kUpSlots + zero_rows) * row_bytes + ends_bytes;
  auto kernel = upsample_bulk_kernel<MODE, OUT_BF16>;
  if (smem > 48 * 1024) {
    static int attr_smem[64] = {};   // per instantiation and device: the largest size asked for so far
    int device = 0;
    MG_CUDA_OK(cudaGetDevice(&device));
    if (static_cast<int>(smem) > attr_smem[device & 63]) {
      MG_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
      attr_smem[device & 63] = static_cast<int>(smem);
    }
  }
  dim3 grid(static_cast<unsigned>((T + rows - 1) / rows), static_cast<unsigned>(B));
  MG_CUDA_OK(mg_launch_pdl(kernel, grid, dim3(kUpThreads), smem, stream, x, x_sb, x_sp, ends, p0, p1, p_sb, out, P,
                           static_cast<int>(in_row_bytes / 16), T, rows, zero_rows, item_ends, total_items));
  MG_LAUNCH_OK();
  return MG_OK;
}

template <typename Vec, int MODE>
int launch_direct(const unsigned char* x, int64_t x_sb, int64_t x_sp, const int32_t* ends, const float* p0,
                  const float* p1, int64_t p_sb, unsigned char* out, int B, int P, int64_t row_bytes, int64_t T,
                  cudaStream_t stream, const int32_t* item_ends = nullptr, int64_t total_items = 0) {
  const int rows = mg_rows_per_cta(T, B, row_bytes);
  const size_t smem = (P <= kMaxEndsSmem) ? static_cast<size_t>(P) * 4 : 0;
  dim3 grid(static_cast<unsigned>((T + rows - 1) / rows), static_cast<unsigned>(B));
  upsample_direct_kernel<Vec, MODE><<<grid, kUpThreads, smem, stream>>>(
      x, x_sb, x_sp, ends, p0, p1, p_sb, out, P, static_cast<int>(row_bytes / sizeof(Vec)), T, rows, item_ends, total_items);
  MG_LAUNCH_OK();
  return MG_OK;
}

bool bulk_eligible(const void* x, int64_t x_sb, int64_t x_sp, const void* out, int64_t row_bytes, int P) {
  if (row_bytes % 16 != 0 || !mg_aligned(out, 16) || !mg_aligned(x, 16) || x_sb % 16 != 0 || x_sp % 16 != 0) return false;
  const size_t ends_bytes = (P <= kMaxEndsSmem) ? static_cast<size_t>(P) * 4 : 0;
  return static_cast<size_t>(kUpWarps * kUpSlots + 1) * row_bytes + ends_bytes <= 160 * 1024;
}

}  // namespace

extern "C" int mg_upsample_norm_f32(const float* x, int64_t x_stride_b, int64_t x_stride_p, const int32_t* ends,
                                    const float* p0, const float* p1, int64_t param_stride_b, int norm_mode, float* out,
                                    int B, int P, int D, int64_t T, int path, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(B >= 0 && P >= 0 && D >= 0 && T >= 0, "mg_upsample_norm_f32: negative shape");
  MG_REQUIRE(norm_mode >= MG_NORM_NONE && norm_mode <= MG_NORM_MINMAX, "mg_upsample_norm_f32: bad norm_mode %d", norm_mode);
  MG_REQUIRE(B <= 65535, "mg_upsample_norm_f32: B=%d exceeds 65535 utterances per call", B);
  if (B == 0 || T == 0 || D == 0) return MG_OK;
  MG_REQUIRE(out != nullptr && (P == 0 || (x != nullptr && ends != nullptr)), "mg_upsample_norm_f32: NULL buffer");
  MG_REQUIRE(norm_mode == MG_NORM_NONE || (p0 != nullptr && p1 != nullptr), "mg_upsample_norm_f32: NULL parameters");
  if (norm_mode == MG_NORM_NONE) { p0 = p1 = nullptr; param_stride_b = 0; }

  const auto* xb = reinterpret_cast<const unsigned char*>(x);
  auto* ob = reinterpret_cast<unsigned char*>(out);
  const int64_t row_bytes = static_cast<int64_t>(D) * 4, x_sb = x_stride_b * 4, x_sp = x_stride_p * 4;
  const bool params_vec_ok = norm_mode == MG_NORM_NONE ||
                             (mg_aligned(p0, 16) && mg_aligned(p1, 16) && (param_stride_b % 4) == 0);
  const bool can_bulk = bulk_eligible(x, x_sb, x_sp, out, row_bytes, P) && params_vec_ok;
  MG_REQUIRE(path != MG_PATH_BULK || can_bulk, "mg_upsample_norm_f32: bulk path needs D %% 4 == 0 and 16-byte alignment");

  if (can_bulk && path != MG_PATH_DIRECT) {
    switch (norm_mode) {
      case MG_NORM_NONE: return launch_bulk<MG_NORM_NONE>(xb, x_sb, x_sp, ends, p0, p1, param_stride_b, ob, B, P, row_bytes, T, stream);
      case MG_NORM_MVN: return launch_bulk<MG_NORM_MVN>(xb, x_sb, x_sp, ends, p0, p1, param_stride_b, ob, B, P, row_bytes, T, stream);
      default: return launch_bulk<MG_NORM_MINMAX>(xb, x_sb, x_sp, ends, p0, p1, param_stride_b, ob, B, P, row_bytes, T, stream);
    }
  }
  const bool vec4 = (D % 4 == 0) && mg_aligned(x, 16) && mg_aligned(out, 16) && x_sb % 16 == 0 && x_sp % 16 == 0 && params_vec_ok;
  if (vec4) {
    switch (norm_mode) {
      case MG_NORM_NONE: return launch_direct<float4, MG_NORM_NONE>(xb, x_sb, x_sp, ends, p0, p1, param_stride_b, ob, B, P, row_bytes, T, stream);
      case MG_NORM_MVN: return launch_direct<float4, MG_NORM_MVN>(xb, x_sb, x_sp, ends, p0, p1, param_stride_b, ob, B, P, row_bytes, T, stream);
      default: return launch_direct<float4, MG_NORM_MINMAX>(xb, x_sb, x_sp, ends, p0, p1, param_stride_b, ob, B, P, row_bytes, T, stream);
    }
  }
  switch (norm_mode) {
    case MG_NORM_NONE: return launch_direct<float, MG_NORM_NONE>(xb, x_sb, x_sp, ends, p0, p1, param_stride_b, ob, B, P, row_bytes, T, stream);
    case MG_NORM_MVN: return launch_direct<float, MG_NORM_MVN>(xb, x_sb, x_sp, ends, p0, p1, param_stride_b, ob, B, P, row_bytes, T, stream);
    default: return launch_direct<float, MG_NORM_MINMAX>(xb, x_sb, x_sp, ends, p0, p1, param_stride_b, ob, B, P, row_bytes, T, stream);
  }
}

extern "C" int mg_upsample_packed_norm_f32(const float* x, int64_t x_stride_p, const int32_t* item_ends, int64_t total_items,
                                           const int32_t* ends,
                                           const float* p0, const float* p1, int64_t param_stride_b, int norm_mode,
                                           float* out, int B, int max_items, int D, int64_t T, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(B >= 0 && max_items >= 0 && D >= 0 && T >= 0 && total_items >= 0, "mg_upsample_packed_norm_f32: negative shape");
  MG_REQUIRE(norm_mode >= MG_NORM_NONE && norm_mode <= MG_NORM_MINMAX, "mg_upsample_packed_norm_f32: bad norm_mode %d", norm_mode);
  MG_REQUIRE(B <= 65535, "mg_upsample_packed_norm_f32: B=%d exceeds 65535 utterances per call", B);
  if (B == 0 || T == 0 || D == 0) return MG_OK;
  MG_REQUIRE(out != nullptr && item_ends != nullptr && (max_items == 0 || (x != nullptr && ends != nullptr)),
             "mg_upsample_packed_norm_f32: NULL buffer");
  MG_REQUIRE(norm_mode == MG_NORM_NONE || (p0 != nullptr && p1 != nullptr), "mg_upsample_packed_norm_f32: NULL parameters");
  if (norm_mode == MG_NORM_NONE) { p0 = p1 = nullptr; param_stride_b = 0; }
  const auto* xb = reinterpret_cast<const unsigned char*>(x);
  auto* ob = reinterpret_cast<unsigned char*>(out);
  const int64_t row_bytes = static_cast<int64_t>(D) * 4, x_sp = x_stride_p * 4;
  const bool params_vec_ok = norm_mode == MG_NORM_NONE ||
                             (mg_aligned(p0, 16) && mg_aligned(p1, 16) && (param_stride_b % 4) == 0);
  // `max_items` bounds every utterance's item count (it sizes the shared-memory copy of the scan row)
  if (bulk_eligible(x, 0, x_sp, out, row_bytes, max_items) && params_vec_ok) {
    switch (norm_mode) {
      case MG_NORM_NONE: return launch_bulk<MG_NORM_NONE>(xb, 0, x_sp, ends, p0, p1, param_stride_b, ob, B, max_items, row_bytes, T, stream, item_ends, total_items);
      case MG_NORM_MVN: return launch_bulk<MG_NORM_MVN>(xb, 0, x_sp, ends, p0, p1, param_stride_b, ob, B, max_items, row_bytes, T, stream, item_ends, total_items);
      default: return launch_bulk<MG_NORM_MINMAX>(xb, 0, x_sp, ends, p0, p1, param_stride_b, ob, B, max_items, row_bytes, T, stream, item_ends, total_items);
    }
  }
  switch (norm_mode) {
    case MG_NORM_NONE: return launch_direct<float, MG_NORM_NONE>(xb, 0, x_sp, ends, p0, p1, param_stride_b, ob, B, max_items, row_bytes, T, stream, item_ends, total_items);
    case MG_NORM_MVN: return launch_direct<float, MG_NORM_MVN>(xb, 0, x_sp, ends, p0, p1, param_stride_b, ob, B, max_items, row_bytes, T, stream, item_ends, total_items);
    default: return launch_direct<float, MG_NORM_MINMAX>(xb, 0, x_sp, ends, p0, p1, param_stride_b, ob, B, max_items, row_bytes, T, stream, item_ends, total_items);
  }
}

extern "C" int mg_upsample_norm_f32_bf16out(const float* x, int64_t x_stride_b, int64_t x_stride_p, const int32_t* ends,
                                            const float* p0, const float* p1, int64_t param_stride_b, int norm_mode,
                                            void* out, int B, int P, int D, int64_t T, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(B >= 0 && P >= 0 && D >= 0 && T >= 0, "mg_upsample_norm_f32_bf16out: negative shape");
  MG_REQUIRE(norm_mode >= MG_NORM_NONE && norm_mode <= MG_NORM_MINMAX, "mg_upsample_norm_f32_bf16out: bad norm_mode %d", norm_mode);
  MG_REQUIRE(B <= 65535, "mg_upsample_norm_f32_bf16out: B=%d exceeds 65535 utterances per call", B);
  MG_REQUIRE(D % 8 == 0, "mg_upsample_norm_f32_bf16out: D=%d must be a multiple of 8 (16-byte bf16 rows)", D);
  if (B == 0 || T == 0 || D == 0) return MG_OK;
  MG_REQUIRE(out != nullptr && (P == 0 || (x != nullptr && ends != nullptr)), "mg_upsample_norm_f32_bf16out: NULL buffer");
  MG_REQUIRE(norm_mode == MG_NORM_NONE || (p0 != nullptr && p1 != nullptr), "mg_upsample_norm_f32_bf16out: NULL parameters");
  if (norm_mode == MG_NORM_NONE) { p0 = p1 = nullptr; param_stride_b = 0; }
  const auto* xb = reinterpret_cast<const unsigned char*>(x);
  auto* ob = static_cast<unsigned char*>(out);
  const int64_t row_bytes = static_cast<int64_t>(D) * 4, x_sb = x_stride_b * 4, x_sp = x_stride_p * 4;
  const bool params_ok = norm_mode == MG_NORM_NONE || (mg_aligned(p0, 16) && mg_aligned(p1, 16) && (param_stride_b % 4) == 0);
  MG_REQUIRE(bulk_eligible(x, x_sb, x_sp, out, row_bytes, P) && params_ok,
             "mg_upsample_norm_f32_bf16out: operands must be 16-byte aligned");
  switch (norm_mode) {
    case MG_NORM_NONE: return launch_bulk<MG_NORM_NONE, true>(xb, x_sb, x_sp, ends, p0, p1, param_stride_b, ob, B, P, row_bytes, T, stream);
    case MG_NORM_MVN: return launch_bulk<MG_NORM_MVN, true>(xb, x_sb, x_sp, ends, p0, p1, param_stride_b, ob, B, P, row_bytes, T, stream);
    default: return launch_bulk<MG_NORM_MINMAX, true>(xb, x_sb, x_sp, ends, p0, p1, param_stride_b, ob, B, P, row_bytes, T, stream);
  }
}

extern "C" int mg_upsample_bytes(const void* x, int64_t x_stride_b_bytes, int64_t x_stride_p_bytes, const int32_t* ends,
                                 void* out, int B, int P, int64_t row_bytes, int64_t T, int path, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(B >= 0 && P >= 0 && row_bytes >= 0 && T >= 0, "mg_upsample_bytes: negative shape");
  MG_REQUIRE(B <= 65535, "mg_upsample_bytes: B=%d exceeds 65535 utterances per call", B);
  MG_REQUIRE(row_bytes < (int64_t(1) << 31), "mg_upsample_bytes: row of %lld bytes is too long", (long long)row_bytes);
  if (B == 0 || T == 0 || row_bytes == 0) return MG_OK;
  MG_REQUIRE(out != nullptr && (P == 0 || (x != nullptr && ends != nullptr)), "mg_upsample_bytes: NULL buffer");
  const auto* xb = static_cast<const unsigned char*>(x);
  auto* ob = static_cast<unsigned char*>(out);
  const int64_t x_sb = x_stride_b_bytes, x_sp = x_stride_p_bytes;
  const bool can_bulk = bulk_eligible(x, x_sb, x_sp, out, row_bytes, P);
  MG_REQUIRE(path != MG_PATH_BULK || can_bulk, "mg_upsample_bytes: bulk path needs row_bytes %% 16 == 0 and 16-byte alignment");
  if (can_bulk && path != MG_PATH_DIRECT)
    return launch_bulk<MG_NORM_NONE>(xb, x_sb, x_sp, ends, nullptr, nullptr, 0, ob, B, P, row_bytes, T, stream);

  auto ok = [&](int64_t a) { return row_bytes % a == 0 && mg_aligned(x, a) && mg_aligned(out, a) && x_sb % a == 0 && x_sp % a == 0; };
  if (ok(16)) return launch_direct<uint4, MG_NORM_NONE>(xb, x_sb, x_sp, ends, nullptr, nullptr, 0, ob, B, P, row_bytes, T, stream);
  if (ok(8)) return launch_direct<uint2, MG_NORM_NONE>(xb, x_sb, x_sp, ends, nullptr, nullptr, 0, ob, B, P, row_bytes, T, stream);
  if (ok(4)) return launch_direct<uint32_t, MG_NORM_NONE>(xb, x_sb, x_sp, ends, nullptr, nullptr, 0, ob, B, P, row_bytes, T, stream);
  if (ok(2)) return launch_direct<uint16_t, MG_NORM_NONE>(xb, x_sb, x_sp, ends, nullptr, nullptr, 0, ob, B, P, row_bytes, T, stream);
  return launch_direct<uint8_t, MG_NORM_NONE>(xb, x_sb, x_sp, ends, nullptr, nullptr, 0, ob, B, P, row_bytes, T, stream);
}

extern "C" int mg_upsample_norm_bwd_f32(const float* grad_out, const int32_t* ends, const float* p0, const float* p1,
                                        int64_t param_stride_b, int norm_mode, float* grad_x, int B, int P, int D,
                                        int64_t T, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(B >= 0 && P >= 0 && D >= 0 && T >= 0, "mg_upsample_norm_bwd_f32: negative shape");
  MG_REQUIRE(norm_mode >= MG_NORM_NONE && norm_mode <= MG_NORM_MINMAX, "mg_upsample_norm_bwd_f32: bad norm_mode %d", norm_mode);
  if (B == 0 || P == 0 || D == 0) return MG_OK;
  MG_REQUIRE(ends != nullptr && grad_x != nullptr && (T == 0 || grad_out != nullptr), "mg_upsample_norm_bwd_f32: NULL buffer");
  MG_REQUIRE(norm_mode == MG_NORM_NONE || (p0 != nullptr && p1 != nullptr), "mg_upsample_norm_bwd_f32: NULL parameters");
  if (norm_mode == MG_NORM_NONE) { p0 = p1 = nullptr; param_stride_b = 0; }
  const int64_t n_items = static_cast<int64_t>(B) * P;
  const unsigned grid = static_cast<unsigned>((n_items + kBwdWarps - 1) / kBwdWarps);
  const bool vec4 = (D % 4 == 0) && mg_aligned(grad_out, 16) && mg_aligned(grad_x, 16);
#define MG_BWD(VEC, MODE, NV)                                                                                   \
  upsample_bwd_kernel<VEC, MODE><<<grid, kBwdWarps * 32, 0, stream>>>(grad_out, ends, p0, p1, param_stride_b, \
                                                                       grad_x, n_items, P, NV, T)
  static int rows_form = -1;   // MG_UPSAMPLE_BWD_ROWS=0: the column-strip kernel for every shape (measurements)
  if (rows_form < 0) { const char* env = getenv("MG_UPSAMPLE_BWD_ROWS"); rows_form = env ? atoi(env) : 1; }
  const bool params_vec = norm_mode == MG_NORM_NONE || (mg_aligned(p0, 16) && mg_aligned(p1, 16) && param_stride_b % 4 == 0);
  if (vec4 && rows_form && D / 4 <= 160 && params_vec) {
    const int nvec = D / 4, ns = (nvec + 31) / 32;
#define MG_BWD_ROWS(MODE, NS)                                                                                        \
  upsample_bwd_rows_kernel<MODE, NS><<<grid, kBwdWarps * 32, 0, stream>>>(grad_out, ends, p0, p1, param_stride_b, \
                                                                           grad_x, n_items, P, nvec, T)
#define MG_BWD_ROWS_MODE(NS)                                              \
  do {                                                                    \
    if (norm_mode == MG_NORM_NONE) MG_BWD_ROWS(MG_NORM_NONE, NS);         \
    else if (norm_mode == MG_NORM_MVN) MG_BWD_ROWS(MG_NORM_MVN, NS);      \
    else MG_BWD_ROWS(MG_NORM_MINMAX, NS);                                 \
  } while (0)
    if (ns == 1) MG_BWD_ROWS_MODE(1);
    else if (ns == 2) MG_BWD_ROWS_MODE(2);
    else if (ns == 3) MG_BWD_ROWS_MODE(3);
    else if (ns == 4) MG_BWD_ROWS_MODE(4);
    else MG_BWD_ROWS_MODE(5);
#undef MG_BWD_ROWS_MODE
#undef MG_BWD_ROWS
  } else if (vec4) {
    if (norm_mode == MG_NORM_NONE) MG_BWD(float4, MG_NORM_NONE, D / 4);
    else if (norm_mode == MG_NORM_MVN) MG_BWD(float4, MG_NORM_MVN, D / 4);
    else MG_BWD(float4, MG_NORM_MINMAX, D / 4);
  } else {
    if (norm_mode == MG_NORM_NONE) MG_BWD(float, MG_NORM_NONE, D);
    else if (norm_mode == MG_NORM_MVN) MG_BWD(float, MG_NORM_MVN, D);
    else MG_BWD(float, MG_NORM_MINMAX, D);
  }
#undef MG_BWD
  MG_LAUNCH_OK();
  return MG_OK;
}
