// K4 / K5 -- seq_len-masked losses (forward + optional backward) and streaming-metric accumulators in one pass.
//
// Replaces losses.sequence_loss -> mse / bce (reference morgana/losses.py:9-56: pointwise loss temp, sequence_mask
// temp, masked-product temp, sum over time, divide, mean = 4 full passes forward and ~3 backward) and the accumulate()
// arithmetic of metrics.Mean/RMSE/MAE/Error/Accuracy/F0Distortion/LF0Distortion/Distortion/MelCepDistortion
// (morgana/metrics.py:383-394, 492-495, 520-522, 547-549, 574-576, 597-609, 630-634, 657-665, 690-694), each of which
// builds a (B, T, 1) mask, a masked temp, two reductions and an .item() sync.
//
// The mask is never materialised: CTA (chunk, b, term) touches only rows t < n_b of utterance b, so padded input is
// not read at all.  Algorithmic HBM bytes: 8*D*sum_b n_b per two-operand term (+ 4*D*B*T when a gradient is written).
//
// Determinism: thread partials (fp64) -> fixed shuffle tree -> fixed order over warps -> one slot per CTA -> the last
// CTA to finish (integer ticket) combines slots in index order.  No floating-point atomics anywhere.
#include <string.h>

#include <stdlib.h>

#include "mg_common.cuh"
#include "mg_finish.cuh"

namespace {

constexpr int kRedThreads = 256;
constexpr int kRedWarps = kRedThreads / 32;
constexpr int kMaxChunks = kMgMaxChunks;
constexpr int64_t kTargetElems = 32768; // elements of one operand per CTA

struct ReduceParams {
  mg_term terms[MG_MAX_TERMS];
  MgFinishSlot slots[MG_MAX_TERMS];
  const int64_t* seq_len;
  MgWorkspace ws;
  int64_t T;
  int n_terms;
  int B;
  int slice_mode;   // 1: wide column slices stream flat in the forward pass (MG_RED_SLICE_MODE=0: thread-per-column)
};

// ---- pointwise functions ------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ float elem_fwd(float a, float b) {
  if constexpr (KIND == MG_RED_SQDIFF) {
    const float d = __fsub_rn(a, b);
    return __fmul_rn(d, d);
  } else if constexpr (KIND == MG_RED_ABSDIFF) {
    return fabsf(__fsub_rn(a, b));
  } else if constexpr (KIND == MG_RED_BCE) {
    // ATen binary_cross_entropy: (y - 1) * max(log1p(-p), -100) - y * max(log(p), -100)
    const float log_p = fmaxf(logf(a), -100.f);
    const float log_1mp = fmaxf(log1pf(-a), -100.f);
    return __fsub_rn(__fmul_rn(__fsub_rn(b, 1.f), log_1mp), __fmul_rn(b, log_p));
  } else if constexpr (KIND == MG_RED_SUM) {
    return a;
  } else if constexpr (KIND == MG_RED_SQ) {
    return __fmul_rn(a, a);
  } else if constexpr (KIND == MG_RED_SQDIFF_EXP) {
    const float d = __fsub_rn(expf(a), expf(b));   // full-precision expf, as torch.exp (metrics.py:631-632)
    return __fmul_rn(d, d);
  } else {  // MG_RED_ROOT_SQDIFF accumulates squared differences per frame; the root is taken by the caller
    const float d = __fsub_rn(a, b);
    return __fmul_rn(d, d);
  }
}

template <int KIND>
__device__ __forceinline__ float elem_bwd(float a, float b) {
  if constexpr (KIND == MG_RED_SQDIFF) {
    return __fmul_rn(2.f, __fsub_rn(a, b));
  } else if constexpr (KIND == MG_RED_ABSDIFF) {
    const float d = __fsub_rn(a, b);
    return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  } else if constexpr (KIND == MG_RED_SUM) {   // the masked mean of a caller-supplied per-element loss (sequence_loss, losses.py:9-47)
    return 1.f;
  } else {  // MG_RED_BCE: ATen binary_cross_entropy_backward, (p - y) / max((1 - p) * p, 1e-12)
    return __fdiv_rn(__fsub_rn(a, b), fmaxf(__fmul_rn(__fsub_rn(1.f, a), a), 1e-12f));
  }
}

__host__ __device__ constexpr bool kind_has_b(int kind) { return kind != MG_RED_SUM && kind != MG_RED_SQ; }

// ---- contiguous rows: flat 16-byte stream with an alignment peel -----------------------------------------------------
template <int KIND, bool GRAD>
__device__ __forceinline__ void run_flat(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ g,
                                         int64_t n, float w, double& sum) {
  constexpr bool HAS_B = kind_has_b(KIND);
  const int tid = threadIdx.x;
  const uintptr_t mis = reinterpret_cast<uintptr_t>(a) & 15;
  const bool vec_ok = (!HAS_B || (reinterpret_cast<uintptr_t>(b) & 15) == mis) &&
                      (!GRAD || (reinterpret_cast<uintptr_t>(g) & 15) == mis);
  const int64_t head = vec_ok ? min(n, static_cast<int64_t>(((16 - mis) & 15) >> 2)) : n;

  // scalar head (everything when the operands disagree on alignment)
  {
    int64_t i = tid;
    for (; i + 3 * kRedThreads < head; i += 4 * kRedThreads) {
      float av[4], bv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        av[j] = __ldcs(a + i + j * kRedThreads);
        bv[j] = HAS_B ? __ldcs(b + i + j * kRedThreads) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sum += static_cast<double>(elem_fwd<KIND>(av[j], bv[j]));
        if constexpr (GRAD) g[i + j * kRedThreads] = __fmul_rn(elem_bwd<KIND>(av[j], bv[j]), w);
      }
    }
    for (; i < head; i += kRedThreads) {
      const float av = __ldcs(a + i), bv = HAS_B ? __ldcs(b + i) : 0.f;
      sum += static_cast<double>(elem_fwd<KIND>(av, bv));
      if constexpr (GRAD) g[i] = __fmul_rn(elem_bwd<KIND>(av, bv), w);
    }
  }
  if (!vec_ok) return;

  const int64_t nvec = (n - head) >> 2;
  const float4* a4 = reinterpret_cast<const float4*>(a + head);
  const float4* b4 = reinterpret_cast<const float4*>(b + head);
  float4* g4 = reinterpret_cast<float4*>(g + head);
  auto one = [&](float4 av, float4 bv, int64_t i) {
    // Per-vector partial in fp32 order x,y,z,w, folded into the fp64 thread accumulator.
    const float f0 = elem_fwd<KIND>(av.x, bv.x), f1 = elem_fwd<KIND>(av.y, bv.y);
    const float f2 = elem_fwd<KIND>(av.z, bv.z), f3 = elem_fwd<KIND>(av.w, bv.w);
    sum += (static_cast<double>(f0) + static_cast<double>(f1)) + (static_cast<double>(f2) + static_cast<double>(f3));
    if constexpr (GRAD) {
      float4 gv;
      gv.x = __fmul_rn(elem_bwd<KIND>(av.x, bv.x), w);
      gv.y = __fmul_rn(elem_bwd<KIND>(av.y, bv.y), w);
      gv.z = __fmul_rn(elem_bwd<KIND>(av.z, bv.z), w);
      gv.w = __fmul_rn(elem_bwd<KIND>(av.w, bv.w), w);
      g4[i] = gv;
    }
  };
  int64_t i = tid;
  for (; i + 3 * kRedThreads < nvec; i += 4 * kRedThreads) {
    float4 av[4], bv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      av[j] = __ldcs(a4 + i + j * kRedThreads);
      bv[j] = HAS_B ? __ldcs(b4 + i + j * kRedThreads) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) one(av[j], bv[j], i + j * kRedThreads);
  }
  for (; i < nvec; i += kRedThreads) one(__ldcs(a4 + i), HAS_B ? __ldcs(b4 + i) : make_float4(0.f, 0.f, 0.f, 0.f), i);

  for (int64_t k = head + (nvec << 2) + tid; k < n; k += kRedThreads) {
    const float av = __ldcs(a + k), bv = HAS_B ? __ldcs(b + k) : 0.f;
    sum += static_cast<double>(elem_fwd<KIND>(av, bv));
    if constexpr (GRAD) g[k] = __fmul_rn(elem_bwd<KIND>(av, bv), w);
  }
}

// ---- wide column slices (at least half of the row, e.g. the 180 mcep columns of a 187-wide tensor, RNN_SPSS.py:134): the
// span from the first to the last slice element is streamed flat, 16 bytes at a time, like contiguous rows; the columns between
// two slice rows are read too (7 of 187: 4 % more bytes) and masked out; a vector's column advances without division.
// Forward only: with a gradient, written to a (B, T, D) tensor of its own, scalar stores from this traversal lose to the
// thread-per-column loop (0.489 vs 0.451 ms at config-3 scale).  0.276 -> 0.197 ms = 5.9 TB/s for the 180-of-187 slice.
template <int KIND>
__device__ __forceinline__ void run_flat_masked(const float* __restrict__ a, const float* __restrict__ b, int64_t st,
                                                int64_t n_rows, int D, double& sum) {
  constexpr bool HAS_B = kind_has_b(KIND);
  const int tid = threadIdx.x;
  const int ist = static_cast<int>(st);
  const int64_t n = (n_rows - 1) * st + D;
  const uintptr_t mis = reinterpret_cast<uintptr_t>(a) & 15;      // b has the same misalignment (checked by the caller)
  const int64_t head = min(n, static_cast<int64_t>(((16 - mis) & 15) >> 2));
  auto one = [&](float av, float bv) { sum += static_cast<double>(elem_fwd<KIND>(av, bv)); };
  auto scalar_at = [&](int64_t j) {
    const int64_t r = j / st;
    const int c = static_cast<int>(j - r * st);
    if (c < D) one(__ldcs(a + j), HAS_B ? __ldcs(b + j) : 0.f);
  };
  if (tid < head) scalar_at(tid);
  const int64_t nvec = (n - head) >> 2;
  const float4* a4 = reinterpret_cast<const float4*>(a + head);
  const float4* b4 = reinterpret_cast<const float4*>(b + head);
  const int64_t j0 = head + 4 * static_cast<int64_t>(tid);
  int col = static_cast<int>(j0 % st);                     // column of the vector's first element
  const int d_col = (4 * kRedThreads) % ist;               // one stride of the thread's vectors, in columns
  auto advance = [&](int& c) {
    c += d_col;
    if (c >= ist) c -= ist;
  };
  auto vec = [&](const float4& av, const float4& bv, int c) {
    const float af[4] = {av.x, av.y, av.z, av.w}, bf[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (c < D) one(af[k], bf[k]);
      if (++c == ist) c = 0;
    }
  };
  int64_t i = tid;
  for (; i + 3 * kRedThreads < nvec; i += 4 * kRedThreads) {
    float4 av[4], bv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      av[j] = __ldcs(a4 + i + j * kRedThreads);
      bv[j] = HAS_B ? __ldcs(b4 + i + j * kRedThreads) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      vec(av[j], bv[j], col);
      advance(col);
    }
  }
  for (; i < nvec; i += kRedThreads) {
    vec(__ldcs(a4 + i), HAS_B ? __ldcs(b4 + i) : make_float4(0.f, 0.f, 0.f, 0.f), col);
    advance(col);
  }
  const int64_t tail = head + (nvec << 2) + tid;
  if (tail < n) scalar_at(tail);
}

// ---- strided rows (column slices of a wider tensor) -------------------------------------------------------------------
// Thread -> (row group, column): consecutive threads read consecutive columns of a row (coalesced 4-byte loads) and each
// thread keeps 8 rows in flight.  No division in the loop.
template <int KIND, bool GRAD>
__device__ __forceinline__ void run_strided(const float* __restrict__ a, int64_t a_st, const float* __restrict__ b,
                                            int64_t b_st, float* __restrict__ g, int64_t g_st, int64_t n_rows, int D,
                                            float w, double& sum) {
  constexpr bool HAS_B = kind_has_b(KIND);
  constexpr int U = 8;
  const int groups = D >= kRedThreads ? 1 : kRedThreads / D;
  const int grp = D >= kRedThreads ? 0 : threadIdx.x / D;
  if (grp >= groups) return;
  for (int d = D >= kRedThreads ? threadIdx.x : threadIdx.x - grp * D; d < D; d += kRedThreads) {
    const float* pa = a + grp * a_st + d;
    const float* pb = HAS_B ? b + grp * b_st + d : nullptr;
    float* pg = GRAD ? g + grp * g_st + d : nullptr;
    const int64_t sa = groups * a_st, sb = groups * b_st, sg = groups * g_st;
    int64_t r = grp;
    for (; r + static_cast<int64_t>(U - 1) * groups < n_rows; r += static_cast<int64_t>(U) * groups) {
      float av[U], bv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        av[u] = __ldcs(pa + u * sa);
        bv[u] = HAS_B ? __ldcs(pb + u * sb) : 0.f;
      }
      float part = 0.f;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        part = __fadd_rn(part, elem_fwd<KIND>(av[u], bv[u]));
        if constexpr (GRAD) pg[u * sg] = __fmul_rn(elem_bwd<KIND>(av[u], bv[u]), w);
      }
      sum += static_cast<double>(part);
      pa += U * sa;
      if (HAS_B) pb += U * sb;
      if (GRAD) pg += U * sg;
    }
    for (; r < n_rows; r += groups) {
      const float av = __ldcs(pa), bv = HAS_B ? __ldcs(pb) : 0.f;
      sum += static_cast<double>(elem_fwd<KIND>(av, bv));
      if constexpr (GRAD) *pg = __fmul_rn(elem_bwd<KIND>(av, bv), w);
      pa += sa;
      if (HAS_B) pb += sb;
      if (GRAD) pg += sg;
    }
  }
}

// ---- one value per frame: Distortion's root, and every kind that carries a per-frame weight m ------------------------
__device__ __forceinline__ float load_weight(const void* m, int m_dtype, int flags, int64_t idx) {
  if (m_dtype == MG_DT_U8) return static_cast<float>(__ldg(static_cast<const unsigned char*>(m) + idx));
  const float w = __ldg(static_cast<const float*>(m) + idx);
  if (flags & MG_FLAG_M_GT_HALF) return w > 0.5f ? 1.f : 0.f;   // `vuv = output_features['vuv'] > 0.5`, RNN_SPSS.py:122
  return w;
}

template <int KIND>
__device__ __forceinline__ void run_per_frame(const float* __restrict__ a, int64_t a_st, const float* __restrict__ b,
                                              int64_t b_st, const void* m, int64_t m_st, int m_dtype, int flags,
                                              int64_t m_off, int64_t n_rows, int D, double& sum, double& cnt) {
  constexpr bool HAS_B = kind_has_b(KIND);
  for (int64_t r = threadIdx.x; r < n_rows; r += kRedThreads) {
    const float weight = m ? load_weight(m, m_dtype, flags, m_off + r * m_st) : 1.f;
    float acc = 0.f;   // the reference sums the feature axis in fp32 (metrics.py:661)
    for (int d = 0; d < D; ++d) {
      const float av = __ldg(a + r * a_st + d);
      const float bv = HAS_B ? __ldg(b + r * b_st + d) : 0.f;
      acc = __fadd_rn(acc, elem_fwd<KIND>(av, bv));
    }
    if constexpr (KIND == MG_RED_ROOT_SQDIFF) acc = sqrtf(acc);   // metrics.py:662
    sum += static_cast<double>(__fmul_rn(acc, weight));
    cnt += static_cast<double>(weight);
  }
}

// ---- cross-entropy over the feature axis (losses.ce, morgana/losses.py:59-61): ATen's log_softmax + nll_loss per frame,
// x[target] - max - log(sum exp(x - max)) negated; one thread per frame, two passes over its D logits ---------------------
__device__ __forceinline__ void run_cross_entropy(const float* __restrict__ a, int64_t a_st, const int64_t* __restrict__ target,
                                                  int64_t t_st, float* __restrict__ g, int64_t g_st, float w, int64_t n_rows,
                                                  int D, double& sum) {
  for (int64_t r = threadIdx.x; r < n_rows; r += kRedThreads) {
    const float* x = a + r * a_st;
    const int64_t cls = __ldg(target + r * t_st);
    float mx = -INFINITY;
    for (int d = 0; d < D; ++d) mx = fmaxf(mx, __ldg(x + d));
    float total = 0.f;
    for (int d = 0; d < D; ++d) total = __fadd_rn(total, expf(__fsub_rn(__ldg(x + d), mx)));
    const float log_total = logf(total);
    const bool valid_cls = cls >= 0 && cls < D;
    const float picked = valid_cls ? __ldg(x + cls) : 0.f;
    if (valid_cls) sum += static_cast<double>(-__fsub_rn(__fsub_rn(picked, mx), log_total));
    if (g != nullptr) {
      float* gr = g + r * g_st;
      for (int d = 0; d < D; ++d) {
        const float soft = valid_cls ? expf(__fsub_rn(__fsub_rn(__ldg(x + d), mx), log_total)) : 0.f;
        gr[d] = __fmul_rn(__fsub_rn(soft, d == cls ? 1.f : 0.f), w);
      }
    }
  }
}

// ---- integer / comparison kinds on uint8 (bool) or float operands -----------------------------------------------------
__device__ __forceinline__ float load_as_float(const void* p, bool is_u8, int64_t idx) {
  if (is_u8) return static_cast<float>(__ldg(static_cast<const unsigned char*>(p) + idx));
  return __ldg(static_cast<const float*>(p) + idx);
}

__device__ __forceinline__ void run_discrete(const mg_term& tm, int64_t a_off, int64_t b_off, int64_t n_rows,
                                             double& sum) {
  const int D = tm.D;
  const bool a_u8 = tm.ab_dtype == MG_DT_U8;
  const bool b_u8 = a_u8 || tm.b_is_u8;
  int64_t r = threadIdx.x / D;
  int d = threadIdx.x % D;
  const int64_t step_r = kRedThreads / D;
  const int step_d = kRedThreads % D;
  long long acc = 0;
  while (r < n_rows) {
    const int64_t ia = a_off + r * tm.a_st + d, ib = b_off + r * tm.b_st + d;
    if (tm.kind == MG_RED_EQ) {
      float av = load_as_float(tm.a, a_u8, ia), bv = load_as_float(tm.b, b_u8, ib);
      if (tm.flags & MG_FLAG_A_GT_HALF) { av = av > 0.5f ? 1.f : 0.f; bv = bv != 0.f ? 1.f : 0.f; }
      acc += av == bv ? 1 : 0;
    } else {
      const unsigned av = __ldg(static_cast<const unsigned char*>(tm.a) + ia);
      if (tm.kind == MG_RED_SUM) acc += av;
      else {
        const unsigned bv = __ldg(static_cast<const unsigned char*>(tm.b) + ib);
        acc += (tm.kind == MG_RED_XOR) ? (av ^ bv) : (av & bv);
      }
    }
    r += step_r;
    d += step_d;
    if (d >= D) { d -= D; ++r; }
  }
  sum += static_cast<double>(acc);   // exact: |acc| < 2^53
}

// Zero rows [row_begin, row_end) of the gradient (the padding of this chunk).
__device__ __forceinline__ void zero_grad_rows(float* g, int64_t g_st, int D, int64_t row_begin, int64_t row_end) {
  if (row_end <= row_begin) return;
  if (g_st == D) {
    float* z = g + row_begin * g_st;
    const int64_t n = (row_end - row_begin) * D;
    const int64_t head = min(n, static_cast<int64_t>(((16 - (reinterpret_cast<uintptr_t>(z) & 15)) & 15) >> 2));
    for (int64_t i = threadIdx.x; i < head; i += kRedThreads) z[i] = 0.f;
    const int64_t nvec = (n - head) >> 2;
    float4* z4 = reinterpret_cast<float4*>(z + head);
    for (int64_t i = threadIdx.x; i < nvec; i += kRedThreads) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = head + (nvec << 2) + threadIdx.x; i < n; i += kRedThreads) z[i] = 0.f;
  } else {
    const int64_t n = (row_end - row_begin) * D;
    for (int64_t i = threadIdx.x; i < n; i += kRedThreads) g[(row_begin + i / D) * g_st + (i % D)] = 0.f;
  }
}

template <int KIND>
__device__ __forceinline__ void run_float_term(const mg_term& tm, int b, int64_t r0, int64_t n_valid, int64_t r1,
                                               float w, double& sum, double& cnt, int slice_mode) {
  const float* a = static_cast<const float*>(tm.a) + b * tm.a_sb + r0 * tm.a_st;
  const float* bb = kind_has_b(KIND) ? static_cast<const float*>(tm.b) + b * tm.b_sb + r0 * tm.b_st : nullptr;
  const int D = tm.D;
  if (tm.m != nullptr || KIND == MG_RED_ROOT_SQDIFF) {
    run_per_frame<KIND>(a, tm.a_st, bb, tm.b_st, tm.m, tm.m_st, tm.m_dtype, tm.flags, b * tm.m_sb + r0 * tm.m_st, n_valid, D, sum, cnt);
    return;
  }
  constexpr bool CAN_GRAD = KIND == MG_RED_SQDIFF || KIND == MG_RED_ABSDIFF || KIND == MG_RED_BCE || KIND == MG_RED_SUM;
  const bool contiguous = tm.a_st == D && (!kind_has_b(KIND) || tm.b_st == D);
  // a wide slice of a row-major tensor: both operands with the same row stride and the same position inside a 16-byte line
  const bool wide_slice = !contiguous && n_valid > 0 && D >= 4 && 2 * static_cast<int64_t>(D) >= tm.a_st && tm.a_st <= 4 * kRedThreads &&
                          (!kind_has_b(KIND) || (tm.b_st == tm.a_st && ((reinterpret_cast<uintptr_t>(a) ^ reinterpret_cast<uintptr_t>(bb)) & 15) == 0));
  if constexpr (CAN_GRAD) {
    if (tm.grad != nullptr) {
      float* g = tm.grad + b * tm.g_sb + r0 * tm.g_st;
      if (contiguous && tm.g_st == D) run_flat<KIND, true>(a, bb, g, n_valid * D, w, sum);
      else run_strided<KIND, true>(a, tm.a_st, bb, tm.b_st, g, tm.g_st, n_valid, D, w, sum);
      zero_grad_rows(g, tm.g_st, D, n_valid, r1 - r0);
      return;
    }
  }
  if (contiguous) run_flat<KIND, false>(a, bb, nullptr, n_valid * D, 0.f, sum);
  else if (wide_slice && slice_mode) run_flat_masked<KIND>(a, bb, tm.a_st, n_valid, D, sum);
  else run_strided<KIND, false>(a, tm.a_st, bb, tm.b_st, nullptr, 0, n_valid, D, 0.f, sum);
}

__global__ void __launch_bounds__(kRedThreads, 4)
masked_reduce_kernel(const __grid_constant__ ReduceParams prm) {
  __shared__ double s_red[96];
  __shared__ bool s_is_last;
  mg_pdl_wait();                 // programmatic dependent launch (mg_common.cuh): nothing above touches global memory
  mg_pdl_launch_dependents();
  double* s_a = s_red;
  double* s_b = s_red + 32;

  const int term_idx = blockIdx.z, b = blockIdx.y, chunk = blockIdx.x;
  const mg_term& tm = prm.terms[term_idx];
  const int64_t T = prm.T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (chunk < prm.slots[term_idx].n_chunks) {
    const int64_t n_b = mg_valid_frames(prm.seq_len, b, T);
    const int64_t R = prm.slots[term_idx].rows_per_cta;
    const int64_t r0 = chunk * R;
    const int64_t r1 = min(r0 + R, T);
    const int64_t n_valid = max(static_cast<int64_t>(0), min(r1, n_b) - r0);   // valid rows of this chunk
    double sum = 0., cnt = 0.;

    float w = 0.f;
    if (tm.grad != nullptr) {
      double scale = static_cast<double>(tm.grad_scale);
      if (tm.grad_scale_dev != nullptr) scale *= static_cast<double>(__ldg(tm.grad_scale_dev));
      w = static_cast<float>(scale / (static_cast<double>(n_b) * prm.B * (tm.kind == MG_RED_CE ? 1 : tm.D)));
    }

    if (n_valid > 0 || tm.grad != nullptr) {
      if (tm.ab_dtype == MG_DT_U8 || tm.kind == MG_RED_EQ) {
        run_discrete(tm, b * tm.a_sb + r0 * tm.a_st, b * tm.b_sb + r0 * tm.b_st, n_valid, sum);
      } else if (tm.kind == MG_RED_CE) {
        float* g = tm.grad != nullptr ? tm.grad + b * tm.g_sb + r0 * tm.g_st : nullptr;
        run_cross_entropy(static_cast<const float*>(tm.a) + b * tm.a_sb + r0 * tm.a_st, tm.a_st,
                          static_cast<const int64_t*>(tm.b) + b * tm.b_sb + r0 * tm.b_st, tm.b_st, g, tm.g_st, w, n_valid,
                          tm.D, sum);
        if (g != nullptr) zero_grad_rows(g, tm.g_st, tm.D, n_valid, r1 - r0);
      } else {
        switch (tm.kind) {
          case MG_RED_SQDIFF: run_float_term<MG_RED_SQDIFF>(tm, b, r0, n_valid, r1, w, sum, cnt, prm.slice_mode); break;
          case MG_RED_ABSDIFF: run_float_term<MG_RED_ABSDIFF>(tm, b, r0, n_valid, r1, w, sum, cnt, prm.slice_mode); break;
          case MG_RED_BCE: run_float_term<MG_RED_BCE>(tm, b, r0, n_valid, r1, w, sum, cnt, prm.slice_mode); break;
          case MG_RED_SUM: run_float_term<MG_RED_SUM>(tm, b, r0, n_valid, r1, w, sum, cnt, prm.slice_mode); break;
          case MG_RED_ROOT_SQDIFF: run_float_term<MG_RED_ROOT_SQDIFF>(tm, b, r0, n_valid, r1, w, sum, cnt, prm.slice_mode); break;
          case MG_RED_SQ: run_float_term<MG_RED_SQ>(tm, b, r0, n_valid, r1, w, sum, cnt, prm.slice_mode); break;
          default: run_float_term<MG_RED_SQDIFF_EXP>(tm, b, r0, n_valid, r1, w, sum, cnt, prm.slice_mode); break;
        }
      }
    }

    if (n_valid > 0) {
      sum = mg_warp_sum(sum);
      cnt = mg_warp_sum(cnt);
      if (lane == 0) { s_a[warp] = sum; s_b[warp] = cnt; }
      __syncthreads();
      if (threadIdx.x == 0) {
        double s = 0., c = 0.;
#pragma unroll
        for (int i = 0; i < kRedWarps; ++i) { s += s_a[i]; c += s_b[i]; }
        prm.ws.partials[(static_cast<int64_t>(term_idx) * prm.B + b) * kMaxChunks + chunk] = make_double2(s, c);
      }
    }
  }

  // ---- ticket: the last CTA of the grid combines all slots in index order --------------------------------------
  mg_finish(prm.slots, prm.n_terms, prm.seq_len, prm.B, T, prm.ws, b, gridDim.x * gridDim.z, s_red, &s_is_last);
}

int rows_per_cta_for(int D, int B, int64_t T, int sms) {
  int64_t rows = kTargetElems / (D > 0 ? D : 1);
  if (rows < 16) rows = 16;
  if (rows > T) rows = T;
  // Small batches: split further so the grid covers the machine a few times over.
  while (rows > 16 && static_cast<int64_t>(B) * ((T + rows - 1) / rows) < 4 * static_cast<int64_t>(sms)) rows = (rows + 1) / 2;
  const int64_t min_rows = (T + kMaxChunks - 1) / kMaxChunks;
  if (rows < min_rows) rows = min_rows;
  if (rows < 1) rows = 1;
  return static_cast<int>(rows);
}

}  // namespace

extern "C" int64_t mg_masked_reduce_workspace_bytes(int n_terms, int B, int64_t T) {
  (void)T;
  if (n_terms < 0 || B < 0) return MG_ERR_INVALID_ARG;
  const int64_t chunked = mg_workspace_bytes(n_terms, B);
  const int64_t streamed = kMgTicketBytes + 1024 * static_cast<int64_t>(n_terms) * 32;   // per-CTA partials of the row-stream form
  return chunked > streamed ? chunked : streamed;
}

extern "C" int mg_masked_reduce(const mg_term* terms, int n_terms, const int64_t* seq_len, int B, int64_t T,
                                void* workspace, int64_t workspace_bytes, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(n_terms >= 1 && n_terms <= MG_MAX_TERMS, "mg_masked_reduce: n_terms=%d outside [1, %d]", n_terms, MG_MAX_TERMS);
  MG_REQUIRE(B >= 1 && B <= 65535, "mg_masked_reduce: B=%d outside [1, 65535]", B);
  MG_REQUIRE(T >= 0 && T < (int64_t(1) << 31), "mg_masked_reduce: bad T");
  MG_REQUIRE(terms != nullptr && workspace != nullptr, "mg_masked_reduce: NULL buffer");
  MG_REQUIRE(workspace_bytes >= mg_masked_reduce_workspace_bytes(n_terms, B, T), "mg_masked_reduce: workspace too small");
  MG_REQUIRE(mg_aligned(workspace, 16), "mg_masked_reduce: workspace must be 16-byte aligned");

  ReduceParams prm;
  memset(&prm, 0, sizeof(prm));
  const int sms = mg_cached_sm_count();
  int max_chunks = 1;
  for (int i = 0; i < n_terms; ++i) {
    const mg_term& tm = terms[i];
    MG_REQUIRE(tm.D >= 1, "mg_masked_reduce: term %d has D=%d", i, tm.D);
    MG_REQUIRE(tm.kind >= MG_RED_SQDIFF && tm.kind <= MG_RED_CE, "mg_masked_reduce: term %d has unknown kind %d", i, tm.kind);
    MG_REQUIRE(tm.a != nullptr || T == 0, "mg_masked_reduce: term %d has a NULL operand", i);
    MG_REQUIRE(tm.result != nullptr && mg_aligned(tm.result, 16), "mg_masked_reduce: term %d needs a 16-byte aligned result record", i);
    const bool discrete = tm.ab_dtype == MG_DT_U8 || tm.kind == MG_RED_EQ;
    if (discrete) {
      MG_REQUIRE(tm.kind == MG_RED_XOR || tm.kind == MG_RED_AND || tm.kind == MG_RED_SUM || tm.kind == MG_RED_EQ,
                 "mg_masked_reduce: term %d: kind %d is not defined for uint8 operands", i, tm.kind);
      MG_REQUIRE(tm.m == nullptr && tm.grad == nullptr, "mg_masked_reduce: term %d: uint8 kinds take no weight / gradient", i);
    } else {
      MG_REQUIRE(tm.kind != MG_RED_XOR && tm.kind != MG_RED_AND, "mg_masked_reduce: term %d: XOR / AND need uint8 operands", i);
    }
    MG_REQUIRE(!kind_has_b(tm.kind) || tm.b != nullptr || T == 0, "mg_masked_reduce: term %d needs a second operand", i);
    if (tm.grad != nullptr) {
      MG_REQUIRE(tm.kind == MG_RED_SQDIFF || tm.kind == MG_RED_ABSDIFF || tm.kind == MG_RED_BCE || tm.kind == MG_RED_CE || tm.kind == MG_RED_SUM,
                 "mg_masked_reduce: term %d: kind %d has no gradient", i, tm.kind);
      MG_REQUIRE(tm.m == nullptr, "mg_masked_reduce: term %d: weighted terms have no gradient", i);
    }
    if (tm.kind == MG_RED_CE) {
      MG_REQUIRE(tm.m == nullptr && tm.ab_dtype == MG_DT_F32, "mg_masked_reduce: term %d: cross-entropy takes float32 logits and no weight", i);
    }
    prm.terms[i] = tm;
    const int rows = rows_per_cta_for(tm.D, B, T, sms);
    MgFinishSlot& sl = prm.slots[i];
    sl.result = tm.result;
    sl.D = tm.kind == MG_RED_CE ? 1 : tm.D;   // the cross-entropy collapses the class axis: its loss has one feature
    sl.rows_per_cta = rows;
    sl.n_chunks = T > 0 ? static_cast<int>((T + rows - 1) / rows) : 1;
    sl.per_frame = !discrete && (tm.m != nullptr || tm.kind == MG_RED_ROOT_SQDIFF || tm.kind == MG_RED_CE);
    sl.weighted = !discrete && tm.m != nullptr;
    sl.accumulate = tm.accumulate;
    sl.in_total = (tm.flags & MG_FLAG_IN_TOTAL) != 0;
    sl.weight = tm.grad_scale;
    if (sl.n_chunks > max_chunks) max_chunks = sl.n_chunks;
  }
  prm.seq_len = seq_len;
  prm.ws = mg_carve_workspace(workspace, n_terms, B);
  prm.T = T;
  prm.n_terms = n_terms;
  prm.B = B;
  prm.slice_mode = 1;
  { const char* e = getenv("MG_RED_SLICE_MODE"); if (e) prm.slice_mode = atoi(e); }

  dim3 grid(static_cast<unsigned>(max_chunks), static_cast<unsigned>(B), static_cast<unsigned>(n_terms));
  MG_CUDA_OK(mg_launch_pdl(masked_reduce_kernel, grid, dim3(kRedThreads), 0, stream, prm));
  MG_LAUNCH_OK();
  return MG_OK;
}
