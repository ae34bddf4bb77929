// K7 backward -- the weight gradient of the dense layers on the tcgen05 tensor cores, and the fused activation backward that
// feeds it.
//
// Replaces what autograd runs for nn.Linear (+ nn.Sigmoid) in the example models' training step (reference README.rst:65-73,
// 86-99; models/RNN_SPSS.py:33,38,41; experiment_builder.py:470-479 loss.backward()): SigmoidBackward, the bias gradient's
// column sum and the cuBLAS sgemm `grad_y^T @ x`.
//
//   K7g  act_grad_kernel      g = grad_y * (1 - y) * y   (or g = grad_y), written as bf16 rows padded to 8 columns -- the
//                             operand of both backward GEMMs -- and the bias gradient sum_m g[m, :] accumulated from the fp32
//                             values in the same pass (fp64 per-thread partials, fixed-order finish).  HBM-bound: every
//                             grad_y / y element is read once, 2 bytes written.
//   K7w  wgrad_tcgen05_kernel dW[n, k] = sum_m g[m, n] * x[m, k].  The reduction runs over FRAMES, so both operands are
//                             "MN-major" in tensor-core terms: a TMA box of 128 frames x 64 features (128-byte swizzle) IS
//                             the canonical MN-major tile -- 8-frame groups 1024 bytes apart (stride byte offset), 64-feature
//                             atoms one box apart (leading byte offset) -- so neither tensor is transposed in memory.
//                             The output is tiny (N x K <= 512 x 640) and the reduction is ~10^5 long: the frames are split
//                             over the SMs, each CTA owns one 128 x BLOCK_K accumulator in tensor memory for its slice, and
//                             the slices are summed in a fixed order by a second, small kernel (deterministic).
#include <stdlib.h>
#include <string.h>

#include "mg_common.cuh"
#include "mg_tcgen05.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------------------
// K7g: activation backward + bf16 cast + bias gradient
// ------------------------------------------------------------------------------------------------------------------
constexpr int kActThreads = 256;
constexpr int kActMinRowsPerCta = 128; // rows of one CTA at least; the grid is one wave of 3 CTAs per SM when there are enough rows

struct ActGradParams {
  const void* grad_y;     // (M, N) fp32 or bf16, row stride ldg
  const void* y;          // (M, N) fp32 or bf16 forward output (sigmoid), or NULL: g = grad_y
  __nv_bfloat16* out;     // (M, ld_out) bf16, columns >= N zero
  double* partial;        // (n_ctas, ld_out) column sums of this CTA's rows, or NULL
  int64_t ldg, ldy, ld_out, M, rows_per_cta;
  int N, grad_is_bf16, y_is_bf16;
};

__device__ __forceinline__ float load_elem(const void* base, int64_t idx, int is_bf16) {
  return is_bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(base)[idx]) : __ldcs(static_cast<const float*>(base) + idx);
}

// Thread t owns the 8-column group t % G of rows t / G, t / G + R, ... (G = ld_out / 8 groups per row, R = rows per pass):
// always the same columns, so the bias gradient is 8 per-thread accumulators.
// 8 consecutive elements of a row as floats: two 16-byte loads of fp32 or one of bf16 (read-once data: no L1 allocation).
__device__ __forceinline__ void load8(const void* base, int64_t idx, int is_bf16, float (&v)[8]) {
  if (is_bf16) {
    uint4 raw;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w)
                 : "l"(static_cast<const __nv_bfloat16*>(base) + idx));
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {   // bf16 -> fp32 is a 16-bit shift
      v[2 * j] = __uint_as_float(w[j] << 16);
      v[2 * j + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
    }
  } else {
    const float4* p = reinterpret_cast<const float4*>(static_cast<const float*>(base) + idx);
    const float4 a = mg_ld_stream_f4(p), b = mg_ld_stream_f4(p + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}

template <bool VEC>
__device__ __forceinline__ void act_grad_load(const ActGradParams& prm, int64_t r, int c0, float (&g)[8], float (&yv)[8]) {
  if (VEC) {   // rows 16-byte aligned in both operands, N a multiple of 8
    load8(prm.grad_y, r * prm.ldg + c0, prm.grad_is_bf16, g);
    if (prm.y != nullptr) load8(prm.y, r * prm.ldy + c0, prm.y_is_bf16, yv);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      g[j] = 0.f;
      yv[j] = 0.f;
      if (c0 + j < prm.N) {
        g[j] = load_elem(prm.grad_y, r * prm.ldg + c0 + j, prm.grad_is_bf16);
        if (prm.y != nullptr) yv[j] = load_elem(prm.y, r * prm.ldy + c0 + j, prm.y_is_bf16);
      }
    }
  }
}

__device__ __forceinline__ void act_grad_finish_row(const ActGradParams& prm, int64_t r, int c0, float (&g)[8], const float (&yv)[8],
                                                    double (&acc)[8]) {
  if (prm.y != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = __fmul_rn(__fmul_rn(g[j], __fsub_rn(1.f, yv[j])), yv[j]);   // ATen sigmoid_backward: grad * (1 - y) * y
  }
  __align__(16) __nv_bfloat16 v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { v[j] = __float2bfloat16_rn(g[j]); acc[j] += static_cast<double>(g[j]); }
  *reinterpret_cast<uint4*>(prm.out + r * prm.ld_out + c0) = *reinterpret_cast<const uint4*>(v);
}

template <bool VEC>
__global__ void __launch_bounds__(kActThreads, 3)
act_grad_kernel(const ActGradParams prm) {
  __shared__ double s_sum[kActThreads][9];   // padded: the finish reads a column of it
  mg_pdl_wait();                 // programmatic dependent launch (mg_common.cuh): nothing above touches global memory
  mg_pdl_launch_dependents();
  const int G = static_cast<int>(prm.ld_out / 8);
  const int R = kActThreads / G;             // G <= 256 is checked by the host
  const int tid = threadIdx.x;
  const int grp = tid % G, lane_row = tid / G;
  const bool active = lane_row < R;
  const int c0 = grp * 8;
  const int64_t r_begin = static_cast<int64_t>(blockIdx.x) * prm.rows_per_cta;
  const int64_t r_end = min(prm.M, r_begin + prm.rows_per_cta);
  double acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.;
  if (active) {
    int64_t r = r_begin + lane_row;
    for (; r + R < r_end; r += 2 * R) {       // two rows in flight per thread (128 bytes of loads with a sigmoid)
      float g0[8], y0[8], g1[8], y1[8];
      act_grad_load<VEC>(prm, r, c0, g0, y0);
      act_grad_load<VEC>(prm, r + R, c0, g1, y1);
      act_grad_finish_row(prm, r, c0, g0, y0, acc);
      act_grad_finish_row(prm, r + R, c0, g1, y1, acc);
    }
    if (r < r_end) {
      float g0[8], y0[8];
      act_grad_load<VEC>(prm, r, c0, g0, y0);
      act_grad_finish_row(prm, r, c0, g0, y0, acc);
    }
  }
  if (prm.partial == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) s_sum[tid][j] = acc[j];
  __syncthreads();
  // column c of this CTA = sum over the R row-lanes that own group c / 8, in lane order
  for (int c = tid; c < prm.ld_out; c += kActThreads) {
    double s = 0.;
    for (int l = 0; l < R; ++l) s += s_sum[l * G + c / 8][c % 8];
    prm.partial[static_cast<int64_t>(blockIdx.x) * prm.ld_out + c] = s;
  }
}

// bias_grad[c] = sum of the per-CTA column sums (fp64), rounded once.  Block = 32 columns x 8 lanes: lane l adds CTAs l, l + 8, ...
// in order (a warp reads 32 consecutive doubles), the 8 lane sums are added in lane order.
__global__ void __launch_bounds__(256)
bias_grad_finish_kernel(const double* __restrict__ partial, int n_ctas, int64_t ld, int N, float* __restrict__ bias_grad) {
  __shared__ double s_part[8][33];
  mg_pdl_wait();                 // programmatic dependent launch (mg_common.cuh): nothing above touches global memory
  mg_pdl_launch_dependents();
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), l = threadIdx.x >> 5;
  double s = 0.;
  if (col < N)
    for (int i = l; i < n_ctas; i += 8) s += partial[static_cast<int64_t>(i) * ld + col];
  s_part[l][threadIdx.x & 31] = s;
  __syncthreads();
  if (l == 0 && col < N) {
    double t = 0.;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += s_part[j][threadIdx.x];
    bias_grad[col] = static_cast<float>(t);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K7w: weight gradient
// ------------------------------------------------------------------------------------------------------------------
constexpr int kWgTileN = 128;             // output rows (out_features) of one accumulator = TMEM lanes
constexpr int kWgFrames = 128;            // frames per shared-memory stage (8 MMAs of 16 per barrier round trip; 64: 0.231 vs 0.214 ms at 512 x 600); MG_WGRAD_FRAMES=64
constexpr int kWgAtom = 64;               // features per TMA box / swizzle atom (128 bytes of bf16)
constexpr int kWgMaxStages = 6;
constexpr uint32_t kWgRingBytes = 192 * 1024;
constexpr int kWgThreads = 192;           // TMA producer, MMA issuer, 4 epilogue warps
constexpr size_t kWgSmem = kWgRingBytes + 1024;
constexpr int kWgTmemCols = 256;

struct WgradParams {
  float* partial;          // (splits, n_tiles * 128, k_tiles * tile_k) fp32
  int64_t M;
  int N, K, tile_k, n_tiles, k_tiles, splits, blocks_per_split, n_fblocks, n_stages, frames;
  uint32_t stage_bytes, box_bytes;   // box = frames x 64 features of bf16
};

// MN-major operand tile, 128-byte swizzle: 64-feature atoms one box apart (leading byte offset), 8-frame groups 1024 bytes
// apart (stride byte offset).  cute/atom/mma_traits_sm100.hpp: ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units.
__device__ __forceinline__ uint64_t umma_smem_desc_mn(uint32_t smem_addr, uint32_t box_bytes) {
  uint64_t desc = 0;
  desc |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  desc |= static_cast<uint64_t>(box_bytes >> 4) << 16;
  desc |= static_cast<uint64_t>(1024 >> 4) << 32;
  desc |= static_cast<uint64_t>(1) << 46;
  desc |= static_cast<uint64_t>(2) << 61;
  return desc;
}

// PAIR: the two CTAs of a cluster (one TPC) share a 256 x tile_k accumulator: each loads its own 128 out_features of g and HALF
// of the x tile, the leader issues tcgen05.mma.cta_group::2 (M = 256), each CTA's tensor memory receives its own 128 rows.
// Per output element a third less crosses L2 -> shared memory, which is what bounds the single-CTA form at 600 -> 512
// (12 tiles x 48 KB per 64 frames = 3.1 GB in 0.255 ms: the chip's ~12 TB/s L2 limit).
template <bool PAIR>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap map_g, const __grid_constant__ CUtensorMap map_x,
                     const __grid_constant__ WgradParams prm) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t s_full[kWgMaxStages], s_empty[kWgMaxStages], s_acc_full;
  __shared__ uint32_t s_tmem_base;
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta_rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
  const bool leader = cta_rank == 0;
  const int unit = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);   // (tile, frame slice) of this CTA / pair
  const int tile = unit % (prm.n_tiles * prm.k_tiles), split = unit / (prm.n_tiles * prm.k_tiles);
  constexpr int kRows = PAIR ? 2 * kWgTileN : kWgTileN;        // out_features of one (pair-)tile
  const int n0 = (tile / prm.k_tiles) * kRows + cta_rank * kWgTileN, k0 = (tile % prm.k_tiles) * prm.tile_k;
  const int fb_begin = split * prm.blocks_per_split;
  const int fb_end = min(prm.n_fblocks, fb_begin + prm.blocks_per_split);
  const int n_blocks = fb_end - fb_begin;       // >= 1 by construction of the grid
  const int kStages = prm.n_stages;
  const int k_boxes = prm.tile_k / kWgAtom / (PAIR ? 2 : 1);   // boxes of x this CTA loads per stage
  const int k_load0 = k0 + cta_rank * k_boxes * kWgAtom;
  const uint32_t kBox = prm.box_bytes;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    for (int s = 0; s < kStages; ++s) { mg_mbar_init(&s_full[s], 1); mg_mbar_init(&s_empty[s], 1); }
    mg_mbar_init(&s_acc_full, 1);
    mg_mbar_fence_init();
  }
  if (warp == 2) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(mg_smem_addr(&s_tmem_base)), "n"(kWgTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(mg_smem_addr(&s_tmem_base)), "n"(kWgTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  mg_pdl_wait();                 // programmatic dependent launch (mg_common.cuh): nothing above touches global memory
  mg_pdl_launch_dependents();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (PAIR) cluster_sync_all(); else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    if (lane == 0) {   // ===== TMA producer: per stage 2 boxes of g (this CTA's 128 out_features) and its boxes of x
      const uint32_t stage_tx = static_cast<uint32_t>(2 + k_boxes) * kBox;
      for (int it = 0; it < n_blocks; ++it) {
        const int s = it % kStages;
        if (it >= kStages) mg_mbar_wait(&s_empty[s], static_cast<uint32_t>(((it / kStages) - 1) & 1));
        unsigned char* a_tile = ring + static_cast<size_t>(s) * prm.stage_bytes;
        const int f0 = (fb_begin + it) * prm.frames;
        if (PAIR) {
          if (leader) mg_mbar_expect_tx(&s_full[s], 2 * stage_tx);     // both CTAs' bytes are counted on the leader's barrier
          tma_load_2d_pair(a_tile, &map_g, n0, f0, &s_full[s]);
          tma_load_2d_pair(a_tile + kBox, &map_g, n0 + kWgAtom, f0, &s_full[s]);
          for (int j = 0; j < k_boxes; ++j)
            tma_load_2d_pair(a_tile + (2 + j) * kBox, &map_x, k_load0 + j * kWgAtom, f0, &s_full[s]);
        } else {
          mg_mbar_expect_tx(&s_full[s], stage_tx);
          tma_load_2d(a_tile, &map_g, n0, f0, &s_full[s]);
          tma_load_2d(a_tile + kBox, &map_g, n0 + kWgAtom, f0, &s_full[s]);
          for (int j = 0; j < k_boxes; ++j)
            tma_load_2d(a_tile + (2 + j) * kBox, &map_x, k_load0 + j * kWgAtom, f0, &s_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {   // ===== MMA issuer: D[rows x tile_k] += g_tile^T (MN-major A) * x_tile (MN-major B), 16 frames per MMA
      uint32_t idesc = umma_instr_desc(prm.tile_k, kRows);
      idesc |= (1u << 15) | (1u << 16);         // A and B are MN-major
      for (int it = 0; it < n_blocks; ++it) {
        const int s = it % kStages;
        mg_mbar_wait(&s_full[s], static_cast<uint32_t>((it / kStages) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = mg_smem_addr(ring + static_cast<size_t>(s) * prm.stage_bytes);
        const uint32_t b_addr = a_addr + 2 * kBox;
        for (int k = 0; k < prm.frames / kUmmaK; ++k) {   // 16 frames = two 8-frame groups = 2048 bytes further into every atom
          if (PAIR) umma_f16_pair(tmem_base, umma_smem_desc_mn(a_addr + k * 2048, kBox), umma_smem_desc_mn(b_addr + k * 2048, kBox), idesc, (it | k) != 0 ? 1u : 0u);
          else umma_f16(tmem_base, umma_smem_desc_mn(a_addr + k * 2048, kBox), umma_smem_desc_mn(b_addr + k * 2048, kBox), idesc, (it | k) != 0 ? 1u : 0u);
        }
        if (PAIR) umma_commit_pair(&s_empty[s]); else umma_commit(&s_empty[s]);
      }
      if (PAIR) umma_commit_pair(&s_acc_full); else umma_commit(&s_acc_full);
    }
  } else {
    // ===== epilogue: warp w reads TMEM lanes 32 * (w % 4) ..; lane = out_feature row, 32 in_feature columns per load
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    mg_mbar_wait(&s_acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int64_t ld = static_cast<int64_t>(prm.k_tiles) * prm.tile_k;
    float* dst_row = prm.partial + (static_cast<int64_t>(split) * prm.n_tiles * kRows + n0 + row) * ld + k0;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    for (int c0 = 0; c0 < prm.tile_k; c0 += 32) {
      uint32_t a[32];
      tmem_ld_32x32(taddr + static_cast<uint32_t>(c0), a);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<uint4*>(dst_row + c0 + 4 * j) = make_uint4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kWgTmemCols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kWgTmemCols) : "memory");
  }
}

// dW[n, k] = sum over frame slices in slice order (fp32, as the tensor cores accumulate).  One thread per 4 consecutive k
// (the slices' rows are 16-byte aligned; the output row only when ldw % 4 == 0), four slices' loads in flight, added in order.
__global__ void __launch_bounds__(256)
wgrad_finish_kernel(const float* __restrict__ partial, int splits, int64_t slice_elems, int64_t ld, int N, int K,
                    float* __restrict__ grad_w, int64_t ldw, int vec_out) {
  mg_pdl_wait();                 // programmatic dependent launch (mg_common.cuh): nothing above touches global memory
  mg_pdl_launch_dependents();
  const int groups = (K + 3) / 4;
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= static_cast<int64_t>(N) * groups) return;
  const int n = static_cast<int>(idx / groups), k = static_cast<int>(idx - static_cast<int64_t>(n) * groups) * 4;
  const float4* p = reinterpret_cast<const float4*>(partial + static_cast<int64_t>(n) * ld + k);   // ld >= round_up(K, 64)
  const int64_t step = slice_elems / 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  int i = 0;
  for (; i + 4 <= splits; i += 4) {
    const float4 a = __ldcg(p + i * step), b = __ldcg(p + (i + 1) * step), c = __ldcg(p + (i + 2) * step), d = __ldcg(p + (i + 3) * step);
    s.x = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s.x, a.x), b.x), c.x), d.x);
    s.y = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s.y, a.y), b.y), c.y), d.y);
    s.z = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s.z, a.z), b.z), c.z), d.z);
    s.w = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(s.w, a.w), b.w), c.w), d.w);
  }
  for (; i < splits; ++i) {
    const float4 a = __ldcg(p + i * step);
    s.x = __fadd_rn(s.x, a.x); s.y = __fadd_rn(s.y, a.y); s.z = __fadd_rn(s.z, a.z); s.w = __fadd_rn(s.w, a.w);
  }
  float* out = grad_w + static_cast<int64_t>(n) * ldw + k;
  if (vec_out && k + 4 <= K) {
    *reinterpret_cast<float4*>(out) = s;
  } else {
    const float v[4] = {s.x, s.y, s.z, s.w};
    for (int j = 0; j < 4 && k + j < K; ++j) out[j] = v[j];
  }
}

// The same sum for many slices (small layers: one or two tiles, up to #SMs slices): a warp per output group, lane l adds slices
// l, l + 32, ... in order, then a fixed shuffle tree -- the chain per thread is splits / 32 loads instead of splits.
__global__ void __launch_bounds__(256)
wgrad_finish_warp_kernel(const float* __restrict__ partial, int splits, int64_t slice_elems, int64_t ld, int N, int K,
                         float* __restrict__ grad_w, int64_t ldw, int vec_out) {
  mg_pdl_wait();                 // programmatic dependent launch (mg_common.cuh): nothing above touches global memory
  mg_pdl_launch_dependents();
  const int groups = (K + 3) / 4;
  const int64_t idx = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (idx >= static_cast<int64_t>(N) * groups) return;      // whole warps leave together
  const int n = static_cast<int>(idx / groups), k = static_cast<int>(idx - static_cast<int64_t>(n) * groups) * 4;
  const float4* p = reinterpret_cast<const float4*>(partial + static_cast<int64_t>(n) * ld + k);
  const int64_t step = slice_elems / 4;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = lane; i < splits; i += 32) {
    const float4 a = __ldcg(p + i * step);
    s.x = __fadd_rn(s.x, a.x); s.y = __fadd_rn(s.y, a.y); s.z = __fadd_rn(s.z, a.z); s.w = __fadd_rn(s.w, a.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s.x = __fadd_rn(s.x, __shfl_xor_sync(MG_FULL_MASK, s.x, o));
    s.y = __fadd_rn(s.y, __shfl_xor_sync(MG_FULL_MASK, s.y, o));
    s.z = __fadd_rn(s.z, __shfl_xor_sync(MG_FULL_MASK, s.z, o));
    s.w = __fadd_rn(s.w, __shfl_xor_sync(MG_FULL_MASK, s.w, o));
  }
  if (lane != 0) return;
  float* out = grad_w + static_cast<int64_t>(n) * ldw + k;
  if (vec_out && k + 4 <= K) {
    *reinterpret_cast<float4*>(out) = s;
  } else {
    const float v[4] = {s.x, s.y, s.z, s.w};
    for (int j = 0; j < 4 && k + j < K; ++j) out[j] = v[j];
  }
}

struct WgradPlan {
  int tile_k, n_tiles, k_tiles, splits, blocks_per_split, n_fblocks, tile_rows, frames;
  bool pair;
  int64_t partial_elems;
};

WgradPlan plan_wgrad(int64_t M, int N, int K) {
  WgradPlan p;
  p.tile_k = K <= 64 ? 64 : (K <= 128 ? 128 : (K <= 192 ? 192 : 256));
  { const char* e = getenv("MG_WGRAD_TILE_K"); if (e && (atoi(e) == 64 || atoi(e) == 128 || atoi(e) == 192 || atoi(e) == 256) && atoi(e) < p.tile_k) p.tile_k = atoi(e); }
  // CTA pairs (256-row tiles) where there are at least two 128-row tiles to pair and the x tile splits into whole atoms
  const int sms = mg_cached_sm_count();
  p.pair = N > kWgTileN && p.tile_k % (2 * kWgAtom) == 0 && sms % 2 == 0;
  { const char* e = getenv("MG_WGRAD_PAIR"); if (e) p.pair = atoi(e) != 0 && p.tile_k % (2 * kWgAtom) == 0 && sms % 2 == 0; }
  p.tile_rows = p.pair ? 2 * kWgTileN : kWgTileN;
  p.n_tiles = (N + p.tile_rows - 1) / p.tile_rows;
  p.k_tiles = (K + p.tile_k - 1) / p.tile_k;
  p.frames = kWgFrames;
  { const char* e = getenv("MG_WGRAD_FRAMES"); if (e && atoi(e) == 64) p.frames = 64; }
  p.n_fblocks = static_cast<int>((M + p.frames - 1) / p.frames);
  const int tiles = p.n_tiles * p.k_tiles;
  int splits = (p.pair ? sms / 2 : sms) / tiles;
  if (splits < 1) splits = 1;
  if (splits > p.n_fblocks) splits = p.n_fblocks > 0 ? p.n_fblocks : 1;
  p.blocks_per_split = p.n_fblocks > 0 ? (p.n_fblocks + splits - 1) / splits : 1;
  p.splits = p.n_fblocks > 0 ? (p.n_fblocks + p.blocks_per_split - 1) / p.blocks_per_split : 1;
  p.partial_elems = static_cast<int64_t>(p.splits) * p.n_tiles * p.tile_rows * p.k_tiles * p.tile_k;
  return p;
}

// One resident wave: 3 CTAs per SM (<= 85 registers x 256 threads, two rows in flight per thread), each with an equal share of the rows.
int64_t act_grad_rows_per_cta(int64_t M) {
  const int64_t slots = static_cast<int64_t>(mg_cached_sm_count()) * 3;
  int64_t rows = (M + slots - 1) / slots;
  return rows < kActMinRowsPerCta ? kActMinRowsPerCta : rows;
}
int act_grad_ctas(int64_t M) { const int64_t rows = act_grad_rows_per_cta(M); return static_cast<int>((M + rows - 1) / rows); }

}  // namespace

extern "C" int64_t mg_act_grad_workspace_bytes(int64_t M, int N) {
  if (M <= 0 || N <= 0) return 0;
  const int64_t ld = (static_cast<int64_t>(N) + 7) / 8 * 8;
  return static_cast<int64_t>(act_grad_ctas(M)) * ld * static_cast<int64_t>(sizeof(double));
}

extern "C" int mg_act_grad_bf16(const void* grad_y, int grad_is_bf16, int64_t ldg, const void* y, int y_is_bf16, int64_t ldy,
                                void* out, int64_t ld_out, float* bias_grad, int64_t M, int N, void* workspace,
                                int64_t workspace_bytes, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(M >= 0 && N >= 1, "mg_act_grad_bf16: bad shape (M=%lld, N=%d)", static_cast<long long>(M), N);
  MG_REQUIRE(ld_out == (static_cast<int64_t>(N) + 7) / 8 * 8, "mg_act_grad_bf16: ld_out must be N rounded up to 8 (got %lld)",
             static_cast<long long>(ld_out));
  MG_REQUIRE(ld_out / 8 <= kActThreads, "mg_act_grad_bf16: N=%d exceeds %d features", N, kActThreads * 8);
  MG_REQUIRE(ldg >= N && (y == nullptr || ldy >= N), "mg_act_grad_bf16: row strides must cover the row");
  if (M == 0) {
    if (bias_grad != nullptr) MG_CUDA_OK(cudaMemsetAsync(bias_grad, 0, sizeof(float) * N, stream));
    return MG_OK;
  }
  MG_REQUIRE(grad_y != nullptr && out != nullptr && mg_aligned(out, 16), "mg_act_grad_bf16: NULL or misaligned buffer");
  const int n_ctas = act_grad_ctas(M);
  if (bias_grad != nullptr)
    MG_REQUIRE(workspace != nullptr && workspace_bytes >= mg_act_grad_workspace_bytes(M, N),
               "mg_act_grad_bf16: workspace of %lld bytes needed", static_cast<long long>(mg_act_grad_workspace_bytes(M, N)));
  ActGradParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.grad_y = grad_y; prm.y = y; prm.out = static_cast<__nv_bfloat16*>(out);
  prm.partial = bias_grad != nullptr ? static_cast<double*>(workspace) : nullptr;
  prm.ldg = ldg; prm.ldy = ldy; prm.ld_out = ld_out; prm.M = M; prm.N = N; prm.rows_per_cta = act_grad_rows_per_cta(M);
  prm.grad_is_bf16 = grad_is_bf16; prm.y_is_bf16 = y_is_bf16;
  // 16-byte row segments in both operands: strides of 4 fp32 / 8 bf16 elements
  const bool vec = N % 8 == 0 && ldg % (grad_is_bf16 ? 8 : 4) == 0 && mg_aligned(grad_y, 16) &&
                   (y == nullptr || (ldy % (y_is_bf16 ? 8 : 4) == 0 && mg_aligned(y, 16)));
  if (vec) MG_CUDA_OK(mg_launch_pdl(act_grad_kernel<true>, dim3(n_ctas), dim3(kActThreads), 0, stream, prm));
  else MG_CUDA_OK(mg_launch_pdl(act_grad_kernel<false>, dim3(n_ctas), dim3(kActThreads), 0, stream, prm));
  MG_LAUNCH_OK();
  if (bias_grad != nullptr) {
    MG_CUDA_OK(mg_launch_pdl(bias_grad_finish_kernel, dim3((N + 31) / 32), dim3(256), 0, stream, static_cast<const double*>(prm.partial), static_cast<int>(n_ctas), ld_out, N, bias_grad));
    MG_LAUNCH_OK();
  }
  return MG_OK;
}

extern "C" int64_t mg_linear_wgrad_workspace_bytes(int64_t M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  return plan_wgrad(M, N, K).partial_elems * static_cast<int64_t>(sizeof(float));
}

// Host-only: the launch plan of mg_linear_wgrad_bf16 as 8 integers (tile_k, n_tiles, k_tiles, splits, blocks_per_split, n_fblocks,
// tile_rows, pair) -- lets the CPU test-suite check its invariants (every frame slice non-empty, workspace size) over many shapes.
extern "C" int mg_linear_wgrad_plan(int64_t M, int N, int K, int64_t* out8) {
  MG_REQUIRE(M >= 1 && N >= 1 && K >= 1 && out8 != nullptr, "mg_linear_wgrad_plan: bad argument");
  const WgradPlan p = plan_wgrad(M, N, K);
  out8[0] = p.tile_k; out8[1] = p.n_tiles; out8[2] = p.k_tiles; out8[3] = p.splits; out8[4] = p.blocks_per_split;
  out8[5] = p.n_fblocks; out8[6] = p.tile_rows; out8[7] = p.pair ? 1 : 0;
  return MG_OK;
}

extern "C" int mg_linear_wgrad_bf16(const void* g, int64_t ldg, const void* x, int64_t ldx, float* grad_w, int64_t ldw,
                                    int64_t M, int N, int K, void* workspace, int64_t workspace_bytes, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(M >= 0 && N >= 1 && K >= 1, "mg_linear_wgrad_bf16: bad shape (M=%lld, N=%d, K=%d)", static_cast<long long>(M), N, K);
  MG_REQUIRE(M < (int64_t(1) << 31) - 128, "mg_linear_wgrad_bf16: too many frames");
  MG_REQUIRE(grad_w != nullptr && ldw >= K, "mg_linear_wgrad_bf16: bad output");
  if (M == 0) {
    MG_CUDA_OK(cudaMemset2DAsync(grad_w, ldw * sizeof(float), 0, K * sizeof(float), N, stream));
    return MG_OK;
  }
  MG_REQUIRE(g != nullptr && x != nullptr, "mg_linear_wgrad_bf16: NULL buffer");
  MG_REQUIRE(ldg % 8 == 0 && ldx % 8 == 0 && ldg >= N && ldx >= K && mg_aligned(g, 16) && mg_aligned(x, 16),
             "mg_linear_wgrad_bf16: operand rows must be 16-byte aligned and cover the row (ldg=%lld, ldx=%lld)",
             static_cast<long long>(ldg), static_cast<long long>(ldx));
  const WgradPlan plan = plan_wgrad(M, N, K);
  MG_REQUIRE(workspace != nullptr && workspace_bytes >= plan.partial_elems * static_cast<int64_t>(sizeof(float)) && mg_aligned(workspace, 16),
             "mg_linear_wgrad_bf16: workspace of %lld bytes needed", static_cast<long long>(plan.partial_elems * sizeof(float)));

  CUtensorMap map_g, map_x;
  int rc = make_map(&map_g, g, M, N, ldg, kWgAtom, plan.frames, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  if (rc != MG_OK) return rc;
  rc = make_map(&map_x, x, M, K, ldx, kWgAtom, plan.frames, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  if (rc != MG_OK) return rc;

  WgradParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.partial = static_cast<float*>(workspace);
  prm.M = M; prm.N = N; prm.K = K; prm.tile_k = plan.tile_k; prm.n_tiles = plan.n_tiles; prm.k_tiles = plan.k_tiles;
  prm.splits = plan.splits; prm.blocks_per_split = plan.blocks_per_split; prm.n_fblocks = plan.n_fblocks;
  prm.frames = plan.frames;
  prm.box_bytes = static_cast<uint32_t>(plan.frames) * kWgAtom * 2;
  prm.stage_bytes = static_cast<uint32_t>(2 + plan.tile_k / kWgAtom / (plan.pair ? 2 : 1)) * prm.box_bytes;     // a multiple of 8 KB
  prm.n_stages = static_cast<int>(kWgRingBytes / prm.stage_bytes);
  if (prm.n_stages > kWgMaxStages) prm.n_stages = kWgMaxStages;

  static bool attr_done[64] = {};
  int device = 0;
  MG_CUDA_OK(cudaGetDevice(&device));
  if (!attr_done[device & 63]) {
    MG_CUDA_OK(cudaFuncSetAttribute(wgrad_tcgen05_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kWgSmem)));
    MG_CUDA_OK(cudaFuncSetAttribute(wgrad_tcgen05_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kWgSmem)));
    attr_done[device & 63] = true;
  }
  const unsigned units = static_cast<unsigned>(plan.splits * plan.n_tiles * plan.k_tiles);
  if (plan.pair) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * units);
    cfg.blockDim = dim3(kWgThreads);
    cfg.dynamicSmemBytes = kWgSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = mg_pdl_enabled() ? 2 : 1;
    MG_CUDA_OK(cudaLaunchKernelEx(&cfg, wgrad_tcgen05_kernel<true>, map_g, map_x, prm));
  } else {
    MG_CUDA_OK(mg_launch_pdl(wgrad_tcgen05_kernel<false>, dim3(units), dim3(kWgThreads), kWgSmem, stream, map_g, map_x, prm));
  }
  MG_LAUNCH_OK();
  const int64_t slice = static_cast<int64_t>(plan.n_tiles) * plan.tile_rows * plan.k_tiles * plan.tile_k;
  const int64_t elems = static_cast<int64_t>(N) * ((K + 3) / 4);
  const int vec_out = (ldw % 4 == 0 && mg_aligned(grad_w, 16)) ? 1 : 0;
  const int64_t ld_partial = static_cast<int64_t>(plan.k_tiles) * plan.tile_k;
  if (plan.splits > 16)
    MG_CUDA_OK(mg_launch_pdl(wgrad_finish_warp_kernel, dim3(static_cast<unsigned>((elems * 32 + 255) / 256)), dim3(256), 0, stream,
                             static_cast<const float*>(prm.partial), plan.splits, slice, ld_partial, N, K, grad_w, ldw, vec_out));
  else
    MG_CUDA_OK(mg_launch_pdl(wgrad_finish_kernel, dim3(static_cast<unsigned>((elems + 255) / 256)), dim3(256), 0, stream,
                             static_cast<const float*>(prm.partial), plan.splits, slice, ld_partial, N, K, grad_w, ldw, vec_out));
  MG_LAUNCH_OK();
  return MG_OK;
}
