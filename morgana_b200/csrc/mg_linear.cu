// K7 -- dense layer on the 5th-generation tensor cores: y = act(x @ w^T + bias), bf16 operands, fp32 accumulation in TMEM.
//
// Replaces nn.Linear (+ nn.Sigmoid) of the example models (reference README.rst:65-73; models/RNN_SPSS.py:33,38,41;
// models/f0_test_model.py:29,41,44), which run as cuBLAS sgemm + separate bias / sigmoid kernels today.
//
// Persistent kernel, one CTA per SM walking 128 x BLOCK_N output tiles, six warps with fixed roles:
//   warp 0  TMA producer   one lane issues cp.async.bulk.tensor (SASS UTMALDG) for the A (128 x 64) and B (BLOCK_N x 64)
//                          bf16 tiles of each K block into a 3-stage ring, 128-byte swizzle, completion on mbarriers;
//                          rows / K columns outside the tensors are zero-filled by the TMA unit, so K = 600 needs no padding
//   warp 1  MMA issuer     one lane issues 4 x tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = BLOCK_N <= 256, K = 16) per K block
//                          (SASS UTCHMMA); tcgen05.commit releases the smem stage and publishes the accumulator
//   warps 2-5 epilogue     tcgen05.ld 32 lanes x 32 columns at a time (SASS LDTM), + bias, optional sigmoid, then a
//                          128-byte-swizzled staging tile in shared memory and one TMA store (SASS UTMASTG) per 128 x 128-byte
//                          chunk (direct stores when the output row stride is not a 16-byte multiple, e.g. N = 187)
// Two accumulators (2 x up to 256 fp32 columns: all 512 TMEM columns) live in tensor memory, so the epilogue of tile i overlaps the loads and MMAs of
// tile i + 1; nothing is kept in registers across K.
//
// At the model's shapes (M = frames ~ 10^5, N <= 512, K <= 640) the layer is HBM-bound: arithmetic intensity
// ~ 2*N*K / (2*K + 4*N) FLOP/B < the ~255 FLOP/B ridge, so the roofline that matters is bytes.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "mg_common.cuh"
#include "mg_tcgen05.cuh"

namespace {

constexpr int kCastThreads = 256;

// fp32 (rows, K) with row stride ldx -> bf16 (rows, ld_out), columns >= K zero-filled.  8 outputs (16 bytes) per thread.
__global__ void __launch_bounds__(kCastThreads)
cast_pad_kernel(const float* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ out, int64_t ld_out, int64_t rows,
                int K, int groups_per_row) {
  mg_pdl_wait();                 // programmatic dependent launch (mg_common.cuh): nothing above touches global memory
  mg_pdl_launch_dependents();
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

// ------------------------------------------------------------------------------------------------------------------
// GEMM
// ------------------------------------------------------------------------------------------------------------------
constexpr int kMaxBlockN = 256;
constexpr int kMaxStages = 6;          // smem ring of (A, B) K-blocks: 3 stages of 48 KB at BLOCK_N = 256 ... 6 of 24 KB at 64
constexpr uint32_t kRingBytes = 144 * 1024;
constexpr int kAccStages = 2;          // accumulators in TMEM: the epilogue of tile i overlaps the MMAs of tile i + 1
constexpr int kTmemCols = kAccStages * kMaxBlockN;   // 512: all of tensor memory (one CTA per SM)
constexpr int kGemmThreads = 320;      // 10 warps: TMA producer, MMA issuer, 2 groups of 4 epilogue warps
constexpr int kEpilogueThreads = 128;
constexpr uint32_t kATileBytes = kBlockM * kBlockK * 2;       // 16 KB
constexpr uint32_t kOutChunkBytes = kBlockM * 128;            // 128 rows x 128 bytes of output (32 fp32 / 64 bf16 columns)
constexpr int kOutBuffers = 2;
constexpr int kMaxN = 2048;
constexpr int kMaxBias = kMaxN + 2 * kMaxBlockN;             // bias staged in shared memory, zero-padded to the widest tile
constexpr int kEpilogueGroups = 2;
constexpr size_t kGemmSmem = kRingBytes + kEpilogueGroups * kOutBuffers * kOutChunkBytes + 1024;   // + slack for 1024-byte alignment

struct GemmParams {
  const float* bias;
  void* y;
  int64_t ldy;
  int M, N, K, block_n, act, y_is_bf16, tma_store;
  int n_stages;            // ring depth for this BLOCK_N
  uint32_t stage_bytes;    // A tile + B tile of one K block (a multiple of 1024)
};

// bias + activation on 32 accumulator columns; `bias32` points at this chunk's 32 biases in shared memory (zero-padded).
// Sigmoid as ex2 + rcp (two MUFU ops): the layer's operands are bf16, so approximate-division accuracy (~1e-7) is ample.
// `coarse`: the result is about to be rounded to bf16 (8 bits of mantissa), so the sigmoid may be 0.5 * tanh(0.5 t) + 0.5 with the
// hardware tanh (ONE MUFU op, relative error 2^-11) instead of ex2 + rcp (two).
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void finish_columns(const uint32_t (&acc)[32], float (&v)[32], const float* bias32, int act, bool coarse = false) {
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float4 b4 = *reinterpret_cast<const float4*>(bias32 + j);
    const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float t = __uint_as_float(acc[j + q]) + b[q];
      if (act == MG_ACT_SIGMOID) t = coarse ? fmaf(0.5f, tanh_approx(0.5f * t), 0.5f) : __fdividef(1.f, 1.f + __expf(-t));
      v[j + q] = t;
    }
  }
}

// Persistent kernel: each CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... (n fastest, so CTAs that run together
// share A tiles through L2).
//
// PAIR: the CTAs of a 2-CTA cluster (one TPC) share one 256 x BLOCK_N tile: each loads its own 128 rows of A and HALF of the
// B tile, the leader issues tcgen05.mma.cta_group::2 (M = 256) reading both halves, and each CTA's tensor memory receives its
// own 128 rows.  Per output element only half of B crosses L2 -> shared memory: at N = 512, K = 600 the single-CTA kernel
// pulls 3.6 GB through L2 for 0.78 GB of HBM traffic and sits at the L2 throughput limit (~6300 B/clk chip-wide).
// MODE 2 (WIDE): a pair whose tile is 256 x 512: the accumulator fills tensor memory (one stage, 512 fp32 columns), every K
// block carries this CTA's halves of two 256-wide B sub-tiles, and the A tile is fetched once for all 512 columns -- another
// 25 % less L2 traffic per output element than 256-wide pair tiles, at the price of not overlapping the epilogue with the
// next tile's MMAs.
template <int MODE>
__global__ void __launch_bounds__(kGemmThreads, 1)
linear_tcgen05_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ CUtensorMap map_y, const __grid_constant__ GemmParams prm) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t s_full[kMaxStages], s_empty[kMaxStages], s_acc_full[kAccStages], s_acc_empty[kAccStages];
  __shared__ uint32_t s_tmem_base;
  __shared__ __align__(16) float s_bias[kMaxBias];   // bias, zero-padded to whole tiles (zeros when there is no bias)

  // 128-byte swizzle wants the tiles on 1024-byte boundaries.
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* out_stage = ring + kRingBytes;
  const int kStages = prm.n_stages;
  const uint32_t kStageBytes = prm.stage_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool PAIR = MODE >= 1, WIDE = MODE == 2;
  constexpr int kAccCount = WIDE ? 1 : kAccStages;              // accumulators in flight
  const int n_tiles = (prm.N + prm.block_n - 1) / prm.block_n;
  constexpr int kTileM = PAIR ? 2 * kBlockM : kBlockM;          // rows of one (pair-)tile
  const int m_tiles = (prm.M + kTileM - 1) / kTileM;
  const int total_tiles = n_tiles * m_tiles;
  const int n_kblocks = (prm.K + kBlockK - 1) / kBlockK;
  const int cta_rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
  const bool leader = cta_rank == 0;
  const int first_tile = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int b_rows = PAIR ? prm.block_n / 2 : prm.block_n;      // rows of W this CTA loads per K block (WIDE: two boxes of 128)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    if (prm.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
    for (int s = 0; s < kStages; ++s) { mg_mbar_init(&s_full[s], 1); mg_mbar_init(&s_empty[s], 1); }
    for (int a = 0; a < kAccStages; ++a) { mg_mbar_init(&s_acc_full[a], 1); mg_mbar_init(&s_acc_empty[a], (PAIR ? 2 : 1) * kEpilogueGroups); }
    mg_mbar_fence_init();
  }
  if (warp == 2) {   // one warp owns the TMEM allocation (and frees it at the end); in a pair, the same warp of both CTAs
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(mg_smem_addr(&s_tmem_base)), "n"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(mg_smem_addr(&s_tmem_base)), "n"(kTmemCols) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  // Programmatic dependent launch: barriers, tensor-map prefetch and the tensor-memory allocation above overlap the previous
  // kernel's tail; the first global read (the bias) comes after the wait.
  mg_pdl_wait();
  mg_pdl_launch_dependents();
  for (int i = threadIdx.x; i < kMaxBias; i += kGemmThreads)
    s_bias[i] = (prm.bias != nullptr && i < prm.N) ? __ldg(prm.bias + i) : 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything is signalled across CTAs
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const uint32_t stage_tx = kATileBytes + static_cast<uint32_t>(b_rows) * kBlockK * 2;   // bytes this CTA loads per stage
      int it = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step) {
        const int m0 = (tile / n_tiles) * kTileM + cta_rank * kBlockM, n0 = (tile % n_tiles) * prm.block_n + cta_rank * b_rows;
        for (int kb = 0; kb < n_kblocks; ++kb, ++it) {
          const int s = it % kStages;
          if (it >= kStages) mg_mbar_wait(&s_empty[s], static_cast<uint32_t>(((it / kStages) - 1) & 1));
          unsigned char* a_tile = ring + static_cast<size_t>(s) * kStageBytes;
          if (PAIR) {
            if (leader) mg_mbar_expect_tx(&s_full[s], 2 * stage_tx);    // both CTAs' bytes land on the leader's barrier
            tma_load_2d_pair(a_tile, &map_x, kb * kBlockK, m0, &s_full[s]);
            if (WIDE) {   // this CTA's 128 rows of each 256-wide sub-tile
              const int w0 = (tile % n_tiles) * prm.block_n + cta_rank * (kMaxBlockN / 2);
              tma_load_2d_pair(a_tile + kATileBytes, &map_w, kb * kBlockK, w0, &s_full[s]);
              tma_load_2d_pair(a_tile + 2 * kATileBytes, &map_w, kb * kBlockK, w0 + kMaxBlockN, &s_full[s]);
            } else {
              tma_load_2d_pair(a_tile + kATileBytes, &map_w, kb * kBlockK, n0, &s_full[s]);
            }
          } else {
            mg_mbar_expect_tx(&s_full[s], stage_tx);
            tma_load_2d(a_tile, &map_x, kb * kBlockK, m0, &s_full[s]);
            tma_load_2d(a_tile + kATileBytes, &map_w, kb * kBlockK, n0, &s_full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0 && leader) {
      const uint32_t idesc = umma_instr_desc(WIDE ? kMaxBlockN : prm.block_n, kTileM);
      int it = 0, t = 0;
      for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++t) {
        const int acc = t % kAccCount;
        if (t >= kAccCount) mg_mbar_wait(&s_acc_empty[acc], static_cast<uint32_t>(((t / kAccCount) - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * kMaxBlockN);
        for (int kb = 0; kb < n_kblocks; ++kb, ++it) {
          const int s = it % kStages;
          mg_mbar_wait(&s_full[s], static_cast<uint32_t>((it / kStages) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = mg_smem_addr(ring + static_cast<size_t>(s) * kStageBytes);
          const uint32_t b_addr = a_addr + kATileBytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // advancing K by 16 elements = 32 bytes inside the 128-byte swizzle row
            if (WIDE) {
              umma_f16_pair(tmem_d, umma_smem_desc(a_addr + k * kUmmaK * 2), umma_smem_desc(b_addr + k * kUmmaK * 2), idesc,
                            (kb | k) != 0 ? 1u : 0u);
              umma_f16_pair(tmem_d + kMaxBlockN, umma_smem_desc(a_addr + k * kUmmaK * 2),
                            umma_smem_desc(b_addr + kATileBytes + k * kUmmaK * 2), idesc, (kb | k) != 0 ? 1u : 0u);
            } else if (PAIR) umma_f16_pair(tmem_d, umma_smem_desc(a_addr + k * kUmmaK * 2), umma_smem_desc(b_addr + k * kUmmaK * 2), idesc,
                                           (kb | k) != 0 ? 1u : 0u);
            else umma_f16(tmem_d, umma_smem_desc(a_addr + k * kUmmaK * 2), umma_smem_desc(b_addr + k * kUmmaK * 2), idesc,
                          (kb | k) != 0 ? 1u : 0u);
          }
          if (PAIR) umma_commit_pair(&s_empty[s]); else umma_commit(&s_empty[s]);   // frees the stage (in both CTAs) when the MMAs that read it are done
        }
        if (PAIR) umma_commit_pair(&s_acc_full[acc]); else umma_commit(&s_acc_full[acc]);   // accumulator of this tile complete
      }
    }
  } else {
    // ===== epilogue: two groups of four warps (warps 2..5 and 6..9); warp w reads TMEM lanes 32 * (w % 4) .. + 31, the
    // groups take alternate column chunks of a tile, each with its own staging buffers, named barrier and store issuer, so
    // one group's tcgen05.ld / MUFU / store latencies are covered by the other's =====
    const int group = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const int tile_row = quarter * 32 + lane;
    const int group_tid = static_cast<int>(threadIdx.x) - 64 - group * kEpilogueThreads;
    const bool issuer = group_tid == 0;               // first thread of the group issues its TMA stores
    const int cols_per_chunk = prm.y_is_bf16 ? 64 : 32;   // 128 bytes of output per row per chunk
    unsigned char* group_stage = out_stage + static_cast<size_t>(group) * kOutBuffers * kOutChunkBytes;
    auto group_barrier = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(kEpilogueThreads) : "memory"); };
    int t = 0, chunk_count = 0;
    auto release_accumulator = [&](int acc_stage) {   // this group's share of this CTA's accumulator rows has been read
      if (PAIR) mbar_arrive_leader(&s_acc_empty[acc_stage]); else mbar_arrive(&s_acc_empty[acc_stage]);
    };
    for (int tile = first_tile; tile < total_tiles; tile += tile_step, ++t) {
      const int m0 = (tile / n_tiles) * kTileM + cta_rank * kBlockM, n0 = (tile % n_tiles) * prm.block_n;
      const int acc = t % kAccCount;
      mg_mbar_wait(&s_acc_full[acc], static_cast<uint32_t>((t / kAccCount) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * kMaxBlockN);
      const int step = prm.tma_store ? cols_per_chunk : 32;           // columns per chunk on this path
      const int n_chunks = (prm.block_n + step - 1) / step;
      const int last_chunk = ((n_chunks - 1 - group) / 2) * 2 + group;   // this group's last chunk (< group: it has none)
      if (n_chunks <= group) {            // nothing to read for this group in such a narrow tile
        if (issuer) release_accumulator(acc);
        continue;
      }

      if (prm.tma_store) {
        // registers -> 128-byte-swizzled staging tile in shared memory -> one TMA store per 128 x (32 | 64) chunk;
        // rows / columns outside the tensor are clipped by the TMA unit.
        for (int ci = group; ci < n_chunks; ci += 2, ++chunk_count) {
          const int c0 = ci * cols_per_chunk;
          unsigned char* stage = group_stage + static_cast<size_t>(chunk_count % kOutBuffers) * kOutChunkBytes;
          if (chunk_count >= kOutBuffers) {   // the store that last read this buffer must have drained
            if (issuer) mg_bulk_wait_read<kOutBuffers - 1>();
            group_barrier();
          }
          uint4* dst_row = reinterpret_cast<uint4*>(stage + tile_row * 128);
          uint32_t a0[32], a1[32];
          float v[32];
          const bool second_half = prm.y_is_bf16 && c0 + 32 < prm.block_n;
          tmem_ld_32x32(taddr + static_cast<uint32_t>(c0), a0);
          if (second_half) tmem_ld_32x32(taddr + static_cast<uint32_t>(c0 + 32), a1);   // both loads in flight, one wait
          tmem_ld_wait();
          finish_columns(a0, v, s_bias + n0 + c0, prm.act, prm.y_is_bf16 != 0);
          if (!prm.y_is_bf16) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              dst_row[j ^ (tile_row & 7)] = make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                                       __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
          } else {
            float v1[32];
            if (second_half) {
              finish_columns(a1, v1, s_bias + n0 + c0 + 32, prm.act, true);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v1[j] = 0.f;
            }
            auto pack = [](float lo, float hi) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
              return *reinterpret_cast<const uint32_t*>(&h);
            };
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              dst_row[j ^ (tile_row & 7)] = make_uint4(pack(v[8 * j], v[8 * j + 1]), pack(v[8 * j + 2], v[8 * j + 3]),
                                                       pack(v[8 * j + 4], v[8 * j + 5]), pack(v[8 * j + 6], v[8 * j + 7]));
              dst_row[(j + 4) ^ (tile_row & 7)] = make_uint4(pack(v1[8 * j], v1[8 * j + 1]), pack(v1[8 * j + 2], v1[8 * j + 3]),
                                                             pack(v1[8 * j + 4], v1[8 * j + 5]), pack(v1[8 * j + 6], v1[8 * j + 7]));
            }
          }
          if (ci == last_chunk) {   // this group's last TMEM read of the tile: hand its share of the accumulator back
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          }
          mg_fence_proxy_async_smem();
          group_barrier();
          if (issuer) {
            if (ci == last_chunk) release_accumulator(acc);
            tma_store_2d(&map_y, n0 + c0, m0, stage);
            mg_bulk_commit();
          }
        }
      } else {
        // Output rows that are not 16-byte multiples (N = 187, 199, 1, ...) cannot go through TMA: transpose each
        // 128 x 32 chunk through a padded shared-memory tile so that a warp writes 32 consecutive columns of one row.
        float* tile_f = reinterpret_cast<float*>(group_stage);     // [128][33]
        for (int ci = group; ci < n_chunks; ci += 2) {
          const int c0 = ci * 32;
          uint32_t a0[32];
          float v[32];
          tmem_ld_32x32(taddr + static_cast<uint32_t>(c0), a0);
          tmem_ld_wait();
          finish_columns(a0, v, s_bias + n0 + c0, prm.act);
          if (ci == last_chunk) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          if (prm.N <= 8) {
            // a handful of output features (the N = 1 head): lane = row, neighbouring lanes write neighbouring rows
            group_barrier();
            if (issuer && ci == last_chunk) release_accumulator(acc);
            const int row = m0 + tile_row;
            if (row < prm.M) {
              for (int j = 0; j < prm.N - n0 - c0 && j < 32; ++j) {
                const int64_t off = static_cast<int64_t>(row) * prm.ldy + n0 + c0 + j;
                if (prm.y_is_bf16) static_cast<__nv_bfloat16*>(prm.y)[off] = __float2bfloat16_rn(v[j]);
                else static_cast<float*>(prm.y)[off] = v[j];
              }
            }
            continue;
          }
          group_barrier();                                          // the previous chunk has been drained from the tile
#pragma unroll
          for (int j = 0; j < 32; ++j) tile_f[tile_row * 33 + j] = v[j];
          group_barrier();
          if (issuer && ci == last_chunk) release_accumulator(acc);
          const int n_valid = min(32, prm.N - (n0 + c0));
          const int col = group_tid & 31;
          if (col < n_valid) {
            for (int r = group_tid >> 5; r < kBlockM; r += kEpilogueThreads / 32) {
              if (m0 + r >= prm.M) break;
              const float val = tile_f[r * 33 + col];
              const int64_t off = static_cast<int64_t>(m0 + r) * prm.ldy + n0 + c0 + col;
              if (prm.y_is_bf16) static_cast<__nv_bfloat16*>(prm.y)[off] = __float2bfloat16_rn(val);
              else static_cast<float*>(prm.y)[off] = val;
            }
          }
        }
      }
    }
    if (issuer && prm.tma_store) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (PAIR) cluster_sync_all();   // neither CTA frees tensor memory (or exits) while the peer may still use it
  else __syncthreads();
  if (warp == 2) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// W (N, K) fp32 row-major -> W^T (K, ld_out >= N) bf16, columns N.. zero: the B operand of the input-gradient GEMM
// (grad_x = g @ W runs through the forward kernel as g @ (W^T)^T).  32 x 32 tiles through shared memory, both sides coalesced.
__global__ void __launch_bounds__(256)
cast_transpose_kernel(const float* __restrict__ w, int64_t ldw, __nv_bfloat16* __restrict__ out, int64_t ld_out, int N, int K) {
  __shared__ float tile[32][33];
  mg_pdl_wait();                 // programmatic dependent launch (mg_common.cuh): nothing above touches global memory
  mg_pdl_launch_dependents();
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8 threads
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, k = k0 + tx;
    tile[r][tx] = (n < N && k < K) ? __ldg(w + static_cast<int64_t>(n) * ldw + k) : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, n = n0 + tx;
    if (k < K && n < ld_out) out[static_cast<int64_t>(k) * ld_out + n] = __float2bfloat16_rn(tile[tx][r]);
  }
}

}  // namespace

extern "C" int mg_cast_transpose_bf16(const float* w, int64_t ldw, void* out, int64_t ld_out, int N, int K, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(N >= 0 && K >= 0 && ldw >= K && ld_out >= N, "mg_cast_transpose_bf16: bad shape");
  MG_REQUIRE(ld_out % 8 == 0 && mg_aligned(out, 16), "mg_cast_transpose_bf16: output rows must be 16-byte aligned");
  if (K == 0 || ld_out == 0) return MG_OK;
  MG_REQUIRE(out != nullptr && (N == 0 || w != nullptr), "mg_cast_transpose_bf16: NULL buffer");
  dim3 grid(static_cast<unsigned>((K + 31) / 32), static_cast<unsigned>((ld_out + 31) / 32));
  MG_CUDA_OK(mg_launch_pdl(cast_transpose_kernel, grid, dim3(256), 0, stream, w, ldw, static_cast<__nv_bfloat16*>(out), ld_out, N, K));
  MG_LAUNCH_OK();
  return MG_OK;
}

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
  MG_CUDA_OK(mg_launch_pdl(cast_pad_kernel, dim3(static_cast<unsigned>(blocks)), dim3(kCastThreads), 0, stream, x, ldx,
                           static_cast<__nv_bfloat16*>(out), ld_out, rows, K, groups));
  MG_LAUNCH_OK();
  return MG_OK;
}

extern "C" int mg_linear_bf16(const void* x, int64_t ldx, const void* w, int64_t ldw, const float* bias, void* y,
                              int64_t ldy, int y_is_bf16, int M, int N, int K, int act, mg_stream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(M >= 0 && N >= 1 && K >= 1, "mg_linear_bf16: bad shape (M=%d, N=%d, K=%d)", M, N, K);
  MG_REQUIRE(act == MG_ACT_NONE || act == MG_ACT_SIGMOID, "mg_linear_bf16: unknown activation %d", act);
  MG_REQUIRE(N <= kMaxN, "mg_linear_bf16: N=%d exceeds %d output features", N, kMaxN);
  if (M == 0) return MG_OK;
  MG_REQUIRE(x != nullptr && w != nullptr && y != nullptr, "mg_linear_bf16: NULL buffer");
  MG_REQUIRE(ldx % 8 == 0 && ldw % 8 == 0 && ldx >= K && ldw >= K && ldy >= N,
             "mg_linear_bf16: row strides must be multiples of 8 elements (16 bytes) and cover the row (ldx=%lld, ldw=%lld, ldy=%lld)",
             static_cast<long long>(ldx), static_cast<long long>(ldw), static_cast<long long>(ldy));
  MG_REQUIRE(mg_aligned(x, 16) && mg_aligned(w, 16), "mg_linear_bf16: operands must be 16-byte aligned");

  // Widest tile that N fills: a 256-wide tile halves how often an A tile is fetched (the layer is bound by operand
  // delivery, not by MMA time), at the price of idle MMA columns when N is just above 128.
  int block_n = 256;
  if (N <= 16) block_n = 16; else if (N <= 32) block_n = 32; else if (N <= 64) block_n = 64; else if (N <= 128) block_n = 128;
  { const char* e = getenv("MG_GEMM_BLOCK_N"); if (e && atoi(e) > 0 && atoi(e) < block_n) block_n = atoi(e); }
  // CTA pairs (256-row tiles, half of the B tile per CTA) where the B tile dominates the L2 -> shared-memory traffic, i.e.
  // several 256-wide N tiles (N > 256), and there are enough rows to give every pair several tiles.  Measured on B200 at
  // M = 349k: 600 -> 512 0.291 -> 0.260 ms with pairs, but 512 -> 128 0.081 -> 0.095 and 256 -> 187 0.251 -> 0.268 (those
  // layers are HBM-bound, and a pair halves the number of independent tile streams).  MG_GEMM_PAIR=0 / 1 overrides.
  const int64_t sms = mg_cached_sm_count();
  bool pair = N > 256 && sms % 2 == 0 && static_cast<int64_t>(M) >= 2 * kBlockM * sms;
  { const char* e = getenv("MG_GEMM_PAIR"); if (e) pair = atoi(e) != 0 && block_n >= 32 && sms % 2 == 0; }
  // 512-wide pair tiles: opt-in (MG_GEMM_WIDE=1).  They move 25 % fewer bytes from L2 than two 256-wide tiles, but with the
  // accumulator filling tensor memory the epilogue no longer overlaps the next tile's MMAs: 0.306 vs 0.252 ms at 600 -> 512.
  bool wide = false;
  { const char* e = getenv("MG_GEMM_WIDE"); if (e) wide = pair && atoi(e) != 0 && N > kMaxBlockN; }
  if (wide) block_n = 2 * kMaxBlockN;
  CUtensorMap map_x, map_w, map_y;
  int rc = make_map(&map_x, x, M, K, ldx, kBlockK, kBlockM, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  if (rc != MG_OK) return rc;
  rc = make_map(&map_w, w, N, K, ldw, kBlockK, wide ? kMaxBlockN / 2 : (pair ? block_n / 2 : block_n), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
  if (rc != MG_OK) return rc;
  // The epilogue stores through TMA when the output rows are 16-byte multiples; otherwise (N = 187, 1, ...) directly.
  const int y_elem = y_is_bf16 ? 2 : 4;
  const bool tma_store = (ldy * y_elem) % 16 == 0 && mg_aligned(y, 16);
  if (tma_store) {
    rc = make_map(&map_y, y, M, N, ldy, 128 / y_elem, kBlockM,
                  y_is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, y_elem);
    if (rc != MG_OK) return rc;
  } else {
    map_y = map_x;   // unused
  }

  GemmParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.bias = bias; prm.y = y; prm.ldy = ldy;
  prm.M = M; prm.N = N; prm.K = K; prm.block_n = block_n; prm.act = act; prm.y_is_bf16 = y_is_bf16; prm.tma_store = tma_store ? 1 : 0;
  prm.stage_bytes = kATileBytes + static_cast<uint32_t>(pair ? block_n / 2 : block_n) * kBlockK * 2;
  if (prm.stage_bytes % 1024) prm.stage_bytes = (prm.stage_bytes / 1024 + 1) * 1024;
  prm.n_stages = static_cast<int>(kRingBytes / prm.stage_bytes);
  if (prm.n_stages > kMaxStages) prm.n_stages = kMaxStages;

  static bool attr_done[64] = {};   // the attribute is per device (several GPUs in one process are allowed)
  int device = 0;
  MG_CUDA_OK(cudaGetDevice(&device));
  bool& attr_set = attr_done[device & 63];
  if (!attr_set) {
    MG_CUDA_OK(cudaFuncSetAttribute(linear_tcgen05_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmem)));
    MG_CUDA_OK(cudaFuncSetAttribute(linear_tcgen05_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmem)));
    MG_CUDA_OK(cudaFuncSetAttribute(linear_tcgen05_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmem)));
    attr_set = true;
  }
  // n fastest: the CTAs that share an A tile are neighbours in launch order, so A is re-read from L2, not HBM.
  const int tile_m = pair ? 2 * kBlockM : kBlockM;
  const int64_t n_tiles_total = static_cast<int64_t>((N + block_n - 1) / block_n) * ((M + tile_m - 1) / tile_m);
  MG_REQUIRE(n_tiles_total < (int64_t(1) << 31), "mg_linear_bf16: too many tiles");
  if (pair) {
    // persistent: one CTA per SM, launched as 2-CTA clusters (a pair shares a TPC)
    const int64_t n_pairs = n_tiles_total < sms / 2 ? n_tiles_total : sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(2 * n_pairs));
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = kGemmSmem;
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
    if (wide) MG_CUDA_OK(cudaLaunchKernelEx(&cfg, linear_tcgen05_kernel<2>, map_x, map_w, map_y, prm));
    else MG_CUDA_OK(cudaLaunchKernelEx(&cfg, linear_tcgen05_kernel<1>, map_x, map_w, map_y, prm));
  } else {
    const unsigned n_ctas = static_cast<unsigned>(n_tiles_total < sms ? n_tiles_total : sms);   // persistent: one CTA per SM
    MG_CUDA_OK(mg_launch_pdl(linear_tcgen05_kernel<0>, dim3(n_ctas), dim3(kGemmThreads), kGemmSmem, stream, map_x, map_w, map_y, prm));
  }
  MG_LAUNCH_OK();
  return MG_OK;
}
