// K7 -- dense layer on the 5th-generation tensor cores: y = act(x @ w^T + bias), bf16 operands, fp32 accumulation in TMEM.
//
// Replaces nn.Linear (+ nn.Sigmoid) of the example models (reference README.rst:65-73; models/RNN_SPSS.py:33,38,41;
// models/f0_test_model.py:29,41,44), which run as cuBLAS sgemm + separate bias / sigmoid kernels today.
//
// One CTA per 128 x BLOCK_N output tile, six warps with fixed roles:
//   warp 0  TMA producer   one lane issues cp.async.bulk.tensor (SASS UTMALDG) for the A (128 x 64) and B (BLOCK_N x 64)
//                          bf16 tiles of each K block into a 3-stage ring, 128-byte swizzle, completion on mbarriers;
//                          rows / K columns outside the tensors are zero-filled by the TMA unit, so K = 600 needs no padding
//   warp 1  MMA issuer     one lane issues 4 x tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = BLOCK_N, K = 16) per K block
//                          (SASS UTCHMMA); tcgen05.commit releases the smem stage and, at the end, publishes the accumulator
//   warps 2-5 epilogue     tcgen05.ld 32 lanes x 32 columns at a time (SASS LDTM), + bias, optional sigmoid, stores
// The accumulator (128 lanes x BLOCK_N fp32 columns) lives in tensor memory; nothing is kept in registers across K.
// Two CTAs fit an SM (96 KB of stages each), so one CTA's epilogue overlaps the other's main loop.
//
// At the model's shapes (M = frames ~ 10^5, N <= 512, K <= 640) the layer is HBM-bound: arithmetic intensity
// ~ 2*N*K / (2*K + 4*N) FLOP/B < the ~255 FLOP/B ridge, so the roofline that matters is bytes.
#include <cuda.h>
#include <cuda_bf16.h>
#include <string.h>

#include "mg_common.cuh"

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

// ------------------------------------------------------------------------------------------------------------------
// GEMM
// ------------------------------------------------------------------------------------------------------------------
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;            // 64 bf16 = 128 bytes = one swizzle row
constexpr int kUmmaK = 16;             // K of one tcgen05.mma for 16-bit operands
constexpr int kMaxBlockN = 128;
constexpr int kStages = 3;
constexpr int kTmemCols = 128;         // power of two >= 32
constexpr int kGemmThreads = 192;      // 6 warps
constexpr uint32_t kATileBytes = kBlockM * kBlockK * 2;       // 16 KB
constexpr uint32_t kBTileBytes = kMaxBlockN * kBlockK * 2;    // 16 KB (BLOCK_N <= 128)
constexpr uint32_t kStageBytes = kATileBytes + kBTileBytes;
constexpr size_t kGemmSmem = kStages * kStageBytes + 1024;    // + slack to align the ring to 1024 bytes (128B swizzle)

struct GemmParams {
  const float* bias;
  void* y;
  int64_t ldy;
  int M, N, K, block_n, act, y_is_bf16;
};

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   mg_smem_addr(smem_dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(mg_smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { mg_mbar_expect_tx(bar, bytes); }

// K-major operand tile in shared memory, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr) {
  uint64_t desc = 0;
  desc |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);        // start address, 16-byte units
  desc |= static_cast<uint64_t>(1) << 16;                           // leading byte offset (unused for swizzled K-major)
  desc |= static_cast<uint64_t>(1024 >> 4) << 32;                   // stride byte offset: 8 rows x 128 bytes
  desc |= static_cast<uint64_t>(1) << 46;                           // descriptor version (Blackwell)
  desc |= static_cast<uint64_t>(2) << 61;                           // SWIZZLE_128B
  return desc;
}

// kind::f16 instruction descriptor: bf16 x bf16 -> fp32, both operands K-major, M = 128.
__device__ __forceinline__ uint32_t umma_instr_desc(int n) {
  uint32_t d = 0;
  d |= 1u << 4;                                   // D format: F32
  d |= 1u << 7;                                   // A format: BF16
  d |= 1u << 10;                                  // B format: BF16
  d |= static_cast<uint32_t>(n >> 3) << 17;       // N / 8
  d |= static_cast<uint32_t>(kBlockM >> 4) << 24; // M / 16
  return d;
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mg_smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kGemmThreads, 2)
linear_tcgen05_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                      const __grid_constant__ GemmParams prm) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages], s_accum;
  __shared__ uint32_t s_tmem_base;

  // 128-byte swizzle wants the tiles on 1024-byte boundaries.
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (prm.N + prm.block_n - 1) / prm.block_n;
  const int m0 = static_cast<int>(blockIdx.x / n_tiles) * kBlockM, n0 = static_cast<int>(blockIdx.x % n_tiles) * prm.block_n;
  const int n_kblocks = (prm.K + kBlockK - 1) / kBlockK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < kStages; ++s) { mg_mbar_init(&s_full[s], 1); mg_mbar_init(&s_empty[s], 1); }
    mg_mbar_init(&s_accum, 1);
    mg_mbar_fence_init();
  }
  if (warp == 2) {   // one warp owns the TMEM allocation (and frees it at the end)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(mg_smem_addr(&s_tmem_base)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = s_tmem_base;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const uint32_t stage_tx = kATileBytes + static_cast<uint32_t>(prm.block_n) * kBlockK * 2;
      for (int kb = 0; kb < n_kblocks; ++kb) {
        const int s = kb % kStages;
        if (kb >= kStages) mg_mbar_wait(&s_empty[s], static_cast<uint32_t>(((kb / kStages) - 1) & 1));
        unsigned char* a_tile = ring + static_cast<size_t>(s) * kStageBytes;
        unsigned char* b_tile = a_tile + kATileBytes;
        mbar_arrive_expect_tx(&s_full[s], stage_tx);
        tma_load_2d(a_tile, &map_x, kb * kBlockK, m0, &s_full[s]);
        tma_load_2d(b_tile, &map_w, kb * kBlockK, n0, &s_full[s]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = umma_instr_desc(prm.block_n);
      for (int kb = 0; kb < n_kblocks; ++kb) {
        const int s = kb % kStages;
        mg_mbar_wait(&s_full[s], static_cast<uint32_t>((kb / kStages) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_addr = mg_smem_addr(ring + static_cast<size_t>(s) * kStageBytes);
        const uint32_t b_addr = a_addr + kATileBytes;
#pragma unroll
        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
          // advancing K by 16 elements = 32 bytes inside the 128-byte swizzle row
          const uint64_t adesc = umma_smem_desc(a_addr + k * kUmmaK * 2);
          const uint64_t bdesc = umma_smem_desc(b_addr + k * kUmmaK * 2);
          umma_f16(tmem_base, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&s_empty[s]);   // frees the stage when the MMAs that read it are done
      }
      umma_commit(&s_accum);        // accumulator complete
    }
  } else {
    // ===== epilogue: warps 2..5 own TMEM lanes 32 * (warp % 4) .. + 31 =====
    const int quarter = warp & 3;
    const int row = m0 + quarter * 32 + lane;
    mg_mbar_wait(&s_accum, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < prm.block_n; c0 += 32) {
      uint32_t acc[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(c0), acc);
      if (row < prm.M) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = n0 + c0 + j;
          float t = __uint_as_float(acc[j]);
          if (col < prm.N) {
            if (prm.bias != nullptr) t += __ldg(prm.bias + col);
            if (prm.act == MG_ACT_SIGMOID) t = 1.f / (1.f + __expf(-t));
          }
          v[j] = t;
        }
        const int n_valid = min(32, prm.N - (n0 + c0));
        if (prm.y_is_bf16) {
          __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(prm.y) + static_cast<int64_t>(row) * prm.ldy + n0 + c0;
          if (n_valid == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              __align__(16) __nv_bfloat16 h[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) h[q] = __float2bfloat16_rn(v[j + q]);
              *reinterpret_cast<uint4*>(dst + j) = *reinterpret_cast<const uint4*>(h);
            }
          } else {
            for (int j = 0; j < n_valid; ++j) dst[j] = __float2bfloat16_rn(v[j]);
          }
        } else {
          float* dst = static_cast<float*>(prm.y) + static_cast<int64_t>(row) * prm.ldy + n0 + c0;
          if (n_valid == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            for (int j = 0; j < n_valid; ++j) dst[j] = v[j];
          }
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
  }
}

// ---- host: tensor maps through the driver entry point (the library links only the CUDA runtime) ----------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult status;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &status) == cudaSuccess &&
        status == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// (rows, K) bf16 row-major with row stride ld -> 2-D map, box = 64 K-elements x box_rows rows, 128-byte swizzle, zero OOB fill.
int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
  EncodeTiledFn encode = get_encode_fn();
  if (encode == nullptr) { mg_set_error("mg_linear_bf16: cuTensorMapEncodeTiled is not available from the driver"); return MG_ERR_CUDA; }
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUresult rc = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, elem_strides,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { mg_set_error("mg_linear_bf16: cuTensorMapEncodeTiled failed with code %d", static_cast<int>(rc)); return MG_ERR_CUDA; }
  return MG_OK;
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
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MG_REQUIRE(M >= 0 && N >= 1 && K >= 1, "mg_linear_bf16: bad shape (M=%d, N=%d, K=%d)", M, N, K);
  MG_REQUIRE(act == MG_ACT_NONE || act == MG_ACT_SIGMOID, "mg_linear_bf16: unknown activation %d", act);
  if (M == 0) return MG_OK;
  MG_REQUIRE(x != nullptr && w != nullptr && y != nullptr, "mg_linear_bf16: NULL buffer");
  MG_REQUIRE(ldx % 8 == 0 && ldw % 8 == 0 && ldx >= K && ldw >= K && ldy >= N,
             "mg_linear_bf16: row strides must be multiples of 8 elements (16 bytes) and cover the row (ldx=%lld, ldw=%lld, ldy=%lld)",
             static_cast<long long>(ldx), static_cast<long long>(ldw), static_cast<long long>(ldy));
  MG_REQUIRE(mg_aligned(x, 16) && mg_aligned(w, 16), "mg_linear_bf16: operands must be 16-byte aligned");

  int block_n = 128;
  if (N <= 16) block_n = 16; else if (N <= 32) block_n = 32; else if (N <= 64) block_n = 64;
  CUtensorMap map_x, map_w;
  int rc = make_map(&map_x, x, M, K, ldx, kBlockM);
  if (rc != MG_OK) return rc;
  rc = make_map(&map_w, w, N, K, ldw, block_n);
  if (rc != MG_OK) return rc;

  GemmParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.bias = bias; prm.y = y; prm.ldy = ldy;
  prm.M = M; prm.N = N; prm.K = K; prm.block_n = block_n; prm.act = act; prm.y_is_bf16 = y_is_bf16;

  static bool attr_set = false;
  if (!attr_set) {
    MG_CUDA_OK(cudaFuncSetAttribute(linear_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kGemmSmem)));
    attr_set = true;
  }
  // n fastest: the CTAs that share an A tile are neighbours in launch order, so A is re-read from L2, not HBM.
  const int64_t n_ctas = static_cast<int64_t>((N + block_n - 1) / block_n) * ((M + kBlockM - 1) / kBlockM);
  MG_REQUIRE(n_ctas < (int64_t(1) << 31), "mg_linear_bf16: too many tiles");
  linear_tcgen05_kernel<<<static_cast<unsigned>(n_ctas), kGemmThreads, kGemmSmem, stream>>>(map_x, map_w, prm);
  MG_LAUNCH_OK();
  return MG_OK;
}
