// tcgen05 / TMA primitives shared by the K7 kernels (forward: mg_linear.cu, weight gradient: mg_linear_bwd.cu).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>

#include "mg_common.cuh"

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;            // 64 bf16 = 128 bytes = one swizzle row
constexpr int kUmmaK = 16;             // K of one tcgen05.mma for 16-bit operands

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   mg_smem_addr(smem_dst)),
               "l"(map), "r"(c0), "r"(c1), "r"(mg_smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, const void* smem_src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1),
               "r"(mg_smem_addr(smem_src))
               : "memory");
}
// ---- CTA-pair (cta_group::2) forms: two CTAs of a cluster on one TPC run ONE 256-row MMA; rank 0 (the leader) issues it ----
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared-window address: "the leader's copy"
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Both CTAs load into their OWN shared memory; the bytes are counted on the LEADER's barrier.
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          mg_smem_addr(smem_dst)),
      "l"(map), "r"(c0), "r"(c1), "r"(mg_smem_addr(bar) & kPeerBitMask)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrives on the barrier at this offset in BOTH CTAs once the leader's MMAs so far are complete.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   mg_smem_addr(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {   // arrive on the leader CTA's copy of `bar`
  asm volatile(
      "{\n"
      ".reg .b32 remote;\n"
      "mapa.shared::cluster.u32 remote, %0, 0;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [remote];\n"
      "}\n" ::"r"(mg_smem_addr(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mg_smem_addr(bar)) : "memory");
}

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
__device__ __forceinline__ uint32_t umma_instr_desc(int n, int m = kBlockM) {
  uint32_t d = 0;
  d |= 1u << 4;                                   // D format: F32
  d |= 1u << 7;                                   // A format: BF16
  d |= 1u << 10;                                  // B format: BF16
  d |= static_cast<uint32_t>(n >> 3) << 17;       // N / 8
  d |= static_cast<uint32_t>(m >> 4) << 24;       // M / 16 (256 for a CTA pair)
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
// Issues the load only: tmem_ld_wait() must follow before the registers are read (several loads may share one wait).
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
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


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

// (rows, cols) row-major with row stride ld elements -> 2-D map, box = box_cols x box_rows (box_cols * elem = 128 bytes),
// 128-byte swizzle, zero fill / clipping outside the tensor.
int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
             CUtensorMapDataType dtype, int elem_bytes) {
  EncodeTiledFn encode = get_encode_fn();
  if (encode == nullptr) { mg_set_error("mg_linear_bf16: cuTensorMapEncodeTiled is not available from the driver"); return MG_ERR_CUDA; }
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUresult rc = encode(map, dtype, 2, const_cast<void*>(base), dims, strides, box, elem_strides,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { mg_set_error("mg_linear_bf16: cuTensorMapEncodeTiled failed with code %d", static_cast<int>(rc)); return MG_ERR_CUDA; }
  return MG_OK;
}


}  // namespace
