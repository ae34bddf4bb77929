// Shared helpers for the morgana_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "morgana_b200.h"

#ifndef __CUDA_ARCH__
#define MG_HOST_ONLY 1
#endif

// ---------------------------------------------------------------------------------------------------------------
// Error reporting: thread-local message + status code, never an exception across the C boundary.
// ---------------------------------------------------------------------------------------------------------------
void mg_set_error(const char* fmt, ...);

#define MG_REQUIRE(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      mg_set_error(__VA_ARGS__);         \
      return MG_ERR_INVALID_ARG;         \
    }                                    \
  } while (0)

#define MG_CUDA_OK(expr)                                                                       \
  do {                                                                                         \
    cudaError_t mg_err__ = (expr);                                                             \
    if (mg_err__ != cudaSuccess) {                                                             \
      mg_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(mg_err__), __FILE__, __LINE__); \
      return MG_ERR_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

#define MG_LAUNCH_OK()                                                                  \
  do {                                                                                  \
    cudaError_t mg_err__ = cudaGetLastError();                                          \
    if (mg_err__ != cudaSuccess) {                                                      \
      mg_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(mg_err__), __FILE__, __LINE__); \
      return MG_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

int mg_cached_sm_count();
bool mg_pdl_enabled();   // programmatic dependent launch for the kernels of the path (MG_PDL=0 switches it off)

static inline bool mg_aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---------------------------------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------------------------------
#define MG_FULL_MASK 0xffffffffu

__device__ __forceinline__ uint32_t mg_smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Streaming (read-once) 16-byte load that does not pollute L1.
__device__ __forceinline__ float4 mg_ld_stream_f4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float mg_ld_stream_f1(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// Streaming store (evict-first in L2: the frame-rate output is far larger than L2 and is not re-read by us).
__device__ __forceinline__ void mg_st_stream_f4(float4* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// --- bulk async copy (TMA engine, non-tensor form): shared::cta -> global -----------------------------------------
// SASS: UBLKCP.  dst, src 16-byte aligned; bytes a multiple of 16.
__device__ __forceinline__ void mg_bulk_store(void* gdst, uint32_t smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mg_bulk_store_hint(void* gdst, uint32_t smem_src, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
               "r"(smem_src), "r"(bytes), "l"(policy)
               : "memory");
}
__device__ __forceinline__ uint64_t mg_policy_evict_first() {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}
// Small, re-used data (tables, lengths, per-CTA records) asks L2 to keep it: the streams around it are hundreds of MB per launch
// and would otherwise push it to DRAM between two uses.
__device__ __forceinline__ uint64_t mg_policy_evict_last() {
  uint64_t policy;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
  return policy;
}
__device__ __forceinline__ uint32_t mg_ld_keep_u32(const void* p, uint64_t policy) {
  uint32_t v;
  asm volatile("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ int64_t mg_ld_keep_s64(const void* p, uint64_t policy) {
  int64_t v;
  asm volatile("ld.global.L2::cache_hint.s64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ double2 mg_ld_keep_cg_f64x2(const void* p, uint64_t policy) {   // from L2, never a stale L1 line
  double2 v;
  asm volatile("ld.global.cg.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ void mg_st_keep_f64x2(void* p, double2 v, uint64_t policy) {
  asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(policy) : "memory");
}
__device__ __forceinline__ void mg_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// Wait until at most N of this thread's bulk groups still have their SOURCE (shared memory) unread.
template <int N>
__device__ __forceinline__ void mg_bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// Make generic-proxy writes to shared memory visible to the async proxy (the bulk-copy engine).
__device__ __forceinline__ void mg_fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// --- bulk async copy global -> shared::cta, completion counted on an mbarrier (TMA engine; SASS UBLKCP) ---------------
// src and dst 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void mg_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mg_smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mg_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mg_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mg_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mg_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   mg_smem_addr(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(mg_smem_addr(bar))
               : "memory");
}
// The same with an L2 eviction-priority hint (a read-once stream should not push the rest of the working set out of L2).
__device__ __forceinline__ void mg_bulk_load_hint(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   mg_smem_addr(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(mg_smem_addr(bar)), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void mg_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(mg_smem_addr(bar)),
      "r"(parity)
      : "memory");
}

// --- programmatic dependent launch ------------------------------------------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may become resident while the kernel before it in
// the stream is still running; it MUST execute mg_pdl_wait() before its first access to global memory -- the wait returns when
// the prerequisite grid has completed and its memory operations are visible, so stream order is kept for everything after it.
// What runs before the wait (barrier initialisation, shared-memory fills) overlaps the previous kernel's tail, and the launch
// latency is hidden.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void mg_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void mg_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t mg_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mg_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// --- warp / CTA reductions in a fixed order ------------------------------------------------------------------------
__device__ __forceinline__ double mg_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MG_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ long long mg_warp_sum(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MG_FULL_MASK, v, o);
  return v;
}
