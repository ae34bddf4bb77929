// Library plumbing: version, last-error buffer, device queries.
#include <stdlib.h>

#include "mg_common.cuh"

static thread_local char g_last_error[512] = "";

void mg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

int mg_cached_sm_count() {
  static thread_local int cached_device = -1;
  static thread_local int cached_count = 0;
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess) return 148;
  if (device != cached_device) {
    int count = 0;
    if (cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || count <= 0) return 148;
    cached_device = device;
    cached_count = count;
  }
  return cached_count;
}

bool mg_pdl_enabled() {
  static const bool enabled = []() { const char* e = getenv("MG_PDL"); return !(e && *e == '0'); }();
  return enabled;
}

extern "C" int mg_abi_version(void) { return MG_ABI_VERSION; }

extern "C" const char* mg_last_error(void) { return g_last_error; }

extern "C" int mg_sm_count(void) {
  int device = 0, count = 0;
  MG_CUDA_OK(cudaGetDevice(&device));
  MG_CUDA_OK(cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, device));
  return count;
}
