// Error reporting, version and device check for the prompt_tts_b200 C ABI.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";
unsigned long long g_pt_launches = 0;
int g_pt_sm_reserve = 0;

void pt_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int pt_version(void) { return 1; }
extern "C" const char* pt_last_error(void) { return g_err; }

extern "C" int pt_check_device(int dev) {
  cudaDeviceProp prop;
  PT_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    pt_set_error("device %d is sm_%d%d; this library only contains sm_100a code", dev, prop.major, prop.minor);
    return PT_EARCH;
  }
  return PT_OK;
}

extern "C" unsigned long long pt_launch_count(void) { return g_pt_launches; }

// Persistent GEMM / attention grids claim every SM's shared memory; a concurrently running collective (NCCL's channel CTAs
// during the overlapped gradient all-reduce) then cannot become resident and stalls until a kernel boundary.  Reserving a few
// SMs keeps the exchange moving.  Takes effect for subsequent launches (CUDA-graph captures keep what they were captured with).
extern "C" int pt_set_sm_reserve(int n) {
  PT_REQUIRE(n >= 0 && n < 64, "pt_set_sm_reserve: n=%d", n);
  g_pt_sm_reserve = n;
  return PT_OK;
}
