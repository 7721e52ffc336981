// Shared helpers for the C-ABI translation units: error slot, launch counter, status macros.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace oai {

char* last_error_buf();  // thread-local, 512 bytes
extern std::atomic<long long> g_launches;

inline int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return 1;
}

inline int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  return fail("%s: %s", what, cudaGetErrorString(e));
}

inline int launched(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaGetLastError(), what);
}

inline int num_sms() {
  static int cached[64] = {0};   // per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
  int n = 0;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (n <= 0) n = 148;
  if (dev >= 0 && dev < 64) cached[dev] = n;
  return n;
}

}  // namespace oai

#define OAI_REQUIRE(cond, ...) \
  do {                         \
    if (!(cond)) return ::oai::fail(__VA_ARGS__); \
  } while (0)
