#include <atomic>
#include <cstdio>
#include <string>
#include "common.cuh"
#include "../../include/go2_b200.h"

namespace go2 {
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};
int set_error(int code, const char* msg) { g_err = msg; return code; }
int set_cuda_error(cudaError_t err, const char* file, int line) {
  char b[512];
  snprintf(b, sizeof b, "CUDA error %d (%s) at %s:%d", (int)err, cudaGetErrorString(err), file, line);
  g_err = b;
  return 100 + (int)err;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace go2

extern "C" {
const char* go2_last_error(void) { return go2::g_err.c_str(); }
long long go2_kernel_launch_count(void) { return go2::g_launches.load(); }
}
