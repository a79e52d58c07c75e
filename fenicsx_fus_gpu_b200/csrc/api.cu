// Error reporting, launch counting and device queries shared by the C ABI.
#include <atomic>
#include <cstdio>

#include "fus_common.cuh"

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace

int fus_set_error(int code, const char* what) {
  if (code > 0 && code < 100000) {
    snprintf(g_err, sizeof(g_err), "%s: %s (%s)", what, cudaGetErrorString((cudaError_t)code),
             cudaGetErrorName((cudaError_t)code));
  } else {
    snprintf(g_err, sizeof(g_err), "%s", what);
  }
  return code;
}

void fus_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int fus_num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (sms[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    sms[dev] = v;
  }
  return sms[dev];
}

extern "C" {
int fus_abi_version(void) { return FUS_ABI_VERSION; }
const char* fus_last_error(void) { return g_err; }
int64_t fus_launch_count(void) { return g_launches.load(); }
void fus_reset_launch_count(void) { g_launches.store(0); }
}
