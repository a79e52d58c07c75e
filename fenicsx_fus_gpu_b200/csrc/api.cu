// Error reporting, launch counting and device queries shared by the C ABI.
#include <atomic>
#include <cstdio>

#include "halo_internal.cuh"

namespace {
thread_local char g_err[512] = "";
thread_local const FusHaloDev* g_armed_halo = nullptr;
thread_local long long g_armed_from = 0;
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

bool fus_take_armed_wait(const FusHaloDev** h, long long* first_interface_cell) {
  if (g_armed_halo == nullptr) return false;
  *h = g_armed_halo;
  *first_interface_cell = g_armed_from;
  g_armed_halo = nullptr;
  return true;
}

extern "C" {
int fus_stiffness_arm_halo_wait(const fus_halo_t* halo, int64_t first_interface_cell) {
  g_armed_halo = halo == nullptr ? nullptr : fus_halo_dev_of(halo);
  g_armed_from = first_interface_cell;
  return 0;
}
int fus_abi_version(void) { return FUS_ABI_VERSION; }
const char* fus_last_error(void) { return g_err; }
int64_t fus_launch_count(void) { return g_launches.load(); }
void fus_reset_launch_count(void) { g_launches.store(0); }
}
