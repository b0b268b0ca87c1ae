// Library-wide state of libcnmnet_b200.so: error string, version, launch counter.
#include <cstdlib>

#include "common.cuh"

namespace cnm {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace cnm

namespace cnm {
static int g_range_check = -1;                      // -1: not decided yet (environment), 0 off, 1 on
static unsigned long long* g_range_ctr[64] = {};   // one counter per device
unsigned long long* range_counter() {
  if (g_range_check < 0) {
    const char* e = getenv("CNM_CHECK_RANGE");
    g_range_check = (e && e[0] == '1') ? 1 : 0;
  }
  if (!g_range_check) return nullptr;
  int dev = 0;
  cudaGetDevice(&dev);
  unsigned long long*& p = g_range_ctr[dev & 63];
  if (!p) {  // first use on this device (never inside a stream capture: plans run once un-captured first)
    if (cudaMalloc(&p, sizeof(unsigned long long)) != cudaSuccess) { p = nullptr; cudaGetLastError(); return nullptr; }
    cudaMemset(p, 0, sizeof(unsigned long long));
  }
  return p;
}
}  // namespace cnm

namespace cnm {
static int g_pdl = -1;  // -1: environment (CNM_PDL, default on)
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("CNM_PDL");
    g_pdl = (e && e[0] == '0') ? 0 : 1;
  }
  return g_pdl == 1;
}
}  // namespace cnm

extern "C" void cnm_set_pdl(int enable) { cnm::g_pdl = enable ? 1 : 0; }
extern "C" void cnm_set_range_check(int enable) { cnm::g_range_check = enable ? 1 : 0; }
extern "C" int64_t cnm_range_overflow_count(int reset) {
  if (cnm::g_range_check != 1) return -1;
  unsigned long long* p = cnm::range_counter();
  if (!p) return -1;
  unsigned long long v = 0;
  if (cudaMemcpy(&v, p, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (reset) cudaMemset(p, 0, sizeof(v));
  return (int64_t)v;
}

extern "C" const char* cnm_last_error(void) { return cnm::g_err; }
extern "C" const char* cnm_version(void) { return "cnmnet_b200 0.1 sm_100a"; }
extern "C" int64_t cnm_launch_count(void) { return (int64_t)cnm::g_launches.load(); }
