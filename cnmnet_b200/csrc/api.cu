// Library-wide state of libcnmnet_b200.so: error string, version, launch counter.
#include "common.cuh"

namespace cnm {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace cnm

extern "C" const char* cnm_last_error(void) { return cnm::g_err; }
extern "C" const char* cnm_version(void) { return "cnmnet_b200 0.1 sm_100a"; }
extern "C" int64_t cnm_launch_count(void) { return (int64_t)cnm::g_launches.load(); }
