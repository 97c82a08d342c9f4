// Library-wide state of libe2enet_b200.so: last-error string, launch counter, version.
#include "common.cuh"

#include <atomic>
#include <string.h>

namespace {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace

void e2e_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void e2e_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

extern "C" const char* e2e_last_error(void) { return g_err; }
extern "C" int e2e_version(void) { return 100; }
extern "C" const char* e2e_precision(void) { return E2E_PRECISION_NAME; }
extern "C" long long e2e_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
