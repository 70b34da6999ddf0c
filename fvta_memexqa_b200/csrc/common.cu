// Library-wide helpers: version, last-error string.
#include "common.cuh"
#include <stdio.h>

namespace fvta {
static thread_local char g_last_error[256] = "";
void set_last_error(const char* what, cudaError_t e) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
}
static unsigned long long g_launches = 0;  // diagnostic only; not synchronised across threads
void count_launch() { ++g_launches; }
}  // namespace fvta

extern "C" int fvta_version(void) { return 100; }
extern "C" const char* fvta_last_error(void) { return fvta::g_last_error; }
extern "C" unsigned long long fvta_launch_count(void) { return fvta::g_launches; }
