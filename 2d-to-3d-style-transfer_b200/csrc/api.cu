// api.cu -- error reporting and version for the libst3d C ABI.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

static thread_local char g_err[512] = "";

void st3d_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* st3d_last_error(void) { return g_err; }
extern "C" int st3d_version(void) { return 100; }

// Kernel launches issued by this library in this process (a statistic, never read by the kernels).
static std::atomic<unsigned long long> g_launches{0};
void st3d_count_launch(void) { g_launches.fetch_add(1, std::memory_order_relaxed); }
extern "C" unsigned long long st3d_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
