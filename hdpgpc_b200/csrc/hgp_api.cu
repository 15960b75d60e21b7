// Library-level entry points of the C ABI (include/hdpgpc_b200.h): version, error text, launch counter.
#include "hgp_common.cuh"
#include <stdarg.h>

std::atomic<int64_t> g_hgp_launches{0};
static thread_local char g_hgp_error[512] = "";

void hgp_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_hgp_error, sizeof(g_hgp_error), fmt, ap);
    va_end(ap);
}

extern "C" int hgp_version(void) { return 1; }

extern "C" const char* hgp_build_info(void) {
    return "hdpgpc_b200 sm_100a nvcc " __DATE__ " cuda " HGP_STR(__CUDACC_VER_MAJOR__) "." HGP_STR(__CUDACC_VER_MINOR__);
}

extern "C" const char* hgp_last_error(void) { return g_hgp_error; }

extern "C" int64_t hgp_launch_count(void) { return g_hgp_launches.load(std::memory_order_relaxed); }
