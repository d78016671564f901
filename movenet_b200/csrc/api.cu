// error plumbing + version of the C ABI
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include "../../include/movenet_b200.h"

static thread_local char g_err[512] = "";

void mvn_set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

#include <atomic>
static std::atomic<unsigned long long> g_launches{0};

int mvn_sm_count() {
    static std::atomic<int> cache[MVN_MAX_DEVICES];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MVN_MAX_DEVICES) return 148;
    int n = cache[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cache[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

int mvn_check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    // MOVENET_B200_SYNC=1 (debugging): wait for every kernel so that an asynchronous fault is attributed to the launch that caused it
    static const bool sync_debug = getenv("MOVENET_B200_SYNC") != nullptr;
    if (e == cudaSuccess && sync_debug) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        mvn_set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

extern "C" const char* mvn_last_error(void) { return g_err; }
extern "C" int mvn_version(void) { return 200; }
extern "C" unsigned long long mvn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
