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

#include <mutex>
namespace {
struct SideSet { bool ready; cudaStream_t s[MVN_SIDE_STREAMS]; cudaEvent_t ev; };
SideSet g_side[MVN_MAX_DEVICES];
std::mutex g_side_mu;
SideSet* side_set() {
    static const bool off = getenv("MOVENET_B200_SIDE_STREAMS") && atoi(getenv("MOVENET_B200_SIDE_STREAMS")) == 0;
    int dev = 0;
    if (off || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MVN_MAX_DEVICES) return nullptr;
    std::lock_guard<std::mutex> lk(g_side_mu);
    SideSet& x = g_side[dev];
    if (!x.ready) {
        for (int i = 0; i < MVN_SIDE_STREAMS; ++i)
            if (cudaStreamCreateWithFlags(&x.s[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&x.ev, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        x.ready = true;
    }
    return &x;
}
}  // namespace

cudaStream_t mvn_side_stream(int i) {
    SideSet* x = side_set();
    return x && i >= 0 && i < MVN_SIDE_STREAMS ? x->s[i] : nullptr;
}
// (one event per device is enough: cudaStreamWaitEvent captures the record that precedes it, a later re-record does not move it)
int mvn_stream_after(cudaStream_t waiter, cudaStream_t signaller) {
    if (waiter == signaller) return 0;
    SideSet* x = side_set();
    MVN_REQUIRE(x != nullptr, "side streams are not available");
    std::lock_guard<std::mutex> lk(g_side_mu);
    MVN_CUDA(cudaEventRecord(x->ev, signaller));
    MVN_CUDA(cudaStreamWaitEvent(waiter, x->ev, 0));
    return 0;
}

extern "C" const char* mvn_last_error(void) { return g_err; }
extern "C" int mvn_version(void) { return 200; }
extern "C" unsigned long long mvn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
