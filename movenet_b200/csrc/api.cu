// error plumbing + version of the C ABI
#include "common.cuh"
#include <stdarg.h>
#include <string.h>
#include "../../include/movenet_b200.h"

static thread_local char g_err[512] = "";

void mvn_set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;

int mvn_check_launch(const char* what) {
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        mvn_set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

extern "C" const char* mvn_last_error(void) { return g_err; }
extern "C" int mvn_version(void) { return 100; }
extern "C" unsigned long long mvn_launch_count(void) { return g_launches; }
