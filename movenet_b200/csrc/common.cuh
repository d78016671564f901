// Shared helpers for the movenet_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define MVN_F32 0
#define MVN_BF16 1

#define MVN_LRELU_SLOPE 0.01f   // F.leaky_relu default (movenet/modules.py:140-141)

// error plumbing: every extern "C" entry returns 0 or a non-zero code and
// leaves a message for mvn_last_error(); nothing throws across the C boundary.
void mvn_set_error(const char* fmt, ...);
int mvn_check_launch(const char* what);

#define MVN_REQUIRE(cond, ...)                      \
    do {                                            \
        if (!(cond)) {                              \
            mvn_set_error(__VA_ARGS__);             \
            return -1;                              \
        }                                           \
    } while (0)

#define MVN_CUDA(call)                                                         \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) {                                              \
            mvn_set_error("%s failed: %s", #call, cudaGetErrorString(e__));    \
            return (int)e__;                                                   \
        }                                                                      \
    } while (0)

// Programmatic dependent launch: a kernel launched with mvn_launch_pdl() may be scheduled while the kernel before it in
// the stream is still draining its last CTAs (its launch latency and prologue overlap that tail); it must call
// mvn_griddep_wait() before it touches anything that kernel wrote.  mvn_griddep_launch() (issued by every CTA at its
// start: the trigger fires once ALL CTAs of the grid are running or done) lets the NEXT kernel's CTAs be scheduled as
// SM resources free up.  Every kernel launched through mvn_launch_pdl starts with MVN_PDL_PROLOGUE().
__device__ __forceinline__ void mvn_griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void mvn_griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#define MVN_PDL_PROLOGUE() do { mvn_griddep_launch(); mvn_griddep_wait(); } while (0)

template <typename... KArgs, typename... Args>
inline cudaError_t mvn_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float mvn_ld(const void* p, int dtype, long long i) {
    return dtype == MVN_BF16 ? __bfloat162float(((const __nv_bfloat16*)p)[i]) : ((const float*)p)[i];
}
__device__ __forceinline__ void mvn_st(void* p, int dtype, long long i, float v) {
    if (dtype == MVN_BF16) ((__nv_bfloat16*)p)[i] = __float2bfloat16(v);
    else ((float*)p)[i] = v;
}
__device__ __forceinline__ float mvn_lrelu(float v) { return v > 0.f ? v : MVN_LRELU_SLOPE * v; }
__device__ __forceinline__ float mvn_lrelu_grad(float pre) { return pre > 0.f ? 1.f : MVN_LRELU_SLOPE; }
__device__ __forceinline__ float mvn_sigmoid(float v) { return 1.f / (1.f + expf(-v)); }

static inline int mvn_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a function: remember what was set per device so
// a process that drives several GPUs opts every one of them in (`slot` is a zero-initialised static of the call site).
#define MVN_MAX_DEVICES 64
struct MvnSmemAttr { int set[MVN_MAX_DEVICES]; };
template <typename F>
inline cudaError_t mvn_ensure_smem(F func, int bytes, MvnSmemAttr& slot) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < MVN_MAX_DEVICES && slot.set[dev] >= bytes) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev >= 0 && dev < MVN_MAX_DEVICES) slot.set[dev] = bytes;   // (a racing thread sets the same value)
    return e;
}
// SMs of the current device (persistent grids are sized from it); cached per device
int mvn_sm_count();
// Side streams of the current device (created on first use, non-blocking) for small follow-up kernels -- partial-sum
// reductions -- that need not sit between two persistent kernels on the caller's stream.  mvn_stream_after(w, s): everything
// enqueued on `w` from now on waits for everything enqueued on `s` so far.  mvn_side_stream(i) returns nullptr when side
// streams are switched off ($MOVENET_B200_SIDE_STREAMS=0): the caller then uses its own stream.
#define MVN_SIDE_STREAMS 3
cudaStream_t mvn_side_stream(int i);
int mvn_stream_after(cudaStream_t waiter, cudaStream_t signaller);
