// mu-law companding (torchaudio.functional.mu_law_encoding / mu_law_decoding as called at
// movenet/dataset.py:284 and movenet/callbacks.py:66-76) and the one-hot expansion of
// movenet/dataset.py:285-288.
//
// Encoding is monotone in x, so inside [-1, 1] the integer code is the number of decision
// thresholds <= x.  The A-1 thresholds are found on the host by bisection over the very fp32 (or
// fp64) operation sequence the CPU function runs (movenet_b200/mulaw.py), which makes the codes
// bit-identical to the CPU result without depending on the GPU's log1p rounding.  Outside [-1, 1]
// (and for NaN/inf) the formula is evaluated directly, with the x86 float->int64 conversion rule
// (out-of-range and NaN give INT64_MIN).
#include "common.cuh"
#include "../../include/movenet_b200.h"
#include <math.h>

template <typename T>
__device__ __forceinline__ long long mulaw_direct(T x, T mu) {
    T s = (x > T(0)) ? T(1) : ((x < T(0)) ? T(-1) : T(0));
    T v = s * log1p(mu * fabs(x)) / log1p(mu);
    v = (v + T(1)) / T(2) * mu + T(0.5);
    if (!(v == v) || fabs(v) >= T(9.2233720368547758e18)) return (long long)0x8000000000000000ULL;
    return (long long)v;   // truncation toward zero, like .to(torch.int64)
}

template <typename T>
__global__ void mulaw_encode_kernel(const T* __restrict__ x, const T* __restrict__ thr, int A, long long* __restrict__ codes,
                                    long long n) {
    extern __shared__ unsigned char smem_raw[];
    T* sthr = (T*)smem_raw;
    for (int i = threadIdx.x; i < A - 1; i += blockDim.x) sthr[i] = thr[i];
    __syncthreads();
    const T mu = (T)(A - 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const T v = x[i];
        long long code;
        if (v >= T(-1) && v <= T(1)) {
            int lo = 0, hi = A - 1;          // count of thresholds <= v
            while (lo < hi) { int mid = (lo + hi) >> 1; if (sthr[mid] <= v) lo = mid + 1; else hi = mid; }
            code = lo;
        } else {
            code = mulaw_direct<T>(v, mu);
        }
        codes[i] = code;
    }
}

__global__ void mulaw_decode_kernel(const long long* __restrict__ codes, const float* __restrict__ lut, int A,
                                    float* __restrict__ x, long long n) {
    extern __shared__ unsigned char smem_raw[];
    float* slut = (float*)smem_raw;
    for (int i = threadIdx.x; i < A; i += blockDim.x) slut[i] = lut[i];
    __syncthreads();
    const float mu = (float)(A - 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long q = codes[i];
        float v;
        if (q >= 0 && q < A) v = slut[q];
        else {
            float t = ((float)q / mu) * 2.f - 1.f;
            float s = (t > 0.f) ? 1.f : ((t < 0.f) ? -1.f : 0.f);
            v = s * (expf(fabsf(t) * log1pf(mu)) - 1.f) / mu;
        }
        x[i] = v;
    }
}

__global__ void one_hot_kernel(const long long* __restrict__ codes, float* __restrict__ audio, int A, int T) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.z;
    if (t >= T) return;
    const long long q = codes[(size_t)b * T + t];
    for (int a = blockIdx.y; a < A; a += gridDim.y) audio[((size_t)b * A + a) * T + t] = (q == a) ? 1.f : 0.f;
}

extern "C" int mvn_mulaw_encode(const void* x, int x_is_f64, const void* thresholds, int n_channels, int64_t* codes,
                                int64_t n, void* stream) {
    if (n == 0) return 0;
    MVN_REQUIRE(x && thresholds && codes && n_channels >= 2 && n > 0, "mvn_mulaw_encode: bad arguments");
    const int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    cudaStream_t st = (cudaStream_t)stream;
    if (x_is_f64)
        mulaw_encode_kernel<double><<<grid, 256, (size_t)n_channels * 8, st>>>((const double*)x, (const double*)thresholds,
                                                                               n_channels, (long long*)codes, n);
    else
        mulaw_encode_kernel<float><<<grid, 256, (size_t)n_channels * 4, st>>>((const float*)x, (const float*)thresholds,
                                                                              n_channels, (long long*)codes, n);
    return mvn_check_launch("mulaw_encode");
}

extern "C" int mvn_mulaw_decode(const int64_t* codes, const float* lut, int n_channels, float* x, int64_t n, void* stream) {
    if (n == 0) return 0;
    MVN_REQUIRE(codes && lut && x && n_channels >= 2 && n > 0, "mvn_mulaw_decode: bad arguments");
    const int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    mulaw_decode_kernel<<<grid, 256, (size_t)n_channels * 4, (cudaStream_t)stream>>>((const long long*)codes, lut, n_channels, x, n);
    return mvn_check_launch("mulaw_decode");
}

extern "C" int mvn_one_hot(const int64_t* codes, float* audio, int B, int A, int T, void* stream) {
    MVN_REQUIRE(codes && audio && B > 0 && A > 0 && T > 0, "mvn_one_hot: bad arguments");
    dim3 grid(mvn_cdiv(T, 256), A < 64 ? A : 64, B);
    one_hot_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const long long*)codes, audio, A, T);
    return mvn_check_launch("one_hot");
}
