// The reference's training loss (movenet/pytorch_lightning_trainer.py:62-65): F.cross_entropy applied to the
// PROBABILITIES forward() returns -- i.e. a second softmax over the channel axis:
//     loss = mean_{b,t} [ logsumexp_c p[b,c,t] - p[b,target[b,t],t] ]
//     d loss / d p[b,c,t] = g / N * ( softmax_c(p)[b,c,t] - [c == target[b,t]] )
// One thread per (b, t) column: consecutive lanes read consecutive t of the channels-first tensor, so every
// load and store is coalesced.  Forward writes one partial sum per block; a second single-block kernel adds
// them in a fixed order (deterministic).
#include "common.cuh"
#include "../../include/movenet_b200.h"

__global__ void __launch_bounds__(256) ce_fwd_kernel(const float* __restrict__ p, const long long* __restrict__ target,
                                                     int A, int T, float* __restrict__ partial) {
    MVN_PDL_PROLOGUE();
    const int t = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    float loss = 0.f;
    if (t < T) {
        const float* col = p + (size_t)b * A * T + t;
        // one pass (online max / sum): the tensor is read once; eight independent loads in flight per thread
        const long long tg = target[(size_t)b * T + t];
        float m = -INFINITY, s = 0.f, pt = 0.f;
        int c = 0;
        for (; c + 8 <= A; c += 8) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = col[(size_t)(c + i) * T];
            float m8 = v[0];
#pragma unroll
            for (int i = 1; i < 8; ++i) m8 = fmaxf(m8, v[i]);
            if (m8 > m) { s *= expf(m - m8); m = m8; }        // exp(-inf) = 0 on the first group
#pragma unroll
            for (int i = 0; i < 8; ++i) { s += expf(v[i] - m); if (c + i == tg) pt = v[i]; }
        }
        for (; c < A; ++c) {
            const float v = col[(size_t)c * T];
            if (v > m) { s *= expf(m - v); m = v; }
            s += expf(v - m);
            if (c == tg) pt = v;
        }
        loss = (m + logf(s)) - pt;
    }
    __shared__ float red[8];
    for (int o = 16; o; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float acc = 0.f;
        for (int i = 0; i < 8; ++i) acc += red[i];
        partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = acc;
    }
}

__global__ void ce_finish_kernel(const float* __restrict__ partial, int n, float inv_count, float* __restrict__ loss) {
    MVN_PDL_PROLOGUE();
    __shared__ double red[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) acc += (double)partial[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) *loss = (float)(red[0] * (double)inv_count);
}

__global__ void __launch_bounds__(256) ce_bwd_kernel(const float* __restrict__ p, const long long* __restrict__ target,
                                                     const float* __restrict__ gout, float inv_count, int A, int T,
                                                     float* __restrict__ dp) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (t >= T) return;
    const float g = gout[0] * inv_count;
    const float* col = p + (size_t)b * A * T + t;
    float* dcol = dp + (size_t)b * A * T + t;
    float m = -INFINITY;
    for (int c = 0; c < A; ++c) m = fmaxf(m, col[(size_t)c * T]);
    float s = 0.f;
    for (int c = 0; c < A; ++c) s += expf(col[(size_t)c * T] - m);
    const float inv = g / s;
    const int tg = (int)target[(size_t)b * T + t];
    for (int c = 0; c < A; ++c) dcol[(size_t)c * T] = expf(col[(size_t)c * T] - m) * inv - (c == tg ? g : 0.f);
}

extern "C" size_t mvn_softmax_ce_partials(int B, int T) { return (size_t)B * ((T + 255) / 256); }

extern "C" int mvn_softmax_ce_fwd(const float* probs, const int64_t* target, int B, int A, int T, float* partials,
                                  float* loss, void* stream) {
    MVN_REQUIRE(probs && target && partials && loss && B > 0 && A > 0 && T > 0, "mvn_softmax_ce_fwd: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((T + 255) / 256, B);
    MVN_CUDA(mvn_launch_pdl(ce_fwd_kernel, dim3(grid), dim3(256), (size_t)(0), st, probs, (const long long*)target, A, T, partials));
    int rc = mvn_check_launch("ce_fwd");
    if (rc) return rc;
    MVN_CUDA(mvn_launch_pdl(ce_finish_kernel, dim3(1), dim3(256), (size_t)(0), st, partials, (int)(grid.x * grid.y), 1.f / ((float)B * (float)T), loss));
    return mvn_check_launch("ce_finish");
}

extern "C" int mvn_softmax_ce_bwd(const float* probs, const int64_t* target, const float* grad_loss, int B, int A, int T,
                                  float* dprobs, void* stream) {
    MVN_REQUIRE(probs && target && grad_loss && dprobs && B > 0 && A > 0 && T > 0, "mvn_softmax_ce_bwd: bad arguments");
    dim3 grid((T + 255) / 256, B);
    ce_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(probs, (const long long*)target, grad_loss,
                                                          1.f / ((float)B * (float)T), A, T, dprobs);
    return mvn_check_launch("ce_bwd");
}
