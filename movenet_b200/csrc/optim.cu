// AdamW over all parameter tensors in ONE launch: the trainer's
// `getattr(torch.optim, config.optimizer)(params, lr)` step (movenet/pytorch_lightning_trainer.py:128-202) and its optional
// gradient-norm clip (`gradient_clip_val`, :233-243) for a model of 10 N + 13 small tensors, where per-tensor (or
// per-list-chunk) launches cost more than the arithmetic.  The host passes a device table of segments (parameter, gradient,
// first and second moment pointers and a length) and a chunk map; one CTA updates one 2048-element chunk.
//
// Update (torch.optim.AdamW, amsgrad = False, maximize = False):
//   p <- p (1 - lr wd);  m <- b1 m + (1 - b1) g;  v <- b2 v + (1 - b2) g^2;  p <- p - (lr / bc1) m / (sqrt(v) / sqrt(bc2) + eps)
// with bc1 = 1 - b1^t, bc2 = 1 - b2^t.  With max_grad_norm > 0 the gradient is first scaled by
// min(1, max_grad_norm / (||g||_2 + 1e-6)) (torch.nn.utils.clip_grad_norm_), the norm taken over ALL segments from
// per-chunk partial sums in a fixed order (deterministic); nothing is read back to the host.
#include <cmath>
#include "common.cuh"
#include "../../include/movenet_b200.h"

namespace {

constexpr int CHUNK = 2048, THREADS = 256;

struct Seg { float* p; const float* g; float* m; float* v; long long n; };
struct Chunk { int seg; int first; };     // elements [first * CHUNK, ...) of segment seg

__global__ void __launch_bounds__(THREADS) grad_sqnorm_kernel(const Seg* __restrict__ segs, const Chunk* __restrict__ chunks,
                                                              float* __restrict__ partials) {
    MVN_PDL_PROLOGUE();
    const Chunk c = chunks[blockIdx.x];
    const Seg s = segs[c.seg];
    const long long i0 = (long long)c.first * CHUNK;
    float acc = 0.f;
    for (int i = threadIdx.x; i < CHUNK && i0 + i < s.n; i += THREADS) { const float g = s.g[i0 + i]; acc = fmaf(g, g, acc); }
    __shared__ float red[THREADS / 32];
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < THREADS / 32; ++w) t += red[w]; partials[blockIdx.x] = t; }
}

__global__ void __launch_bounds__(THREADS) adamw_kernel(const Seg* __restrict__ segs, const Chunk* __restrict__ chunks, int n_chunks,
                                                        float step, float decay, float b1, float omb1, float b2, float omb2,
                                                        float eps, float sqrt_bc2, float max_norm, const float* __restrict__ partials,
                                                        float* __restrict__ norm_out) {
    MVN_PDL_PROLOGUE();
    float coef = 1.f;
    if (max_norm > 0.f) {      // every CTA re-reduces the few hundred partials in the same order
        __shared__ float red[THREADS / 32];
        float acc = 0.f;
        for (int i = threadIdx.x; i < n_chunks; i += THREADS) acc += partials[i];
        for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
        __syncthreads();
        float t = 0.f;
        for (int w = 0; w < THREADS / 32; ++w) t += red[w];
        const float norm = sqrtf(t);
        coef = fminf(1.f, max_norm / (norm + 1e-6f));
        if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) norm_out[0] = norm;
    }
    const Chunk c = chunks[blockIdx.x];
    const Seg s = segs[c.seg];
    const long long i0 = (long long)c.first * CHUNK;
    for (int i = threadIdx.x; i < CHUNK && i0 + i < s.n; i += THREADS) {
        const long long j = i0 + i;
        const float g = s.g[j] * coef;
        const float m = b1 * s.m[j] + omb1 * g;
        const float v = b2 * s.v[j] + omb2 * g * g;
        s.m[j] = m; s.v[j] = v;
        s.p[j] = s.p[j] * decay - step * (m / (sqrtf(v) / sqrt_bc2 + eps));
    }
}

}  // namespace

extern "C" size_t mvn_adamw_segment_bytes(void) { return sizeof(Seg); }
extern "C" int mvn_adamw_chunk_elems(void) { return CHUNK; }

extern "C" int mvn_adamw_step(const void* segments_dev, const void* chunks_dev, int n_chunks, double lr, double beta1, double beta2,
                              double eps, double weight_decay, double bias_correction1, double bias_correction2,
                              double max_grad_norm, float* sq_partials, float* grad_norm_out, void* stream) {
    MVN_REQUIRE(segments_dev && chunks_dev && n_chunks > 0, "mvn_adamw_step: empty table");
    MVN_REQUIRE(bias_correction1 > 0. && bias_correction2 > 0., "mvn_adamw_step: bias corrections must be positive (step >= 1)");
    MVN_REQUIRE(!(max_grad_norm > 0.) || sq_partials, "mvn_adamw_step: clipping needs the partial-sum buffer (n_chunks floats)");
    cudaStream_t st = (cudaStream_t)stream;
    if (max_grad_norm > 0.) {
        MVN_CUDA(mvn_launch_pdl(grad_sqnorm_kernel, dim3(n_chunks), dim3(THREADS), (size_t)0, st, (const Seg*)segments_dev,
                                (const Chunk*)chunks_dev, sq_partials));
        int rc = mvn_check_launch("grad_sqnorm"); if (rc) return rc;
    }
    MVN_CUDA(mvn_launch_pdl(adamw_kernel, dim3(n_chunks), dim3(THREADS), (size_t)0, st, (const Seg*)segments_dev,
                            (const Chunk*)chunks_dev, n_chunks,
                            // the scalars are formed in double like torch forms them in Python (1 - 0.999 in fp32 is off by 5e-5)
                            (float)(lr / bias_correction1), (float)(1.0 - lr * weight_decay), (float)beta1, (float)(1.0 - beta1),
                            (float)beta2, (float)(1.0 - beta2), (float)eps, (float)sqrt(bias_correction2), (float)max_grad_norm,
                            (const float*)sq_partials, grad_norm_out));
    return mvn_check_launch("adamw");
}
