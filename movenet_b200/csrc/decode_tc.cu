// Cached autoregressive decoding on tensor cores: the throughput mode of WaveNet.generate
// (movenet/wavenet.py:193-239) for models whose weights fit in shared memory (the receptive-field
// configuration, experiments/04: C = 16, S = 8, A = 128, 14 layers).
//
// A group of 4 warps advances 128 clips in lock-step: clip = MMA row = TMEM lane = thread.  Per layer
//   A = [x_l[t-d] | x_l[t]]  (queue pop from HBM, current activation from registers)  -> bf16 tile
//   tcgen05.mma  D1[128 x 2C] = A . Wz^T   -> gate in-thread -> bf16 tile
//   tcgen05.mma  D2[128 x (C+S)] = gated . [Wr|Ws]^T -> residual / skip update in registers
// One barrier round per layer, not two: the residual update h_{l+1} = h_l + Wr g_l + br sits between the two products, but
//   Wz_{l+1} [old | h_{l+1}] = Wz_{l+1} [old | h_l] + (Wz1_{l+1} Wr_l) g_l + Wz1_{l+1} br_l
// so layer l+1's pre-activation is issued TOGETHER with layer l's out product, from the tile that already holds bf16(h_l), the
// gated tile and a pre-multiplied weight (Wc = Wz1_{l+1} Wr_l, f16, built by the pack kernel).  h_{l+1} itself is still formed
// in fp32 registers from D2 for the queue push and the layers above.  The residual biases never appear in the step loop:
// registers, tiles and queues hold h^_l = h_l - beta_l with beta_l = br_0 + .. + br_{l-1} (h^_{l+1} = h^_l + Wr g_l exactly), the
// pre-activation gets the constant (Wz0_l + Wz1_l) beta_l as a gate bias, and the prefill kernel subtracts beta_l from the
// training forward's activations.
// then the dense head (two more MMAs) and the next token is chosen by the thread that owns the clip:
// argmax (lowest index on ties) or a draw from softmax(softmax(z)/temperature) need no cross-thread
// traffic at all.  Two such groups share one CTA (and one copy of the weights in shared memory) and
// interleave, so one group's MMA / barrier latency hides behind the other's epilogue.
//
// The gated tile and [Wr|Ws] are f16 (gate math in packed f16x2), everything else bf16.
// Queues are bf16, laid out (layer, slot, clip, channel): a group's pop and push of one layer are two
// contiguous 128 x C x 2-byte blocks.  Operand tiles use the un-swizzled K-major core-matrix layout
// (8 rows x 16 bytes contiguous), which is compact for any K and conflict-free for one-row-per-thread
// epilogue writes.
#include "tc_common.cuh"
#include "layer_tc.h"

using namespace tc;

namespace {

constexpr int DS = 8;            // skip_channels handled
template <int C> struct Cfg { static constexpr int GROUPS = C <= 16 ? 4 : 2; };   // 128-clip groups per CTA (register budget)

struct DecTcArgs {
    const uint8_t* img; int img_bytes;
    const float* win;      // [2][A][C] fp32 input-conv rows, gathered from global memory (L1/L2 resident)
    __nv_bfloat16* queues; int* last2; int* out_codes_t; float* out_logits; const int* forced;
    int B, N, A, t_start, n_new;
    float temperature; unsigned seed;
    int layer_stride, oWrs, oBrs, oW1, oB1, oW2, oB2, oWin;
    long long qoff[MVN_MAX_LAYERS];
    int dil[MVN_MAX_LAYERS];
};

// byte offset of element (r, k) in a K-major, un-swizzled tile whose rows hold K elements (bf16)
__host__ __device__ inline int core_off(int r, int k, int K) { return (r >> 3) * (K * 16) + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2; }
__device__ __forceinline__ uint64_t desc_k_plain(uint32_t saddr, int K) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(128 >> 4) << 16;                  // LBO: next 8-element core along K
    d |= (uint64_t)(((K * 16) >> 4) & 0x3FFF) << 32;  // SBO: next 8-row group
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

template <int C>
struct Img {   // image layout shared by the pack kernel and the decode kernel
    static constexpr int N2 = ((C + DS + 15) / 16) * 16;
    static constexpr int wz = 0, wz_bytes = 2 * C * 2 * C * 2;
    static constexpr int wrs = wz + wz_bytes, wrs_bytes = N2 * C * 2;
    static constexpr int brs = wrs + wrs_bytes;
    static constexpr int wc = brs + N2 * 4, wc_bytes = 2 * C * C * 2;      // Wc^T[n][k] = (Wz1_l Wr_{l-1})[n][k], f16 (layers >= 1)
    static constexpr int zb = wc + wc_bytes;                              // (Wz0_l + Wz1_l) beta_l as f16x2 pairs: C/2 filter | C/2 gate
    static constexpr int layer_bytes = zb + 2 * C * 2;
};

template <int C>
__global__ void decode_tc_pack_kernel(const float* __restrict__ packed, PackedLayout P, int N, int A, uint8_t* __restrict__ img,
                                      int oW1, int oB1, int oW2, int oB2, int oWin) {
    using I = Img<C>;
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (int l = 0; l < N; ++l) {
        const float* lw = packed + P.layer0 + (size_t)l * P.layer_stride;
        uint8_t* li = img + (size_t)l * I::layer_bytes;
        for (int i = i0; i < 2 * C * 2 * C; i += stride) {          // Wz^T[n][k], n: filter c | gate c ; k: tap0 | tap1
            const int n = i / (2 * C), k = i % (2 * C);
            const float v = lw[P.oWz + (size_t)k * 2 * C + 2 * (n % C) + n / C];
            *(__nv_bfloat16*)(li + I::wz + core_off(n, k, 2 * C)) = __float2bfloat16(v);
        }
        for (int i = i0; i < I::N2 * C; i += stride) {               // [Wr|Ws]^T[n][k]
            const int n = i / C, k = i % C;
            const float v = n < C + DS ? lw[P.oWrs + (size_t)k * (C + DS) + n] : 0.f;
            *(__half*)(li + I::wrs + core_off(n, k, C)) = __float2half_rn(v);      // f16: the gated tile it multiplies is f16
        }
        for (int i = i0; i < I::N2; i += stride) ((float*)(li + I::brs))[i] = i < C + DS ? lw[P.obrs + i] : 0.f;
        if (l > 0) {                                                  // the pre-multiplied hand-over from layer l - 1
            const float* pw = packed + P.layer0 + (size_t)(l - 1) * P.layer_stride;
            for (int i = i0; i < 2 * C * C; i += stride) {
                const int n = i / C, k = i % C;
                float acc = 0.f;
                for (int j = 0; j < C; ++j)
                    acc += lw[P.oWz + (size_t)(C + j) * 2 * C + 2 * (n % C) + n / C] * pw[P.oWrs + (size_t)k * (C + DS) + j];
                *(__half*)(li + I::wc + core_off(n, k, C)) = __float2half_rn(acc);
            }
            for (int i = i0; i < 2 * C; i += stride) {             // (Wz0_l + Wz1_l) beta_l, beta_l = sum of the residual biases below
                float acc = 0.f;
                for (int j = 0; j < C; ++j) {
                    float beta = 0.f;
                    for (int m = 0; m < l; ++m) beta += packed[P.layer0 + (size_t)m * P.layer_stride + P.obrs + j];
                    const int col = 2 * (i % C) + i / C;
                    acc += (lw[P.oWz + (size_t)j * 2 * C + col] + lw[P.oWz + (size_t)(C + j) * 2 * C + col]) * beta;
                }
                ((__half*)(li + I::zb))[i] = __float2half_rn(acc);
            }
        }
    }
    for (int i = i0; i < A * 16; i += stride) {                      // W1^T[n][k], k padded 8 -> 16
        const int n = i / 16, k = i % 16;
        // k = DS carries the bias: the A tile holds 1.0 there (b1 rounded to bf16 like the weights)
        *(__nv_bfloat16*)(img + oW1 + core_off(n, k, 16)) = __float2bfloat16(k < DS ? packed[P.w1p + (size_t)k * A + n] : k == DS ? packed[P.b1 + n] : 0.f);
    }
    for (int i = i0; i < A * A; i += stride) {                       // W2^T[n][k]
        const int n = i / A, k = i % A;
        *(__nv_bfloat16*)(img + oW2 + core_off(n, k, A)) = __float2bfloat16(packed[P.w2p + (size_t)k * A + n]);
    }
    for (int i = i0; i < A; i += stride) { ((float*)(img + oB1))[i] = packed[P.b1 + i]; ((float*)(img + oB2))[i] = packed[P.b2 + i]; }
    for (int i = i0; i < DS; i += stride) {          // sum over the layers of the skip biases
        float acc = 0.f;
        for (int l = 0; l < N; ++l) acc += packed[P.layer0 + (size_t)l * P.layer_stride + P.obrs + C + i];
        ((float*)(img + oWin))[i] = acc;
    }
}

template <int C>
__global__ void __launch_bounds__(128 * Cfg<C>::GROUPS, 1) decode_tc_kernel(const DecTcArgs a) {
    using I = Img<C>;
    constexpr int GROUPS = Cfg<C>::GROUPS;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* simg = smem;
    const int A = a.A;
    const int tid = threadIdx.x, grp = tid >> 7, r = tid & 127, warp = tid >> 5;
    // the first warp of each group issues its MMAs from warp-uniform code through one elected lane (tc_common.cuh)
    const bool issue_warp = __shfl_sync(0xffffffffu, (tid >> 5) & 3, 0) == 0;
    // per-group tiles after the image: layer A [128 x 2C] | gated [128 x C], and -- aliased onto them, the layers are
    // finished when the head runs -- the head's A tile [128 x max(16, A)]
    const int layer_tiles = 128 * 2 * C * 2 + 128 * C * 2, head_tile = 128 * A * 2;
    const int tiles_per_group = layer_tiles > head_tile ? layer_tiles : head_tile;
    uint8_t* gbase = smem + ((a.img_bytes + 1023) & ~1023) + grp * ((tiles_per_group + 1023) & ~1023);
    uint8_t* sA = gbase;
    uint8_t* sG = sA + 128 * 2 * C * 2;
    uint8_t* sH = gbase;
    uint64_t* bars = (uint64_t*)(smem + ((a.img_bytes + 1023) & ~1023) + GROUPS * ((tiles_per_group + 1023) & ~1023));
    uint64_t* mma_bar = bars + grp;
    uint32_t* tmem_slot = (uint32_t*)(bars + GROUPS);

    for (int i = tid; i < a.img_bytes / 16; i += blockDim.x) ((uint4*)simg)[i] = ((const uint4*)a.img)[i];
    if (r == 0) mbar_init(mma_bar, 1);
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot + grp * (512 / GROUPS);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    // TMEM columns of the group's window: D1 [0, 2C) ; D2 [2C, 2C + N2) ; the head reuses [0, A) once the layers are done
    constexpr int D1 = 0, D2 = 2 * C, DH = 0;
    static_assert(2 * C + I::N2 <= 512 / GROUPS, "TMEM window");
    const uint32_t i1 = umma_idesc_major(128, 2 * C, 0, 0), i1h = i1 & ~((1u << 7) | (1u << 10)) /* f16 operands */, i2 = umma_idesc_major(128, I::N2, 0, 0) & ~((1u << 7) | (1u << 10)) /* f16 operands */, ih = umma_idesc_major(128, A, 0, 0);

    const int b = (blockIdx.x * GROUPS + grp) * 128 + r;
    const bool live = b < a.B;
    const float* win = a.win;
    int code_prev = live ? a.last2[2 * b] : -1, code_cur = live ? a.last2[2 * b + 1] : -1;
    uint32_t phase = 0;

    for (int i = a.t_start; i < a.t_start + a.n_new; ++i) {
        const int tau = i - 1;
        float h[C], skip[DS];
#pragma unroll
        for (int c = 0; c < C; c += 4) {
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 w0 = code_prev >= 0 ? *(const float4*)(win + (size_t)code_prev * C + c) : z4;
            const float4 w1 = code_cur >= 0 ? *(const float4*)(win + ((size_t)A + code_cur) * C + c) : z4;
            h[c] = w0.x + w1.x; h[c + 1] = w0.y + w1.y; h[c + 2] = w0.z + w1.z; h[c + 3] = w0.w + w1.w;
        }
#pragma unroll
        for (int s = 0; s < DS; ++s) skip[s] = 0.f;

        // the queue rows x_l[tau - d] do not depend on this step's arithmetic: the row of layer l + 1 is fetched while layer l
        // computes (one global-memory latency per layer would otherwise sit on the token's critical path)
        // (dilations are powers of two -- movenet/modules.py:113 builds 2**x -- so the ring slot is a mask, not a division: the
        // kernel is bound by instruction issue, and two runtime modulos per layer were a tenth of its instructions)
        auto ring_of = [&](int l) { return a.queues + (a.qoff[l] + (long long)(tau & (a.dil[l] - 1)) * C) * a.B + (size_t)(live ? b : 0) * C; };
        uint4 old_next[C / 8];
#pragma unroll
        for (int q = 0; q < C / 8; ++q) {
            old_next[q] = make_uint4(0, 0, 0, 0);
            if (live && tau - a.dil[0] >= 0) old_next[q] = ((const uint4*)ring_of(0))[q];
        }
        // prologue: A = [x_0[tau - d] | x_0[tau]], push x_0[tau], pre-activation of layer 0
        auto stage_old = [&]() {
#pragma unroll
            for (int q = 0; q < C / 8; ++q) *(uint4*)(sA + core_off(r, 8 * q, 2 * C)) = old_next[q];
        };
        auto stage_push_h = [&](int l) {
            __nv_bfloat16* ring = ring_of(l);
#pragma unroll
            for (int q = 0; q < C / 8; ++q) {
                const uint4 cur = make_uint4(pack_bf16(h[8 * q], h[8 * q + 1]), pack_bf16(h[8 * q + 2], h[8 * q + 3]),
                                             pack_bf16(h[8 * q + 4], h[8 * q + 5]), pack_bf16(h[8 * q + 6], h[8 * q + 7]));
                *(uint4*)(sA + core_off(r, C + 8 * q, 2 * C)) = cur;
                if (live) ((uint4*)ring)[q] = cur;
            }
        };
        auto prefetch_old = [&](int l) {
            if (l < a.N) {
                const __nv_bfloat16* nring = ring_of(l);
                const bool has = live && tau - a.dil[l] >= 0;
#pragma unroll
                for (int q = 0; q < C / 8; ++q) old_next[q] = has ? ((const uint4*)nring)[q] : make_uint4(0, 0, 0, 0);
            }
        };
        stage_old();
        stage_push_h(0);
        fence_proxy_async();
        tc_fence_before();
        group_sync(grp);
        if (issue_warp) {
            tc_fence_after();
            const uint64_t dA = desc_k_plain(smem_u32(sA), 2 * C), dW = desc_k_plain(smem_u32(simg + I::wz), 2 * C);
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 2 * C / 16; ++k) umma(tmem_u + D1, desc_adv(dA, k * 256), desc_adv(dW, k * 256), i1, k != 0);
                umma_commit(mma_bar);
            }
            __syncwarp();
        }
        prefetch_old(1);
        mbar_wait(mma_bar, phase); phase ^= 1;
        tc_fence_after();
        for (int l = 0; l < a.N; ++l) {
            // here D1 holds layer l's pre-activation and (l > 0) D2 holds layer l - 1's out product
            const uint8_t* li = simg + (size_t)l * I::layer_bytes;
            if (l > 0) {
                // (TMEM reads are what this kernel is made of -- 64 B/clk per SM: only the C + 8 live columns of D2 are read)
                uint32_t v[C + DS];
#pragma unroll
                for (int q = 0; q < C / 16; ++q) tmem_ld16(tmem + lane_base + D2 + 16 * q, v + 16 * q);
                tmem_ld8(tmem + lane_base + D2 + C, v + C);
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < C; ++c) h[c] += __uint_as_float(v[c]);            // h^ (bias-free, see the header)
#pragma unroll
                for (int s = 0; s < DS; ++s) skip[s] += __uint_as_float(v[C + s]);       // (the skip biases are added once, below)
                stage_push_h(l);                 // x_l[tau]: queue push, and the second half of the next A tile
            }
            {
                uint32_t f[C], g[C];
#pragma unroll
                for (int q = 0; q < C / 16; ++q) { tmem_ld16(tmem + lane_base + D1 + 16 * q, f + 16 * q); tmem_ld16(tmem + lane_base + D1 + C + 16 * q, g + 16 * q); }
                tmem_ld_wait();
                const uint32_t* zb = (const uint32_t*)(li + I::zb);
#pragma unroll
                for (int q = 0; q < C / 8; ++q) {
                    // tanh(f) * sigmoid(g) on channel pairs in packed f16x2 (one MUFU op per two tanh, as in layer_tc.cu): the
                    // 11-bit intermediates are above the bf16 queues' precision; the out GEMM runs on f16 operands
                    const uint32_t h05 = 0x38003800u;     // (0.5, 0.5)
                    uint32_t o[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int c = 8 * q + 2 * e;
                        uint32_t fh = f16x2(__uint_as_float(f[c]), __uint_as_float(f[c + 1]));
                        uint32_t gh = f16x2(__uint_as_float(g[c]), __uint_as_float(g[c + 1]));
                        if (l > 0) { fh = hadd2(fh, zb[c >> 1]); gh = hadd2(gh, zb[(C + c) >> 1]); }
                        o[e] = hmul2(htanh2(fh), hfma2(htanh2(hmul2(gh, h05)), h05, h05));
                    }
                    *(uint4*)(sG + core_off(r, 8 * q, C)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            if (l + 1 < a.N) stage_old();        // x_{l+1}[tau - d] (fetched one layer ahead)
            fence_proxy_async();
            tc_fence_before();
            group_sync(grp);
            if (issue_warp) {
                tc_fence_after();
                const uint64_t dG = desc_k_plain(smem_u32(sG), C), dW = desc_k_plain(smem_u32(li + I::wrs), C);
                const uint64_t dA = desc_k_plain(smem_u32(sA), 2 * C), dWz = desc_k_plain(smem_u32(li + I::layer_bytes + I::wz), 2 * C),
                               dWc = desc_k_plain(smem_u32(li + I::layer_bytes + I::wc), C);
                const bool more = l + 1 < a.N;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < C / 16; ++k) umma(tmem_u + D2, desc_adv(dG, k * 256), desc_adv(dW, k * 256), i2, k != 0);
                    if (more) {                  // layer l + 1's pre-activation from [old | h_l] and the gated tile
#pragma unroll
                        for (int k = 0; k < 2 * C / 16; ++k) umma(tmem_u + D1, desc_adv(dA, k * 256), desc_adv(dWz, k * 256), i1, k != 0);
#pragma unroll
                        for (int k = 0; k < C / 16; ++k) umma(tmem_u + D1, desc_adv(dG, k * 256), desc_adv(dWc, k * 256), i1h, 1);
                    }
                    umma_commit(mma_bar);
                }
                __syncwarp();
            }
            prefetch_old(l + 2);
            mbar_wait(mma_bar, phase); phase ^= 1;
            tc_fence_after();
        }
        {   // the last layer's out product: only its skip rows are used (the residual output is discarded)
            const float* bs = (const float*)(simg + a.oWin);       // sum over the layers of the skip biases (pack kernel)
            uint32_t v[DS];
            tmem_ld8(tmem + lane_base + D2 + C, v);
            tmem_ld_wait();
#pragma unroll
            for (int s = 0; s < DS; ++s) skip[s] += __uint_as_float(v[s]) + bs[s];
            tc_fence_before();
        }
        // ---- dense head: a1 = W1 lrelu(skip) + b1 ; z = W2 lrelu(a1) + b2 -----------------------------
        {
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float x0 = skip[2 * e], x1 = skip[2 * e + 1];
                o[e] = pack_bf16(x0 > 0.f ? x0 : MVN_LRELU_SLOPE * x0, x1 > 0.f ? x1 : MVN_LRELU_SLOPE * x1);
            }
            *(uint4*)(sH + core_off(r, 0, 16)) = make_uint4(o[0], o[1], o[2], o[3]);
            *(uint4*)(sH + core_off(r, 8, 16)) = make_uint4(0x3F80u, 0, 0, 0);      // 1.0 at k = 8: b1 comes out of the MMA
        }
        fence_proxy_async();
        tc_fence_before();
        group_sync(grp);
        if (issue_warp) {
            tc_fence_after();
            const uint64_t dH = desc_k_plain(smem_u32(sH), 16), dW = desc_k_plain(smem_u32(simg + a.oW1), 16);
            if (elect_one()) {
                umma(tmem_u + DH, dH, dW, ih, 0);
                umma_commit(mma_bar);
            }
            __syncwarp();
        }
        mbar_wait(mma_bar, phase); phase ^= 1;
        tc_fence_after();
        {
            for (int q = 0; q < A / 16; ++q) {
                uint32_t v[16];
                tmem_ld16(tmem + lane_base + DH + 16 * q, v);
                tmem_ld_wait();
                uint32_t o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {       // leaky ReLU as max(x, slope * x): two instructions per element
                    const float x0 = __uint_as_float(v[2 * e]), x1 = __uint_as_float(v[2 * e + 1]);
                    o[e] = pack_bf16(fmaxf(x0, MVN_LRELU_SLOPE * x0), fmaxf(x1, MVN_LRELU_SLOPE * x1));
                }
                *(uint4*)(sH + core_off(r, 16 * q, A)) = make_uint4(o[0], o[1], o[2], o[3]);
                *(uint4*)(sH + core_off(r, 16 * q + 8, A)) = make_uint4(o[4], o[5], o[6], o[7]);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        group_sync(grp);          // also: every thread has finished reading a1 from TMEM before it is overwritten
        if (issue_warp) {
            tc_fence_after();
            const uint64_t dH = desc_k_plain(smem_u32(sH), A), dW = desc_k_plain(smem_u32(simg + a.oW2), A);
            if (elect_one()) {
                for (int k = 0; k < A / 16; ++k) umma(tmem_u + DH, desc_adv(dH, k * 256), desc_adv(dW, k * 256), ih, k != 0);
                umma_commit(mma_bar);
            }
            __syncwarp();
        }
        mbar_wait(mma_bar, phase); phase ^= 1;
        tc_fence_after();
        // ---- next token, chosen by the thread that owns the clip --------------------------------------
        int arg = 0;
        {
            const float* b2 = (const float*)(simg + a.oB2);
            float best = -INFINITY;
            // (the kernel is bound by instruction issue: the plain path carries no logit-store code, and the biases come as float4)
            float* zout = (a.out_logits && live) ? a.out_logits + ((size_t)b * a.n_new + (i - a.t_start)) * A : nullptr;
            const bool any_out = a.out_logits != nullptr;
            for (int q = 0; q < A / 16; ++q) {
                uint32_t v[16];
                tmem_ld16(tmem + lane_base + DH + 16 * q, v);
                float bb[16];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 t4 = ((const float4*)(b2 + 16 * q))[e];
                    bb[4 * e] = t4.x; bb[4 * e + 1] = t4.y; bb[4 * e + 2] = t4.z; bb[4 * e + 3] = t4.w;
                }
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float z = __uint_as_float(v[e]) + bb[e];
                    if (z > best) { best = z; arg = 16 * q + e; }
                    bb[e] = z;
                }
                if (any_out && zout) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) ((float4*)(zout + 16 * q))[e] = make_float4(bb[4 * e], bb[4 * e + 1], bb[4 * e + 2], bb[4 * e + 3]);
                }
            }
            if (a.temperature > 0.f) {      // draw from softmax(softmax(z) / temperature) (movenet/wavenet.py:227-231)
                float s = 0.f;
                for (int q = 0; q < A / 16; ++q) {
                    uint32_t v[16];
                    tmem_ld16(tmem + lane_base + DH + 16 * q, v); tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) s += __expf(__uint_as_float(v[e]) + b2[16 * q + e] - best);
                }
                const float inv_s = 1.f / s, inv_t = 1.f / a.temperature, pmax = inv_s * inv_t;
                float qs = 0.f;
                for (int q = 0; q < A / 16; ++q) {
                    uint32_t v[16];
                    tmem_ld16(tmem + lane_base + DH + 16 * q, v); tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) qs += __expf(__expf(__uint_as_float(v[e]) + b2[16 * q + e] - best) * inv_s * inv_t - pmax);
                }
                unsigned long long x = ((unsigned long long)a.seed << 32) ^ ((unsigned long long)(unsigned)b * 0x9E3779B97F4A7C15ULL) ^ (unsigned long long)(unsigned)i;
                x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 27; x *= 0x94D049BB133111EBULL; x ^= x >> 31;
                const float target = (float)(x >> 40) * (1.f / 16777216.f) * qs;
                float run = 0.f; int pick = -1;
                for (int q = 0; q < A / 16; ++q) {
                    uint32_t v[16];
                    tmem_ld16(tmem + lane_base + DH + 16 * q, v); tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        run += __expf(__expf(__uint_as_float(v[e]) + b2[16 * q + e] - best) * inv_s * inv_t - pmax);
                        if (pick < 0 && target < run) pick = 16 * q + e;
                    }
                }
                arg = pick < 0 ? A - 1 : pick;
            }
        }
        if (a.forced && live) arg = a.forced[(size_t)b * a.n_new + (i - a.t_start)];
        if (live) a.out_codes_t[(size_t)(i - a.t_start) * a.B + b] = arg;
        code_prev = code_cur; code_cur = arg;
        tc_fence_before();
        group_sync(grp);          // the head's TMEM columns and tiles are free for the next step
    }
    if (live) { a.last2[2 * b] = code_prev; a.last2[2 * b + 1] = code_cur; }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "n"(512) : "memory");
    }
}

// rings (layer, slot, clip, channel) bf16 from the layer inputs of a forward over the T-column prompt:
// ring_l[tau % d] = x_l[tau] for tau in [T-1-d, T-1)  (the first decode step re-evaluates time T-1 itself)
__global__ void decode_tc_prefill_kernel(const void* __restrict__ x, int adt, int B, int T, int C, int d, __nv_bfloat16* __restrict__ ring,
                                         const float* __restrict__ packed, long long brs0, long long layer_stride, int layer) {
    const long long n = (long long)B * d * C;
    const int Tend = T - 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); const long long rr = i / C; const int b = (int)(rr % B); const int slot = (int)(rr / B);
        int tau = (Tend / d) * d + slot; if (tau >= Tend) tau -= d;
        float beta = 0.f;                            // the queues hold h^ = h - beta_l (decode_tc_kernel)
        for (int m = 0; m < layer; ++m) beta += packed[brs0 + m * layer_stride + c];
        ring[i] = __float2bfloat16(tau >= 0 ? mvn_ld(x, adt, ((size_t)b * T + tau) * C + c) - beta : 0.f);
    }
}
__global__ void decode_tc_last2_kernel(const int* __restrict__ codes, int B, int T, int* __restrict__ last2) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    last2[2 * b] = T >= 2 ? codes[(size_t)b * T + T - 2] : -1;
    last2[2 * b + 1] = T >= 1 ? codes[(size_t)b * T + T - 1] : -1;
}

static size_t queue_elems(const Geo& g, long long* qoff) {
    long long o = 0;
    for (int l = 0; l < g.N; ++l) { if (qoff) qoff[l] = o; o += (long long)g.dil[l] * g.C; }
    return (size_t)o;
}

template <int C>
int image_offsets(const Geo& g, DecTcArgs& a) {
    using I = Img<C>;
    int o = g.N * I::layer_bytes;
    a.layer_stride = I::layer_bytes; a.oWrs = I::wrs; a.oBrs = I::brs;
    a.oW1 = o; o += g.A * 16 * 2;
    a.oB1 = o; o += g.A * 4;
    a.oW2 = o; o += g.A * g.A * 2;
    a.oB2 = o; o += g.A * 4;
    a.oWin = o; o += DS * 4;                       // (the slot holds the summed skip biases)
    a.img_bytes = (o + 15) & ~15;
    return a.img_bytes;
}

template <int C>
int smem_bytes(const Geo& g, int img_bytes) {
    const int layer_tiles = 128 * 2 * C * 2 + 128 * C * 2, head_tile = 128 * g.A * 2;
    const int tiles = layer_tiles > head_tile ? layer_tiles : head_tile;
    return ((img_bytes + 1023) & ~1023) + Cfg<C>::GROUPS * ((tiles + 1023) & ~1023) + 64 + 1024;
}

template <int C>
int run_steps(const Geo& g, DecTcArgs& a, const float* packed, const PackedLayout& P, uint8_t* img, cudaStream_t st) {
    decode_tc_pack_kernel<C><<<64, 256, 0, st>>>(packed, P, g.N, g.A, img, a.oW1, a.oB1, a.oW2, a.oB2, a.oWin);
    int rc = mvn_check_launch("decode_tc_pack");
    if (rc) return rc;
    const int smem = smem_bytes<C>(g, a.img_bytes);
    MVN_CUDA(cudaFuncSetAttribute(decode_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    decode_tc_kernel<C><<<mvn_cdiv(g.B, 128 * Cfg<C>::GROUPS), 128 * Cfg<C>::GROUPS, smem, st>>>(a);
    return mvn_check_launch("decode_tc_steps");
}

}  // namespace

int mvn_tc_decode_supported(const Geo& g) {
    if (g.video || g.S != DS || (g.C != 16 && g.C != 32) || g.A % 16 || g.A < 16 || g.A > 128) return 0;
    for (int l = 0; l < g.N; ++l) if (g.dil[l] & (g.dil[l] - 1)) return 0;      // ring slots are taken with a mask
    DecTcArgs a;
    const int img = g.C == 16 ? image_offsets<16>(g, a) : image_offsets<32>(g, a);
    const int smem = g.C == 16 ? smem_bytes<16>(g, img) : smem_bytes<32>(g, img);
    return smem <= 227 * 1024;
}

extern "C" size_t mvn_decode_tc_state_bytes(const mvn_shape_t* s) {
    Geo g; if (geo_init(g, s)) return 0;
    DecTcArgs a;
    const int img = g.C == 16 ? image_offsets<16>(g, a) : image_offsets<32>(g, a);
    return al256(queue_elems(g, nullptr) * (size_t)g.B * 2) + al256((size_t)g.B * 2 * 4) + al256(img);
}

extern "C" int mvn_decode_tc_supported(const mvn_shape_t* s) {
    Geo g; if (geo_init(g, s)) return 0;
    return mvn_tc_decode_supported(g);
}

extern "C" int mvn_decode_tc_prefill(const mvn_shape_t* s, const void* packed, const void* acts, void* state, void* stream) {
    Geo g; MVN_REQUIRE(s && geo_init(g, s) == 0, "mvn_decode_tc_prefill: bad shape");
    MVN_REQUIRE(packed && acts && state && mvn_tc_decode_supported(g), "mvn_decode_tc_prefill: unsupported shape");
    PackedLayout P; packed_layout(g, P);
    ActsLayout AL; acts_layout(g, AL);
    long long qoff[MVN_MAX_LAYERS];
    const size_t qe = queue_elems(g, qoff);
    __nv_bfloat16* queues = (__nv_bfloat16*)state;
    int* last2 = (int*)((char*)state + al256(qe * (size_t)g.B * 2));
    cudaStream_t st = (cudaStream_t)stream;
    for (int l = 0; l < g.N; ++l) {
        const void* x = (const char*)acts + AL.x0 + (size_t)l * AL.x_stride;
        const long long n = (long long)g.B * g.dil[l] * g.C;
        decode_tc_prefill_kernel<<<mvn_cdiv(n, 256) < 1184 ? mvn_cdiv(n, 256) : 1184, 256, 0, st>>>(x, g.adt, g.B, g.T, g.C, g.dil[l],
                                                                                                 queues + qoff[l] * g.B,
                                                                                                 (const float*)packed, (long long)(P.layer0 + P.obrs),
                                                                                                 (long long)P.layer_stride, l);
    }
    decode_tc_last2_kernel<<<mvn_cdiv(g.B, 128), 128, 0, st>>>((const int*)((const char*)acts + AL.codes), g.B, g.T, last2);
    return mvn_check_launch("decode_tc_prefill");
}

extern "C" int mvn_decode_tc_steps(const mvn_shape_t* s, const void* packed, void* state, int t_start, int n_new,
                                   int* out_codes_t, float* out_logits, const int* forced, float temperature, unsigned seed,
                                   void* stream) {
    Geo g; MVN_REQUIRE(s && geo_init(g, s) == 0, "mvn_decode_tc_steps: bad shape");
    MVN_REQUIRE(packed && state && out_codes_t && n_new >= 0 && t_start >= 1 && mvn_tc_decode_supported(g), "mvn_decode_tc_steps: bad arguments");
    // every queue row the steps pop was pushed by a real time step (no zero padding inside the queues: they hold h - beta)
    for (int l = 0; l < g.N; ++l) MVN_REQUIRE(t_start - 1 >= g.dil[l], "mvn_decode_tc_steps: t_start must be past every dilation (prompt >= receptive field)");
    if (n_new == 0) return 0;
    DecTcArgs a; memset(&a, 0, sizeof(a));
    PackedLayout P; packed_layout(g, P);
    const size_t qe = queue_elems(g, a.qoff);
    for (int l = 0; l < g.N; ++l) a.dil[l] = g.dil[l];
    a.queues = (__nv_bfloat16*)state;
    a.last2 = (int*)((char*)state + al256(qe * (size_t)g.B * 2));
    uint8_t* img = (uint8_t*)state + al256(qe * (size_t)g.B * 2) + al256((size_t)g.B * 2 * 4);
    a.img = img;
    a.out_codes_t = out_codes_t; a.out_logits = out_logits; a.forced = forced; a.win = (const float*)packed + P.win;
    a.B = g.B; a.N = g.N; a.A = g.A; a.t_start = t_start; a.n_new = n_new; a.temperature = temperature; a.seed = seed;
    cudaStream_t st = (cudaStream_t)stream;
    if (g.C == 16) { image_offsets<16>(g, a); return run_steps<16>(g, a, (const float*)packed, P, img, st); }
    image_offsets<32>(g, a); return run_steps<32>(g, a, (const float*)packed, P, img, st);
}
