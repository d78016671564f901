// DenseConv head (movenet/modules.py:133-142) + drop-last + softmax (movenet/wavenet.py:183-191) and its
// backward on tcgen05 tensor cores, for input_channels == 64.
//
// One thread owns one time row (TMEM lane = time), so the softmax over the 64 channels needs no
// shuffles and the channels-first (B, A, T) API tensors are read/written with the warp's 32 lanes on
// 32 consecutive time steps: coalesced without a transpose.
//
// forward : a1 = W1 lrelu(skip) + b1 on CUDA cores (K = S is tiny) -> lrelu -> bf16 tile (K-major,
//           128B swizzle) -> tcgen05.mma z = . W2^T -> softmax -> probabilities (or logits).
// backward: recompute a1; dz = p (dp - <dp, p>) in-thread -> bf16 tile -> tcgen05.mma da1 = dz . W2
//           (B = the SAME W2 image read MN-major) -> lrelu' -> dskip on CUDA cores;
//           weight / bias gradients: [dz | da1]^T . [lrelu(a1) | lrelu(skip)] with K = time, accumulated
//           in TMEM over the CTA's tiles, written as per-CTA partials and reduced in a fixed order.
#include "tc_common.cuh"
#include "layer_tc.h"

using namespace tc;

namespace {

constexpr int HA = 64;                       // input_channels handled
constexpr int HPART = 128 * 128 + 128;       // per-CTA partial: D_w[128][128] + bias sums[128]

struct HeadArgs {
    const float* w1p;    // [S][A] fp32
    const float* b1;     // [A]
    const float* b2;     // [A]
    const void* img;     // W2 image: [A n][A k] bf16, K-major, 128B swizzle
    const float* skip;   // (B, Tout, S)
    float* out;          // forward: (B, A, Tn)
    const float* probs;  // backward
    const float* dout;   // backward
    float* dskip;        // backward: (B, Tout, S)
    float* partial;      // backward
    int B, Tout, Tn, S, logits, tiles_per_clip, n_tiles;
};

__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : MVN_LRELU_SLOPE * v; }

// a1[n] for n in [n0, n0+32) of one row, from the row's S skip values (already leaky-ReLU'd)
template <int S>
__device__ __forceinline__ void head_a1(const float* sw1, const float* sb1, const float* ls, int n0, float* a1) {
#pragma unroll
    for (int i = 0; i < 32; ++i) a1[i] = sb1[n0 + i];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const float x = ls[s];
        const float4* w = (const float4*)(sw1 + s * HA + n0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 wv = w[q];
            a1[4 * q] = fmaf(wv.x, x, a1[4 * q]); a1[4 * q + 1] = fmaf(wv.y, x, a1[4 * q + 1]);
            a1[4 * q + 2] = fmaf(wv.z, x, a1[4 * q + 2]); a1[4 * q + 3] = fmaf(wv.w, x, a1[4 * q + 3]);
        }
    }
}

// 32 fp32 values of row r, channels [32*half, +32) -> bf16 chunks of a 128B-swizzled [128 x 64] tile
__device__ __forceinline__ void store_half_row(uint8_t* tile, int r, int half, const float* v) {
    const int sw = r & 7;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *(uint4*)(tile + r * 128 + (((4 * half + q) ^ sw) << 4)) =
            make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                       pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
}

template <int S>
__global__ void __launch_bounds__(256, 2) head_fwd_tc_kernel(const HeadArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sW2 = smem;                          // 8 KB image
    uint8_t* sA = smem + 8192;                    // 16 KB: lrelu(a1) tile
    float* sw1 = (float*)(sA + TILE_BYTES);       // [S][64]
    float* sb1 = sw1 + S * HA;
    float* sb2 = sb1 + HA;
    float* sx = sb2 + HA;                         // [2][128] softmax exchange
    uint64_t* mma_bar = (uint64_t*)(sx + 256);
    uint32_t* tmem_slot = (uint32_t*)(mma_bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, r = tid & 127, half = tid >> 7, n0 = 32 * half;

    for (int i = tid; i < 8192 / 16; i += 256) ((uint4*)sW2)[i] = ((const uint4*)a.img)[i];
    for (int i = tid; i < S * HA; i += 256) sw1[i] = a.w1p[i];
    if (tid < HA) { sb1[tid] = a.b1[tid]; sb2[tid] = a.b2[tid]; }
    if (tid == 0) { mbar_init(mma_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t idesc = umma_idesc_major(TILE_T, HA, 0, 0);

    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int b = tile / a.tiles_per_clip, j = (tile - b * a.tiles_per_clip) * TILE_T + r;
        const bool live = j < a.Tn;
        float ls[S];
        {
            const float* src = a.skip + ((size_t)b * a.Tout + (live ? j : 0)) * S;
            #pragma unroll
            for (int s = 0; s < S; s += 4) {
                const float4 v = live ? *(const float4*)(src + s) : make_float4(0.f, 0.f, 0.f, 0.f);
                ls[s] = lrelu(v.x); ls[s + 1] = lrelu(v.y); ls[s + 2] = lrelu(v.z); ls[s + 3] = lrelu(v.w);
            }
        }
        float v[32];
        head_a1<S>(sw1, sb1, ls, n0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = lrelu(v[i]);
        store_half_row(sA, r, half, v);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k) umma(tmem, umma_desc(smem_u32(sA) + k * 32), umma_desc(smem_u32(sW2) + k * 32), idesc, k != 0);
            umma_commit(mma_bar);
        }
        mbar_wait(mma_bar, it & 1);
        tc_fence_after();
        uint32_t z0[16], z1[16];
        tmem_ld16(tmem + lane_base + n0, z0);
        tmem_ld16(tmem + lane_base + n0 + 16, z1);
        tmem_ld_wait();
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            v[i] = __uint_as_float(z0[i]) + sb2[n0 + i]; v[16 + i] = __uint_as_float(z1[i]) + sb2[n0 + 16 + i];
            m = fmaxf(m, fmaxf(v[i], v[16 + i]));
        }
        if (!a.logits) {      // softmax over all 64 channels: the two halves of a row live in threads r and r + 128
            sx[half * 128 + r] = m;
            __syncthreads();
            m = fmaxf(sx[r], sx[128 + r]);
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) { v[i] = __expf(v[i] - m); sum += v[i]; }
            __syncthreads();
            sx[half * 128 + r] = sum;
            __syncthreads();
            const float inv = 1.f / (sx[r] + sx[128 + r]);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= inv;
        }
        if (live) {
            float* dst = a.out + ((size_t)b * HA + n0) * a.Tn + j;
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[(size_t)i * a.Tn] = v[i];
        }
        tc_fence_before();
        __syncthreads();       // every thread has read its TMEM row and the A tile is free again
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64) : "memory");
    }
}

template <int S>
__global__ void __launch_bounds__(256, 2) head_bwd_tc_kernel(const HeadArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sW2 = smem;                          // 8 KB
    uint8_t* sDZ = smem + 8192;                   // [DZ | DA1] adjacent: the M = 128 operand of the weight-gradient MMA
    uint8_t* sDA = sDZ + TILE_BYTES;
    uint8_t* sLA = sDA + TILE_BYTES;              // [LA | LS] adjacent: its N = 128 operand
    uint8_t* sLS = sLA + TILE_BYTES;
    uint8_t* sONES = sLS + TILE_BYTES;            // 1 KB
    float* sw1 = (float*)(sONES + 1024);          // [S][64]
    float* sb1 = sw1 + S * HA;
    float* sx = sb1 + HA;                         // [2][128] exchange: <dp,p> halves
    float* sds = sx + 256;                        // [128][S+1] exchange: dskip partial of half 1
    uint64_t* mma_bar = (uint64_t*)(sds + 128 * (S + 1));
    uint64_t* w_bar = mma_bar + 1;
    uint32_t* tmem_slot = (uint32_t*)(mma_bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5, r = tid & 127, half = tid >> 7, n0 = 32 * half, sw = r & 7;

    for (int i = tid; i < 8192 / 16; i += 256) ((uint4*)sW2)[i] = ((const uint4*)a.img)[i];
    for (int i = tid; i < S * HA; i += 256) sw1[i] = a.w1p[i];
    if (tid < HA) sb1[tid] = a.b1[tid];
    for (int i = tid; i < TILE_BYTES / 16; i += 256) ((uint4*)sLS)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 256; i += 256) ((uint32_t*)sONES)[i] = 0x3F803F80u;
    if (tid == 0) { mbar_init(mma_bar, 1); mbar_init(w_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t iG = umma_idesc_major(TILE_T, HA, 0, 1);        // da1 = dz . W2 : B MN-major
    const uint32_t iW = umma_idesc_major(TILE_T, 128, 1, 1);       // [dz|da1]^T . [la|ls]
    const uint32_t iB = umma_idesc_major(TILE_T, 16, 1, 1);
    constexpr int DA_COL = 0, W_COL = 64, B_COL = 192;

    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int b = tile / a.tiles_per_clip, j = (tile - b * a.tiles_per_clip) * TILE_T + r;
        const bool live = j < a.Tn;
        // ---- loads first (their latency overlaps the previous tile's weight-gradient MMAs) ----------
        float ls[S];
        unsigned long long skip_pos = 0;
        {
            const float* src = a.skip + ((size_t)b * a.Tout + (live ? j : 0)) * S;
            #pragma unroll
            for (int s = 0; s < S; s += 4) {
                const float4 v = live ? *(const float4*)(src + s) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) { skip_pos |= (unsigned long long)(x[e] > 0.f) << (s + e); ls[s + e] = lrelu(x[e]); }
            }
        }
        float dz[32];
        {
            const size_t o = ((size_t)b * HA + n0) * a.Tn + (live ? j : 0);
            float dot = 0.f;
            float p[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                dz[i] = live ? a.dout[o + (size_t)i * a.Tn] : 0.f;
                p[i] = (live && !a.logits) ? a.probs[o + (size_t)i * a.Tn] : 0.f;
                dot = fmaf(dz[i], p[i], dot);
            }
            if (!a.logits) {
                sx[half * 128 + r] = dot;
                __syncthreads();
                dot = sx[r] + sx[128 + r];
#pragma unroll
                for (int i = 0; i < 32; ++i) dz[i] = p[i] * (dz[i] - dot);
            }
        }
        float a1[32];
        head_a1<S>(sw1, sb1, ls, n0, a1);
        uint32_t a1_pos = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) { a1_pos |= (uint32_t)(a1[i] > 0.f) << i; a1[i] = lrelu(a1[i]); }
        if (it) { mbar_wait(w_bar, (it - 1) & 1); tc_fence_after(); }      // previous tile's MMAs are done with the tiles
        store_half_row(sDZ, r, half, dz);
        store_half_row(sLA, r, half, a1);
        if (half == 0) {
#pragma unroll
            for (int s = 0; s < S; s += 8)
                *(uint4*)(sLS + r * 128 + (((s >> 3) ^ sw) << 4)) =
                    make_uint4(pack_bf16(ls[s], ls[s + 1]), pack_bf16(ls[s + 2], ls[s + 3]), pack_bf16(ls[s + 4], ls[s + 5]),
                               pack_bf16(ls[s + 6], ls[s + 7]));
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma(tmem + DA_COL, umma_desc(smem_u32(sDZ) + k * 32), umma_desc_mn(smem_u32(sW2) + k * 2048, TILE_BYTES), iG, k != 0);
            umma_commit(mma_bar);
        }
        mbar_wait(mma_bar, it & 1);
        tc_fence_after();
        // ---- da1 = (dz W2) * lrelu'(a1); dskip = (da1 W1) * lrelu'(skip) ------------------------------
        float da[32];
        {
            uint32_t v0[16], v1[16];
            tmem_ld16(tmem + lane_base + DA_COL + n0, v0);
            tmem_ld16(tmem + lane_base + DA_COL + n0 + 16, v1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                da[i] = __uint_as_float(v0[i]) * ((a1_pos >> i) & 1 ? 1.f : MVN_LRELU_SLOPE);
                da[16 + i] = __uint_as_float(v1[i]) * ((a1_pos >> (16 + i)) & 1 ? 1.f : MVN_LRELU_SLOPE);
            }
        }
        store_half_row(sDA, r, half, da);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const float4* w = (const float4*)(sw1 + s * HA + n0);
            float acc = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 wv = w[q];
                acc = fmaf(wv.x, da[4 * q], acc); acc = fmaf(wv.y, da[4 * q + 1], acc);
                acc = fmaf(wv.z, da[4 * q + 2], acc); acc = fmaf(wv.w, da[4 * q + 3], acc);
            }
            ls[s] = acc;                         // reuse: partial dskip over this half's 32 channels
        }
        if (half == 1) {
#pragma unroll
            for (int s = 0; s < S; ++s) sds[r * (S + 1) + s] = ls[s];
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t acc0 = it != 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint64_t am = umma_desc_mn(smem_u32(sDZ) + k * 2048, TILE_BYTES);
                umma(tmem + W_COL, am, umma_desc_mn(smem_u32(sLA) + k * 2048, TILE_BYTES), iW, acc0 | (k != 0));
                umma(tmem + B_COL, am, umma_desc_mn_plain(smem_u32(sONES), 256, 128), iB, acc0 | (k != 0));
            }
            umma_commit(w_bar);
        }
        if (half == 0 && live) {
            float* dst = a.dskip + ((size_t)b * a.Tout + j) * a.S;
            #pragma unroll
            for (int s = 0; s < S; s += 4) {
                float o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    o[e] = (ls[s + e] + sds[r * (S + 1) + s + e]) * ((skip_pos >> (s + e)) & 1 ? 1.f : MVN_LRELU_SLOPE);
                *(float4*)(dst + s) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
        __syncthreads();           // sds / sx are reused by the next tile
    }
    if (it) { mbar_wait(w_bar, (it - 1) & 1); }
    tc_fence_after();
    float* part = a.partial + (size_t)blockIdx.x * HPART;
#pragma unroll 1
    for (int jj = half; jj < 8; jj += 2) {
        uint32_t v[16];
        tmem_ld16(tmem + lane_base + W_COL + 16 * jj, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q)
            ((float4*)(part + (size_t)r * 128 + 16 * jj))[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                                          __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
    }
    {
        uint32_t v[8];
        tmem_ld8(tmem + lane_base + B_COL, v);
        tmem_ld_wait();
        if (half == 0) part[128 * 128 + r] = __uint_as_float(v[0]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
    }
}

// D_w[m][n]: m < 64: dz channel, m >= 64: da1 channel ; n < 64: lrelu(a1) channel, n >= 64: lrelu(skip) channel
__global__ void head_reduce_kernel(const float* __restrict__ partial, int n_cta, float* __restrict__ pg, PackedLayout P, int S) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HPART; i += gridDim.x * blockDim.x) {
        float* dst = nullptr;
        if (i < 128 * 128) {
            const int m = i >> 7, n = i & 127;
            if (m < 64 && n < 64) dst = pg + P.w2p + (size_t)n * HA + m;                 // dw2p[k = n][n_out = m]
            else if (m >= 64 && n >= 64 && n - 64 < S) dst = pg + P.w1p + (size_t)(n - 64) * HA + (m - 64);   // dw1p[s][a]
        } else {
            const int m = i - 128 * 128;
            dst = m < 64 ? pg + P.b2 + m : pg + P.b1 + (m - 64);
        }
        if (!dst) continue;
        float acc = 0.f;
#pragma unroll 8
        for (int c = 0; c < n_cta; ++c) acc += partial[(size_t)c * HPART + i];
        *dst = acc;
    }
}

// W2 image: [n][k] = dense_conv.conv2.weight[n][k], bf16, 128B swizzle
__global__ void head_pack_kernel(const float* __restrict__ w2, uint8_t* __restrict__ img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HA * HA) return;
    const int n = i >> 6, k = i & 63;
    *(__nv_bfloat16*)(img + n * 128 + ((((k >> 3) ^ (n & 7)) << 4) | ((k & 7) << 1))) = __float2bfloat16(w2[i]);
}

int fwd_smem(int S) { return 8192 + TILE_BYTES + (S * HA + 2 * HA + 256) * 4 + 64 + 1024; }
int bwd_smem(int S) { return 8192 + 4 * TILE_BYTES + 1024 + (S * HA + HA + 256 + 128 * (S + 1)) * 4 + 64 + 1024; }

}  // namespace

int mvn_tc_head_supported(int A, int S) { return A == HA && (S == 8 || S == 16 || S == 32 || S == 64); }
size_t mvn_tc_head_partial_bytes() { return (size_t)2 * 148 * HPART * 4; }

int mvn_tc_head_pack(const float* w2_ref, float* packed, const PackedLayout& P, cudaStream_t st) {
    head_pack_kernel<<<(HA * HA + 255) / 256, 256, 0, st>>>(w2_ref, (uint8_t*)(packed + P.tc_head));
    return mvn_check_launch("head_pack");
}

static void fill_args(HeadArgs& a, const float* packed, const PackedLayout& P, const Geo& g) {
    a.w1p = packed + P.w1p; a.b1 = packed + P.b1; a.b2 = packed + P.b2; a.img = packed + P.tc_head;
    a.B = g.B; a.Tout = g.Tout; a.Tn = g.Tn; a.S = g.S; a.logits = g.logits;
    a.tiles_per_clip = (g.Tn + TILE_T - 1) / TILE_T; a.n_tiles = a.tiles_per_clip * g.B;
}

int mvn_tc_head_fwd(const float* packed, const PackedLayout& P, const Geo& g, const float* skip, float* out, cudaStream_t st) {
    HeadArgs a; memset(&a, 0, sizeof(a)); fill_args(a, packed, P, g);
    a.skip = skip; a.out = out;
    if (a.n_tiles <= 0) return 0;
    const int smem = fwd_smem(g.S);
    const int grid = a.n_tiles < 4 * 148 ? a.n_tiles : 4 * 148;
#define HEAD_FWD_CASE(SS)                                                                                                   \
    case SS:                                                                                                                \
        MVN_CUDA(cudaFuncSetAttribute(head_fwd_tc_kernel<SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));         \
        head_fwd_tc_kernel<SS><<<grid, 256, smem, st>>>(a);                                                                 \
        break;
    switch (g.S) { HEAD_FWD_CASE(8) HEAD_FWD_CASE(16) HEAD_FWD_CASE(32) HEAD_FWD_CASE(64) default: mvn_set_error("head: unsupported skip_channels"); return -1; }
    return mvn_check_launch("head_fwd_tc");
}

int mvn_tc_head_bwd(const float* packed, const PackedLayout& P, const Geo& g, const float* skip, const float* probs,
                    const float* dout, float* dskip, float* pg, float* partial, cudaStream_t st) {
    HeadArgs a; memset(&a, 0, sizeof(a)); fill_args(a, packed, P, g);
    a.skip = skip; a.probs = probs; a.dout = dout; a.dskip = dskip; a.partial = partial;
    if (a.n_tiles <= 0) return 0;
    const int smem = bwd_smem(g.S);
    const int grid = a.n_tiles < 2 * 148 ? a.n_tiles : 2 * 148;
#define HEAD_BWD_CASE(SS)                                                                                                   \
    case SS:                                                                                                                \
        MVN_CUDA(cudaFuncSetAttribute(head_bwd_tc_kernel<SS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));         \
        head_bwd_tc_kernel<SS><<<grid, 256, smem, st>>>(a);                                                                 \
        break;
    switch (g.S) { HEAD_BWD_CASE(8) HEAD_BWD_CASE(16) HEAD_BWD_CASE(32) HEAD_BWD_CASE(64) default: mvn_set_error("head: unsupported skip_channels"); return -1; }
    int rc = mvn_check_launch("head_bwd_tc");
    if (rc) return rc;
    head_reduce_kernel<<<(HPART + 255) / 256, 256, 0, st>>>(partial, grid, pg, P, g.S);
    return mvn_check_launch("head_reduce");
}
