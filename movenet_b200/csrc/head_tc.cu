// DenseConv head (movenet/modules.py:133-142) + drop-last + softmax (movenet/wavenet.py:183-191) and its
// backward on tcgen05 tensor cores, for input_channels A == 64 or 128 (template parameter).
//
// One thread owns one time row (TMEM lane = time) and 32 of the A channels (A/32 groups of 128 threads share
// the rows), so the softmax needs no shuffles -- only a tiny shared-memory exchange between the groups -- and
// the channels-first (B, A, T) API tensors are read/written with a warp's 32 lanes on 32 consecutive time
// steps: coalesced without a transpose.
//
// forward : a1 = W1 lrelu(skip) + b1 on CUDA cores (K = S is tiny) -> lrelu -> bf16 tiles (K-major,
//           128B swizzle) -> tcgen05.mma z = . W2^T -> softmax -> probabilities (or logits).
// backward: recompute a1; dz = p (dp - <dp, p>) in-thread -> bf16 tiles -> tcgen05.mma da1 = dz . W2
//           (B = the SAME W2 image read MN-major) -> lrelu' -> dskip on CUDA cores;
//           weight / bias gradients with K = time, accumulated in TMEM over the CTA's tiles, written as
//           per-CTA partials and reduced in a fixed order:
//             A == 64 : one chain  [dz | da1]^T . [lrelu(a1) | lrelu(skip)]   (M = 128 = 64 + 64)
//             A == 128: two chains  dz^T . lrelu(a1)   and   da1^T . lrelu(skip)
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
#include "layer_tc.h"

using namespace tc;

namespace {

// Instrumented build (MOVENET_B200_NVCC_EXTRA=-DMVN_PHASE_CLOCKS=1): clock64() stamps of tiles 2..4 of CTA 0, thread 0, in the
// head forward (role 0) and backward (role 1) kernels; printed by the host wrappers when MVN_PROF is set.
#ifndef MVN_PHASE_CLOCKS
#define MVN_PHASE_CLOCKS 0
#endif
#if MVN_PHASE_CLOCKS
__device__ unsigned long long g_clkh[2][3][12];
#define CLK(role, i) do { if (blockIdx.x == 0 && tid == 0 && it >= 2 && it < 5) g_clkh[role][it - 2][i] = clock64(); } while (0)
static void print_clocks(int role, const char* name) {
    if (!getenv("MVN_PROF")) return;
    cudaDeviceSynchronize();
    unsigned long long h[2][3][12];
    cudaMemcpyFromSymbol(h, g_clkh, sizeof(h));
    for (int i = 0; i < 3; ++i) {
        fprintf(stderr, "CLKH %s tile%d:", name, i + 2);
        for (int j = 0; j < 12; ++j) fprintf(stderr, " %lld", h[role][i][j] ? (long long)(h[role][i][j] - h[role][0][0]) : -1LL);
        fprintf(stderr, "\n");
    }
}
#else
#define CLK(role, i) do {} while (0)
static void print_clocks(int, const char*) {}
#endif

template <int A> struct HP {      // per-CTA partial gradients
    // A == 64 : D_w[128][128] (rows dz|da1, cols lrelu(a1)|lrelu(skip)) + 128 bias sums
    // A == 128: D_w2[128][128] (dz x lrelu(a1)) | D_w1[128][64] (da1 x lrelu(skip)) | 256 bias sums (dz | da1)
    static constexpr int floats = A == 64 ? 128 * 128 + 128 : 128 * 128 + 128 * 64 + 256;
};

struct HeadArgs {
    const float* w1p;    // [S][A] fp32
    const float* b1;     // [A]
    const float* b2;     // [A]
    const void* img;     // W2 image: A/64 chunks of [A n][64 k] bf16, K-major, 128B swizzle
    const float* skip;   // (B, Tout, S)
    float* out;          // forward: (B, A, Tn)
    const float* probs;  // backward
    const float* dout;   // backward: d(out), or null when the loss gradient is formed in the kernel:
    const long long* target;   //   (B, Tn) class indices and
    const float* gloss;        //   d(loss) (1 float): d(out) = gloss / (B Tn) * (softmax_c(probs) - onehot(target)),
                               //   the backward of the trainer's cross_entropy(probabilities, target) (loss.py)
    float* dskip;        // backward: (B, Tout, S)
    float* partial;      // backward
    int B, Tout, Tn, S, logits, tiles_per_clip, n_tiles;
};

__device__ __forceinline__ float lrelu(float v) { return v > 0.f ? v : MVN_LRELU_SLOPE * v; }

// a1[n] for n in [n0, n0+32) of one row, from the row's S skip values (already leaky-ReLU'd)

// 32 fp32 values of row r, channels [32*part, +32) -> bf16 chunks of the 128B-swizzled [128 x 64] tile(s) at `tiles`
__device__ __forceinline__ void store_part_row(uint8_t* tiles, int r, int part, const float* v) {
    uint8_t* tile = tiles + (part >> 1) * TILE_BYTES;
    const int sw = r & 7, half = part & 1;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        *(uint4*)(tile + r * 128 + (((4 * half + q) ^ sw) << 4)) =
            make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                       pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
}

// The first 1x1 conv (K = S <= 64) also runs on the tensor core.  Its input lrelu(skip_sum) is fp32 and is NOT rounded to
// bf16: the row is split into a bf16 head and a bf16 remainder, hi = bf16(x), lo = bf16(x - hi), stored side by side in one
// [128 x 64] tile (columns [0, S) and [LO, LO + S), LO = max(16, S)), and multiplied by an image that holds W1 twice, so the
// product sees x to 2^-17.  (S = 64 has no room for the remainder: plain bf16 there.)
template <int S> struct SkipTile {
    static constexpr int LO = S < 16 ? 16 : S;
    static constexpr bool SPLIT = 2 * LO <= 64;
    static constexpr int K = SPLIT ? 2 * LO : (S < 16 ? 16 : S);     // contraction length of the a1 GEMM (multiple of 16)
};
// row r of the skip tile from this thread's S fp32 values
template <int S>
__device__ __forceinline__ void store_skip_row(uint8_t* tile, int r, const float* ls) {
    const int sw = r & 7;
#pragma unroll
    for (int s = 0; s < S; s += 8) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            hi[e] = pack_bf16(ls[s + 2 * e], ls[s + 2 * e + 1]);
            const float2 h = unpack_bf16(hi[e]);
            lo[e] = pack_bf16(ls[s + 2 * e] - h.x, ls[s + 2 * e + 1] - h.y);
        }
        *(uint4*)(tile + r * 128 + (((s >> 3) ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        if (SkipTile<S>::SPLIT)
            *(uint4*)(tile + r * 128 + ((((SkipTile<S>::LO + s) >> 3) ^ sw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}
// W1 image for a1 = skip_tile . W1^T: [A rows n][64 k] bf16 K-major, 128B swizzle; k in [0,S) and [LO, LO+S) hold w1p[k][n]
template <int A, int S>
__device__ __forceinline__ void build_w1_image(uint8_t* img, const float* sw1, int tid, int nt) {
    for (int i = tid; i < A * 64; i += nt) {
        const int n = i >> 6, k = i & 63;
        const int s = k < S ? k : (SkipTile<S>::SPLIT && k >= SkipTile<S>::LO && k < SkipTile<S>::LO + S ? k - SkipTile<S>::LO : -1);
        const float v = s >= 0 ? sw1[s * A + n] : 0.f;
        *(__nv_bfloat16*)(img + n * 128 + ((((k >> 3) ^ (n & 7)) << 4) | ((k & 7) << 1))) = __float2bfloat16(v);
    }
}

template <int A, int S>
__global__ void __launch_bounds__(A * 4, A == 64 ? 4 : 1) head_fwd_tc_kernel(const HeadArgs a) {
    MVN_PDL_PROLOGUE();
    constexpr int PARTS = A / 32, KC = A / 64, NT = A * 4, W2_BYTES = A * A * 2;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sW2 = smem;
    uint8_t* sA = smem + W2_BYTES;                // KC tiles: lrelu(a1)
    uint8_t* sLS = sA + KC * TILE_BYTES;          // skip tile (hi | lo)
    uint8_t* sW1 = sLS + TILE_BYTES;              // W1 image [A][64]
    float* sw1 = (float*)(sW1 + A * 128);         // [S][A] (only to build the image)
    float* sb1 = sw1 + S * A;
    float* sb2 = sb1 + A;
    float* sx = sb2 + A;                          // [PARTS][128] softmax exchange
    uint64_t* mma_bar = (uint64_t*)(sx + PARTS * 128);
    uint64_t* a1_bar = mma_bar + 1;
    uint32_t* tmem_slot = (uint32_t*)(mma_bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5, r = tid & 127, part = tid >> 7, n0 = 32 * part;

    for (int i = tid; i < W2_BYTES / 16; i += NT) ((uint4*)sW2)[i] = ((const uint4*)a.img)[i];
    for (int i = tid; i < S * A; i += NT) sw1[i] = a.w1p[i];
    for (int i = tid; i < TILE_BYTES / 16; i += NT) ((uint4*)sLS)[i] = make_uint4(0, 0, 0, 0);
    if (tid < A) { sb1[tid] = a.b1[tid]; sb2[tid] = a.b2[tid]; }
    if (tid == 0) { mbar_init(mma_bar, 1); mbar_init(a1_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    build_w1_image<A, S>(sW1, sw1, tid, NT);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(2 * A) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t idesc = umma_idesc_major(TILE_T, A, 0, 0);
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    constexpr int A1_COL = A;                     // TMEM: z [0, A) | a1 pre-activation [A, 2A)

    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int b = tile / a.tiles_per_clip, j = (tile - b * a.tiles_per_clip) * TILE_T + r;
        const bool live = j < a.Tn;
        CLK(0, 0);
        if (part == 0) {             // one thread per row builds the skip tile
            {   // this CTA's next tile: its skip rows towards L2 now (they are the head of that tile's dependency chain)
                const int nt = tile + gridDim.x;
                if (nt < a.n_tiles) {
                    const int nb = nt / a.tiles_per_clip, nj = (nt - nb * a.tiles_per_clip) * TILE_T + r;
                    if (nj < a.Tn) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.skip + ((size_t)nb * a.Tout + nj) * S));
                }
            }
            float ls[S];
            const float* src = a.skip + ((size_t)b * a.Tout + (live ? j : 0)) * S;
#pragma unroll
            for (int s = 0; s < S; s += 4) {
                const float4 v = live ? *(const float4*)(src + s) : make_float4(0.f, 0.f, 0.f, 0.f);
                ls[s] = lrelu(v.x); ls[s + 1] = lrelu(v.y); ls[s + 2] = lrelu(v.z); ls[s + 3] = lrelu(v.w);
            }
            store_skip_row<S>(sLS, r, ls);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (warp_u == 0) {             // a1_pre = skip_tile . W1^T
            tc_fence_after();
            const uint64_t kLS = umma_desc(smem_u32(sLS)), kW1 = umma_desc(smem_u32(sW1));
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < SkipTile<S>::K / 16; ++k)
                    umma(tmem_u + A1_COL, desc_adv(kLS, k * 32), desc_adv(kW1, k * 32), idesc, k != 0);
                umma_commit(a1_bar);
            }
            __syncwarp();
        }
        mbar_wait(a1_bar, it & 1);
        tc_fence_after();
        float v[32];
        {
            uint32_t p0[16], p1[16];
            tmem_ld16(tmem + lane_base + A1_COL + n0, p0);
            tmem_ld16(tmem + lane_base + A1_COL + n0 + 16, p1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                v[i] = lrelu(__uint_as_float(p0[i]) + sb1[n0 + i]);
                v[16 + i] = lrelu(__uint_as_float(p1[i]) + sb1[n0 + 16 + i]);
            }
        }
        CLK(0, 1);
        store_part_row(sA, r, part, v);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        CLK(0, 2);
        if (warp_u == 0) {             // warp-uniform issue through one elected lane (tc_common.cuh)
            tc_fence_after();
            const uint64_t kA = umma_desc(smem_u32(sA)), kW2 = umma_desc(smem_u32(sW2));
            if (elect_one()) {
#pragma unroll
                for (int kc = 0; kc < KC; ++kc)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma(tmem_u, desc_adv(kA, kc * TILE_BYTES + k * 32), desc_adv(kW2, kc * A * 128 + k * 32), idesc, (kc | k) != 0);
                umma_commit(mma_bar);
            }
            __syncwarp();
        }
        CLK(0, 3);
        mbar_wait(mma_bar, it & 1);
        CLK(0, 4);
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
        CLK(0, 5);
        if (!a.logits) {      // softmax over all A channels: a row's channel groups live in threads r, r + 128, ...
            sx[part * 128 + r] = m;
            __syncthreads();
#pragma unroll
            for (int p = 0; p < PARTS; ++p) m = fmaxf(m, sx[p * 128 + r]);
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) { v[i] = __expf(v[i] - m); sum += v[i]; }
            __syncthreads();
            sx[part * 128 + r] = sum;
            __syncthreads();
            float tot = 0.f;
#pragma unroll
            for (int p = 0; p < PARTS; ++p) tot += sx[p * 128 + r];
            const float inv = 1.f / tot;
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] *= inv;
        }
        CLK(0, 6);
        if (live) {
            float* dst = a.out + ((size_t)b * A + n0) * a.Tn + j;
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[(size_t)i * a.Tn] = v[i];
        }
        CLK(0, 7);
        tc_fence_before();
        __syncthreads();       // every thread has read its TMEM row and the A tiles are free again
        CLK(0, 8);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * A) : "memory");
    }
}

template <int A, int S>
__global__ void __launch_bounds__(A * 4, A == 64 ? 2 : 1) head_bwd_tc_kernel(const HeadArgs a) {
    MVN_PDL_PROLOGUE();
    constexpr int PARTS = A / 32, KC = A / 64, NT = A * 4, W2_BYTES = A * A * 2;
    constexpr int TMEM_COLS = A == 64 ? 256 : 512;
    // TMEM columns: a1_pre, then da_pre [0, A) | weight-gradient accumulators | bias sums | dskip_pre [NS]
    constexpr int DA_COL = 0, W_COL = A, W1_COL = 2 * A /* A == 128 only */, B_COL = A == 64 ? 192 : 320;
    constexpr int NS = S < 16 ? 16 : S;           // columns of the dskip GEMM
    constexpr int DSK_COL = A == 64 ? 224 : 480;
    static_assert(DSK_COL + NS <= TMEM_COLS, "no TMEM room for the dskip accumulator");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sW2 = smem;
    uint8_t* sDZ = smem + W2_BYTES;               // A == 64: [DZ | DA] adjacent form the M = 128 operand
    uint8_t* sDA = sDZ + KC * TILE_BYTES;
    uint8_t* sLA = sDA + KC * TILE_BYTES;         // A == 64: [LA | LS] adjacent form the N = 128 operand
    uint8_t* sLS = sLA + KC * TILE_BYTES;
    uint8_t* sONES = sLS + TILE_BYTES;            // 1 KB
    uint8_t* sW1 = sONES + 1024;                  // W1 image [A n][64 k] for a1 = skip_tile . W1^T
    uint8_t* sW1T = sW1 + A * 128;                // W1 as [NS s][A k] (KC chunks) for dskip = da1 . W1
    float* sw1 = (float*)(sW1T + KC * NS * 128);  // [S][A] (only to build the images)
    float* sb1 = sw1 + S * A;
    float* sx = sb1 + A;                          // [3][PARTS][128] exchange: <dp,p> partial sums (fused loss: Z, <e,p>, p[target])
    uint64_t* mma_bar = (uint64_t*)(sx + 3 * PARTS * 128);
    uint64_t* w_bar = mma_bar + 1;
    uint64_t* a1_bar = mma_bar + 2;
    uint64_t* dsk_bar = mma_bar + 3;
    uint32_t* tmem_slot = (uint32_t*)(mma_bar + 4);
    const int tid = threadIdx.x, warp = tid >> 5, r = tid & 127, part = tid >> 7, n0 = 32 * part, sw = r & 7;

    for (int i = tid; i < W2_BYTES / 16; i += NT) ((uint4*)sW2)[i] = ((const uint4*)a.img)[i];
    for (int i = tid; i < S * A; i += NT) sw1[i] = a.w1p[i];
    if (tid < A) sb1[tid] = a.b1[tid];
    for (int i = tid; i < TILE_BYTES / 16; i += NT) ((uint4*)sLS)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 256; i += NT) ((uint32_t*)sONES)[i] = 0x3F803F80u;
    if (tid == 0) {
        mbar_init(mma_bar, 1); mbar_init(w_bar, 1); mbar_init(a1_bar, 1); mbar_init(dsk_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    build_w1_image<A, S>(sW1, sw1, tid, NT);
    for (int i = tid; i < KC * NS * 64; i += NT) {      // [kc][s][k]: w1p[s][64 kc + k], rows s >= S are zero
        const int kc = i / (NS * 64), s_ = (i / 64) % NS, k = i & 63;
        const float v = s_ < S ? sw1[s_ * A + 64 * kc + k] : 0.f;
        *(__nv_bfloat16*)(sW1T + kc * NS * 128 + s_ * 128 + ((((k >> 3) ^ (s_ & 7)) << 4) | ((k & 7) << 1))) = __float2bfloat16(v);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t iG = umma_idesc_major(TILE_T, A, 0, 1);         // da_pre = dz . W2 : B MN-major
    const uint32_t iW = umma_idesc_major(TILE_T, 128, 1, 1);
    const uint32_t iW1 = umma_idesc_major(TILE_T, 64, 1, 1);
    const uint32_t iB = umma_idesc_major(TILE_T, 16, 1, 1);
    const uint32_t iA1 = umma_idesc_major(TILE_T, A, 0, 0);        // a1_pre = skip_tile . W1^T
    const uint32_t iDS = umma_idesc_major(TILE_T, NS, 0, 0);       // dskip_pre = da1 . W1
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);

    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int b = tile / a.tiles_per_clip, j = (tile - b * a.tiles_per_clip) * TILE_T + r;
        const bool live = j < a.Tn;
        {   // the next tile's probabilities / d(out) rows: towards L2 now (32 channel rows, one 128-byte line per warp each)
            // (one instruction per warp: lane i asks for channel row n0 + i, each line covers the warp's 32 time steps)
            const int nt = tile + gridDim.x;
            if (nt < a.n_tiles) {
                const int nb = nt / a.tiles_per_clip, nj = (nt - nb * a.tiles_per_clip) * TILE_T + (r & ~31);
                if (part == 0 && nj + (tid & 31) < a.Tn)       // ... and its skip rows (one 32 S-byte row per thread of the first group)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(a.skip + ((size_t)nb * a.Tout + nj + (tid & 31)) * S));
                if (nj < a.Tn) {
                    const size_t o = ((size_t)nb * A + n0 + (tid & 31)) * a.Tn + nj;
                    // (Tn is not a multiple of 32: the 32 time steps straddle two 128-byte lines)
                    if (!a.logits) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a.probs + o)); asm volatile("prefetch.global.L2 [%0];" ::"l"(a.probs + o + 31)); }
                    if (a.dout) { asm volatile("prefetch.global.L2 [%0];" ::"l"(a.dout + o)); asm volatile("prefetch.global.L2 [%0];" ::"l"(a.dout + o + 31)); }
                }
            }
        }
        // ---- loads first (their latency overlaps the previous tile's weight-gradient MMAs) ----------
        CLK(1, 0);
        unsigned long long skip_pos = 0;
        if (it) { mbar_wait(w_bar, (it - 1) & 1); tc_fence_after(); }      // previous tile's MMAs are done with the tiles
        if (part == 0) {             // one thread per row: the skip tile (hi | lo) for the a1 GEMM and the weight gradient
            float ls[S];
            const float* src = a.skip + ((size_t)b * a.Tout + (live ? j : 0)) * S;
#pragma unroll
            for (int s = 0; s < S; s += 4) {
                const float4 v = live ? *(const float4*)(src + s) : make_float4(0.f, 0.f, 0.f, 0.f);
                const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) { skip_pos |= (unsigned long long)(x[e] > 0.f) << (s + e); ls[s + e] = lrelu(x[e]); }
            }
            store_skip_row<S>(sLS, r, ls);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (warp_u == 0) {             // a1_pre = skip_tile . W1^T into the (still free) da_pre columns; runs under the loads below
            tc_fence_after();
            const uint64_t kLS = umma_desc(smem_u32(sLS)), kW1 = umma_desc(smem_u32(sW1));
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < SkipTile<S>::K / 16; ++k)
                    umma(tmem_u + DA_COL, desc_adv(kLS, k * 32), desc_adv(kW1, k * 32), iA1, k != 0);
                umma_commit(a1_bar);
            }
            __syncwarp();
        }
        float dz[32];
        {
            const size_t o = ((size_t)b * A + n0) * a.Tn + (live ? j : 0);
            float dot = 0.f;
            float p[32];
            if (a.dout) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    dz[i] = live ? a.dout[o + (size_t)i * a.Tn] : 0.f;
                    p[i] = (live && !a.logits) ? a.probs[o + (size_t)i * a.Tn] : 0.f;
                    dot = fmaf(dz[i], p[i], dot);
                }
            } else {
                // fused loss gradient: d(out)_i = gs (e_i / Z - [i == target]), e = exp(probs) (probabilities: no overflow);
                // the three row sums it needs travel through one exchange
                const int tg = live ? (int)a.target[(size_t)b * a.Tn + j] - n0 : -1;
                float z = 0.f, ep = 0.f, pt = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    p[i] = live ? a.probs[o + (size_t)i * a.Tn] : 0.f;
                    dz[i] = __expf(p[i]);
                    z += dz[i];
                    ep = fmaf(dz[i], p[i], ep);
                    if (i == tg) pt = p[i];
                }
                sx[part * 128 + r] = z; sx[(PARTS + part) * 128 + r] = ep; sx[(2 * PARTS + part) * 128 + r] = pt;
                __syncthreads();
                z = 0.f; ep = 0.f; pt = 0.f;
#pragma unroll
                for (int q = 0; q < PARTS; ++q) { z += sx[q * 128 + r]; ep += sx[(PARTS + q) * 128 + r]; pt += sx[(2 * PARTS + q) * 128 + r]; }
                const float gs = live ? a.gloss[0] / ((float)a.B * (float)a.Tn) : 0.f, inv = gs / z;
                dot = ep * inv - gs * pt;
#pragma unroll
                for (int i = 0; i < 32; ++i) dz[i] = p[i] * (dz[i] * inv - (i == tg ? gs : 0.f) - dot);
            }
            if (a.dout && !a.logits) {
                sx[part * 128 + r] = dot;
                __syncthreads();
                dot = 0.f;
#pragma unroll
                for (int q = 0; q < PARTS; ++q) dot += sx[q * 128 + r];
#pragma unroll
                for (int i = 0; i < 32; ++i) dz[i] = p[i] * (dz[i] - dot);
            }
        }
        CLK(1, 1);
        float a1[32];
        uint32_t a1_pos = 0;
        mbar_wait(a1_bar, it & 1);
        tc_fence_after();
        {
            uint32_t p0[16], p1[16];
            tmem_ld16(tmem + lane_base + DA_COL + n0, p0);
            tmem_ld16(tmem + lane_base + DA_COL + n0 + 16, p1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                a1[i] = __uint_as_float(p0[i]) + sb1[n0 + i];
                a1[16 + i] = __uint_as_float(p1[i]) + sb1[n0 + 16 + i];
            }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) { a1_pos |= (uint32_t)(a1[i] > 0.f) << i; a1[i] = lrelu(a1[i]); }
        CLK(1, 2);
        CLK(1, 3);
        store_part_row(sDZ, r, part, dz);
        store_part_row(sLA, r, part, a1);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (warp_u == 0) {
            tc_fence_after();
            const uint64_t kDZ = umma_desc(smem_u32(sDZ)), mW2 = umma_desc_mn(smem_u32(sW2), A * 128);
            if (elect_one()) {
                // da_pre[t][k] = sum_n dz[t][n] W2[n][k] : contraction over the image's ROWS, 16 per step
#pragma unroll
                for (int s = 0; s < A / 16; ++s)
                    umma(tmem_u + DA_COL, desc_adv(kDZ, (s >> 2) * TILE_BYTES + (s & 3) * 32), desc_adv(mW2, s * 2048), iG, s != 0);
                umma_commit(mma_bar);
            }
            __syncwarp();
        }
        CLK(1, 4);
        mbar_wait(mma_bar, it & 1);
        CLK(1, 5);
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
        store_part_row(sDA, r, part, da);
        CLK(1, 6);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        CLK(1, 7);
        if (warp_u == 0) {
            tc_fence_after();
            const uint32_t acc0 = it != 0;
            const uint64_t ones = umma_desc_mn_plain(smem_u32(sONES), 256, 128);
            const uint64_t mDZ = umma_desc_mn(smem_u32(sDZ), TILE_BYTES), mLA = umma_desc_mn(smem_u32(sLA), TILE_BYTES),
                           mDA = umma_desc_mn(smem_u32(sDA), TILE_BYTES), mLS = umma_desc_mn(smem_u32(sLS), TILE_BYTES);
            const uint64_t kDA = umma_desc(smem_u32(sDA)), kW1T = umma_desc(smem_u32(sW1T));
            if (elect_one()) {
                // dskip_pre[t][s] = sum_a da1[t][a] W1[a][s]  (first: one thread per row is waiting for it)
#pragma unroll
                for (int kc = 0; kc < KC; ++kc)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma(tmem_u + DSK_COL, desc_adv(kDA, kc * TILE_BYTES + k * 32), desc_adv(kW1T, kc * NS * 128 + k * 32), iDS, (kc | k) != 0);
                umma_commit(dsk_bar);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t acc = acc0 | (k != 0);
                    if constexpr (A == 64) {          // [dz|da1]^T . [la|ls]
                        umma(tmem_u + W_COL, desc_adv(mDZ, k * 2048), desc_adv(mLA, k * 2048), iW, acc);
                        umma(tmem_u + B_COL, desc_adv(mDZ, k * 2048), ones, iB, acc);
                    } else {                          // dz^T . la ; da1^T . ls ; bias sums of both
                        umma(tmem_u + W_COL, desc_adv(mDZ, k * 2048), desc_adv(mLA, k * 2048), iW, acc);
                        umma(tmem_u + W1_COL, desc_adv(mDA, k * 2048), desc_adv(mLS, k * 2048), iW1, acc);
                        umma(tmem_u + B_COL, desc_adv(mDZ, k * 2048), ones, iB, acc);
                        umma(tmem_u + B_COL + 16, desc_adv(mDA, k * 2048), ones, iB, acc);
                    }
                }
                umma_commit(w_bar);
            }
            __syncwarp();
        }
        mbar_wait(dsk_bar, it & 1);
        tc_fence_after();
        if (part == 0) {             // dskip = dskip_pre * lrelu'(skip): one thread per row (warp-uniform: TMEM loads are collective)
            float* dst = a.dskip + ((size_t)b * a.Tout + (live ? j : 0)) * S;
#pragma unroll
            for (int s = 0; s < S; s += 8) {
                uint32_t v[8];
                tmem_ld8(tmem + lane_base + DSK_COL + s, v);
                tmem_ld_wait();
                if (live) {
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[e]) * ((skip_pos >> (s + e)) & 1 ? 1.f : MVN_LRELU_SLOPE);
                    *(float4*)(dst + s) = make_float4(o[0], o[1], o[2], o[3]);
                    *(float4*)(dst + s + 4) = make_float4(o[4], o[5], o[6], o[7]);
                }
            }
        }
        CLK(1, 8);
        tc_fence_before();
        __syncthreads();           // sx is reused by the next tile; the dskip accumulator has been read
        CLK(1, 9);
    }
    if (it) { mbar_wait(w_bar, (it - 1) & 1); }
    tc_fence_after();
    float* part_out = a.partial + (size_t)blockIdx.x * HP<A>::floats;
    // D_w (A == 64) / D_w2 (A == 128): 128 columns
#pragma unroll 1
    for (int jj = part; jj < 8; jj += PARTS) {
        uint32_t v[16];
        tmem_ld16(tmem + lane_base + W_COL + 16 * jj, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q)
            ((float4*)(part_out + (size_t)r * 128 + 16 * jj))[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                                              __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
    }
    if constexpr (A == 128) {
#pragma unroll 1
        for (int jj = part; jj < 4; jj += PARTS) {
            uint32_t v[16];
            tmem_ld16(tmem + lane_base + W1_COL + 16 * jj, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q)
                ((float4*)(part_out + 128 * 128 + (size_t)r * 64 + 16 * jj))[q] =
                    make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        }
    }
    {
        uint32_t v[8], v2[8];
        tmem_ld8(tmem + lane_base + B_COL, v);
        tmem_ld8(tmem + lane_base + B_COL + 16, v2);      // A == 128 only: da1 sums (harmless extra read otherwise)
        tmem_ld_wait();
        if (part == 0) {
            if constexpr (A == 64) part_out[128 * 128 + r] = __uint_as_float(v[0]);
            else { part_out[128 * 128 + 128 * 64 + r] = __uint_as_float(v[0]); part_out[128 * 128 + 128 * 64 + 128 + r] = __uint_as_float(v2[0]); }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
    }
}

template <int A>
__global__ void head_reduce_kernel(const float* __restrict__ partial, int n_cta, float* __restrict__ pg, PackedLayout P, int S) {
    MVN_PDL_PROLOGUE();
    constexpr int HPART = HP<A>::floats;
    {
        const int i = blockIdx.x * 32 + threadIdx.x;
        float* dst = nullptr;
        if (i < HPART) {
        if constexpr (A == 64) {      // D_w[m][n]: m < 64 dz, m >= 64 da1 ; n < 64 lrelu(a1), n >= 64 lrelu(skip)
            if (i < 128 * 128) {
                const int m = i >> 7, n = i & 127;
                if (m < 64 && n < 64) dst = pg + P.w2p + (size_t)n * A + m;                       // dw2p[k = n][n_out = m]
                else if (m >= 64 && n >= 64 && n - 64 < S) dst = pg + P.w1p + (size_t)(n - 64) * A + (m - 64);   // dw1p[s][a]
            } else {
                const int m = i - 128 * 128;
                dst = m < 64 ? pg + P.b2 + m : pg + P.b1 + (m - 64);
            }
        } else {
            if (i < 128 * 128) { const int m = i >> 7, n = i & 127; dst = pg + P.w2p + (size_t)n * A + m; }
            else if (i < 128 * 128 + 128 * 64) { const int q = i - 128 * 128, m = q >> 6, n = q & 63; if (n < S) dst = pg + P.w1p + (size_t)n * A + m; }
            else { const int m = i - 128 * 128 - 128 * 64; dst = m < 128 ? pg + P.b2 + m : pg + P.b1 + (m - 128); }
        }
        }
        const float acc = column_sum(partial, n_cta, HPART, i, dst != nullptr);
        if (dst && threadIdx.y == 0) *dst = acc;
    }
}

// W2 image: chunk kc = k / 64 holds [n][k % 64] = dense_conv.conv2.weight[n][k], bf16, 128B swizzle, A rows per chunk
__global__ void head_pack_kernel(const float* __restrict__ w2, uint8_t* __restrict__ img, int A) {
    MVN_PDL_PROLOGUE();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A * A) return;
    const int n = i / A, k = i % A, kc = k >> 6, kk = k & 63;
    *(__nv_bfloat16*)(img + (size_t)kc * A * 128 + n * 128 + ((((kk >> 3) ^ (n & 7)) << 4) | ((kk & 7) << 1))) = __float2bfloat16(w2[i]);
}

template <int A> int fwd_smem(int S) { return A * A * 2 + (A / 64 + 1) * TILE_BYTES + A * 128 + (S * A + 2 * A + (A / 32) * 128) * 4 + 64 + 1024; }
template <int A> int bwd_smem(int S) {
    return A * A * 2 + (3 * (A / 64) + 1) * TILE_BYTES + 1024 + A * 128 + (A / 64) * (S < 16 ? 16 : S) * 128 + (S * A + A + 3 * (A / 32) * 128 + 2) * 4 + 64 + 1024;
}

template <int A, int S>
int launch_fwd(const HeadArgs& a, int grid, cudaStream_t st) {
    const int smem = fwd_smem<A>(S);
    MVN_CUDA(cudaFuncSetAttribute(head_fwd_tc_kernel<A, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    MVN_CUDA(mvn_launch_pdl(head_fwd_tc_kernel<A, S>, dim3(grid), dim3(A * 4), (size_t)(smem), st, a));
    return 0;
}
template <int A, int S>
int launch_bwd(const HeadArgs& a, int grid, cudaStream_t st) {
    const int smem = bwd_smem<A>(S);
    MVN_CUDA(cudaFuncSetAttribute(head_bwd_tc_kernel<A, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    MVN_CUDA(mvn_launch_pdl(head_bwd_tc_kernel<A, S>, dim3(grid), dim3(A * 4), (size_t)(smem), st, a));
    return 0;
}

}  // namespace

int mvn_tc_head_supported(int A, int S) {
    if (A == 64) return S == 8 || S == 16 || S == 32;
    if (A == 128) return S == 8 || S == 16 || S == 32;
    return 0;
}
size_t mvn_tc_head_partial_bytes() { return (size_t)2 * 148 * HP<128>::floats * 4; }

int mvn_tc_head_pack(const float* w2_ref, float* packed, const PackedLayout& P, int A, cudaStream_t st) {
    MVN_CUDA(mvn_launch_pdl(head_pack_kernel, dim3((A * A + 255) / 256), dim3(256), (size_t)(0), st, w2_ref, (uint8_t*)(packed + P.tc_head), A));
    return mvn_check_launch("head_pack");
}

static void fill_args(HeadArgs& a, const float* packed, const PackedLayout& P, const Geo& g) {
    a.w1p = packed + P.w1p; a.b1 = packed + P.b1; a.b2 = packed + P.b2; a.img = packed + P.tc_head;
    a.B = g.B; a.Tout = g.Tout; a.Tn = g.Tn; a.S = g.S; a.logits = g.logits;
    a.tiles_per_clip = (g.Tn + TILE_T - 1) / TILE_T; a.n_tiles = a.tiles_per_clip * g.B;
}

#define HEAD_DISPATCH(FN, ...)                                                                          \
    do {                                                                                                \
        int rc_ = -1;                                                                                   \
        if (g.A == 64) {                                                                                \
            if (g.S == 8) rc_ = FN<64, 8>(__VA_ARGS__); else if (g.S == 16) rc_ = FN<64, 16>(__VA_ARGS__); \
            else if (g.S == 32) rc_ = FN<64, 32>(__VA_ARGS__);                                          \
        } else if (g.A == 128) {                                                                        \
            if (g.S == 8) rc_ = FN<128, 8>(__VA_ARGS__); else if (g.S == 16) rc_ = FN<128, 16>(__VA_ARGS__); \
            else if (g.S == 32) rc_ = FN<128, 32>(__VA_ARGS__);                                         \
        }                                                                                               \
        if (rc_ < 0) { mvn_set_error("head: unsupported channel counts"); return -1; }                  \
        if (rc_) return rc_;                                                                            \
    } while (0)

int mvn_tc_head_fwd(const float* packed, const PackedLayout& P, const Geo& g, const float* skip, float* out, cudaStream_t st) {
    HeadArgs a; memset(&a, 0, sizeof(a)); fill_args(a, packed, P, g);
    a.skip = skip; a.out = out;
    if (a.n_tiles <= 0) return 0;
    const int per_sm = g.A == 64 ? 4 : 1;
    const int grid = a.n_tiles < per_sm * 148 ? a.n_tiles : per_sm * 148;
    HEAD_DISPATCH(launch_fwd, a, grid, st);
    print_clocks(0, "fwd");
    return mvn_check_launch("head_fwd_tc");
}

int mvn_tc_head_bwd(const float* packed, const PackedLayout& P, const Geo& g, const float* skip, const float* probs,
                    const float* dout, const long long* target, const float* grad_loss, float* dskip, float* pg, float* partial,
                    cudaStream_t st, int defer_reduce) {
    HeadArgs a; memset(&a, 0, sizeof(a)); fill_args(a, packed, P, g);
    a.skip = skip; a.probs = probs; a.dout = dout; a.target = target; a.gloss = grad_loss; a.dskip = dskip; a.partial = partial;
    MVN_REQUIRE(dout || (target && grad_loss && probs && !g.logits), "head backward: d(out) or (target, d(loss)) on probabilities is required");
    if (a.n_tiles <= 0) return 0;
    const int per_sm = g.A == 64 ? 2 : 1;
    const int grid = a.n_tiles < per_sm * 148 ? a.n_tiles : per_sm * 148;
    HEAD_DISPATCH(launch_bwd, a, grid, st);
    print_clocks(1, "bwd");
    int rc = mvn_check_launch("head_bwd_tc");
    if (rc || defer_reduce) return rc;
    return mvn_tc_head_reduce(P, g, pg, partial, st);
}

// fixed-order sum of the per-CTA partials of mvn_tc_head_bwd into the packed gradients (may run on another stream, later)
int mvn_tc_head_reduce(const PackedLayout& P, const Geo& g, float* pg, const float* partial, cudaStream_t st) {
    const int n_tiles = ((g.Tn + TILE_T - 1) / TILE_T) * g.B;      // (the grid mvn_tc_head_bwd launched)
    if (n_tiles <= 0) return 0;
    const int per_sm = g.A == 64 ? 2 : 1;
    const int grid = n_tiles < per_sm * 148 ? n_tiles : per_sm * 148;
    if (g.A == 64) MVN_CUDA(mvn_launch_pdl(head_reduce_kernel<64>, dim3((HP<64>::floats + 31) / 32), dim3(32, RED_SPLIT), (size_t)(0), st, partial, grid, pg, P, g.S));
    else MVN_CUDA(mvn_launch_pdl(head_reduce_kernel<128>, dim3((HP<128>::floats + 31) / 32), dim3(32, RED_SPLIT), (size_t)(0), st, partial, grid, pg, P, g.S));
    return mvn_check_launch("head_reduce");
}
