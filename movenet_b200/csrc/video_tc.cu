// The video encoder on tensor cores (bf16 mode, C = 64): Conv3d with a (1, 64, 64) kernel = one 4096 Cin -> C linear map per
// frame (movenet/wavenet.py:94-98,152), i.e.  enc[B 160 rows][64] = video[rows][K = 4096 Cin] . Wv[K][64] + b  and its weight
// gradient  dWv[K][64] = video^T . d(enc).  Few rows (480 at the benchmark's three clips), long K: both are cut into
// [128 rows x 128 k] tiles of the fp32 video, one CTA each (forward: 4 row tiles x 32 K slices = 128 CTAs), converted to bf16
// on the way into shared memory in the 128-byte-swizzled layout that is BOTH the K-major [M = rows, K = k] operand of the
// forward and the MN-major [K = rows, M = k] operand of the weight gradient.  The second operand (the fp32 weight slice
// [128 k x 64 c], or the fp32 d(enc) tile [128 rows x 64 c]) is channel-contiguous = MN-major as it lies in memory.  Every
// CTA writes its fp32 partial product; a second kernel adds the partials in a fixed order (deterministic, no atomics) on top
// of the bias.  Replaces the split-K FFMA kernels of the exact mode (42 + 20 us per step at cfg01).
#include "tc_common.cuh"
#include "layer_tc.h"

using namespace tc;

namespace {

constexpr int VK = 128;                 // k per CTA
constexpr int V_THREADS = 256;

struct VcArgs {
    const float* video;                 // [rows][K] fp32
    const float* other;                 // forward: Wv [K][64] fp32 ; backward: d(enc) [rows][64] fp32
    float* part;                        // forward: [K / 128][rows][64] ; backward: [row splits][K][64]
    int rows, K, m_tiles;
};

// [128 x 64] fp32 block (row stride ld floats, rows >= n_valid read as zero) -> bf16 tile, 128-byte rows, 128B swizzle
__device__ __forceinline__ void stage_tile(uint8_t* tile, const float* src, size_t ld, int n_valid) {
    for (int i = threadIdx.x; i < 128 * 8; i += V_THREADS) {
        const int r = i >> 3, q = i & 7;
        uint4 z = make_uint4(0, 0, 0, 0);
        if (r < n_valid) {
            const float4 lo = *(const float4*)(src + (size_t)r * ld + 8 * q), hi = *(const float4*)(src + (size_t)r * ld + 8 * q + 4);
            z = make_uint4(pack_bf16(lo.x, lo.y), pack_bf16(lo.z, lo.w), pack_bf16(hi.x, hi.y), pack_bf16(hi.z, hi.w));
        }
        *(uint4*)(tile + r * 128 + ((q ^ (r & 7)) << 4)) = z;
    }
}

template <bool BWD>
__global__ void __launch_bounds__(V_THREADS) video_conv_tc_kernel(const VcArgs a) {
    MVN_PDL_PROLOGUE();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = smem;                          // video tile: two [128 rows x 64 k] blocks
    uint8_t* sB = smem + 2 * TILE_BYTES;         // forward: weight slice [128 k x 64 c] ; backward: d(enc) tile [128 rows x 64 c]
    uint64_t* mma_bar = (uint64_t*)(sB + TILE_BYTES);
    uint32_t* tmem_slot = (uint32_t*)(mma_bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(mma_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const int k0 = blockIdx.x * VK;
    // forward: operand A K-major (M = rows), B MN-major ; backward: A MN-major (M = k, two 64-k blocks), B MN-major
    const uint32_t idesc = umma_idesc_major(TILE_T, 64, BWD ? 1 : 0, 1);

    if (!BWD) stage_tile(sB, a.other + (size_t)k0 * 64, 64, 128);
    uint32_t it = 0;
    for (int mt = blockIdx.y; mt < a.m_tiles; mt += gridDim.y, ++it) {
        const int row0 = mt * TILE_T, n_valid = a.rows - row0 < TILE_T ? a.rows - row0 : TILE_T;
        if (it) { mbar_wait(mma_bar, (it - 1) & 1); tc_fence_after(); }       // the previous tile's MMAs are done with the tiles
        stage_tile(sA, a.video + (size_t)row0 * a.K + k0, a.K, n_valid);
        stage_tile(sA + TILE_BYTES, a.video + (size_t)row0 * a.K + k0 + 64, a.K, n_valid);
        if (BWD) stage_tile(sB, a.other + (size_t)row0 * 64, 64, n_valid);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (warp_u == 0) {
            tc_fence_after();
            const uint64_t kA = umma_desc(smem_u32(sA)), mA = umma_desc_mn(smem_u32(sA), TILE_BYTES), mB = umma_desc_mn(smem_u32(sB), TILE_BYTES);
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (BWD) umma(tmem_u, desc_adv(mA, j * 2048), desc_adv(mB, j * 2048), idesc, (it != 0) | (j != 0));
                    else umma(tmem_u, desc_adv(kA, (j >> 2) * TILE_BYTES + (j & 3) * 32), desc_adv(mB, j * 2048), idesc, j != 0);
                }
                umma_commit(mma_bar);
            }
            __syncwarp();
        }
        if (!BWD) break;          // forward: one row tile per CTA (gridDim.y == m_tiles)
    }
    mbar_wait(mma_bar, BWD ? (it - 1) & 1 : 0);
    tc_fence_after();
    {
        // warp w reads TMEM lanes 32 (w % 4) .. + 31 (its sub-partition); warps 0-3 take columns 0-31, warps 4-7 columns 32-63
        const int r = (warp & 3) * 32 + (tid & 31), c0 = (warp >> 2) * 32;
        uint32_t v[32];
        tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0, v);
        tmem_ld_wait();
        float* dst = nullptr;
        if (BWD) dst = a.part + ((size_t)blockIdx.y * a.K + k0 + r) * 64 + c0;
        else if (blockIdx.y * TILE_T + r < a.rows) dst = a.part + ((size_t)blockIdx.x * a.rows + blockIdx.y * TILE_T + r) * 64 + c0;
        if (dst) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                ((float4*)dst)[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64) : "memory");
    }
}

// enc[row][c] = b[c] + sum over the K slices, in slice order; optionally also as bf16 (the tensor-core upsampler's input)
__global__ void __launch_bounds__(256) video_fwd_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bias,
                                                               float* __restrict__ enc, __nv_bfloat16* __restrict__ enc16, int rows, int nk) {
    MVN_PDL_PROLOGUE();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;       // one float4 each
    if (i >= rows * 16) return;
    float4 acc = ((const float4*)bias)[i & 15];
    const float4* p = (const float4*)part + i;
    const size_t slice = (size_t)rows * 16;
    int k = 0;
    for (; k + 8 <= nk; k += 8) {
        float4 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = p[(size_t)(k + e) * slice];
#pragma unroll
        for (int e = 0; e < 8; ++e) { acc.x += v[e].x; acc.y += v[e].y; acc.z += v[e].z; acc.w += v[e].w; }
    }
    for (; k < nk; ++k) { const float4 v = p[(size_t)k * slice]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    ((float4*)enc)[i] = acc;
    if (enc16) ((uint2*)enc16)[i] = make_uint2(pack_bf16(acc.x, acc.y), pack_bf16(acc.z, acc.w));
}

// dWv[k][c] = sum over the row splits, in order ; the last block row: d(bias)[c] = sum over the rows of d(enc), fixed order
__global__ void __launch_bounds__(256) video_bwd_reduce_kernel(const float* __restrict__ part, const float* __restrict__ denc,
                                                               float* __restrict__ dwv, float* __restrict__ dbv, int rows, int K, int ns) {
    MVN_PDL_PROLOGUE();
    if (blockIdx.y == 1) {
        if (blockIdx.x) return;
        __shared__ float red[4][64];
        const int c = threadIdx.x & 63, ph = threadIdx.x >> 6;
        float acc = 0.f;
        for (int r = ph; r < rows; r += 4) acc += denc[(size_t)r * 64 + c];
        red[ph][c] = acc;
        __syncthreads();
        if (ph == 0) dbv[c] = (red[0][c] + red[1][c]) + (red[2][c] + red[3][c]);
        return;
    }
    const size_t n4 = (size_t)K * 16;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 acc = ((const float4*)part)[i];
        for (int s = 1; s < ns; ++s) { const float4 v = ((const float4*)part)[(size_t)s * n4 + i]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        ((float4*)dwv)[i] = acc;
    }
}

constexpr int V_SMEM = 3 * TILE_BYTES + 64 + 1024;

}  // namespace

int mvn_tc_video_supported(int C, int K) { return C == 64 && K % VK == 0; }
// row splits of the weight gradient (each writes one [K][64] fp32 partial)
static int video_row_splits(int rows) { const int m = (rows + TILE_T - 1) / TILE_T; return m < 4 ? m : 4; }
size_t mvn_tc_video_partial_floats(int rows, int K) {
    const size_t f = (size_t)(K / VK) * rows * 64, b = (size_t)video_row_splits(rows) * K * 64;
    return f > b ? f : b;
}

int mvn_tc_video_fwd(const float* video, const float* wv, const float* bv, float* part, float* enc, void* enc16, int rows, int K,
                     cudaStream_t st) {
    VcArgs a; a.video = video; a.other = wv; a.part = part; a.rows = rows; a.K = K; a.m_tiles = (rows + TILE_T - 1) / TILE_T;
    static MvnSmemAttr attr;
    MVN_CUDA(mvn_ensure_smem(video_conv_tc_kernel<false>, V_SMEM, attr));
    MVN_CUDA(mvn_launch_pdl(video_conv_tc_kernel<false>, dim3(K / VK, a.m_tiles), dim3(V_THREADS), (size_t)V_SMEM, st, a));
    int rc = mvn_check_launch("video_conv_tc"); if (rc) return rc;
    MVN_CUDA(mvn_launch_pdl(video_fwd_reduce_kernel, dim3(mvn_cdiv((long long)rows * 16, 256)), dim3(256), (size_t)0, st,
                            (const float*)part, bv, enc, (__nv_bfloat16*)enc16, rows, K / VK));
    return mvn_check_launch("video_fwd_reduce");
}

int mvn_tc_video_reduce(const float* denc, const float* part, float* dwv, float* dbv, int rows, int K, cudaStream_t st) {
    MVN_CUDA(mvn_launch_pdl(video_bwd_reduce_kernel, dim3(2 * mvn_sm_count(), 2), dim3(256), (size_t)0, st, part, denc, dwv,
                            dbv, rows, K, video_row_splits(rows)));
    return mvn_check_launch("video_bwd_reduce");
}

int mvn_tc_video_bwd(const float* video, const float* denc, float* part, float* dwv, float* dbv, int rows, int K, cudaStream_t st,
                     int defer_reduce) {
    VcArgs a; a.video = video; a.other = denc; a.part = part; a.rows = rows; a.K = K; a.m_tiles = (rows + TILE_T - 1) / TILE_T;
    const int ns = video_row_splits(rows);
    static MvnSmemAttr attr;
    MVN_CUDA(mvn_ensure_smem(video_conv_tc_kernel<true>, V_SMEM, attr));
    MVN_CUDA(mvn_launch_pdl(video_conv_tc_kernel<true>, dim3(K / VK, ns), dim3(V_THREADS), (size_t)V_SMEM, st, a));
    int rc = mvn_check_launch("video_wgrad_tc");
    if (rc || defer_reduce) return rc;
    return mvn_tc_video_reduce(denc, part, dwv, dbv, rows, K, st);
}
