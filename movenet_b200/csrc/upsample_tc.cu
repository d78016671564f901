// Last level of the learned video upsampler (ConvTranspose1d k = stride = 10, 16000 -> 160000 frames,
// movenet/wavenet.py:100-118,154) on tensor cores, C = 64.  With W re-laid as [c_in][j*C + c_out] the level is
//     ctx[rows = B*16000][640] = u2[rows][64] . W[64][640] + bias          (row-major output == time-major ctx)
// forward : one TMA tile of u2, five 128-wide output blocks through double-buffered TMEM, bf16 TMA stores.
// backward: d(u2) = d(ctx) . W^T  and  dW^T[n][c_in] = sum_rows d(ctx)[row][n] u2[row][c_in] (K = rows) and the
//           bias gradient, from ONE pass over d(ctx): each [128 x 64] tile of it is used K-major for the
//           data gradient and MN-major for the weight / bias gradients; accumulators stay in TMEM across
//           the CTA's tiles.  The forward weight image (K-major [640 n][64 k]) is read MN-major for d(u2).
#include "tc_common.cuh"
#include "layer_tc.h"

using namespace tc;

namespace {

constexpr int UN = 640;                       // 10 * C output columns
constexpr int UPART = UN * 64 + UN;           // per-CTA partial: dW^T[640][64] + bias sums[640]
constexpr int IMG_BYTES = 5 * TILE_BYTES;     // 5 blocks of [128 n][64 k]

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// [rows][cols] bf16 row-major matrix, box {64 cols, 128 rows}, 128B swizzle
int make_wide_map(CUtensorMap* map, const void* ptr, long long rows, int cols) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    MVN_REQUIRE(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p,
                "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)TILE_T};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncFn)p)(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MVN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (wide) failed (%d)", (int)r);
    return 0;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// image: [n][k] = W[k][n] (packed wt layout [C][10C]) as 5 K-major 128B-swizzled blocks of 128 n-rows, then the bias
__global__ void up_pack_kernel(const float* __restrict__ wt, const float* __restrict__ bt, uint8_t* __restrict__ img) {
    MVN_PDL_PROLOGUE();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < UN * 64 + UN; i += gridDim.x * blockDim.x) {
        if (i < UN * 64) {
            const int n = i >> 6, k = i & 63, blk = n >> 7, nr = n & 127;
            *(__nv_bfloat16*)(img + blk * TILE_BYTES + nr * 128 + ((((k >> 3) ^ (nr & 7)) << 4) | ((k & 7) << 1))) =
                __float2bfloat16(wt[(size_t)k * UN + n]);
        } else ((float*)(img + IMG_BYTES))[i - UN * 64] = bt[i - UN * 64];
    }
}

struct UpArgs { const void* img; float* du2; float* partial; long long rows; int n_tiles; int du_bf16; };

__global__ void __launch_bounds__(256, 1)
up_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_u2, const __grid_constant__ CUtensorMap map_ctx, const UpArgs a) {
    MVN_PDL_PROLOGUE();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sImg = smem;
    float* sbias = (float*)(smem + IMG_BYTES);                   // 640 floats (2560 B) -> pad to 3 KB
    uint8_t* sA = smem + IMG_BYTES + 3072;
    uint8_t* sOut = sA + TILE_BYTES;                             // [buf 2][half 2] x 16 KB
    uint64_t* a_bar = (uint64_t*)(sOut + 4 * TILE_BYTES);
    uint64_t* mma_bar = a_bar + 1;                               // [2]
    uint32_t* tmem_slot = (uint32_t*)(a_bar + 3);
    const int tid = threadIdx.x, warp = tid >> 5, r = tid & 127, half = tid >> 7, sw = r & 7;

    if (tid == 0) {
        mbar_init(a_bar, 1); mbar_init(mma_bar, 1); mbar_init(mma_bar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(a_bar, IMG_BYTES + 2560);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem)), "l"(a.img), "r"(IMG_BYTES + 2560), "r"(smem_u32(a_bar)) : "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mbar_wait(a_bar, 0);
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t idesc = umma_idesc_major(TILE_T, 128, 0, 0);
    uint32_t it = 0, ph0 = 0, ph1 = 0, nstore = 0;

    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int row0 = tile * TILE_T;
        if (tid == 0) {
            mbar_expect_tx(a_bar, TILE_BYTES);
            tma_load_2d(sA, &map_u2, a_bar, 0, row0);
        }
        mbar_wait(a_bar, (it + 1) & 1);
        auto issue = [&](int nb) {
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma(tmem + (nb & 1) * 128, umma_desc(smem_u32(sA) + k * 32), umma_desc(smem_u32(sImg + nb * TILE_BYTES) + k * 32), idesc, k != 0);
            umma_commit(mma_bar + (nb & 1));
        };
        if (tid == 0) issue(0);
        for (int nb = 0; nb < 5; ++nb) {
            if (tid == 0 && nb + 1 < 5) issue(nb + 1);          // the other TMEM buffer was drained two blocks ago
            if (nb & 1) { mbar_wait(mma_bar + 1, ph1); ph1 ^= 1; } else { mbar_wait(mma_bar, ph0); ph0 ^= 1; }
            tc_fence_after();
            uint8_t* stage = sOut + ((nstore & 1) * 2 + half) * TILE_BYTES;
            if (tid == 0) tma_wait_read1();                      // the store that last used this staging pair has been read
            __syncthreads();
#pragma unroll 1
            for (int j = 0; j < 4; ++j) {
                uint32_t v[16];
                tmem_ld16(tmem + lane_base + (nb & 1) * 128 + 64 * half + 16 * j, v);
                tmem_ld_wait();
                const float* bb = sbias + nb * 128 + 64 * half + 16 * j;
                uint32_t o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = pack_bf16(__uint_as_float(v[2 * e]) + bb[2 * e], __uint_as_float(v[2 * e + 1]) + bb[2 * e + 1]);
                *(uint4*)(stage + r * 128 + (((2 * j) ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                *(uint4*)(stage + r * 128 + (((2 * j + 1) ^ sw) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tma_store_2d(&map_ctx, sOut + ((nstore & 1) * 2) * TILE_BYTES, nb * 128, row0);
                tma_store_2d(&map_ctx, sOut + ((nstore & 1) * 2 + 1) * TILE_BYTES, nb * 128 + 64, row0);
                tma_commit();
            }
            ++nstore;
        }
    }
    if (tid == 0) tma_wait_all0();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
    }
}

__global__ void __launch_bounds__(256, 1)
up_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_u2, const __grid_constant__ CUtensorMap map_dctx, const UpArgs a) {
    MVN_PDL_PROLOGUE();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sImg = smem;
    uint8_t* sU2 = smem + IMG_BYTES;
    uint8_t* sD = sU2 + TILE_BYTES;                              // [pair buffer 2][chunk 2] x 16 KB
    uint8_t* sONES = sD + 4 * TILE_BYTES;
    uint64_t* full = (uint64_t*)(sONES + 1024);                  // [2] pair buffers
    uint64_t* done = full + 2;                                   // [2]
    uint64_t* u_bar = full + 4;
    uint64_t* fin_bar = full + 5;
    uint32_t* tmem_slot = (uint32_t*)(full + 6);
    const int tid = threadIdx.x, warp = tid >> 5, r = tid & 127, half = tid >> 7;

    if (tid == 0) {
        for (int i = 0; i < 6; ++i) mbar_init(full + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(u_bar, IMG_BYTES);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem)), "l"(a.img), "r"(IMG_BYTES), "r"(smem_u32(u_bar)) : "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < 256; i += 256) ((uint32_t*)sONES)[i] = 0x3F803F80u;
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mbar_wait(u_bar, 0);
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    constexpr int DU_COL = 0, W_COL = 64, B_COL = 64 + 5 * 64;    // du2 [0,64) | dW^T blocks 5 x 64 | bias 5 x 16
    const uint32_t iDU = umma_idesc_major(TILE_T, 64, 0, 1), iW = umma_idesc_major(TILE_T, 64, 1, 1), iB = umma_idesc_major(TILE_T, 16, 1, 1);

    uint32_t it = 0;
    uint32_t fph[2] = {0, 0}, dph[2] = {0, 0};
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int row0 = tile * TILE_T;
        if (tid == 0) {
            auto load_pair = [&](int g) {
                uint8_t* dst = sD + (g & 1) * 2 * TILE_BYTES;
                mbar_expect_tx(full + (g & 1), 2 * TILE_BYTES);
                tma_load_2d(dst, &map_dctx, full + (g & 1), g * 128, row0);
                tma_load_2d(dst + TILE_BYTES, &map_dctx, full + (g & 1), g * 128 + 64, row0);
            };
            // every buffer of the previous tile has been released (its `done` waits below), the u2 tile too
            mbar_expect_tx(u_bar, TILE_BYTES);
            tma_load_2d(sU2, &map_u2, u_bar, 0, row0);
            load_pair(0); load_pair(1);
            mbar_wait(u_bar, (it + 1) & 1);
            for (int g = 0; g < 5; ++g) {
                const int p = g & 1;
                mbar_wait(full + p, fph[p]); fph[p] ^= 1;
                tc_fence_after();
                uint8_t* pair = sD + p * 2 * TILE_BYTES;
                // d(u2)[r][c_in] += sum_n d(ctx)[r][n] W[c_in][n] : A K-major (K = n), B = forward image read MN-major
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma(tmem + DU_COL, umma_desc(smem_u32(pair + q * TILE_BYTES) + k * 32),
                             umma_desc_mn(smem_u32(sImg + g * TILE_BYTES) + (q * 64 + k * 16) * 128, TILE_BYTES), iDU, (g | q | k) != 0);
                // dW^T[n][c_in] += sum_rows d(ctx)[row][n] u2[row][c_in] ; bias: the same A times ones
                const uint32_t acc0 = it != 0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint64_t am = umma_desc_mn(smem_u32(pair) + k * 2048, TILE_BYTES);
                    umma(tmem + W_COL + g * 64, am, umma_desc_mn(smem_u32(sU2) + k * 2048, TILE_BYTES), iW, acc0 | (k != 0));
                    umma(tmem + B_COL + g * 16, am, umma_desc_mn_plain(smem_u32(sONES), 256, 128), iB, acc0 | (k != 0));
                }
                umma_commit(done + p);
                if (g + 2 < 5) { mbar_wait(done + p, dph[p]); dph[p] ^= 1; load_pair(g + 2); }
            }
            umma_commit(fin_bar);
            // release the buffers for the next tile: the last two pairs' MMAs
            mbar_wait(done + 1, dph[1]); dph[1] ^= 1;     // g = 3
            mbar_wait(done + 0, dph[0]); dph[0] ^= 1;     // g = 4
        }
        mbar_wait(fin_bar, it & 1);
        tc_fence_after();
        {   // d(input) rows -> (rows, 64), fp32 or (when the level below also runs here and reads it as its d(output)) bf16
            float* dst = a.du2 + (size_t)(row0 + r) * 64 + 32 * half;
            __nv_bfloat16* dsth = (__nv_bfloat16*)a.du2 + (size_t)(row0 + r) * 64 + 32 * half;
#pragma unroll 1
            for (int j = 0; j < 2; ++j) {
                uint32_t v[16];
                tmem_ld16(tmem + lane_base + DU_COL + 32 * half + 16 * j, v);
                tmem_ld_wait();
                if (row0 + r < a.rows) {
                    if (a.du_bf16) {
                        uint32_t o[8];
#pragma unroll
                        for (int q = 0; q < 8; ++q) o[q] = pack_bf16(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
                        ((uint4*)(dsth + 16 * j))[0] = make_uint4(o[0], o[1], o[2], o[3]);
                        ((uint4*)(dsth + 16 * j))[1] = make_uint4(o[4], o[5], o[6], o[7]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            ((float4*)(dst + 16 * j))[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                                       __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    tc_fence_after();
    float* part = a.partial + (size_t)blockIdx.x * UPART;
    for (int g = 0; g < 5; ++g) {
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
            uint32_t v[16];
            tmem_ld16(tmem + lane_base + W_COL + g * 64 + 32 * half + 16 * j, v);
            tmem_ld_wait();
            float* dst = part + (size_t)(g * 128 + r) * 64 + 32 * half + 16 * j;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                ((float4*)dst)[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        }
        uint32_t vb[8];
        tmem_ld8(tmem + lane_base + B_COL + g * 16, vb);
        tmem_ld_wait();
        if (half == 0) part[UN * 64 + g * 128 + r] = __uint_as_float(vb[0]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

// dwt[c_in][n] = sum_cta part[n][c_in] ; dbt[n] = sum_cta bias[n]
__global__ void up_reduce_kernel(const float* __restrict__ partial, int n_cta, float* __restrict__ dwt, float* __restrict__ dbt) {
    MVN_PDL_PROLOGUE();
    const int i = blockIdx.x * 32 + threadIdx.x;
    const bool valid = i < UPART;
    const float acc = column_sum(partial, n_cta, UPART, i, valid);
    if (valid && threadIdx.y == 0) { if (i < UN * 64) dwt[(size_t)(i & 63) * UN + (i >> 6)] = acc; else dbt[i - UN * 64] = acc; }
}

}  // namespace

int mvn_tc_upsample_supported(int C) { return C == 64; }
size_t mvn_tc_upsample_img_floats() { return (IMG_BYTES + 3072) / 4; }

int mvn_tc_upsample_pack(const float* wt, const float* bt, float* img, cudaStream_t st) {
    MVN_CUDA(mvn_launch_pdl(up_pack_kernel, dim3(64), dim3(256), (size_t)(0), st, wt, bt, (uint8_t*)img));
    return mvn_check_launch("upsample_pack");
}

int mvn_tc_upsample_fwd(const float* img, const void* u2_bf16, void* ctx_bf16, long long rows, cudaStream_t st) {
    CUtensorMap mu, mc; int rc;
    if ((rc = make_wide_map(&mc, ctx_bf16, rows, UN))) return rc;
    if ((rc = make_wide_map(&mu, u2_bf16, rows, 64))) return rc;
    UpArgs a; a.img = img; a.du2 = nullptr; a.partial = nullptr; a.rows = rows; a.n_tiles = (int)((rows + TILE_T - 1) / TILE_T); a.du_bf16 = 0;
    const int smem = IMG_BYTES + 3072 + 5 * TILE_BYTES + 64 + 1024;
    static MvnSmemAttr attr;
    MVN_CUDA(mvn_ensure_smem(up_fwd_tc_kernel, smem, attr));
    const int grid = a.n_tiles < 148 ? a.n_tiles : 148;
    MVN_CUDA(mvn_launch_pdl(up_fwd_tc_kernel, dim3(grid), dim3(256), (size_t)(smem), st, mu, mc, a));
    return mvn_check_launch("upsample_fwd_tc");
}

int mvn_tc_upsample_bwd(const float* img, const void* u2_bf16, const void* dctx_bf16, void* du2, int du_bf16, float* dwt, float* dbt,
                        float* partial, long long rows, cudaStream_t st, int defer_reduce) {
    CUtensorMap mu, md; int rc;
    if ((rc = make_wide_map(&mu, u2_bf16, rows, 64))) return rc;
    if ((rc = make_wide_map(&md, dctx_bf16, rows, UN))) return rc;
    UpArgs a; a.img = img; a.du2 = (float*)du2; a.partial = partial; a.rows = rows; a.n_tiles = (int)((rows + TILE_T - 1) / TILE_T);
    a.du_bf16 = du_bf16;
    const int smem = IMG_BYTES + 5 * TILE_BYTES + 1024 + 64 + 1024;
    static MvnSmemAttr attr;
    MVN_CUDA(mvn_ensure_smem(up_bwd_tc_kernel, smem, attr));
    const int grid = a.n_tiles < 148 ? a.n_tiles : 148;
    MVN_CUDA(mvn_launch_pdl(up_bwd_tc_kernel, dim3(grid), dim3(256), (size_t)(smem), st, mu, md, a));
    if ((rc = mvn_check_launch("upsample_bwd_tc")) || defer_reduce) return rc;
    return mvn_tc_upsample_reduce(dwt, dbt, partial, rows, st);
}

int mvn_tc_upsample_reduce(float* dwt, float* dbt, const float* partial, long long rows, cudaStream_t st) {
    const int n_tiles = (int)((rows + TILE_T - 1) / TILE_T), grid = n_tiles < 148 ? n_tiles : 148;      // (the grid of the backward kernel)
    MVN_CUDA(mvn_launch_pdl(up_reduce_kernel, dim3((UPART + 31) / 32), dim3(32, RED_SPLIT), (size_t)(0), st, partial, grid, dwt, dbt));
    return mvn_check_launch("upsample_reduce");
}
