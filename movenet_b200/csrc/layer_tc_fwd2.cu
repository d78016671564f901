// Forward residual layer, second generation: one persistent CTA per SM, warp-specialised.
//
//   warp 16 (one lane) : TMA producer   -- a 3-stage ring of {x(t-d), x(t), ctx(t)} tiles, up to 3 tiles ahead
//   warp 17 (one lane) : MMA issuer     -- gate GEMM of tile k, then the residual/skip GEMM of tile k-1
//   warps 0-7, 8-15    : two epilogue groups of 8 warps, each owning every other tile and its own TMEM window
//                        (two warps share a TMEM lane quarter and split the channel range)
//
// so the loads, the tensor-core work and the two epilogues of consecutive tiles all overlap (the first generation
// kept one tile in flight per CTA and relied on two co-resident CTAs to overlap).  Same math, same operand images
// and layouts as layer_tc.cu.
#include "tc_common.cuh"
#include "layer_tc.h"

using namespace tc;

namespace {

constexpr int STAGES = 3;
constexpr int GROUP_COLS = 256;      // TMEM window of one epilogue group: D1 [0,128) | D2 [128, 128 + N2)

struct Fwd2Args {
    const void* img;
    float* skip;
    int B, T, Tout, RF, S, N2, dil, nchunks, has_out, skip_init, tiles_per_clip, n_tiles;
};

__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 256;" ::"r"(g + 1) : "memory"); }

__global__ void __launch_bounds__(576, 1)
layer_fwd_tc2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_ctx,
                     const __grid_constant__ CUtensorMap map_out, const Fwd2Args a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int nc = a.nchunks;
    uint8_t* sBz = smem;
    uint8_t* sBrs = smem + smem_brs_off(nc);
    float* sbz = (float*)(smem + smem_bias_off(nc, a.N2));
    float* sbrs = sbz + 128;
    uint8_t* sStage = smem + smem_a_off(nc, a.N2);
    const int stage_bytes = nc * TILE_BYTES;
    uint64_t* bars = (uint64_t*)(sStage + STAGES * stage_bytes);
    uint64_t* full = bars;                 // [3] TMA -> issuer
    uint64_t* empty = bars + 3;            // [3] epilogue leader -> producer (the x' store has finished reading the stage)
    uint64_t* mma1_done = bars + 6;        // [2] issuer -> group
    uint64_t* g_ready = bars + 8;          // [2] group -> issuer (gated tile written)
    uint64_t* mma2_done = bars + 10;       // [2] issuer -> group
    uint64_t* tmem_free = bars + 12;       // [2] group -> issuer (its TMEM window has been drained)
    uint64_t* img_bar = bars + 14;
    uint32_t* tmem_slot = (uint32_t*)(bars + 15);

    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) {
        for (int i = 0; i < 15; ++i) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t wbytes = (uint32_t)smem_a_off(nc, a.N2);
        mbar_expect_tx(img_bar, wbytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem)), "l"(a.img), "r"(wbytes), "r"(smem_u32(img_bar)) : "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int n_mine = a.n_tiles > (int)blockIdx.x ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);     // warp-uniform copy: role branches stay convergent
    if (warp_u == 16) {
        // ================================ TMA producer ============================================
        if ((tid & 31) == 0) {
            for (int k = 0; k < n_mine; ++k) {
                const int tile = blockIdx.x + k * gridDim.x, st = k % STAGES, j = k / STAGES;
                const int b = tile / a.tiles_per_clip, t0 = (tile - b * a.tiles_per_clip) * TILE_T;
                if (j > 0) mbar_wait(empty + st, (j - 1) & 1);
                uint8_t* dst = sStage + st * stage_bytes;
                mbar_expect_tx(full + st, (uint32_t)stage_bytes);
                tma_load_3d(dst, &map_x, full + st, 0, t0 - a.dil, b);
                tma_load_3d(dst + TILE_BYTES, &map_x, full + st, 0, t0, b);
                if (nc == 3) tma_load_3d(dst + 2 * TILE_BYTES, &map_ctx, full + st, 0, t0, b);
            }
        }
    } else if (warp_u == 17) {
        // ================================ MMA issuer ==============================================
        // the whole warp runs the loop (uniform datapath), one elected lane issues; descriptors = base + constant
        {
            const uint32_t idesc1 = umma_idesc(TILE_T, 128), idesc2 = umma_idesc(TILE_T, a.N2);
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
            const uint64_t kBz = umma_desc(smem_u32(sBz)), kBrs = umma_desc(smem_u32(sBrs));
            mbar_wait(img_bar, 0);
            auto mma2 = [&](int k) {
                const int st = k % STAGES, g = k & 1, j = k >> 1;
                const uint64_t kA0 = umma_desc(smem_u32(sStage + st * stage_bytes));
                mbar_wait(g_ready + g, j & 1);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        umma(tmem_u + g * GROUP_COLS + 128, desc_adv(kA0, kk * 32), desc_adv(kBrs, kk * 32), idesc2, kk != 0);
                    umma_commit(mma2_done + g);
                }
                __syncwarp();
            };
            for (int k = 0; k < n_mine; ++k) {
                const int st = k % STAGES, g = k & 1, j = k >> 1;
                const uint64_t kA0 = umma_desc(smem_u32(sStage + st * stage_bytes));
                mbar_wait(full + st, (k / STAGES) & 1);
                if (j > 0) mbar_wait(tmem_free + g, (j - 1) & 1);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            if (c < nc)
                                umma(tmem_u + g * GROUP_COLS, desc_adv(kA0, c * TILE_BYTES + kk * 32), desc_adv(kBz, c * TILE_BYTES + kk * 32),
                                     idesc1, (c | kk) != 0);
                    umma_commit(mma1_done + g);
                }
                __syncwarp();
                if (k > 0) mma2(k - 1);
            }
            if (n_mine > 0) mma2(n_mine - 1);
        }
    } else {
        // ================================ epilogue groups =========================================
        const int g = warp >> 3, r = tid & 127, sw = r & 7, half = (tid >> 7) & 1;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tm = tmem + g * GROUP_COLS;
        mbar_wait(img_bar, 0);                 // biases live in the image
        for (int k = g; k < n_mine; k += 2) {
            const int tile = blockIdx.x + k * gridDim.x, st = k % STAGES, j = k >> 1;
            const int b = tile / a.tiles_per_clip, t0 = (tile - b * a.tiles_per_clip) * TILE_T;
            uint8_t* sA0 = sStage + st * stage_bytes;
            uint8_t* sA1 = sA0 + TILE_BYTES;
            const int t = t0 + r, js = t - (a.RF - 1);
            const bool live = t < a.T && js >= 0 && js < a.Tout;
            float* skip_dst = a.skip + ((size_t)b * a.Tout + (live ? js : 0)) * a.S;
            float4 old0 = make_float4(0.f, 0.f, 0.f, 0.f), old1 = old0;
            if (half == 0 && live && !a.skip_init) { old0 = ((const float4*)skip_dst)[0]; old1 = ((const float4*)skip_dst)[1]; }

            mbar_wait(mma1_done + g, j & 1);
            tc_fence_after();
#pragma unroll 1
            for (int q = 2 * half; q < 2 * half + 2; ++q) {
                uint32_t f[16], gg[16];
                tmem_ld16(tm + lane_base + 16 * q, f);
                tmem_ld16(tm + lane_base + 64 + 16 * q, gg);
                tmem_ld_wait();
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const int c = 16 * q + i;
                    const float f0 = __uint_as_float(f[i]) + sbz[c], f1 = __uint_as_float(f[i + 1]) + sbz[c + 1];
                    const float g0 = __uint_as_float(gg[i]) + sbz[64 + c], g1 = __uint_as_float(gg[i + 1]) + sbz[64 + c + 1];
                    o[i >> 1] = pack_bf16(tanh_fast(f0) * fmaf(0.5f, tanh_fast(0.5f * g0), 0.5f),
                                          tanh_fast(f1) * fmaf(0.5f, tanh_fast(0.5f * g1), 0.5f));
                }
                *(uint4*)(sA0 + r * 128 + (((2 * q) ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                *(uint4*)(sA0 + r * 128 + (((2 * q + 1) ^ sw) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
            }
            fence_proxy_async();
            tc_fence_before();
            group_bar(g);
            if (r == 0 && half == 0) mbar_arrive(g_ready + g);

            mbar_wait(mma2_done + g, j & 1);
            tc_fence_after();
            if (a.has_out) {
#pragma unroll 1
                for (int q = 2 * half; q < 2 * half + 2; ++q) {
                    uint32_t rr[16];
                    tmem_ld16(tm + lane_base + 128 + 16 * q, rr);
                    tmem_ld_wait();
                    uint4* p0 = (uint4*)(sA1 + r * 128 + (((2 * q) ^ sw) << 4));
                    uint4* p1 = (uint4*)(sA1 + r * 128 + (((2 * q + 1) ^ sw) << 4));
                    const uint4 x0 = *p0, x1 = *p1;
                    const uint32_t xi[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                    uint32_t o[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int c = 16 * q + 2 * i;
                        const float2 xv = unpack_bf16(xi[i]);
                        o[i] = pack_bf16(__uint_as_float(rr[2 * i]) + sbrs[c] + xv.x, __uint_as_float(rr[2 * i + 1]) + sbrs[c + 1] + xv.y);
                    }
                    *p0 = make_uint4(o[0], o[1], o[2], o[3]);
                    *p1 = make_uint4(o[4], o[5], o[6], o[7]);
                }
            }
            for (int s0 = 8 * half; s0 < a.S; s0 += 16) {
                uint32_t sv[8];
                tmem_ld8(tm + lane_base + 128 + CC + s0, sv);
                tmem_ld_wait();
                if (live) {
                    float4 v0 = make_float4(__uint_as_float(sv[0]) + sbrs[CC + s0], __uint_as_float(sv[1]) + sbrs[CC + s0 + 1],
                                            __uint_as_float(sv[2]) + sbrs[CC + s0 + 2], __uint_as_float(sv[3]) + sbrs[CC + s0 + 3]);
                    float4 v1 = make_float4(__uint_as_float(sv[4]) + sbrs[CC + s0 + 4], __uint_as_float(sv[5]) + sbrs[CC + s0 + 5],
                                            __uint_as_float(sv[6]) + sbrs[CC + s0 + 6], __uint_as_float(sv[7]) + sbrs[CC + s0 + 7]);
                    float4* d4 = (float4*)(skip_dst + s0);
                    if (!a.skip_init) {
                        float4 p0, p1;
                        if (s0 == 0) { p0 = old0; p1 = old1; } else { p0 = d4[0]; p1 = d4[1]; }
                        v0.x += p0.x; v0.y += p0.y; v0.z += p0.z; v0.w += p0.w;
                        v1.x += p1.x; v1.y += p1.y; v1.z += p1.z; v1.w += p1.w;
                    }
                    d4[0] = v0; d4[1] = v1;
                }
            }
            fence_proxy_async();
            tc_fence_before();
            group_bar(g);
            if (r == 0 && half == 0) {
                mbar_arrive(tmem_free + g);            // every thread of the group has drained its TMEM rows
                if (a.has_out) {
                    tma_store_3d(&map_out, sA1, 0, t0, b);
                    tma_commit();
                    tma_wait_read0();                  // the stage may be refilled once the store has read it
                }
                mbar_arrive(empty + st);
            }
        }
        if (r == 0 && half == 0) tma_wait_all0();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

}  // namespace

int mvn_tc_layer_fwd2(const void* x_in, const void* ctx, void* x_out, float* skip_sum, const float* lw,
                      const PackedLayout& P, const Geo& g, int layer, cudaStream_t st) {
    CUtensorMap map_x, map_ctx, map_out;
    int rc;
    if ((rc = make_act_map(&map_x, x_in, g.B, g.T))) return rc;
    if ((rc = make_act_map(&map_ctx, g.video ? ctx : x_in, g.B, g.T))) return rc;
    if ((rc = make_act_map(&map_out, x_out ? x_out : x_in, g.B, g.T))) return rc;
    Fwd2Args a;
    a.img = lw + P.oTc; a.skip = skip_sum;
    a.B = g.B; a.T = g.T; a.Tout = g.Tout; a.RF = g.RF; a.S = g.S; a.N2 = ((g.C + g.S + 15) / 16) * 16;
    a.dil = g.dil[layer]; a.nchunks = g.video ? 3 : 2; a.has_out = x_out != nullptr; a.skip_init = layer == 0;
    a.tiles_per_clip = (g.T + TILE_T - 1) / TILE_T; a.n_tiles = a.tiles_per_clip * g.B;
    const int smem = smem_a_off(a.nchunks, a.N2) + STAGES * a.nchunks * TILE_BYTES + 256 + 1024;
    MVN_REQUIRE(smem <= 227 * 1024, "forward layer kernel: shared memory budget exceeded (%d)", smem);
    static int attr_smem = 0;
    if (smem > attr_smem) {
        MVN_CUDA(cudaFuncSetAttribute(layer_fwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_smem = smem;
    }
    int grid = 148;
    if (grid > a.n_tiles) grid = a.n_tiles;
    layer_fwd_tc2_kernel<<<grid, 576, smem, st>>>(map_x, map_ctx, map_out, a);
    return mvn_check_launch("layer_fwd_tc2");
}
