// GatedResidualConv1d.forward (movenet/modules.py:67-93) as ONE fused sm_100a kernel:
//
//   TMA loads the time tile of the layer input twice (rows t-d.. and t..: the two taps of the k=2
//   dilated conv are "the same tensor, d rows earlier", out-of-range rows are zero-filled by TMA)
//   plus the context tile  ->  tcgen05.mma  D1[128 x 2C] = [x(t-d) | x(t) | ctx(t)] . Wz^T  (TMEM)
//   -> epilogue 1 (tcgen05.ld): gated = tanh(f) * sigmoid(g), written as bf16 into shared memory in
//   the K-major SWIZZLE_128B layout the tensor core reads  ->  tcgen05.mma  D2[128 x (C+S)] =
//   gated . [Wr | Ws]^T  ->  epilogue 2: x' = r + b_r + x(t) in place in the tap-1 tile, TMA store;
//   skip_sum += s + b_s (fp32, coalesced 32 B per row).
//
// Activations are time-major (B, T, C) bf16, so the time tile is the MMA's M dimension (TMEM lane =
// time row) and both gate halves of one (t, c) sit in the same thread.  One 128-row tile is in
// flight per CTA; two CTAs share an SM (106 KB shared memory, 256 TMEM columns each) so one CTA's
// epilogue overlaps the other's loads and MMAs.
#include <cstdlib>
#include <mutex>
#include "tc_common.cuh"
#include "layer_tc.h"

using namespace tc;

namespace {

struct TcArgs {
    const void* img;      // shared-memory weight image of this layer (mvn_tc_pack)
    float* skip;          // (B, Tout, S) fp32
    int B, T, Tout, RF, S, N2, dil, nchunks, has_out, skip_init, tiles_per_clip, n_tiles;
};

// shared-memory carve-up (dynamic, 1024-byte aligned): weight image (tc_common.cuh) | A tiles | barriers
__host__ __device__ inline int smem_total(int nchunks, int N2) { return smem_a_off(nchunks, N2) + nchunks * TILE_BYTES + 64 + 256; }


// image writer: one block row per layer
__global__ void tc_pack_kernel(const float* const* __restrict__ ptrs, float* __restrict__ packed, PackedLayout P, int S,
                               int nchunks, int N2, int video, int Cl) {
    MVN_PDL_PROLOGUE();
    const int l = blockIdx.y;
    const float* const* lp = ptrs + MVN_PARAM_LAYER(l, 0);
    const float *wf = lp[0], *wg = lp[1], *vf = lp[2], *bvf = lp[3], *vg = lp[4], *bvg = lp[5], *wr = lp[6], *br = lp[7],
                *ws = lp[8], *bs = lp[9];
    uint8_t* img = (uint8_t*)(packed + P.layer0 + (size_t)l * P.layer_stride + P.oTc);
    const int Kz = nchunks * CC, nz = 128 * Kz, nrs = N2 * CC;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nz + nrs + 256; i += gridDim.x * blockDim.x) {
        if (i < nz) {
            const int n = i / Kz, k = i - n * Kz, c = n & 63, gate = n >> 6;
            float v = 0.f;
            const int chunk = k >> 6, kk = k & 63;
            if (c < Cl && kk < Cl) {          // reference tensors have Cl <= 64 channels; the rest of the image is zero
                if (k < CC) v = (gate ? wg : wf)[((size_t)c * Cl + kk) * 2 + 0];
                else if (k < 2 * CC) v = (gate ? wg : wf)[((size_t)c * Cl + kk) * 2 + 1];
                else v = (gate ? vg : vf)[(size_t)c * Cl + kk];
            }
            *(__nv_bfloat16*)(img + chunk * TILE_BYTES + n * 128 + ((((kk >> 3) ^ (n & 7)) << 4) | ((kk & 7) << 1))) = __float2bfloat16(v);
        } else if (i < nz + nrs) {
            const int j = i - nz, n = j >> 6, k = j & 63;
            float v = 0.f;
            if (k < Cl) { if (n < CC) { if (n < Cl) v = wr[(size_t)n * Cl + k]; } else if (n < CC + S) v = ws[(size_t)(n - CC) * Cl + k]; }
            *(__nv_bfloat16*)(img + smem_brs_off(nchunks) + n * 128 + ((((k >> 3) ^ (n & 7)) << 4) | ((k & 7) << 1))) = __float2bfloat16(v);
        } else {
            const int n = i - nz - nrs;
            float* bias = (float*)(img + smem_bias_off(nchunks, N2));
            if (n < 128) bias[n] = (video && (n & 63) < Cl) ? ((n >> 6) ? bvg : bvf)[n & 63] : 0.f;
            else { const int m = n - 128; bias[n] = m < CC ? (m < Cl ? br[m] : 0.f) : (m < CC + S ? bs[m - CC] : 0.f); }
        }
    }
}

// HALF_GATE: epilogue 1 in packed f16x2 (default).  false: the fp32 gate math (MOVENET_B200_GATE_F32=1), kept as the
// reference point for the accuracy of the packed path.
template <bool HALF_GATE>
__global__ void __launch_bounds__(256, 2)
layer_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_ctx,
                    const __grid_constant__ CUtensorMap map_out, const TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sBz = smem;
    uint8_t* sBrs = smem + smem_brs_off(a.nchunks);
    float* sbz = (float*)(smem + smem_bias_off(a.nchunks, a.N2));
    float* sbrs = sbz + 128;
    uint8_t* sA0 = smem + smem_a_off(a.nchunks, a.N2);   // tap t-d; later the gated tile
    uint8_t* sA1 = sA0 + TILE_BYTES;                      // tap t; later x'
    uint8_t* sA2 = sA1 + TILE_BYTES;                      // context (video only)
    uint64_t* full_bar = (uint64_t*)(sA0 + a.nchunks * TILE_BYTES);
    uint64_t* mma_bar = full_bar + 1;
    uint32_t* tmem_slot = (uint32_t*)(full_bar + 2);
    uint32_t* sbh = (uint32_t*)(full_bar + 8);            // gate biases as f16x2 pairs: [0,32) filter, [32,64) gate

    const int tid = threadIdx.x, warp = tid >> 5;
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);     // warp-uniform copy: the issue branches stay convergent

    // ---- one-time setup: barriers, TMEM, and the weight image (one bulk copy) --------------------
    // INVARIANT (programmatic dependent launch): the image copy below is issued BEFORE griddepcontrol.wait, so it is ordered
    // only after kernels that completed before the PREVIOUS kernel started.  The image is written by tc_pack_kernel
    // (mvn_pack_weights); between that kernel and the first layer kernel of a pass there is always at least one full
    // stream-ordered operation that is not a PDL launch (the cudaMemsetAsync of the skip buffer in mvn_wavenet_forward /
    // mvn_layer_fwd, the memsets at the top of the backward), which drains the pack kernels.  Keep such an operation there
    // (or move this copy below the wait) when reordering launches.
    if (tid == 0) {
        mbar_init(full_bar, 1);
        mbar_init(mma_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t wbytes = (uint32_t)smem_a_off(a.nchunks, a.N2);
        mbar_expect_tx(full_bar, wbytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem)), "l"(a.img), "r"(wbytes), "r"(smem_u32(full_bar)) : "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mvn_griddep_launch();
    mvn_griddep_wait();              // everything above overlapped the previous kernel's tail; its output is read from here on
    mbar_wait(full_bar, 0);
    if (HALF_GATE) {
        if (tid < 64) sbh[tid] = f16x2(sbz[2 * tid], sbz[2 * tid + 1]);
        // kind::f16 wants both operands of one MMA in the same format: this CTA's copy of [Wr|Ws] becomes f16 too (same layout)
        for (int i = tid; i < a.N2 * 32; i += 256) {
            const float2 w = unpack_bf16(((const uint32_t*)sBrs)[i]);
            ((uint32_t*)sBrs)[i] = f16x2(w.x, w.y);
        }
        fence_proxy_async();
        __syncthreads();
    }
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    // the out GEMM runs on f16 operands (formats 0): the gated tile is written in f16 by epilogue 1
    const uint32_t idesc1 = umma_idesc(TILE_T, 128),
                   idesc2 = HALF_GATE ? umma_idesc(TILE_T, a.N2) & ~((1u << 7) | (1u << 10)) : umma_idesc(TILE_T, a.N2);
    const uint32_t load_bytes = (uint32_t)(a.nchunks * TILE_BYTES);
    const int r = tid & 127;         // this thread's row of the tile == its TMEM lane (warps w and w+4 share a lane quarter
    const int half = tid >> 7;       // and split the channel range between them)
    const int sw = r & 7;

    // all TMA / MMA instructions are issued by one elected lane of warp 0 from warp-uniform code (descriptors = base + constant)
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    bool issuer = false;
    if (warp_u == 0) issuer = elect_one();

    uint32_t it = 1;            // phase 0 of full_bar was the weight image
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int b = tile / a.tiles_per_clip, t0 = (tile - b * a.tiles_per_clip) * TILE_T;
        if (warp_u == 0) {
          if (issuer) {
            // (the tap t - d and ctx tiles of this tile were requested as soon as the previous tile's out GEMM had released them,
            // see below; only the first tile of the CTA asks for everything here)
            if (it == 1) {
                mbar_expect_tx(full_bar, load_bytes);
                tma_load_3d(sA0, &map_x, full_bar, 0, t0 - a.dil, b);
                if (a.nchunks == 3) tma_load_3d(sA2, &map_ctx, full_bar, 0, t0, b);
            }
            tma_wait_read0();        // the previous tile's x' store must be done reading A1
            tma_load_3d(sA1, &map_x, full_bar, 0, t0, b);
            const int nt = tile + gridDim.x;       // this CTA's next tile: start pulling it into L2 now
            if (nt < a.n_tiles) {
                const int nb = nt / a.tiles_per_clip, n0 = (nt - nb * a.tiles_per_clip) * TILE_T;
                tma_prefetch_3d(&map_x, 0, n0 - a.dil, nb);
                tma_prefetch_3d(&map_x, 0, n0, nb);
                if (a.nchunks == 3) tma_prefetch_3d(&map_ctx, 0, n0, nb);
            }
          }
          __syncwarp();
        }
        // the running skip sum of this row: fetch it now, it is only needed at the very end of the tile
        const int t = t0 + r, js = t - (a.RF - 1);
        const bool live = t < a.T && js >= 0 && js < a.Tout;
        float* skip_dst = a.skip + ((size_t)b * a.Tout + (live ? js : 0)) * a.S;
        float4 old0 = make_float4(0.f, 0.f, 0.f, 0.f), old1 = old0;
        if (half == 0 && live && !a.skip_init) { old0 = ((const float4*)skip_dst)[0]; old1 = ((const float4*)skip_dst)[1]; }
        mbar_wait(full_bar, it & 1);
        if (warp_u == 0) {
            tc_fence_after();
            const uint64_t kA0 = umma_desc(smem_u32(sA0)), kBz = umma_desc(smem_u32(sBz));
            if (elect_one()) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (c < a.nchunks)
                            umma(tmem_u, desc_adv(kA0, c * TILE_BYTES + k * 32), desc_adv(kBz, c * TILE_BYTES + k * 32), idesc1, (c | k) != 0);
                umma_commit(mma_bar);
            }
            __syncwarp();
        }
        mbar_wait(mma_bar, 0);
        tc_fence_after();
        // ---- epilogue 1: gate, bf16, into A0 as the next MMA's A operand ------------------------
        // (this thread's 32 channels in ONE TMEM round trip: all loads in flight, then 32 independent chains)
        {
            uint32_t f[32], g[32];
            tmem_ld32(tmem + lane_base + 32 * half, f);
            tmem_ld32(tmem + lane_base + 64 + 32 * half, g);
            tmem_ld_wait();
            const uint32_t h05 = 0x38003800u;     // (0.5, 0.5)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = 8 * q + 2 * e;
                    if (HALF_GATE) {
                        // tanh(f) * sigmoid(g) on channel pairs in f16x2: the 11-bit intermediate precision is above the
                        // output's; the tile the out GEMM reads is f16
                        const int pc = 16 * half + 4 * q + e;
                        const uint32_t fh = hadd2(f16x2(__uint_as_float(f[i]), __uint_as_float(f[i + 1])), sbh[pc]);
                        const uint32_t gh = hadd2(f16x2(__uint_as_float(g[i]), __uint_as_float(g[i + 1])), sbh[32 + pc]);
                        o[e] = hmul2(htanh2(fh), hfma2(htanh2(hmul2(gh, h05)), h05, h05));
                    } else {
                        const int c = 32 * half + i;
                        const float f0 = __uint_as_float(f[i]) + sbz[c], f1 = __uint_as_float(f[i + 1]) + sbz[c + 1];
                        const float g0 = __uint_as_float(g[i]) + sbz[64 + c], g1 = __uint_as_float(g[i + 1]) + sbz[64 + c + 1];
                        o[e] = pack_bf16(tanh_fast(f0) * fmaf(0.5f, tanh_fast(0.5f * g0), 0.5f),
                                         tanh_fast(f1) * fmaf(0.5f, tanh_fast(0.5f * g1), 0.5f));
                    }
                }
                *(uint4*)(sA0 + r * 128 + (((4 * half + q) ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (warp_u == 0) {
            tc_fence_after();
            const uint64_t kA0 = umma_desc(smem_u32(sA0)), kBrs = umma_desc(smem_u32(sBrs));
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma(tmem_u + D2_COL, desc_adv(kA0, k * 32), desc_adv(kBrs, k * 32), idesc2, k != 0);
                umma_commit(mma_bar);
            }
            __syncwarp();
        }
        mbar_wait(mma_bar, 1);
        tc_fence_after();
        // the out GEMM was the last reader of the gated tile (A0), the gate GEMM of the ctx tile (A2): the next tile's copies are
        // requested now, an epilogue and a store drain before the loop top could do it (A1 follows there, once x' has left it)
        if (issuer) {
            const int nt = tile + gridDim.x;
            if (nt < a.n_tiles) {
                const int nb = nt / a.tiles_per_clip, n0 = (nt - nb * a.tiles_per_clip) * TILE_T;
                mbar_expect_tx(full_bar, load_bytes);
                tma_load_3d(sA0, &map_x, full_bar, 0, n0 - a.dil, nb);
                if (a.nchunks == 3) tma_load_3d(sA2, &map_ctx, full_bar, 0, n0, nb);
            }
        }
        // ---- epilogue 2: residual in place in the tap-1 tile; skip accumulation ------------------
        if (a.has_out) {
            uint32_t rr[32];
            tmem_ld32(tmem + lane_base + D2_COL + 32 * half, rr);
            uint4 xin[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) xin[q] = *(const uint4*)(sA1 + r * 128 + (((4 * half + q) ^ sw) << 4));
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t xi[4] = {xin[q].x, xin[q].y, xin[q].z, xin[q].w};
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int i = 8 * q + 2 * e, c = 32 * half + i;
                    const float2 xv = unpack_bf16(xi[e]);
                    o[e] = pack_bf16(__uint_as_float(rr[i]) + sbrs[c] + xv.x, __uint_as_float(rr[i + 1]) + sbrs[c + 1] + xv.y);
                }
                *(uint4*)(sA1 + r * 128 + (((4 * half + q) ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        }
        for (int s0 = 8 * half; s0 < a.S; s0 += 16) {
            uint32_t sv[8];
            tmem_ld8(tmem + lane_base + D2_COL + CC + s0, sv);
            tmem_ld_wait();
            if (live) {
                float4 v0 = make_float4(__uint_as_float(sv[0]) + sbrs[CC + s0], __uint_as_float(sv[1]) + sbrs[CC + s0 + 1],
                                        __uint_as_float(sv[2]) + sbrs[CC + s0 + 2], __uint_as_float(sv[3]) + sbrs[CC + s0 + 3]);
                float4 v1 = make_float4(__uint_as_float(sv[4]) + sbrs[CC + s0 + 4], __uint_as_float(sv[5]) + sbrs[CC + s0 + 5],
                                        __uint_as_float(sv[6]) + sbrs[CC + s0 + 6], __uint_as_float(sv[7]) + sbrs[CC + s0 + 7]);
                float4* d4 = (float4*)(skip_dst + s0);
                if (!a.skip_init && s0 != 0) {
                    // channels past the prefetched eight (skip_channels > 8): a vector reduction executed by L2 instead of a
                    // load-add-store of the same line by one thread, which serialises on the store's round trip (measured on
                    // the wide path: 25x; here 233 -> see profiles/ us per layer at S = 64).  One add per element and launch,
                    // launches in layer order: deterministic.
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d4), "f"(v0.x), "f"(v0.y), "f"(v0.z), "f"(v0.w) : "memory");
                    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(d4 + 1), "f"(v1.x), "f"(v1.y), "f"(v1.z), "f"(v1.w) : "memory");
                    continue;
                }
                if (!a.skip_init) {
                    v0.x += old0.x; v0.y += old0.y; v0.z += old0.z; v0.w += old0.w;
                    v1.x += old1.x; v1.y += old1.y; v1.z += old1.z; v1.w += old1.w;
                }
                d4[0] = v0; d4[1] = v1;
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (issuer && a.has_out) {
            tma_store_3d(&map_out, sA1, 0, t0, b);
            tma_commit();
        }
    }
    if (issuer) tma_wait_all0();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
    }
}

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

}  // namespace

// A tensor map is a pure function of (pointer, B, T): the caller-owned buffers of one geometry are reused step after step,
// so the ~100 descriptors a training step needs are encoded once and then served from this table (mutex-guarded, the only
// mutable global state of the library besides the per-thread error string and the launch counter).
namespace {
struct MapSlot { const void* ptr; int B, T, rows; bool used; CUtensorMap map; };
constexpr int MAP_SLOTS = 512;
MapSlot g_maps[MAP_SLOTS];
std::mutex g_maps_mu;
int encode_act_map(CUtensorMap* map, const void* ptr, int B, int T, int rows);
}

int tc::make_act_map(CUtensorMap* map, const void* ptr, int B, int T) { return tc::make_act_map_rows(map, ptr, B, T, TILE_T); }

// box = {64 channels, `rows` time steps, 1 clip}
int tc::make_act_map_rows(CUtensorMap* map, const void* ptr, int B, int T, int rows) {
    const size_t h = ((size_t)(uintptr_t)ptr >> 8) * 0x9E3779B97F4A7C15ull + (size_t)B * 1315423911u + (size_t)T + (size_t)rows * 7919u;
    const int i0 = (int)((h >> 20) % MAP_SLOTS);
    {
        std::lock_guard<std::mutex> lk(g_maps_mu);
        for (int probe = 0; probe < 4; ++probe) {
            const MapSlot& s = g_maps[(i0 + probe) % MAP_SLOTS];
            if (s.used && s.ptr == ptr && s.B == B && s.T == T && s.rows == rows) { *map = s.map; return 0; }
        }
    }
    const int rc = encode_act_map(map, ptr, B, T, rows);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_maps_mu);
    int victim = i0;
    for (int probe = 0; probe < 4; ++probe) if (!g_maps[(i0 + probe) % MAP_SLOTS].used) { victim = (i0 + probe) % MAP_SLOTS; break; }
    g_maps[victim].ptr = ptr; g_maps[victim].B = B; g_maps[victim].T = T; g_maps[victim].rows = rows; g_maps[victim].map = *map; g_maps[victim].used = true;
    return 0;
}

namespace {
int encode_act_map(CUtensorMap* map, const void* ptr, int B, int T, int rows) {
    EncodeTiledFn fn = encode_fn();
    MVN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[3] = {(cuuint64_t)CC, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)CC * 2, (cuuint64_t)T * CC * 2};
    cuuint32_t box[3] = {(cuuint32_t)CC, (cuuint32_t)rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MVN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}
}  // namespace

int mvn_tc_layer_supported(int C, int S, int video) {
    return C == CC && S >= 8 && S % 8 == 0 && S <= (video ? 32 : 64);   // the backward kernel's shared-memory budget
}

int mvn_tc_pack(const float* const* param_ptrs_dev, float* packed, const PackedLayout& P, const Geo& g, cudaStream_t st) {
    const int nchunks = g.video ? 3 : 2, N2 = ((g.C + g.S + 15) / 16) * 16;
    MVN_REQUIRE(smem_a_off(nchunks, N2) <= MVN_TC_IMG_BYTES, "tensor-core weight image does not fit its slot");
    dim3 grid(16, g.N);
    MVN_CUDA(mvn_launch_pdl(tc_pack_kernel, dim3(grid), dim3(256), (size_t)(0), st, param_ptrs_dev, packed, P, g.S, nchunks, N2, g.video, g.Cl));
    return mvn_check_launch("tc_pack");
}

int mvn_tc_layer_fwd(const void* x_in, const void* ctx, void* x_out, float* skip_sum, const float* lw,
                     const PackedLayout& P, const Geo& g, int layer, cudaStream_t st) {
    MVN_REQUIRE(mvn_tc_layer_supported(g.C, g.S, g.video), "tensor-core layer kernel: unsupported channel counts");
    MVN_REQUIRE((((uintptr_t)x_in) & 15) == 0 && (!x_out || (((uintptr_t)x_out) & 15) == 0), "activations must be 16-byte aligned");
    CUtensorMap map_x, map_ctx, map_out;
    int rc;
    if ((rc = make_act_map(&map_x, x_in, g.B, g.T))) return rc;
    if ((rc = make_act_map(&map_ctx, g.video ? ctx : x_in, g.B, g.T))) return rc;
    if ((rc = make_act_map(&map_out, x_out ? x_out : x_in, g.B, g.T))) return rc;
    TcArgs a;
    a.img = lw + P.oTc;
    a.skip = skip_sum;
    a.B = g.B; a.T = g.T; a.Tout = g.Tout; a.RF = g.RF; a.S = g.S; a.N2 = ((g.C + g.S + 15) / 16) * 16;
    a.dil = g.dil[layer]; a.nchunks = g.video ? 3 : 2; a.has_out = x_out != nullptr; a.skip_init = layer == 0;
    a.tiles_per_clip = (g.T + TILE_T - 1) / TILE_T; a.n_tiles = a.tiles_per_clip * g.B;
    const int smem = smem_total(a.nchunks, a.N2) + 1024;
    static const bool gate_f32 = getenv("MOVENET_B200_GATE_F32") != nullptr;
    static MvnSmemAttr attr_a, attr_b;
    MVN_CUDA(mvn_ensure_smem(layer_fwd_tc_kernel<true>, smem, attr_a));
    MVN_CUDA(mvn_ensure_smem(layer_fwd_tc_kernel<false>, smem, attr_b));
    int grid = 2 * mvn_sm_count();
    if (grid > a.n_tiles) grid = a.n_tiles;
    if (gate_f32) MVN_CUDA(mvn_launch_pdl(layer_fwd_tc_kernel<false>, dim3(grid), dim3(256), (size_t)smem, st, map_x, map_ctx, map_out, a));
    else MVN_CUDA(mvn_launch_pdl(layer_fwd_tc_kernel<true>, dim3(grid), dim3(256), (size_t)smem, st, map_x, map_ctx, map_out, a));
    return mvn_check_launch("layer_fwd_tc");
}
