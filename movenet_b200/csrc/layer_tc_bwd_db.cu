// Backward of GatedResidualConv1d (autograd of movenet/modules.py:67-93) for layers with dilation <= 8: the fused kernel of
// layer_tc_bwd.cu with a SECOND x / ctx buffer, so that the next tile's recompute GEMM no longer waits for this tile's weight
// gradient MMAs plus a TMA round trip (profiles/r01_ablation.md, "Tensor-pipe timeline": 2.0 k of the 10.6 k cycles of a tile).
// Same mathematics, same (P, U) gradient pair, same partial-gradient layout and reduce kernel.  What pays for the buffer:
//
//  * both dilation taps come from ONE 136-row TMA box of x (rows t0-8 .. t0+127): a UMMA descriptor may start at any 128-byte row
//    of a 128B-swizzle atom, so x(t) is the box from row 8 and x(t-d) the box from row 8-d -- as the K-major A operand of the
//    recompute GEMM and as the MN-major B operand of the weight gradient, where the two taps are the two 64-wide blocks of ONE
//    N = 128 operand whose block stride (LBO) is d rows = 128 d bytes;
//  * the running sum Q of the context gradient is no longer loaded: the layer's contribution is staged (in the ctx tile the
//    weight gradient has just released) and ADDED in place by a bf16 TMA reduction (cp.reduce.async.bulk.tensor .add); the top
//    layer, whose incoming sum is zero, stores instead.  One tile of shared memory and one load per tile less;
//  * the gate-bias sums ride in the weight gradient: the ctx block of its B operand is followed (block stride = the distance to
//    the d(skip) tile) by 16 columns of ones that live in the d(skip) tile's unused channels, so W1 = [dz^T x(t-d) | dz^T x(t)]
//    (N = 128) + [dz^T ctx | dz^T 1] (N = 80) and the eight N = 16 MMAs that re-read the whole dz operand are gone;
//  * the residual / skip bias sums ride in W2 the same way ([gated | ones], N = 80): no N = 16 MMA is left in the kernel.
//
// The buffer alone bought nothing; the order in which the one service warp issues its work did (profiles/r02_bwd_db.md, measured
// with the phase clocks below): the add-reduction goes LAST (it occupies the TMA unit ~10x longer than a store and must not sit in
// front of the next tile's P / U loads), the next tile's G1 is issued the moment W2 has left the in-order tensor pipe but AFTER the
// P' / U' stores, G2 is committed after wait_group.read 1 (the stores, not the reduction).  130.2 -> 114.7 us per launch at cfg01.
//
// Shared memory: weight image | set 0 | set 1 | G | U | P | DSK | DZ0 | DZ1 | barriers, set = x box (17 KB) + ctx tile.
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
#include "layer_tc.h"

using namespace tc;

namespace {

// Instrumented build (MOVENET_B200_NVCC_EXTRA=-DMVN_PHASE_CLOCKS=1): clock64() stamps of tile iterations 5..7 of CTA 0 for the
// control lane (CLKC) and two worker threads (CLKW, CLKM); printed by launch 20 when MVN_PROF is set.
#ifndef MVN_PHASE_CLOCKS
#define MVN_PHASE_CLOCKS 0
#endif
#if MVN_PHASE_CLOCKS
__device__ unsigned long long g_clk_db[3][3][20];
#define CLK_(role, i, cond) do { if (blockIdx.x == 0 && it >= 5 && it < 8 && (cond)) g_clk_db[role][it - 5][i] = clock64(); } while (0)
#define CLKW(i) CLK_(0, i, tid == 256)
#define CLKC(i) CLK_(1, i, leader)
#define CLKM(i) CLK_(2, i, tid == 511)          // (one more worker: another lane quarter and channel range)
#else
#define CLKW(i) do {} while (0)
#define CLKC(i) do {} while (0)
#define CLKM(i) do {} while (0)
#endif

constexpr int PART_LD = 256;                       // partial row: 192 (dWz^T) + 64 (dWrs^T)          (== layer_tc_bwd.cu)
constexpr int PART_FLOATS = 128 * PART_LD + 256;   // + bias sums
constexpr int XBOX_ROWS = TILE_T + 8, XBOX_BYTES = XBOX_ROWS * 128;       // 17408 = 17 x 1024
constexpr int SET_BYTES = XBOX_BYTES + TILE_BYTES;
constexpr int ONES_COLS = 16;                      // the ones block: logical channels [0, 16) of the DSK tile, d(skip) follows
// TMEM: the tile's columns [0, 192) as in layer_tc_bwd.cu; accumulators that live across the CTA's tiles:
constexpr int W1A_COL = 192, W1B_COL = 320, W2_COL = 400;   // 128 | 64 + 16 | 64 + 16 (the 16: products with the ones block = bias sums)

struct DbArgs {
    const void* img;
    const float* dskip;   // (B, Tout, S) fp32
    float* partial;       // [grid][PART_FLOATS]
    int B, T, Tout, RF, S, N2, dil, dil_up, nchunks, tiles_per_clip, n_tiles;
    int zero_in;          // the incoming stream gradient (P, U) and context-gradient sum are zero (top layer)
};

__host__ __device__ inline int db_sets_off(int nc, int N2) { return smem_a_off(nc, N2); }
__host__ __device__ inline int db_smem_total(int nc, int N2) { return db_sets_off(nc, N2) + 2 * SET_BYTES + 6 * TILE_BYTES + 128; }

__device__ __forceinline__ void warp_arrive(uint64_t* bar) {     // every lane has fenced its own writes; one lane signals
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ void tma_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

constexpr int N_WORKERS = 512, N_THREADS = N_WORKERS + 32;   // 16 worker warps + the control warp

template <bool PAIR_IN>
__global__ void __launch_bounds__(N_THREADS, 1)
layer_bwd_db_kernel(const __grid_constant__ CUtensorMap map_xbox, const __grid_constant__ CUtensorMap map_ctx,
                    const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_u,
                    const __grid_constant__ CUtensorMap map_pout, const __grid_constant__ CUtensorMap map_uout,
                    const __grid_constant__ CUtensorMap map_q, const DbArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int nc = a.nchunks;
    uint8_t* sBz = smem;
    uint8_t* sBrs = smem + smem_brs_off(nc);
    float* sbz = (float*)(smem + smem_bias_off(nc, a.N2));
    uint8_t* sSet = smem + db_sets_off(nc, a.N2);      // set s: x box at sSet + s SET_BYTES, ctx tile XBOX_BYTES further
    uint8_t* sG = sSet + 2 * SET_BYTES;                // gated tile; the ones block of DSK is its second B block (3 tiles further)
    uint8_t* sU = sG + TILE_BYTES;                     // U | P | DSK in this order: [P|DSK] and [U|DSK] are both M = 128 block pairs
    uint8_t* sDXS = sU + TILE_BYTES;                   // the P tile (the stream gradient is P + U, never summed in memory)
    uint8_t* sDSK = sDXS + TILE_BYTES;
    uint8_t* sDZ = sDSK + TILE_BYTES;                  // DZ0 (filter half) | DZ1 (gate half)
    // barriers, one completion per tile each (parity = tile iteration & 1), except IMG (once) and A_IN0/1 (every other tile).
    // E_* are the worker -> control-warp signals (one arrival per worker warp), the rest are TMA / tcgen05.commit completions.
    enum { IMG = 0, A_IN0, A_IN1, P_IN, U_IN, G1, G2, G3, W1, WALL, E_DSK, E_DZ, E_OUT, N_BARS };
    uint64_t* bar = (uint64_t*)(sDZ + 2 * TILE_BYTES);
    uint32_t* tmem_slot = (uint32_t*)(bar + N_BARS);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & 127, sw = r & 7;          // row of the tile == TMEM lane; warps w and w+4 share a lane quarter
    const int half = (tid >> 7) & 3;              // ... and split the channel range between them (4 quarters of 16)
    const int NZ = nc * CC;                        // columns of D4

    if (tid == 0) {
        for (int i = 0; i < N_BARS; ++i) mbar_init(bar + i, i == E_DSK ? 4 : i > E_DSK ? N_WORKERS / 32 : 1);   // E_DSK: the four warps that write the tile
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t wbytes = (uint32_t)smem_a_off(nc, a.N2);
        mbar_expect_tx(bar + IMG, wbytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(smem)), "l"(a.img), "r"(wbytes), "r"(smem_u32(bar + IMG)) : "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // constant tile: DSK = ones in logical channels [0, 16), zero elsewhere (the S live channels are rewritten every tile)
    for (int i = tid; i < TILE_BYTES / 16; i += N_THREADS) {
        const int row = i >> 3, q = (i & 7) ^ (row & 7);        // 16-byte chunk i holds logical channels [8 q, 8 q + 8)
        const uint32_t v = q < ONES_COLS / 8 ? 0x3F803F80u : 0u;
        ((uint4*)sDSK)[i] = make_uint4(v, v, v, v);
    }
    if (a.zero_in)        // U | P are adjacent and stay zero for the whole kernel
        for (int i = tid; i < 2 * TILE_BYTES / 16; i += N_THREADS) ((uint4*)sU)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    mvn_griddep_launch();
    mvn_griddep_wait();              // everything above overlapped the previous kernel's tail; its output is read from here on
    mbar_wait(bar + IMG, 0);
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int top = blockIdx.x, step = gridDim.x;
    const int count = (a.n_tiles - top + step - 1) / step;

    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);   // warp-uniform copy: keeps the role branch convergent
    if (warp_u == N_WORKERS / 32) {
        // ================================ control warp ============================================
        // The whole warp runs the loop (so the code stays on the uniform datapath); one elected lane issues every TMA and MMA.
        const bool leader = elect_one();
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
        const int tap0 = (8 - a.dil) * 128;           // x(t - d) starts this many bytes into the box, x(t) at 1024
        const uint64_t kX = umma_desc(smem_u32(sSet)), kBz = umma_desc(smem_u32(sBz)), kDXS = umma_desc(smem_u32(sDXS)),
                       kU = umma_desc(smem_u32(sU)), kDSK = umma_desc(smem_u32(sDSK)), kDZ = umma_desc(smem_u32(sDZ));
        const uint64_t mBrs = umma_desc_mn(smem_u32(sBrs), TILE_BYTES), mBz = umma_desc_mn(smem_u32(sBz), TILE_BYTES),
                       mDZ = umma_desc_mn(smem_u32(sDZ), TILE_BYTES),
                       mDXS = umma_desc_mn(smem_u32(sDXS), TILE_BYTES), mU = umma_desc_mn(smem_u32(sU), 2 * TILE_BYTES),
                       mG = umma_desc_mn(smem_u32(sG), (uint32_t)(sDSK - sG));      // [gated | ones]
        // weight-gradient B operands per set: [x(t-d) | x(t)] = two 64-wide blocks d rows apart; [ctx | ones] = the ctx tile and,
        // one block stride further, the first 16 channels of the DSK tile
        const uint64_t mX0 = umma_desc_mn(smem_u32(sSet) + tap0, (uint32_t)a.dil * 128),
                       mX1 = umma_desc_mn(smem_u32(sSet + SET_BYTES) + tap0, (uint32_t)a.dil * 128);
        const uint64_t mC0 = umma_desc_mn(smem_u32(sSet + XBOX_BYTES), (uint32_t)(sDSK - (sSet + XBOX_BYTES))),
                       mC1 = umma_desc_mn(smem_u32(sSet + SET_BYTES + XBOX_BYTES), (uint32_t)(sDSK - (sSet + SET_BYTES + XBOX_BYTES)));
        const uint32_t iG1 = umma_idesc_major(TILE_T, 128, 0, 0);
        const uint32_t iG2 = umma_idesc_major(TILE_T, 64, 0, 1);
        const uint32_t iG3 = umma_idesc_major(TILE_T, NZ, 0, 1);
        const uint32_t iW1a = umma_idesc_major(TILE_T, 128, 1, 1);
        const uint32_t iW1b = umma_idesc_major(TILE_T, CC + ONES_COLS, 1, 1);
        const uint32_t iW2 = umma_idesc_major(TILE_T, CC + ONES_COLS, 1, 1);
        auto load_set = [&](int s, int lb, int l0) {       // x box / ctx tile of one time tile -> A_IN[s]
            uint8_t* dst = sSet + s * SET_BYTES;
            mbar_expect_tx(bar + A_IN0 + s, (uint32_t)(XBOX_BYTES + (nc == 3 ? TILE_BYTES : 0)));
            tma_load_3d(dst, &map_xbox, bar + A_IN0 + s, 0, l0 - 8, lb);
            if (nc == 3) tma_load_3d(dst + XBOX_BYTES, &map_ctx, bar + A_IN0 + s, 0, l0, lb);
        };
        auto load_tile = [&](uint8_t* dst, const CUtensorMap* map, int which, int lb, int l0) {
            mbar_expect_tx(bar + which, (uint32_t)TILE_BYTES);
            tma_load_3d(dst, map, bar + which, 0, l0, lb);
        };
        // G1: recompute the gate pre-activations of the tile in set s (its x / ctx have long arrived, except for the first tile)
        auto issue_g1 = [&](uint32_t s, uint32_t use) {
            mbar_wait(bar + A_IN0 + s, use & 1);
            tc_fence_after();
            if (leader) {
                const int set_off = (int)s * SET_BYTES;
#pragma unroll
                for (int k = 0; k < 4; ++k)            // tap t - d
                    umma(tmem_u, desc_adv(kX, set_off + tap0 + k * 32), desc_adv(kBz, k * 32), iG1, k != 0);
#pragma unroll
                for (int k = 0; k < 4; ++k)            // tap t
                    umma(tmem_u, desc_adv(kX, set_off + 1024 + k * 32), desc_adv(kBz, TILE_BYTES + k * 32), iG1, 1);
                if (nc == 3) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma(tmem_u, desc_adv(kX, set_off + XBOX_BYTES + k * 32), desc_adv(kBz, 2 * TILE_BYTES + k * 32), iG1, 1);
                }
                umma_commit(bar + G1);
            }
        };
        if (leader) {
            const int lb = top / a.tiles_per_clip, l0 = (top - lb * a.tiles_per_clip) * TILE_T;
            load_set(0, lb, l0);
            if (!a.zero_in) load_tile(sDXS, &map_p, P_IN, lb, l0);
            if (PAIR_IN) load_tile(sU, &map_u, U_IN, lb, l0 + a.dil_up);
        }
        issue_g1(0, 0);
        for (uint32_t it = 0; it < (uint32_t)count; ++it) {
            const uint32_t ph = it & 1, s = it & 1;
            const int tile = top + (int)it * step;
            const int b = tile / a.tiles_per_clip, t0 = (tile - b * a.tiles_per_clip) * TILE_T;
            const int nt = tile + step;                // this CTA's next tile
            const bool has_next = it + 1 < (uint32_t)count;
            const int nb = nt / a.tiles_per_clip, n0 = (nt - nb * a.tiles_per_clip) * TILE_T;
            const int set_off = (int)s * SET_BYTES;
            // (G1 of this tile was issued at the end of the previous iteration, or before the loop)
            // the previous tile's P' / U' stores have left DZ0 / DZ1 (ordered before G2's commit: the workers write DZ again only after
            // they have seen G2); its context-gradient reduction -- the younger bulk group, and a slow one -- may still be reading
            if (leader) { if (nc == 3) tma_wait_read1(); else tma_wait_read0(); }
            // G2: d(gated) = (P + U) . Wr + dskip . Ws as three accumulating products (no pre-sum pass): contraction over
            // the image's ROWS (c_out | s) -> B is MN-major.  Needs only the loads and the DSK tile, so it runs next to G1.
            CLKC(3);
            if (!a.zero_in) mbar_wait(bar + P_IN, ph);
            if (PAIR_IN) mbar_wait(bar + U_IN, ph);
            CLKC(4);
            mbar_wait(bar + E_DSK, ph);
            CLKC(5);
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma(tmem_u + 128, desc_adv(kDXS, k * 32), desc_adv(mBrs, k * 2048), iG2, k != 0);
                if (PAIR_IN) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma(tmem_u + 128, desc_adv(kU, k * 32), desc_adv(mBrs, k * 2048), iG2, 1);
                }
#pragma unroll
                for (int k = 0; k < 2; ++k)                // the skip channels: rows 64.. of the image, 16 per step; they sit
                    if (k < (a.S + 15) / 16)               // behind the ones block in the DSK tile (one K step further)
                        umma(tmem_u + 128, desc_adv(kDSK, (k + 1) * 32), desc_adv(mBrs, (4 + k) * 2048), iG2, 1);
                umma_commit(bar + G2);
                CLKC(6);
                // every store of the previous tile has left shared memory -> the other set (its ctx tile was the staging tile of the
                // reduction) takes the NEXT tile's x / ctx now, most of a tile before its recompute GEMM needs them
                tma_wait_read0();
                if (has_next) {
                    load_set((int)(s ^ 1), nb, n0);
                    if (!a.zero_in) {                  // ... and the next tile's gradient tiles start towards L2
                        tma_prefetch_3d(&map_p, 0, n0, nb);
                        if (PAIR_IN) tma_prefetch_3d(&map_u, 0, n0 + a.dil_up, nb);
                    }
                }
            }
            // W2 (K = time): d[Wr | Ws]^T += [P|DSK]^T . [gated | 1] + [U|DSK]^T . [gated | 1]: column 64 holds the bias sums.
            // The DSK rows (64..) are accumulated twice and halved at the flush (exact).  It follows W1 (issuing it between G2 and G3, under
            // epilogue 1b, was measured slower: it delays G3 and competes with the epilogue for shared-memory bandwidth).
            const uint32_t acc0 = it != 0;
            auto issue_w2 = [&]() {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma(tmem_u + W2_COL, desc_adv(mDXS, k * 2048), desc_adv(mG, k * 2048), iW2, acc0 | (k != 0));
                if (PAIR_IN) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        umma(tmem_u + W2_COL, desc_adv(mU, k * 2048), desc_adv(mG, k * 2048), iW2, 1);
                }
                umma_commit(bar + WALL);
                CLKC(10);
            };
            // G3: D4[t][kin] = sum_m dz[t][m] Wz[m][kin]  (A = dz tiles K-major, B = the image read MN-major)
            mbar_wait(bar + E_DZ, ph);
            CLKC(7);
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma(tmem_u, desc_adv(kDZ, c * TILE_BYTES + k * 32), desc_adv(mBz, (c * 64 + k * 16) * 128), iG3, (c | k) != 0);
                umma_commit(bar + G3);
                CLKC(8);
                // W1 (K = time): every tile is [time x 64 ch], i.e. an MN-major operand
                const uint64_t mX = s ? mX1 : mX0, mC = s ? mC1 : mC0;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma(tmem_u + W1A_COL, desc_adv(mDZ, k * 2048), desc_adv(mX, k * 2048), iW1a, acc0 | (k != 0));
                if (nc == 3) {       // the context convs and their biases exist only with video
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        umma(tmem_u + W1B_COL, desc_adv(mDZ, k * 2048), desc_adv(mC, k * 2048), iW1b, acc0 | (k != 0));
                }
                umma_commit(bar + W1);
                CLKC(9);
                issue_w2();
            }
            mbar_wait(bar + E_OUT, ph);            // P', U', Q' are staged; nobody reads the P / U tiles or the tile's TMEM columns any more
            // Order of the tail: the P' / U' stores (the DZ tiles must drain before the next epilogue 1b), then -- as soon as W2 has
            // left the tensor pipe -- the next tile's G1, because the workers are idle until it completes; the P / U reloads (needed by
            // G2, half a tile away) and the context-gradient add-reduction (it occupies the TMA unit far longer than a store) follow.
            CLKC(12);
            if (leader) {
                tma_store_3d(&map_pout, sDZ, 0, t0, b);
                tma_store_3d(&map_uout, sDZ + TILE_BYTES, 0, t0, b);
                tma_commit();
            }
            if (has_next) {
                mbar_wait(bar + WALL, ph);         // W2 no longer reads the P and U tiles
                CLKC(15);
                issue_g1(s ^ 1, (it + 1) >> 1);
                if (leader && !a.zero_in) load_tile(sDXS, &map_p, P_IN, nb, n0);
                if (leader && PAIR_IN) load_tile(sU, &map_u, U_IN, nb, n0 + a.dil_up);
            }
            if (leader && nc == 3) {
                if (a.zero_in) tma_store_3d(&map_q, sSet + set_off + XBOX_BYTES, 0, t0, b);
                else tma_reduce_add_3d(&map_q, sSet + set_off + XBOX_BYTES, 0, t0, b);
                tma_commit();
            }
            CLKC(13);
            __syncwarp();
        }
        if (leader) tma_wait_all0();
    } else if (tid < N_WORKERS) {
        // ================================ worker warps ============================================
        // this thread's d(skip) row of tile `tl` (first 8 channels), fetched one tile ahead
        auto load_dskip = [&](int tl, float4& v0, float4& v1) {
            const int lb = tl / a.tiles_per_clip, lt = (tl - lb * a.tiles_per_clip) * TILE_T + r, js = lt - (a.RF - 1);
            v0 = make_float4(0.f, 0.f, 0.f, 0.f); v1 = v0;
            if (tl < a.n_tiles && lt < a.T && js >= 0 && js < a.Tout) {
                const float4* src = (const float4*)(a.dskip + ((size_t)lb * a.Tout + js) * a.S);
                v0 = src[0]; v1 = src[1];
            }
        };
        float4 ds0, ds1;
        load_dskip(top, ds0, ds1);
        const int o0 = r * 128 + (((2 * half) ^ sw) << 4), o1 = r * 128 + (((2 * half + 1) ^ sw) << 4);   // this thread's 16 channels
        for (uint32_t it = 0; it < (uint32_t)count; ++it) {
            const uint32_t ph = it & 1;
            const int tile = top + (int)it * step;
            const bool has_next = it + 1 < (uint32_t)count;
            const int b = tile / a.tiles_per_clip, t0 = (tile - b * a.tiles_per_clip) * TILE_T;
            const int t = t0 + r;
            uint8_t* sQS = sSet + (it & 1) * SET_BYTES + XBOX_BYTES;     // this tile's ctx tile: the staging tile of its Q contribution
            CLKW(0);
            if (it) mbar_wait(bar + WALL, ph ^ 1);      // the previous tile's weight-gradient MMAs are done with DSK and G
            CLKW(1);
            if (half == 0) {   // d(skip) row of this thread -> bf16, logical channels [16, 16 + S) of the DSK tile
                constexpr int Q0 = ONES_COLS / 8;
                *(uint4*)(sDSK + r * 128 + ((Q0 ^ sw) << 4)) =
                    make_uint4(pack_bf16(ds0.x, ds0.y), pack_bf16(ds0.z, ds0.w), pack_bf16(ds1.x, ds1.y), pack_bf16(ds1.z, ds1.w));
                const int js = t - (a.RF - 1);
                const bool live = t < a.T && js >= 0 && js < a.Tout;
                const float* src = a.dskip + ((size_t)b * a.Tout + (live ? js : 0)) * a.S;
                for (int s0 = 8; s0 < a.S; s0 += 8) {
                    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                    if (live) { v0 = ((const float4*)(src + s0))[0]; v1 = ((const float4*)(src + s0))[1]; }
                    *(uint4*)(sDSK + r * 128 + (((Q0 + (s0 >> 3)) ^ sw) << 4)) =
                        make_uint4(pack_bf16(v0.x, v0.y), pack_bf16(v0.z, v0.w), pack_bf16(v1.x, v1.y), pack_bf16(v1.z, v1.w));
                }
                fence_proxy_async();       // (warp-uniform: only the four writing warps fence and signal)
                warp_arrive(bar + E_DSK);
            }
            // ---- epilogue 1a (needs G1 only): th, sg, gated -> G tile ---------------------------------
            float th[16], sg[16];
            CLKW(2);
            mbar_wait(bar + G1, ph);
            CLKW(3);
            tc_fence_after();
            {
                uint32_t f[16], g[16];
                tmem_ld16(tmem + lane_base + 16 * half, f);
                tmem_ld16(tmem + lane_base + 64 + 16 * half, g);
                tmem_ld_wait();
                uint32_t oy[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c = 16 * half + i + e;
                        th[i + e] = tanh_fast(__uint_as_float(f[i + e]) + sbz[c]);
                        sg[i + e] = fmaf(0.5f, tanh_fast(0.5f * (__uint_as_float(g[i + e]) + sbz[64 + c])), 0.5f);
                    }
                    oy[i >> 1] = pack_bf16(th[i] * sg[i], th[i + 1] * sg[i + 1]);
                }
                *(uint4*)(sG + o0) = make_uint4(oy[0], oy[1], oy[2], oy[3]);
                *(uint4*)(sG + o1) = make_uint4(oy[4], oy[5], oy[6], oy[7]);
            }
            // ---- epilogue 1b: gate derivative -> DZ0 | DZ1 ------------------------------------------
            CLKW(4); CLKM(4);
            mbar_wait(bar + G2, ph);
            CLKW(5);
            tc_fence_after();
            {
                uint32_t dg[16];
                tmem_ld16(tmem + lane_base + 128 + 16 * half, dg);
                tmem_ld_wait();
                uint32_t of[8], og[8];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    float zf[2], zg[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float d = __uint_as_float(dg[i + e]), h = th[i + e], s = sg[i + e];
                        zf[e] = d * s * (1.f - h * h);
                        zg[e] = d * (h * s) * (1.f - s);
                    }
                    of[i >> 1] = pack_bf16(zf[0], zf[1]); og[i >> 1] = pack_bf16(zg[0], zg[1]);
                }
                *(uint4*)(sDZ + o0) = make_uint4(of[0], of[1], of[2], of[3]);
                *(uint4*)(sDZ + o1) = make_uint4(of[4], of[5], of[6], of[7]);
                *(uint4*)(sDZ + TILE_BYTES + o0) = make_uint4(og[0], og[1], og[2], og[3]);
                *(uint4*)(sDZ + TILE_BYTES + o1) = make_uint4(og[4], og[5], og[6], og[7]);
            }
            fence_proxy_async();
            tc_fence_before();
            warp_arrive(bar + E_DZ);
            // the next tile's d(skip) row towards L2 now: the load itself (end of the iteration) is the head of the next tile's chain.
            // (Loading into registers here instead was measured slower: 120.5 vs 116.4 us per launch.)
            if (half == 0 && has_next) {
                const int tl = tile + step, lb = tl / a.tiles_per_clip, lt = (tl - lb * a.tiles_per_clip) * TILE_T + r, js = lt - (a.RF - 1);
                if (lt < a.T && js >= 0 && js < a.Tout) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.dskip + ((size_t)lb * a.Tout + js) * a.S));
            }
            CLKW(6); CLKM(6);
            // ---- epilogue 2: U' = W0^T dz, P' = d(x') + W1^T dz, Q contribution = V^T dz: registers until W1 releases the tiles
            mbar_wait(bar + G3, ph);
            CLKW(7);
            tc_fence_after();
            uint32_t po[8], uo[8], qo[8];
            {
                uint32_t w[16], v[16];
                tmem_ld16(tmem + lane_base + 16 * half, w);
                tmem_ld16(tmem + lane_base + 64 + 16 * half, v);
                const uint4 x0 = *(const uint4*)(sDXS + o0), x1 = *(const uint4*)(sDXS + o1);
                const uint4 zz = make_uint4(0, 0, 0, 0);
                const uint4 y0 = PAIR_IN ? *(const uint4*)(sU + o0) : zz, y1 = PAIR_IN ? *(const uint4*)(sU + o1) : zz;
                const uint32_t xi[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                const uint32_t yi[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    uo[i] = pack_bf16(__uint_as_float(w[2 * i]), __uint_as_float(w[2 * i + 1]));
                    const float2 xp = unpack_bf16(xi[i]), xu = unpack_bf16(yi[i]);
                    po[i] = pack_bf16(__uint_as_float(v[2 * i]) + (xp.x + xu.x), __uint_as_float(v[2 * i + 1]) + (xp.y + xu.y));
                }
            }
            if (nc == 3) {
                uint32_t v[16];
                tmem_ld16(tmem + lane_base + 128 + 16 * half, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) qo[i] = pack_bf16(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
            }
            CLKW(8);
            mbar_wait(bar + W1, ph);            // W1 no longer reads the DZ tiles nor this tile's x / ctx set
            CLKW(9);
            *(uint4*)(sDZ + TILE_BYTES + o0) = make_uint4(uo[0], uo[1], uo[2], uo[3]);
            *(uint4*)(sDZ + TILE_BYTES + o1) = make_uint4(uo[4], uo[5], uo[6], uo[7]);
            *(uint4*)(sDZ + o0) = make_uint4(po[0], po[1], po[2], po[3]);
            *(uint4*)(sDZ + o1) = make_uint4(po[4], po[5], po[6], po[7]);
            if (nc == 3) {
                *(uint4*)(sQS + o0) = make_uint4(qo[0], qo[1], qo[2], qo[3]);
                *(uint4*)(sQS + o1) = make_uint4(qo[4], qo[5], qo[6], qo[7]);
            }
            fence_proxy_async();
            tc_fence_before();
            warp_arrive(bar + E_OUT);
            CLKW(10); CLKM(10);
            load_dskip(has_next ? tile + step : a.n_tiles, ds0, ds1);
        }
        mbar_wait(bar + WALL, (uint32_t)(count - 1) & 1);
    }
    // ---- flush this CTA's partial weight / bias gradients (the layout layer_tc_bwd.cu's reduce kernel reads) --------------
    if (tid < N_WORKERS) {
        tc_fence_after();
        float* part = a.partial + (size_t)blockIdx.x * PART_FLOATS;
        float* prow = part + (size_t)r * PART_LD;
#pragma unroll 1
        for (int j = half; j < 8; j += 4) {             // dWz^T[m = r][k]: the two taps
            uint32_t v[16];
            tmem_ld16(tmem + lane_base + W1A_COL + 16 * j, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q)
                ((float4*)(prow + 16 * j))[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                            __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        }
        if (nc == 3) {                                   // ... the context conv, and (one more column) the gate-bias sums
            uint32_t v[16];
            tmem_ld16(tmem + lane_base + W1B_COL + 16 * half, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q)
                ((float4*)(prow + 128 + 16 * half))[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                                     __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        }
        // rows of the [P|DSK]^T products: 0..63 residual channels, 64..79 the ones block (not a gradient), 80.. the skip channels
        const int dr = r < CC ? r : r - ONES_COLS;
        const bool keep = r < CC || r >= CC + ONES_COLS;
        const float sc = (r >= CC && PAIR_IN) ? 0.5f : 1.f;     // pair input: the DSK rows were accumulated once with P and once with U
        {
            uint32_t v[16];
            tmem_ld16(tmem + lane_base + W2_COL + 16 * half, v);
            tmem_ld_wait();
            if (keep) {
                float* drow = part + (size_t)dr * PART_LD + 192 + 16 * half;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    ((float4*)drow)[q] = make_float4(sc * __uint_as_float(v[4 * q]), sc * __uint_as_float(v[4 * q + 1]),
                                                     sc * __uint_as_float(v[4 * q + 2]), sc * __uint_as_float(v[4 * q + 3]));
            }
        }
        {
            uint32_t v1[8], v2[8];
            tmem_ld8(tmem + lane_base + W1B_COL + CC, v1);
            tmem_ld8(tmem + lane_base + W2_COL + CC, v2);
            tmem_ld_wait();
            if (half == 0) {
                if (nc == 3) part[128 * PART_LD + r] = __uint_as_float(v1[0]);
                if (keep) part[128 * PART_LD + 128 + dr] = sc * __uint_as_float(v2[0]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
    }
}

}  // namespace

// which layers run on this kernel: pair output, dilation <= 8 (both taps in one box), the shared-memory budget; MOVENET_B200_BWD_DB=0
// keeps the single-buffer kernel of layer_tc_bwd.cu (the two update the context-gradient sum differently: in place / ping-pong)
int mvn_tc_bwd_db_supported(const Geo& g, int layer) {
    const char* e = getenv("MOVENET_B200_BWD_DB");      // read per call: the tests switch it
    if (e && !atoi(e)) return 0;
    if (!mvn_tc_layer_supported(g.C, g.S, g.video) || g.S > 32 || mvn_tc_bwd_sum_out(g, layer)) return 0;
    const int nc = g.video ? 3 : 2, N2 = ((g.C + g.S + 15) / 16) * 16;
    return g.dil[layer] <= 8 && db_smem_total(nc, N2) + 1024 <= 227 * 1024;
}

int mvn_tc_layer_bwd_db(const void* x_in, const void* ctx, const void* p_in, const void* u_in, void* p_out, void* u_out,
                        const float* dskip, void* q_sum, const float* lw, float* partial, const PackedLayout& P, const Geo& g, int layer,
                        cudaStream_t st) {
    MVN_REQUIRE(mvn_tc_bwd_db_supported(g, layer), "double-buffered tensor-core backward kernel: unsupported layer");
    MVN_REQUIRE(mvn_tc_bwd_partial_bytes() == (size_t)148 * PART_FLOATS * 4, "partial-gradient layouts of the two backward kernels differ");
    MVN_REQUIRE(p_out && u_out && (p_in || !u_in) && (!g.video || q_sum), "double-buffered tensor-core backward kernel: bad buffers");
    CUtensorMap mx, mc, mp, mu, mpo, muo, mq;
    int rc;
    if ((rc = make_act_map_rows(&mx, x_in, g.B, g.T, XBOX_ROWS))) return rc;
    if ((rc = make_act_map(&mc, g.video ? ctx : x_in, g.B, g.T))) return rc;
    if ((rc = make_act_map(&mp, p_in ? p_in : x_in, g.B, g.T))) return rc;      // p_in == u_in == null: zero incoming gradient
    if ((rc = make_act_map(&mu, u_in ? u_in : x_in, g.B, g.T))) return rc;
    if ((rc = make_act_map(&mpo, p_out, g.B, g.T))) return rc;
    if ((rc = make_act_map(&muo, u_out, g.B, g.T))) return rc;
    if ((rc = make_act_map(&mq, g.video ? q_sum : p_out, g.B, g.T))) return rc;
    DbArgs a;
    a.img = lw + P.oTc; a.dskip = dskip; a.partial = partial;
    a.B = g.B; a.T = g.T; a.Tout = g.Tout; a.RF = g.RF; a.S = g.S; a.N2 = ((g.C + g.S + 15) / 16) * 16;
    a.dil = g.dil[layer]; a.dil_up = layer + 1 < g.N ? g.dil[layer + 1] : 0;
    a.zero_in = p_in == nullptr;
    a.nchunks = g.video ? 3 : 2;
    a.tiles_per_clip = (g.T + TILE_T - 1) / TILE_T; a.n_tiles = a.tiles_per_clip * g.B;
    const int smem = db_smem_total(a.nchunks, a.N2) + 1024;
    static MvnSmemAttr attr_a, attr_b;
    MVN_CUDA(mvn_ensure_smem(layer_bwd_db_kernel<false>, smem, attr_a));
    MVN_CUDA(mvn_ensure_smem(layer_bwd_db_kernel<true>, smem, attr_b));
    int grid = mvn_sm_count() < 148 ? mvn_sm_count() : 148;      // the per-CTA partial buffers are sized for 148 CTAs
    if (grid > a.n_tiles) grid = a.n_tiles;
    if (u_in) MVN_CUDA(mvn_launch_pdl(layer_bwd_db_kernel<true>, dim3(grid), dim3(N_THREADS), (size_t)smem, st, mx, mc, mp, mu, mpo, muo, mq, a));
    else MVN_CUDA(mvn_launch_pdl(layer_bwd_db_kernel<false>, dim3(grid), dim3(N_THREADS), (size_t)smem, st, mx, mc, mp, mu, mpo, muo, mq, a));
#if MVN_PHASE_CLOCKS
    if (getenv("MVN_PROF")) {
        static int launches = 0;
        if (++launches == 20) {
            cudaDeviceSynchronize();
            unsigned long long h[3][3][20];
            cudaMemcpyFromSymbol(h, g_clk_db, sizeof(h));
            const char* names[3] = {"worker", "control", "last-worker"};
            for (int w = 0; w < 3; ++w)
                for (int i = 0; i < 3; ++i) {
                    fprintf(stderr, "CLK %-11s it%d:", names[w], i + 5);
                    for (int j = 0; j < 17; ++j) fprintf(stderr, " %lld", h[w][i][j] ? (long long)(h[w][i][j] - h[1][0][0]) : -1LL);
                    fprintf(stderr, "\n");
                }
        }
    }
#endif
    return mvn_check_launch("layer_bwd_db");
}
