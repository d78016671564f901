// Backward of GatedResidualConv1d (autograd of movenet/modules.py:67-93) as ONE fused sm_100a kernel
// per layer: data gradients, weight gradients and bias gradients, all on tcgen05 tensor cores.
//
// The gradient of the residual stream is never materialised as one tensor.  Layer l+1 hands down
//   P[t] = d(x_{l+1})[t] without its tap-0 term        and   U[t] = W0^T dz_{l+1}[t]   (the tap-0 term,
//   which belongs to time t - d_{l+1}),  so  d(x_{l+1})[t] = P[t] + U[t + d_{l+1}]
// and the dilation shift becomes a TMA row coordinate instead of a scatter.  Per 128-row time tile:
//
//   TMA : x(t-d) x(t) ctx(t) P(t) U(t+d_up)          | threads: d(skip) rows -> bf16 tile
//   dxs = P + U is never formed in memory: every product with it is two accumulating MMAs
//   G1  : D1 = [x(t-d)|x(t)|ctx] . Wz^T               (recompute the gate pre-activations)
//   G2  : D3 = [P | dskip] . [Wr ; Ws] + U . Wr       (d gated)          B read MN-major from the SAME image
//   epilogue 1: th, sg, gated, dz_f, dz_g  -> bf16 tiles DZ0 DZ1 G  (128B-swizzled)
//   G3  : D4 = dz . Wz  -> [U' | W1^T dz | V^T dz]    (B = the forward weight image read MN-major)
//   W1  : dWz^T  += dz^T . [x(t-d)|x(t)|ctx]          (K = time: every tile is read MN-major)
//   W2  : dWrs^T += [P|dskip]^T . gated + [U|dskip]^T . gated (skip rows halved at the flush) ;
//         bias grads = the same A operands times an all-ones B
//   epilogue 2: P' = dxs + D4[tap1] ; U' = D4[tap0]  -> TMA stores ; d(ctx) += D4[ctx]
//
// Summed output (opt-in, MOVENET_B200_BWD_SUM=1; every dilation from the top layer down to this one <= 128, <= 32 with video): the
// producing layer adds the two terms itself and writes ONE stream D'[t] = P'[t] + U'[t + d]: row r of the sum needs U' row r + d,
// which is in the same tile (exchanged through the U tile, free because the input is one stream too) or in the first d rows of
// the next tile in time.  A CTA therefore walks a CONTIGUOUS run of tiles backwards in time and keeps those d rows ("carry") in
// shared memory; the run starts with one warm-up tile (the tile after the run, recomputed, nothing stored or accumulated) unless
// the run ends at a clip's end.  The consumer reads one tile instead of two, G2 / W2 / the bias sums lose their second pass, the
// sum is staged in the U tile so epilogue 2 never waits for the weight-gradient MMAs, and the control warp reloads x/ctx itself.
//
// The weight-gradient accumulators stay in TMEM for the CTA's whole tile loop and are written once
// per CTA as partial sums; a small second kernel reduces the partials in a fixed order
// (deterministic, no atomics) into the packed gradient buffer.
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"
#include "layer_tc.h"

using namespace tc;

namespace {

// Instrumented build (MOVENET_B200_NVCC_EXTRA=-DMVN_PHASE_CLOCKS=1): clock64() stamps of tile iterations 5..7 of CTA 0 for
// the control lane (CLKC), one worker thread (CLKW) and the latest worker (CLKM); printed by launch 20 when MVN_PROF is set.
#ifndef MVN_PHASE_CLOCKS
#define MVN_PHASE_CLOCKS 0
#endif
#if MVN_PHASE_CLOCKS
__device__ unsigned long long g_clk[3][3][20];
#define CLK_(role, i, cond) do { if (blockIdx.x == 0 && it >= 5 && it < 8 && (cond)) g_clk[role][it - 5][i] = clock64(); } while (0)
#define CLKW(i) CLK_(0, i, tid == 256)
#define CLKC(i) CLK_(1, i, leader)
#define CLKM(i) do { if (blockIdx.x == 0 && it >= 5 && it < 8) atomicMax(&g_clk[2][it - 5][i], (unsigned long long)clock64()); } while (0)
#else
#define CLKW(i) do {} while (0)
#define CLKC(i) do {} while (0)
#define CLKM(i) do {} while (0)
#endif

constexpr int W1_COL = 192, W2_COL = 384, B1_COL = 448, B2_COL = 464;
constexpr int PART_LD = 256;                       // partial row: 192 (dWz^T) + 64 (dWrs^T)
constexpr int PART_FLOATS = 128 * PART_LD + 256;   // + bias sums of the two A operands

struct BwdArgs {
    const void* img;
    const float* dskip;   // (B, Tout, S) fp32
    float* partial;       // [grid][PART_FLOATS]
    int B, T, Tout, RF, S, N2, dil, dil_up, nchunks, tiles_per_clip, n_tiles;
    int zero_in;          // the incoming stream gradient (P, U) and context-gradient sum (Q) are zero (last layer): never loaded
    int pair_in;          // the incoming gradient is the pair (P, U); otherwise one summed stream (or zero)
    int sum_out;          // write the summed stream D' (dilation <= 128) instead of the pair (P', U')
};

// The order in which a CTA visits its tiles.  Pair output: grid-strided.  Summed output: a contiguous run, backwards in
// time, preceded by a warm-up tile when the run does not end at the end of a clip.
struct TileSeq { int top, step, count, warm; };
__device__ __forceinline__ TileSeq tile_seq(const BwdArgs& a, bool sum_out) {
    TileSeq q;
    if (!sum_out) {
        q.top = blockIdx.x; q.step = gridDim.x; q.warm = 0;
        q.count = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    } else {
        const int lo = (int)((long long)a.n_tiles * blockIdx.x / gridDim.x), hi = (int)((long long)a.n_tiles * (blockIdx.x + 1) / gridDim.x);
        q.warm = hi < a.n_tiles && hi % a.tiles_per_clip != 0;
        q.top = q.warm ? hi : hi - 1; q.step = -1; q.count = q.top - lo + 1;
    }
    return q;
}

__host__ __device__ inline int bwd_tiles_off(int nc, int N2) { return smem_a_off(nc, N2); }
// tiles after the image: A0..A(nc-1) | DXS | DSK | U | DZ0 | DZ1 | G | [Q, video only] | ONES(1 KB) | barriers
// (audio only: a CARRY tile takes the Q tile's place; with video the carry (dilation <= 32) sits after the barriers)
constexpr int CARRY_VIDEO_ROWS = 32;
__host__ __device__ inline int bwd_smem_total(int nc, int N2) {
    return bwd_tiles_off(nc, N2) + (nc + 7) * TILE_BYTES + 1024 + 128 + (nc == 3 ? CARRY_VIDEO_ROWS * 128 : 0);
}

// every lane has fenced its own writes; one lane signals for the warp
__device__ __forceinline__ void warp_arrive(uint64_t* bar) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}

constexpr int N_WORKERS = 512, N_THREADS = N_WORKERS + 32;   // 16 worker warps + the control warp

template <bool SUM_OUT, bool PAIR_IN>   // == a.sum_out, a.pair_in (template parameters: each variant keeps only its own epilogue 2 and MMAs)
__global__ void __launch_bounds__(N_THREADS, 1)
layer_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_ctx,
                    const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_u,
                    const __grid_constant__ CUtensorMap map_pout, const __grid_constant__ CUtensorMap map_uout,
                    const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_qout,
                    const BwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int nc = a.nchunks;
    uint8_t* sBz = smem;
    uint8_t* sBrs = smem + smem_brs_off(nc);
    float* sbz = (float*)(smem + smem_bias_off(nc, a.N2));
    uint8_t* sA = smem + bwd_tiles_off(nc, a.N2);
    uint8_t* sU = sA + nc * TILE_BYTES;           // U | P | DSK in this order: [P|DSK] and [U|DSK] are both M=128 block pairs
    uint8_t* sDXS = sU + TILE_BYTES;              // the P tile (the stream gradient is P + U, never summed in memory)
    uint8_t* sDSK = sDXS + TILE_BYTES;
    uint8_t* sDZ = sDSK + TILE_BYTES;             // DZ0 (filter half) | DZ1 (gate half)
    uint8_t* sG = sDZ + 2 * TILE_BYTES;
    uint8_t* sQ = sG + TILE_BYTES;                // running sum of the context gradient (video only)
    // first d rows of U' of the tile processed before this one (summed output): audio: the spare tile; video: 4 KB after the barriers
    uint8_t* sCARRY = nc == 3 ? sQ + TILE_BYTES + 1024 + 128 : sQ;
    uint8_t* sONES = sQ + TILE_BYTES;
    // barriers, one completion per tile each (parity = tile iteration & 1), except IMG (once).
    // E_* are the worker -> control-warp signals (one arrival per worker warp), the rest are TMA / tcgen05.commit completions.
    enum { IMG = 0, A_IN, P_IN, U_IN, Q_IN, G1, G2, G3, W1, WALL, E_DSK, E_DZ, E_OUT, N_BARS };
    uint64_t* bar = (uint64_t*)(sONES + 1024);
    uint32_t* tmem_slot = (uint32_t*)(bar + N_BARS);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int r = tid & 127, sw = r & 7;          // row of the tile == TMEM lane; warps w and w+4 share a lane quarter
    const int half = (tid >> 7) & 3;              // ... and split the channel range between them (4 quarters of 16)
    const int NZ = nc * CC;                        // columns of D4 / dWz^T

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
    // constant tiles: DSK is zero outside the S live channels, ONES is all bf16 1.0
    for (int i = tid; i < TILE_BYTES / 16; i += N_THREADS) ((uint4*)sDSK)[i] = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 1024 / 4; i += N_THREADS) ((uint32_t*)sONES)[i] = 0x3F803F80u;
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

    // One tile at a time per CTA (TMEM and shared memory are full), but its phases overlap:
    //  * a control thread issues every TMA and MMA, so the 16 worker warps never wait for instruction issue
    //    and never meet at a CTA-wide barrier: the hand-offs are mbarriers in both directions;
    //  * G1 runs as soon as x/ctx are in, and the tanh/sigmoid half of epilogue 1 only needs G1, so it runs while P
    //    is still in flight and while G2 executes;
    //  * the outputs are staged in the DZ tiles (free once W1 is done), so the stores of tile i drain during
    //    tile i+1, and U(i+1), x/ctx(i+1), P(i+1) are loaded as soon as their buffers are free
    //    (after the pre-sum / after W1 / after W2).
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);   // warp-uniform copy: keeps the role branch convergent
    if (warp_u == N_WORKERS / 32) {
        // ================================ control warp ============================================
        // The whole warp runs the loop (so the code stays on the uniform datapath); one elected lane issues.
        const bool leader = elect_one();
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
        // base descriptors of every operand; one MMA's descriptors are these plus a compile-time offset
        const uint64_t kA = umma_desc(smem_u32(sA)), kBz = umma_desc(smem_u32(sBz)), kDXS = umma_desc(smem_u32(sDXS)),
                       kU = umma_desc(smem_u32(sU)), kDSK = umma_desc(smem_u32(sDSK)), kDZ = umma_desc(smem_u32(sDZ));
        const uint64_t mBrs = umma_desc_mn(smem_u32(sBrs), TILE_BYTES), mBz = umma_desc_mn(smem_u32(sBz), TILE_BYTES),
                       mDZ = umma_desc_mn(smem_u32(sDZ), TILE_BYTES), mA = umma_desc_mn(smem_u32(sA), TILE_BYTES),
                       mDXS = umma_desc_mn(smem_u32(sDXS), TILE_BYTES), mU = umma_desc_mn(smem_u32(sU), 2 * TILE_BYTES),
                       mG = umma_desc_mn(smem_u32(sG), TILE_BYTES);
        const uint64_t ones = umma_desc_mn_plain(smem_u32(sONES), 256, 128);
        const uint32_t iG1 = umma_idesc_major(TILE_T, 128, 0, 0);
        const uint32_t iG2 = umma_idesc_major(TILE_T, 64, 0, 1);
        const uint32_t iG3 = umma_idesc_major(TILE_T, NZ, 0, 1);
        const uint32_t iW1 = umma_idesc_major(TILE_T, NZ, 1, 1);
        const uint32_t iW2 = umma_idesc_major(TILE_T, 64, 1, 1);
        const uint32_t iB = umma_idesc_major(TILE_T, 16, 1, 1);
        auto load_a_tiles = [&](int lb, int l0) {       // x / ctx tiles -> A_IN
            mbar_expect_tx(bar + A_IN, (uint32_t)(nc * TILE_BYTES));
            tma_load_3d(sA, &map_x, bar + A_IN, 0, l0 - a.dil, lb);
            tma_load_3d(sA + TILE_BYTES, &map_x, bar + A_IN, 0, l0, lb);
            if (nc == 3) tma_load_3d(sA + 2 * TILE_BYTES, &map_ctx, bar + A_IN, 0, l0, lb);
        };
        auto load_tile = [&](uint8_t* dst, const CUtensorMap* map, int which, int lb, int l0) {
            mbar_expect_tx(bar + which, (uint32_t)TILE_BYTES);
            tma_load_3d(dst, map, bar + which, 0, l0, lb);
        };
        const TileSeq sq = tile_seq(a, SUM_OUT);
        if (leader) {
            const int lb = sq.top / a.tiles_per_clip, l0 = (sq.top - lb * a.tiles_per_clip) * TILE_T;
            load_a_tiles(lb, l0);
            if (!a.zero_in) load_tile(sDXS, &map_p, P_IN, lb, l0);
            if (PAIR_IN) load_tile(sU, &map_u, U_IN, lb, l0 + a.dil_up);
        }
        for (uint32_t it = 0; it < (uint32_t)sq.count; ++it) {
            const uint32_t ph = it & 1;
            const int tile = sq.top + (int)it * sq.step;
            const bool is_warm = sq.warm && it == 0;   // recomputed for its U' rows only: nothing stored, nothing accumulated
            const int b = tile / a.tiles_per_clip, t0 = (tile - b * a.tiles_per_clip) * TILE_T;
            const int nt = tile + sq.step;             // this CTA's next tile
            const bool has_next = it + 1 < (uint32_t)sq.count;
            const int nb = nt / a.tiles_per_clip, n0 = (nt - nb * a.tiles_per_clip) * TILE_T;
            // G1: recompute the gate pre-activations (the tile's TMEM columns are free: E_OUT of the previous tile)
            CLKC(0);
            mbar_wait(bar + A_IN, ph);
            CLKC(1);
            tc_fence_after();
            if (leader) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (c < nc)
                            umma(tmem_u, desc_adv(kA, c * TILE_BYTES + k * 32), desc_adv(kBz, c * TILE_BYTES + k * 32), iG1, (c | k) != 0);
                umma_commit(bar + G1);
            }
            CLKC(2);
            if (leader && has_next) {              // start pulling the next tile into L2 now
                tma_prefetch_3d(&map_x, 0, n0 - a.dil, nb);
                tma_prefetch_3d(&map_x, 0, n0, nb);
                if (nc == 3) tma_prefetch_3d(&map_ctx, 0, n0, nb);
                if (!a.zero_in) {
                    if (nc == 3) tma_prefetch_3d(&map_q, 0, n0, nb);
                    tma_prefetch_3d(&map_p, 0, n0, nb);
                    if (PAIR_IN) tma_prefetch_3d(&map_u, 0, n0 + a.dil_up, nb);
                }
            }
            if (leader) {
                tma_wait_read0();     // the previous tile's P'/U'/Q' stores have left DZ0 / DZ1 / Q  (ordered before G2's
                                      // commit: the workers write DZ again only after they have seen G2)
                if (nc == 3 && !a.zero_in && !is_warm) load_tile(sQ, &map_q, Q_IN, b, t0);       // needed by epilogue 2 only
            }
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
                for (int k = 0; k < 4; ++k)                // the skip channels: rows 64.. of the image, 16 per step
                    if (k < (a.S + 15) / 16) umma(tmem_u + 128, desc_adv(kDSK, k * 32), desc_adv(mBrs, (4 + k) * 2048), iG2, 1);
                umma_commit(bar + G2);
            }
            CLKC(6);
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
            // weight / bias gradients: K = time.  Every tile is [time x 64 ch], i.e. an MN-major operand.
            // First the ones that read the x/ctx and DZ tiles (W1): those buffers are needed first.
            const uint32_t acc0 = it != (uint32_t)sq.warm;      // the first tile that counts starts the accumulators
            if (!is_warm) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    umma(tmem_u + W1_COL, desc_adv(mDZ, k * 2048), desc_adv(mA, k * 2048), iW1, acc0 | (k != 0));
                    if (nc == 3)     // the gate biases are the context convs': no such parameters without video
                        umma(tmem_u + B1_COL, desc_adv(mDZ, k * 2048), ones, iB, acc0 | (k != 0));
                }
            }
            umma_commit(bar + W1);
            CLKC(9);
            // [P|DSK]^T and [U|DSK]^T: the skip rows (64..) are accumulated twice and halved at the flush (exact)
            if (!is_warm) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    umma(tmem_u + W2_COL, desc_adv(mDXS, k * 2048), desc_adv(mG, k * 2048), iW2, acc0 | (k != 0));
                    umma(tmem_u + B2_COL, desc_adv(mDXS, k * 2048), ones, iB, acc0 | (k != 0));
                }
                if (PAIR_IN) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        umma(tmem_u + W2_COL, desc_adv(mU, k * 2048), desc_adv(mG, k * 2048), iW2, 1);
                        umma(tmem_u + B2_COL, desc_adv(mU, k * 2048), ones, iB, 1);
                    }
                }
            }
            umma_commit(bar + WALL);
            CLKC(10);
            }
            // (the x/ctx tiles of the next tile are reloaded by worker thread 0 as soon as it has seen W1: this warp is
            // still blocked issuing W2 then, and the reload is the head of the next tile's dependency chain)
            CLKC(12);
            if (SUM_OUT && has_next) {             // summed output: the workers do not wait for W1, this warp reloads x/ctx
                mbar_wait(bar + W1, ph);
                if (leader) load_a_tiles(nb, n0);
            }
            mbar_wait(bar + E_OUT, ph);            // P', U', Q' are staged; nobody reads DXS or the tile's TMEM columns any more
            CLKC(13);
            if (leader && !is_warm) {
                tma_store_3d(&map_pout, SUM_OUT ? sU : sDZ, 0, t0, b);
                if (!SUM_OUT) tma_store_3d(&map_uout, sDZ + TILE_BYTES, 0, t0, b);
                if (nc == 3) tma_store_3d(&map_qout, sQ, 0, t0, b);
                tma_commit();
            }
            if (has_next) {
                CLKC(14);
                mbar_wait(bar + WALL, ph);         // W2 no longer reads the P and U tiles
                CLKC(15);
                if (leader && !a.zero_in) load_tile(sDXS, &map_p, P_IN, nb, n0);
                if (leader && PAIR_IN) load_tile(sU, &map_u, U_IN, nb, n0 + a.dil_up);
            }
            __syncwarp();
            CLKC(16);
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
        const TileSeq sq = tile_seq(a, SUM_OUT);
        float4 ds0, ds1;
        load_dskip(sq.top, ds0, ds1);
        const int o0 = r * 128 + (((2 * half) ^ sw) << 4), o1 = r * 128 + (((2 * half + 1) ^ sw) << 4);   // this thread's 16 channels
        for (uint32_t it = 0; it < (uint32_t)sq.count; ++it) {
            const uint32_t ph = it & 1;
            const int tile = sq.top + (int)it * sq.step;
            const bool is_warm = sq.warm && it == 0;
            const bool has_next = it + 1 < (uint32_t)sq.count;
            const int b = tile / a.tiles_per_clip, t0 = (tile - b * a.tiles_per_clip) * TILE_T;
            const int t = t0 + r;
            CLKW(0);
            if (it) mbar_wait(bar + WALL, ph ^ 1);      // the previous tile's weight-gradient MMAs are done with DSK and G
            if (half == 0) {   // d(skip) row of this thread -> bf16, logical channels [0, S) of the DSK tile
                *(uint4*)(sDSK + r * 128 + ((0 ^ sw) << 4)) =
                    make_uint4(pack_bf16(ds0.x, ds0.y), pack_bf16(ds0.z, ds0.w), pack_bf16(ds1.x, ds1.y), pack_bf16(ds1.z, ds1.w));
                const int js = t - (a.RF - 1);
                const bool live = t < a.T && js >= 0 && js < a.Tout;
                const float* src = a.dskip + ((size_t)b * a.Tout + (live ? js : 0)) * a.S;
                for (int s0 = 8; s0 < a.S; s0 += 8) {
                    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                    if (live) { v0 = ((const float4*)(src + s0))[0]; v1 = ((const float4*)(src + s0))[1]; }
                    *(uint4*)(sDSK + r * 128 + ((((s0 >> 3)) ^ sw) << 4)) =
                        make_uint4(pack_bf16(v0.x, v0.y), pack_bf16(v0.z, v0.w), pack_bf16(v1.x, v1.y), pack_bf16(v1.z, v1.w));
                }
            }
            if (half == 0) {       // (warp-uniform: only the four writing warps fence and signal)
                fence_proxy_async();
                warp_arrive(bar + E_DSK);
            }
            CLKW(1); CLKM(1);
            // ---- epilogue 1a (needs G1 only): th, sg, gated -> G tile ---------------------------------
            float th[16], sg[16];
            mbar_wait(bar + G1, ph);
            CLKW(2);
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
            CLKW(3); CLKM(3);
            // ---- epilogue 1b: gate derivative -> DZ0 | DZ1 ------------------------------------------
            mbar_wait(bar + G2, ph);
            CLKW(4);
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
            CLKW(5); CLKM(5);
            // ---- epilogue 2: U' = W0^T dz, P' = d(x') + W1^T dz (registers until the DZ tiles are free); Q' = Q + V^T dz in place
            mbar_wait(bar + G3, ph);
            CLKW(6);
            tc_fence_after();
            uint32_t po[8], uo[8];
            auto load_uo = [&]() {
                uint32_t w[16];
                tmem_ld16(tmem + lane_base + 16 * half, w);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) uo[i] = pack_bf16(__uint_as_float(w[2 * i]), __uint_as_float(w[2 * i + 1]));
            };
            if (!SUM_OUT) {
                load_uo();
                uint32_t v[16];
                tmem_ld16(tmem + lane_base + 64 + 16 * half, v);
                const uint4 x0 = *(const uint4*)(sDXS + o0), x1 = *(const uint4*)(sDXS + o1);
                const uint4 zz = make_uint4(0, 0, 0, 0);
                const uint4 y0 = PAIR_IN ? *(const uint4*)(sU + o0) : zz, y1 = PAIR_IN ? *(const uint4*)(sU + o1) : zz;
                const uint32_t xi[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                const uint32_t yi[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 xp = unpack_bf16(xi[i]), xu = unpack_bf16(yi[i]);
                    po[i] = pack_bf16(__uint_as_float(v[2 * i]) + (xp.x + xu.x), __uint_as_float(v[2 * i + 1]) + (xp.y + xu.y));
                }
            }
            if (nc == 3 && !is_warm) {
                uint32_t v[16];
                tmem_ld16(tmem + lane_base + 128 + 16 * half, v);
                CLKW(7);
                if (!a.zero_in) mbar_wait(bar + Q_IN, (it - (uint32_t)sq.warm) & 1);   // no Q load for the warm-up tile
                CLKW(8);
                uint4* p0 = (uint4*)(sQ + o0);
                uint4* p1 = (uint4*)(sQ + o1);
                const uint4 zz = make_uint4(0, 0, 0, 0);
                const uint4 x0 = a.zero_in ? zz : *p0, x1 = a.zero_in ? zz : *p1;
                const uint32_t xi[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                tmem_ld_wait();
                uint32_t o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 xv = unpack_bf16(xi[i]);
                    o[i] = pack_bf16(__uint_as_float(v[2 * i]) + xv.x, __uint_as_float(v[2 * i + 1]) + xv.y);
                }
                *p0 = make_uint4(o[0], o[1], o[2], o[3]);
                *p1 = make_uint4(o[4], o[5], o[6], o[7]);
            }
            CLKW(9);
            if (SUM_OUT) {
                // One summed stream: the U' term of row r is U' row r + d -- of this tile, exchanged through the U tile (nothing
                // else lives there: the input is one stream), or of the tile processed before this one (the carry: its first d
                // rows, written after that tile's first barrier and ordered before this read by E_OUT -> G1 -> G3).  The sum is
                // staged in the U tile too, so this epilogue never waits for the weight-gradient MMAs (they read the DZ tiles).
                load_uo();
                *(uint4*)(sU + o0) = make_uint4(uo[0], uo[1], uo[2], uo[3]);
                *(uint4*)(sU + o1) = make_uint4(uo[4], uo[5], uo[6], uo[7]);
                const int rs = r + a.dil;
                const int cs0 = ((2 * half) ^ (rs & 7)) << 4, cs1 = ((2 * half + 1) ^ (rs & 7)) << 4;
                uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
                if (rs >= TILE_T && it != 0 && tile % a.tiles_per_clip != a.tiles_per_clip - 1) {
                    c0 = *(const uint4*)(sCARRY + (rs - TILE_T) * 128 + cs0);
                    c1 = *(const uint4*)(sCARRY + (rs - TILE_T) * 128 + cs1);
                }
                uint32_t v[16];
                tmem_ld16(tmem + lane_base + 64 + 16 * half, v);
                asm volatile("bar.sync 1, %0;" ::"n"(N_WORKERS) : "memory");      // every U' row of this tile is in the U tile
                if (rs < TILE_T) {
                    c0 = *(const uint4*)(sU + rs * 128 + cs0);
                    c1 = *(const uint4*)(sU + rs * 128 + cs1);
                }
                if (r < a.dil) {                 // this tile's first d rows of U' are the next tile's carry
                    *(uint4*)(sCARRY + o0) = make_uint4(uo[0], uo[1], uo[2], uo[3]);
                    *(uint4*)(sCARRY + o1) = make_uint4(uo[4], uo[5], uo[6], uo[7]);
                }
                const uint4 x0 = *(const uint4*)(sDXS + o0), x1 = *(const uint4*)(sDXS + o1);
                const uint32_t xi[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                const uint32_t yi[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 xp = unpack_bf16(xi[i]), xu = unpack_bf16(yi[i]);
                    po[i] = pack_bf16(__uint_as_float(v[2 * i]) + (xp.x + xu.x), __uint_as_float(v[2 * i + 1]) + (xp.y + xu.y));
                }
                asm volatile("bar.sync 1, %0;" ::"n"(N_WORKERS) : "memory");      // every shifted row has been read
                *(uint4*)(sU + o0) = make_uint4(po[0], po[1], po[2], po[3]);
                *(uint4*)(sU + o1) = make_uint4(po[4], po[5], po[6], po[7]);
            } else {
                mbar_wait(bar + W1, ph);            // W1 no longer reads the DZ tiles, nor the x/ctx tiles:
                if (tid == 0 && has_next) {         // reload those for the next tile right away
                    const int nt = tile + sq.step, nb = nt / a.tiles_per_clip, n0 = (nt - nb * a.tiles_per_clip) * TILE_T;
                    mbar_expect_tx(bar + A_IN, (uint32_t)(nc * TILE_BYTES));
                    tma_load_3d(sA, &map_x, bar + A_IN, 0, n0 - a.dil, nb);
                    tma_load_3d(sA + TILE_BYTES, &map_x, bar + A_IN, 0, n0, nb);
                    if (nc == 3) tma_load_3d(sA + 2 * TILE_BYTES, &map_ctx, bar + A_IN, 0, n0, nb);
                }
                CLKW(10);
                *(uint4*)(sDZ + TILE_BYTES + o0) = make_uint4(uo[0], uo[1], uo[2], uo[3]);
                *(uint4*)(sDZ + TILE_BYTES + o1) = make_uint4(uo[4], uo[5], uo[6], uo[7]);
                *(uint4*)(sDZ + o0) = make_uint4(po[0], po[1], po[2], po[3]);
                *(uint4*)(sDZ + o1) = make_uint4(po[4], po[5], po[6], po[7]);
            }
            fence_proxy_async();
            tc_fence_before();
            warp_arrive(bar + E_OUT);
            CLKW(11); CLKM(11);
            load_dskip(has_next ? tile + sq.step : a.n_tiles, ds0, ds1);
        }
        mbar_wait(bar + WALL, (uint32_t)(sq.count - 1) & 1);
    }
    // ---- flush this CTA's partial weight / bias gradients ----------------------------------------
    if (tid < N_WORKERS) {
    tc_fence_after();
    float* part = a.partial + (size_t)blockIdx.x * PART_FLOATS;
    float* prow = part + (size_t)r * PART_LD;
#pragma unroll 1
    for (int j = half; j < NZ / 16; j += 4) {
        uint32_t v[16];
        tmem_ld16(tmem + lane_base + W1_COL + 16 * j, v);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q)
            ((float4*)(prow + 16 * j))[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                        __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
    }
#pragma unroll 1
    for (int j = half; j < 4; j += 4) {
        uint32_t v[16];
        tmem_ld16(tmem + lane_base + W2_COL + 16 * j, v);
        tmem_ld_wait();
        const float sc = (r >= CC && PAIR_IN) ? 0.5f : 1.f;     // pair input: the skip rows were accumulated once with P and once with U
#pragma unroll
        for (int q = 0; q < 4; ++q)
            ((float4*)(prow + 192 + 16 * j))[q] = make_float4(sc * __uint_as_float(v[4 * q]), sc * __uint_as_float(v[4 * q + 1]),
                                                              sc * __uint_as_float(v[4 * q + 2]), sc * __uint_as_float(v[4 * q + 3]));
    }
    {
        uint32_t v1[8], v2[8];
        tmem_ld8(tmem + lane_base + B1_COL, v1);
        tmem_ld8(tmem + lane_base + B2_COL, v2);
        tmem_ld_wait();
        if (half == 0) {
            part[128 * PART_LD + r] = __uint_as_float(v1[0]);
            part[128 * PART_LD + 128 + r] = ((r >= CC && PAIR_IN) ? 0.5f : 1.f) * __uint_as_float(v2[0]);
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

// Fixed-order reduction of the per-CTA partials into the packed gradient layout (layout.h):
//   dWz[k][2c+gate] = sum_cta part[gate*64+c][k] ; dbz likewise ; dWrs[k][n] = sum_cta part[n'][192+k] ; dbrs.
// blockIdx.y = layer: every layer's partials are reduced by ONE launch after the whole backward sweep
__global__ void tc_bwd_reduce_kernel(const float* __restrict__ partial_all, int n_cta, float* __restrict__ pg, PackedLayout P,
                                     int S, int Kz, int video, int n_layers) {
    MVN_PDL_PROLOGUE();
    const int layer = blockIdx.y;
    const float* partial = partial_all + (size_t)layer * 148 * PART_FLOATS;
    float* lg = pg + P.layer0 + (size_t)layer * P.layer_stride;
    const int has_resid = layer + 1 < n_layers;
    // one thread per SOURCE element so the n_cta reads of a warp are coalesced; the destination is scattered.  (176 MB of
    // partials for nine layers: this one is bandwidth-bound, a plain loop per column is the fastest shape.)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < PART_FLOATS; i += gridDim.x * blockDim.x) {
        float* dst = nullptr;
        if (i < 128 * PART_LD) {
            const int m = i / PART_LD, col = i % PART_LD;
            if (col < 192) {                       // dWz^T[m = gate*64 + c][k = col]
                if (col < Kz) dst = lg + P.oWz + (size_t)col * 128 + 2 * (m & 63) + (m >> 6);
            } else {                               // dWrs^T[m = c_out | 64 + s][k = col - 192]
                const int n = m;                   // C == 64: column n of Wrs is row m
                if (n < CC + S && (n >= CC || has_resid)) dst = lg + P.oWrs + (size_t)(col - 192) * (CC + S) + n;
            }
        } else {
            const int j = i - 128 * PART_LD;
            if (j < 128) { if (video) dst = lg + P.obz + 2 * (j & 63) + (j >> 6); }
            else { const int n = j - 128; if (n < CC + S && (n >= CC || has_resid)) dst = lg + P.obrs + n; }
        }
        if (!dst) continue;
        float acc = 0.f;
#pragma unroll 8
        for (int c = 0; c < n_cta; ++c) acc += partial[(size_t)c * PART_FLOATS + i];
        *dst = acc;
    }
}

}  // namespace

size_t mvn_tc_bwd_partial_bytes() { return (size_t)148 * PART_FLOATS * 4; }   // per layer

int mvn_tc_bwd_reduce_all(const float* partial_all, float* pg, const PackedLayout& P, const Geo& g, cudaStream_t st) {
    int grid_ctas = mvn_sm_count() < 148 ? mvn_sm_count() : 148;
    const int n_tiles = ((g.T + TILE_T - 1) / TILE_T) * g.B;
    if (grid_ctas > n_tiles) grid_ctas = n_tiles;
    dim3 grid((PART_FLOATS + 255) / 256, g.N);
    MVN_CUDA(mvn_launch_pdl(tc_bwd_reduce_kernel, dim3(grid), dim3(256), (size_t)(0), st, partial_all, grid_ctas, pg, P, g.S, g.Kz, g.video, g.N));
    return mvn_check_launch("tc_bwd_reduce");
}

int mvn_tc_layer_bwd(const void* x_in, const void* ctx, const void* p_in, const void* u_in, void* p_out, void* u_out,
                     const float* dskip, const void* q_in, void* q_out, const float* lw, float* lg, float* partial, const PackedLayout& P,
                     const Geo& g, int layer, cudaStream_t st) {
    MVN_REQUIRE(mvn_tc_layer_supported(g.C, g.S, g.video), "tensor-core layer kernel: unsupported channel counts");
    CUtensorMap mx, mc, mp, mu, mpo, muo, mq, mqo;
    int rc;
    if ((rc = make_act_map(&mx, x_in, g.B, g.T))) return rc;
    if ((rc = make_act_map(&mc, g.video ? ctx : x_in, g.B, g.T))) return rc;
    MVN_REQUIRE(p_in || !u_in, "tensor-core backward kernel: U without P");
    const int sum_out = u_out == nullptr;
    MVN_REQUIRE(!sum_out || (g.dil[layer] <= (g.video ? CARRY_VIDEO_ROWS : TILE_T) && !u_in),
                "tensor-core backward kernel: summed output needs a small dilation and a summed (or zero) input");
    if ((rc = make_act_map(&mp, p_in ? p_in : x_in, g.B, g.T))) return rc;      // p_in == u_in == null: zero incoming gradient
    if ((rc = make_act_map(&mu, u_in ? u_in : x_in, g.B, g.T))) return rc;
    if ((rc = make_act_map(&mpo, p_out, g.B, g.T))) return rc;
    if ((rc = make_act_map(&muo, u_out ? u_out : p_out, g.B, g.T))) return rc;
    if ((rc = make_act_map(&mq, g.video ? q_in : x_in, g.B, g.T))) return rc;
    if ((rc = make_act_map(&mqo, g.video ? q_out : p_out, g.B, g.T))) return rc;
    BwdArgs a;
    a.img = lw + P.oTc; a.dskip = dskip; a.partial = partial;
    a.B = g.B; a.T = g.T; a.Tout = g.Tout; a.RF = g.RF; a.S = g.S; a.N2 = ((g.C + g.S + 15) / 16) * 16;
    a.dil = g.dil[layer]; a.dil_up = layer + 1 < g.N ? g.dil[layer + 1] : 0;
    a.zero_in = p_in == nullptr; a.pair_in = u_in != nullptr; a.sum_out = sum_out;
    a.nchunks = g.video ? 3 : 2;
    a.tiles_per_clip = (g.T + TILE_T - 1) / TILE_T; a.n_tiles = a.tiles_per_clip * g.B;
    // (the carry rows behind the barriers exist only for the summed output with video)
    const int smem = bwd_smem_total(a.nchunks, a.N2) - ((a.nchunks == 3 && !sum_out) ? CARRY_VIDEO_ROWS * 128 : 0) + 1024;
    MVN_REQUIRE(smem <= 227 * 1024, "tensor-core backward kernel: shared memory budget exceeded (%d)", smem);
    static MvnSmemAttr attr_a, attr_b, attr_c;
    MVN_CUDA(mvn_ensure_smem(layer_bwd_tc_kernel<false, false>, smem, attr_a));
    MVN_CUDA(mvn_ensure_smem(layer_bwd_tc_kernel<false, true>, smem, attr_b));
    MVN_CUDA(mvn_ensure_smem(layer_bwd_tc_kernel<true, false>, smem, attr_c));
    int grid = mvn_sm_count() < 148 ? mvn_sm_count() : 148;      // the per-CTA partial buffers are sized for 148 CTAs
    if (grid > a.n_tiles) grid = a.n_tiles;
    auto kernel = sum_out ? layer_bwd_tc_kernel<true, false>       // (summed output takes a summed or zero input: checked above)
                          : (a.pair_in ? layer_bwd_tc_kernel<false, true> : layer_bwd_tc_kernel<false, false>);
    MVN_CUDA(mvn_launch_pdl(kernel, dim3(grid), dim3(N_THREADS), (size_t)smem, st, mx, mc, mp, mu, mpo, muo, mq, mqo, a));
    (void)lg;
#if MVN_PHASE_CLOCKS
    if (getenv("MVN_PROF")) {
        static int launches = 0;
        if (++launches == 20) {
            cudaDeviceSynchronize();
            unsigned long long h[3][3][20];
            cudaMemcpyFromSymbol(h, g_clk, sizeof(h));
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
    return mvn_check_launch("layer_bwd_tc");
}

// Opt-in (MOVENET_B200_BWD_SUM=1): layer l writes the summed stream when its dilation fits the carry and the layer above wrote a
// summed stream too (the U tile is the exchange / staging tile, so there is no pair input).  Default: the (P, U) pair everywhere --
// measured on one box, the summed stream moves 24 % fewer HBM bytes and its tile period is the pair's (5.2 us), but every CTA run
// pays one warm-up tile (26 -> 27 tile periods): 140 vs 137 us per launch.  The kernel is bound by the per-tile dependency chain
// (W1 done -> x/ctx reload -> G1), not by HBM or MMA count (profiles/r01_ablation.md).
int mvn_tc_bwd_sum_out(const Geo& g, int layer) {
    const char* sum = getenv("MOVENET_B200_BWD_SUM");      // read per call: the tests switch it
    if (!sum || !atoi(sum)) return 0;
    if (bwd_smem_total(g.video ? 3 : 2, ((g.C + g.S + 15) / 16) * 16) + 1024 > 227 * 1024) return 0;   // (video with 32 skip channels: no room for the carry)
    int above = 1;                                   // the top layer's input is zero
    for (int l = g.N - 1; l >= layer; --l) {
        const int s = g.dil[l] <= (g.video ? CARRY_VIDEO_ROWS : TILE_T) && above;   // (the U tile is the staging tile: no pair input)
        if (l == layer) return s;
        above = s;
    }
    return 0;
}
