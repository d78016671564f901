// Weight gradient of the causal input conv (movenet/modules.py:15-30) on tensor cores, C = 64, A <= 64 (AH = 1), <= 128 (AH = 2), <= 256 (AH = 4):
//   dW[c][a][tap] = sum_t d(h0)[t][c] * x[a][t-1+tap]
// With one-hot audio x is a one-hot matrix, so this is  OneHot^T . d(h0)  with K = time: the one-hot tiles
// [time x 64 codes] are built in shared memory from the integer codes (exact in bf16) and used as the
// MN-major M operand (tap 0 | tap 1 = the two 64-row halves of M = 128), d(h0) arrives as the (P, U) tile
// pair of the tensor-core backward (two accumulating MMA chains).  Columns that are not one-hot put
// their real values into the tile (bf16-rounded).  AH = 2 (experiments 03 / 04: A = 128): a tap's one-hot tile is two 64-code
// blocks = one M = 128 operand, one MMA chain and one 64-column accumulator per tap.
#include "tc_common.cuh"
#include "layer_tc.h"

using namespace tc;

namespace {

constexpr int IPART_MAX = 256 * 64;      // per-CTA partial: [2 taps x 64 AH codes][64 channels]

struct InArgs {
    const float* audio; const int* codes; const unsigned char* dense;
    float* partial;
    int B, T, A, dil0, tiles_per_clip, n_tiles;
    int has_u;            // the gradient comes as the pair (P, U); otherwise one summed stream
};

template <int AH>
__global__ void __launch_bounds__(128, AH == 1 ? 3 : (AH == 2 ? 2 : 1))
input_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_p, const __grid_constant__ CUtensorMap map_u, const InArgs a) {
    MVN_PDL_PROLOGUE();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* sOH = smem;                        // [tap][AH blocks of 64 codes]
    uint8_t* sP = smem + 2 * AH * TILE_BYTES;
    uint8_t* sU = sP + TILE_BYTES;
    uint64_t* full_bar = (uint64_t*)(sU + TILE_BYTES);
    uint64_t* w_bar = full_bar + 1;
    uint32_t* tmem_slot = (uint32_t*)(full_bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5, r = tid, sw = r & 7;

    if (tid == 0) { mbar_init(full_bar, 1); mbar_init(w_bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(64 * AH) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t idesc = umma_idesc_major(TILE_T, 64, 1, 1);
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);

    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int b = tile / a.tiles_per_clip, t0 = (tile - b * a.tiles_per_clip) * TILE_T, t = t0 + r;
        if (it) { mbar_wait(w_bar, (it - 1) & 1); tc_fence_after(); }     // the previous tile's MMAs are done with the tiles
        if (tid == 0) {
            mbar_expect_tx(full_bar, (a.has_u ? 2 : 1) * TILE_BYTES);
            tma_load_3d(sP, &map_p, full_bar, 0, t0, b);
            if (a.has_u) tma_load_3d(sU, &map_u, full_bar, 0, t0 + a.dil0, b);
        }
        // one-hot rows of this time step: tap 1 looks at x[t], tap 0 at x[t-1]
#pragma unroll
        for (int tap = 0; tap < 2; ++tap) {
            const int ts = t - 1 + tap;
            const bool in = t < a.T && ts >= 0;
            const long long gr = (long long)b * a.T + (in ? ts : 0);
            const bool is_dense = in && a.dense[gr];
            const int code = (in && !is_dense) ? a.codes[gr] : -1;
#pragma unroll
            for (int h = 0; h < AH; ++h) {
                uint8_t* row = sOH + (tap * AH + h) * TILE_BYTES + r * 128;
                if (is_dense) {
                    for (int q = 0; q < 8; ++q) {
                        float v[8];
                        for (int e = 0; e < 8; ++e) {
                            const int ch = 64 * h + 8 * q + e;
                            v[e] = ch < a.A ? a.audio[((size_t)b * a.A + ch) * a.T + ts] : 0.f;
                        }
                        *(uint4*)(row + ((q ^ sw) << 4)) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                    }
                } else {
                    const int cl = code - 64 * h;          // the code inside this 64-code block (or outside it)
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        uint4 z = make_uint4(0, 0, 0, 0);
                        if (code >= 0 && (cl >> 3) == q) {
                            const uint32_t one = (cl & 1) ? 0x3F800000u : 0x00003F80u;
                            const int w = (cl & 7) >> 1;
                            if (w == 0) z.x = one; else if (w == 1) z.y = one; else if (w == 2) z.z = one; else z.w = one;
                        }
                        *(uint4*)(row + ((q ^ sw) << 4)) = z;
                    }
                }
            }
        }
        fence_proxy_async();
        mbar_wait(full_bar, it & 1);
        tc_fence_before();
        __syncthreads();
        if (warp_u == 0) {             // warp-uniform issue through one elected lane (tc_common.cuh)
            tc_fence_after();
            const uint32_t acc0 = it != 0;
            const uint64_t mOH = umma_desc_mn(smem_u32(sOH), TILE_BYTES), mP = umma_desc_mn(smem_u32(sP), TILE_BYTES),
                           mU = umma_desc_mn(smem_u32(sU), TILE_BYTES);
            if (elect_one()) {
                // AH = 1: ONE M = 128 operand = (tap 0 | tap 1); AH = 2: one M = 128 operand (two 64-code blocks) per tap
#pragma unroll
                for (int m = 0; m < AH; ++m)
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint64_t dOH = desc_adv(mOH, m * 2 * TILE_BYTES + k * 2048);
                        umma(tmem_u + 64 * m, dOH, desc_adv(mP, k * 2048), idesc, acc0 | (k != 0));
                        if (a.has_u) umma(tmem_u + 64 * m, dOH, desc_adv(mU, k * 2048), idesc, 1);
                    }
                umma_commit(w_bar);
            }
            __syncwarp();
        }
    }
    if (it) mbar_wait(w_bar, (it - 1) & 1);
    tc_fence_after();
    // partial rows: AH = 1: row r = tap * 64 + code; AH = 2: accumulator m holds tap m, row r = code
#pragma unroll 1
    for (int m = 0; m < AH; ++m) {
        float* part = a.partial + (size_t)blockIdx.x * (AH * IPART_MAX / 2) + ((size_t)m * 128 + r) * 64;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            uint32_t v[16];
            tmem_ld16(tmem + lane_base + 64 * m + 16 * j, v);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q)
                ((float4*)(part + 16 * j))[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                            __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64 * AH) : "memory");
    }
}

// dwin[tap][a][c] = sum_cta part[tap * (64 AH) + a][c]
__global__ void input_reduce_kernel(const float* __restrict__ partial, int n_cta, float* __restrict__ dwin, int A, int AH) {
    MVN_PDL_PROLOGUE();
    const int ipart = AH * IPART_MAX / 2, span = 64 * AH;
    const int i = blockIdx.x * 32 + threadIdx.x;
    const int m = i >> 6, col = i & 63, tap = m / span, ch = m - tap * span;
    const bool valid = i < ipart && ch < A;
    const float acc = column_sum(partial, n_cta, (size_t)ipart, (size_t)i, valid);
    if (valid && threadIdx.y == 0) dwin[((size_t)tap * A + ch) * 64 + col] = acc;
}

}  // namespace

int mvn_tc_input_supported(int A, int C) { return C == 64 && A <= 256; }

static int input_grid(const Geo& g, int* AH_out) {
    const int tiles = ((g.T + TILE_T - 1) / TILE_T) * g.B;
    const int AH = g.A <= 64 ? 1 : (g.A <= 128 ? 2 : 4), per_sm = AH == 1 ? 3 : (AH == 2 ? 2 : 1);
    int grid = tiles < per_sm * mvn_sm_count() ? tiles : per_sm * mvn_sm_count();
    const int cap = AH == 1 ? 444 : 296;          // (the partial slot holds 2 x 148 x 33024 floats)
    if (grid > cap) grid = cap;
    *AH_out = AH;
    return grid;
}

int mvn_tc_input_reduce(float* dwin, const float* partial, const Geo& g, cudaStream_t st) {
    int AH; const int grid = input_grid(g, &AH);
    MVN_CUDA(mvn_launch_pdl(input_reduce_kernel, dim3((AH * IPART_MAX / 2 + 31) / 32), dim3(32, RED_SPLIT), (size_t)(0), st, partial, grid, dwin, g.A, AH));
    return mvn_check_launch("input_reduce");
}

int mvn_tc_input_bwd(const float* audio, const int* codes, const unsigned char* dense, const void* p, const void* u,
                     float* dwin, float* partial, const Geo& g, cudaStream_t st, int defer_reduce) {
    CUtensorMap mp, mu;
    int rc;
    if ((rc = make_act_map(&mp, p, g.B, g.T))) return rc;
    if ((rc = make_act_map(&mu, u ? u : p, g.B, g.T))) return rc;
    InArgs a;
    a.audio = audio; a.codes = codes; a.dense = dense; a.partial = partial;
    a.B = g.B; a.T = g.T; a.A = g.A; a.dil0 = g.dil[0]; a.has_u = u != nullptr;
    a.tiles_per_clip = (g.T + TILE_T - 1) / TILE_T; a.n_tiles = a.tiles_per_clip * g.B;
    int AH; const int grid = input_grid(g, &AH);
    const int smem = (2 * AH + 2) * TILE_BYTES + 64 + 1024;
    static MvnSmemAttr attr1, attr2, attr4;
    if (AH == 1) {
        MVN_CUDA(mvn_ensure_smem(input_bwd_tc_kernel<1>, smem, attr1));
        MVN_CUDA(mvn_launch_pdl(input_bwd_tc_kernel<1>, dim3(grid), dim3(128), (size_t)(smem), st, mp, mu, a));
    } else if (AH == 2) {
        MVN_CUDA(mvn_ensure_smem(input_bwd_tc_kernel<2>, smem, attr2));
        MVN_CUDA(mvn_launch_pdl(input_bwd_tc_kernel<2>, dim3(grid), dim3(128), (size_t)(smem), st, mp, mu, a));
    } else {          // A = 256 (the reference's test architecture): four M = 128 operands, 160 KB of one-hot tiles, one CTA per SM
        MVN_CUDA(mvn_ensure_smem(input_bwd_tc_kernel<4>, smem, attr4));
        MVN_CUDA(mvn_launch_pdl(input_bwd_tc_kernel<4>, dim3(grid), dim3(128), (size_t)(smem), st, mp, mu, a));
    }
    if ((rc = mvn_check_launch("input_bwd_tc")) || defer_reduce) return rc;
    return mvn_tc_input_reduce(dwin, partial, g, st);
}
