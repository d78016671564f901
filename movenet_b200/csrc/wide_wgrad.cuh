// Weight gradients of the wide path on tcgen05: dW[M x N] = A^T . B with K = TIME (the reduction runs over all B T rows).
//
// Both operands are time-major activations, i.e. MN-major for the tensor core (the channel index is contiguous, K = time
// strides by a row): a [64 time x 64 channel] TMA box in the 128-byte-swizzled layout is directly an MN-major operand block.
// A "job" is one [256 x <=512] block of a weight-gradient matrix -- 256 rows = a CTA pair (cta_group::2, 128 rows per CTA),
// <=512 fp32 columns = the whole TMEM of a CTA -- and the pairs of the grid split the time axis of every job between them:
// each pair accumulates its time slice in TMEM over ~100 k-blocks with no epilogue in between, then writes one fp32 partial;
// a second kernel adds the partials of a job in a fixed order (deterministic) straight into the packed-gradient layout.
// Per layer at C = S = 256: three jobs (the two taps of dWz, and d[Wr | Ws]), 126 GFLOP.
#pragma once
#include "wide_gemm.cuh"

namespace wide {

constexpr int WG_STAGE_BYTES = 3 * 16384;      // A [64 t x 128 ch] | B chunk 0 [64 t x 128] | B chunk 1 [64 t x 128]
constexpr int WG_STAGES = 4;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + 2048;
constexpr int WG_MAX_JOBS = 12;

struct WgJob {
    int a_map, a_c0, a_shift;      // A: tensor (0..4), first channel of the 256 rows of this block, time shift (dilation tap)
    int n;                         // columns of this block: 256 or 512
    int b_map[2], b_c0[2], b_n0;   // B: columns [0, b_n0) come from tensor b_map[0] at b_c0[0] + n, the rest from b_map[1]
    float* dst; int ld;            // where the reduced block goes (row-major, fp32)
};
struct WgArgs {
    int B, T, kb_per_clip, n_jobs, n_splits;
    float* partial;                // [n_jobs][n_splits][256][512] fp32
    // optional passenger of the reduce kernel: column sums of cs_rows partial rows [cs_cols] (the residual-bias gradient the
    // d(x) GEMM's epilogue left per CTA and staging warp), one extra row of blocks instead of one more launch per layer
    const float* cs_partial; float* cs_out; int cs_rows, cs_cols;
    WgJob job[WG_MAX_JOBS];
};

__global__ void __launch_bounds__(N_THREADS, 1)
wide_wgrad_kernel(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1, const __grid_constant__ CUtensorMap m2,
                  const __grid_constant__ CUtensorMap m3, const __grid_constant__ CUtensorMap m4, const WgArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = (uint64_t*)(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t* full = bars; uint64_t* empty = bars + WG_STAGES; uint64_t* tfull = empty + WG_STAGES;
    uint32_t* tmem_slot = (uint32_t*)(tfull + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster = blockIdx.x / 2;
    if (threadIdx.x == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    mvn_griddep_launch();
    mvn_griddep_wait();
    const uint32_t tmem = *tmem_slot;
    const int j = cluster / a.n_splits, split = cluster - j * a.n_splits;
    const bool active = j < a.n_jobs;
    const WgJob job = a.job[active ? j : 0];
    const int total_kb = a.B * a.kb_per_clip;
    const int per = (total_kb + a.n_splits - 1) / a.n_splits;
    const int kb0 = split * per, kb1 = kb0 + per < total_kb ? kb0 + per : total_kb;
    const int nchunks = job.n / 256;
    const bool work = active && kb0 < kb1;

    auto map_of = [&](int i) -> const CUtensorMap* { return i == 0 ? &m0 : i == 1 ? &m1 : i == 2 ? &m2 : i == 3 ? &m3 : &m4; };

    if (warp == 0) {
        if (work && elect_one()) {
            const CUtensorMap* ma = map_of(job.a_map);
            uint32_t it = 0;
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int b = kb / a.kb_per_clip, t0 = (kb - b * a.kb_per_clip) * 64;
                const int stage = it % WG_STAGES;
                mbar_wait_addr(smem_u32(empty + stage), ((it / WG_STAGES) & 1) ^ 1);
                const uint32_t s0 = smem_u32(smem + stage * WG_STAGE_BYTES);
                const uint32_t fb = mapa_rank(smem_u32(full + stage), 0);
                if (rank == 0) mbar_expect_tx_addr(smem_u32(full + stage), 2 * (16384 + nchunks * 16384));
                tma_a<2>(s0, ma, fb, job.a_c0 + 128 * (int)rank, t0 + job.a_shift, b);
                tma_a<2>(s0 + 8192, ma, fb, job.a_c0 + 128 * (int)rank + 64, t0 + job.a_shift, b);
                for (int c = 0; c < nchunks; ++c)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int n = 256 * c + 128 * (int)rank + 64 * h;
                        const int src = n < job.b_n0 ? 0 : 1;
                        tma_a<2>(s0 + 16384 * (1 + c) + 8192 * h, map_of(job.b_map[src]), fb, job.b_c0[src] + n - (src ? job.b_n0 : 0), t0, b);
                    }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (work && rank == 0) {
            const uint32_t idesc = umma_idesc_major(256, 256, 1, 1);
            uint32_t it = 0;
            for (int kb = kb0; kb < kb1; ++kb, ++it) {
                const int stage = it % WG_STAGES;
                mbar_wait_addr(smem_u32(full + stage), (it / WG_STAGES) & 1);
                tc_fence_after();
                const uint32_t s0 = smem_u32(smem + stage * WG_STAGE_BYTES);
                const uint64_t da = umma_desc_mn(s0, 8192);
                if (elect_one()) {
                    for (int c = 0; c < nchunks; ++c) {
                        const uint64_t db = umma_desc_mn(s0 + 16384 * (1 + c), 8192);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_p<2>(tmem + 256 * c, desc_adv(da, k * 2048), desc_adv(db, k * 2048), idesc, (kb > kb0) || k != 0);
                    }
                    commit_p<2>(smem_u32(empty + stage));
                    if (kb == kb1 - 1) commit_p<2>(smem_u32(tfull));
                }
                __syncwarp();
            }
        }
    } else if (work) {
        // one thread = one row of the block; the two warps of a lane quarter take one 256-column chunk each
        const int q = warp & 3, half = (warp - 2) >> 2;
        mbar_wait_addr(smem_u32(tfull), 0);
        tc_fence_after();
        if (half < nchunks) {
            const int m = 128 * (int)rank + 32 * q + lane;
            float* dst = a.partial + (((size_t)j * a.n_splits + split) * 256 + m) * 512 + 256 * half;
            const uint32_t tm = tmem + ((uint32_t)(32 * q) << 16) + 256 * half;
#pragma unroll 1
            for (int u = 0; u < 8; ++u) {
                uint32_t v[32];
                tmem_ld32(tm + 32 * u, v);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    ((float4*)(dst + 32 * u))[e] = make_float4(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1]),
                                                               __uint_as_float(v[4 * e + 2]), __uint_as_float(v[4 * e + 3]));
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// dst[m][n] = sum over the job's time slices, in slice order; one thread per four consecutive n (16-byte loads, eight slices in
// flight), 128 blocks per job: the 38 MB of partials of a layer stream at HBM speed
__global__ void __launch_bounds__(256) wide_wgrad_reduce_kernel(const WgArgs a) {
    MVN_PDL_PROLOGUE();
    const int j = blockIdx.y;
    if (j == a.n_jobs) {           // column sums: two columns per block, 128 threads per column, rows dealt round-robin, fixed tree
        __shared__ float red[256];
        const int c = 2 * blockIdx.x + (threadIdx.x >> 7), k = threadIdx.x & 127;
        float acc = 0.f;
        if (c < a.cs_cols)
            for (int r = k; r < a.cs_rows; r += 128) acc += a.cs_partial[(size_t)r * a.cs_cols + c];
        red[threadIdx.x] = acc;
        __syncthreads();
        for (int w = 64; w >= 1; w >>= 1) {
            if (k < w) red[threadIdx.x] += red[threadIdx.x + w];
            __syncthreads();
        }
        if (k == 0 && c < a.cs_cols) a.cs_out[c] = red[threadIdx.x];
        return;
    }
    const WgJob job = a.job[j];
    const int total_kb = a.B * a.kb_per_clip, per = (total_kb + a.n_splits - 1) / a.n_splits;
    const int used = (total_kb + per - 1) / per;            // slices that had work
    const int n4 = job.n / 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 256 * n4; i += gridDim.x * blockDim.x) {
        const int m = i / n4, n = (i - m * n4) * 4;
        const float4* p = (const float4*)(a.partial + ((size_t)j * a.n_splits * 256 + m) * 512 + n);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int s = 0;
        for (; s + 8 <= used; s += 8) {
            float4 v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = p[(size_t)(s + e) * (256 * 512 / 4)];
#pragma unroll
            for (int e = 0; e < 8; ++e) { acc.x += v[e].x; acc.y += v[e].y; acc.z += v[e].z; acc.w += v[e].w; }
        }
        for (; s < used; ++s) { const float4 v = p[(size_t)s * (256 * 512 / 4)]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
        *(float4*)(job.dst + (size_t)m * job.ld + n) = acc;
    }
}

}  // namespace wide
