// Data-parallel gradient averaging over NVLink / NVSwitch peer memory, fused in front of the gradient unpack.
//
// The reference averages gradients with DistributedDataParallel's bucketed NCCL all-reduce (movenet/trainer.py:230-234).  Here a
// step's gradients are ONE packed fp32 buffer per rank (2.65 MB at cfg01) that every rank finishes at the same moment -- the video
// encoder's weight gradient, 40 % of the bytes, is the LAST thing the backward produces, so there is nothing to overlap a bucket
// with and the exposed cost is the latency of one small all-reduce.  This file replaces the NCCL call by one kernel over peer
// memory (cudaIpc mappings of every rank's exchange buffer; one process per GPU, one node):
//
//   phase 1   rank r pushes slice j of its gradient regions into rank j's staging slot r, for every peer j
//   phase 2   rank r sums slice r over the staging slots in RANK ORDER (its own from its local buffer) -- every element is summed
//             by exactly one rank, so all ranks end up with bit-identical gradients (deterministic) -- and pushes the sums to
//             every peer
//   phase 3   every rank moves the sums it received to their place in the packed-gradient layout (in place)
//
// followed by the ordinary unpack kernels (packed layout -> the reference's parameter shapes, times 1/world).  Everything that
// crosses NVLink is a 16-byte STORE carrying 8 bytes of payload and the step number twice ("flag-in-data", the low-latency
// protocol of collective libraries): the receiver polls the line until both flags show this step, so there is no barrier, no
// system-scope fence and no dependence between thread blocks.
//
// Measured (B200 x 2 / x 8 on NVSwitch, 2.65 MB, device time of the exchange alone, scripts/dev/dp_exchange_time.py;
// profiles/r02_dp_exchange.md): this kernel 19.5 us at N = 2 and 35.4 us at N = 8; NCCL 2.28 all-reduce 13.4 and 26.3 us; whole
// training step at N = 8: 2.469 ms with this kernel, 2.408 ms with NCCL (N = 1: 2.339 ms).  Earlier versions: pulling with loads
// behind two flag barriers 24 us at N = 2, pushing with stores behind two barriers 32 us (a system-scope fence waits for every
// outstanding remote store).  The flag-in-data protocol doubles the bytes on the wire and NCCL reduces inside the switch;
// NCCL therefore STAYS THE DEFAULT and this exchange is opt-in ($MOVENET_B200_DP=peer), kept for its fixed summation order
// and as the tested starting point for a multimem (in-switch reduction) version.
// No buffer needs a guard: a peer can only push step e+1's slices after it has received ALL of this rank's step-e sums (sent
// after this rank consumed its staging slots), and step e+1's sums only after this rank's step-e+1 slices (sent by a kernel
// that follows step e's phase 3 in stream order).
#include "common.cuh"
#include "layout.h"
#include "../../include/movenet_b200.h"

namespace {

constexpr int MAX_PEERS = MVN_PEER_MAX, MAX_RANGES = 64, PEER_THREADS = 512;

struct PeerArgs {
    float* pg;                        // this rank's packed gradients of this step (local memory): reduced in place
    uint4* stage[MAX_PEERS];          // every rank's staging area: [world slots][per][2] lines, slot = the pushing rank
    uint4* recv[MAX_PEERS];           // every rank's receive area for the sums: [total4][2] lines
    int rank, world; unsigned epoch;
    int n_ranges;
    unsigned start4[MAX_RANGES + 1];  // prefix sums of the ranges' lengths (float4 units): the compact index space
    unsigned off4[MAX_RANGES];        // offset of every range in the packed layout (float4 units)
};

// one float4 travels as two 16-byte lines {value, step, value, step}
__device__ __forceinline__ void push(uint4* dst, const float4 v, unsigned step) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(__float_as_uint(v.x)), "r"(step), "r"(__float_as_uint(v.y)), "r"(step) : "memory");
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 1), "r"(__float_as_uint(v.z)), "r"(step), "r"(__float_as_uint(v.w)), "r"(step) : "memory");
}
__device__ __forceinline__ uint4 ld_line(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// A line that never arrives (a rank that died or skipped its backward) is a hang by construction: after ~5 s the kernel
// traps, which surfaces as a CUDA error on this rank.  `l0`, `l1` = a first look at the two lines (issued early by the caller so
// that several elements' polls are in flight together).
__device__ __forceinline__ float4 pull(uint4 l0, uint4 l1, const uint4* src, unsigned step, const PeerArgs& a, int from) {
    if (l0.y != step || l0.w != step || l1.y != step || l1.w != step) {
        const long long t0 = clock64();
        do {
            l0 = ld_line(src); l1 = ld_line(src + 1);
            if (clock64() - t0 > (10ll << 30)) {
                printf("movenet_b200: rank %d waited 5 s for rank %d in the gradient exchange (step %u)\n", a.rank, from, step);
                __trap();
            }
        } while (l0.y != step || l0.w != step || l1.y != step || l1.w != step);
    }
    return make_float4(__uint_as_float(l0.x), __uint_as_float(l0.z), __uint_as_float(l1.x), __uint_as_float(l1.z));
}
__device__ __forceinline__ int find_range(const unsigned* start4, int n, unsigned v) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (start4[mid] <= v) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <int WORLD>
__global__ void __launch_bounds__(PEER_THREADS) peer_allreduce_kernel(const PeerArgs a) {
    __shared__ unsigned s_start[MAX_RANGES + 1], s_off[MAX_RANGES];
    MVN_PDL_PROLOGUE();                    // (stream order: the backward's kernels are done and their writes visible)
    for (int i = threadIdx.x; i <= a.n_ranges; i += blockDim.x) { s_start[i] = a.start4[i]; if (i < a.n_ranges) s_off[i] = a.off4[i]; }
    __syncthreads();
    const unsigned total4 = s_start[a.n_ranges], per = (total4 + WORLD - 1) / WORLD;
    const unsigned stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lo = a.rank * per, hi = lo + per < total4 ? lo + per : total4;
    float4* pg = (float4*)a.pg;
    // Every loop keeps several independent loads in flight per thread (an L2 or NVLink access is microseconds, the loops are
    // a handful of iterations long): first all addresses and loads of U elements, then the stores / the flag checks.
    constexpr int U = 4;
    // phase 1: every peer's slice of my gradients goes into my slot of its staging area
    for (unsigned base = t0; base < total4; base += U * stride) {
        float4 x[U]; uint4* dst[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned v = base + u * stride, j = v / per;
            dst[u] = nullptr;
            if (v < total4 && j != (unsigned)a.rank) {
                const int r = find_range(s_start, a.n_ranges, v);
                x[u] = pg[(size_t)s_off[r] + (v - s_start[r])];
                dst[u] = a.stage[j] + ((size_t)a.rank * per + (v - j * per)) * 2;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (dst[u]) push(dst[u], x[u], a.epoch);
    }
    // phase 2: my slice, summed in rank order, goes to every peer (and into my own buffer)
    const uint4* my_stage = a.stage[a.rank];
    for (unsigned v = lo + t0; v < hi; v += stride) {
        const int r = find_range(s_start, a.n_ranges, v);
        const size_t o = (size_t)s_off[r] + (v - s_start[r]);
        uint4 l[WORLD][2];
#pragma unroll
        for (int p = 0; p < WORLD; ++p)
            if (p != a.rank) { const uint4* src = my_stage + ((size_t)p * per + (v - lo)) * 2; l[p][0] = ld_line(src); l[p][1] = ld_line(src + 1); }
        const float4 own = pg[o];
        float4 acc;
#pragma unroll
        for (int p = 0; p < WORLD; ++p) {
            const float4 x = p == a.rank ? own : pull(l[p][0], l[p][1], my_stage + ((size_t)p * per + (v - lo)) * 2, a.epoch, a, p);
            if (p == 0) acc = x; else { acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w; }
        }
        pg[o] = acc;
#pragma unroll
        for (int p = 0; p < WORLD; ++p)
            if (p != a.rank) push(a.recv[p] + (size_t)v * 2, acc, a.epoch);
    }
    // phase 3: the other slices' sums, as they arrive (the SAME thread read pg[o] of this element in phase 1: in place is safe)
    const uint4* my_recv = a.recv[a.rank];
    for (unsigned base = t0; base < total4; base += U * stride) {
        uint4 l[U][2]; size_t o[U]; bool on[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned v = base + u * stride;
            on[u] = v < total4 && v / per != (unsigned)a.rank;
            if (on[u]) {
                const int r = find_range(s_start, a.n_ranges, v);
                o[u] = (size_t)s_off[r] + (v - s_start[r]);
                l[u][0] = ld_line(my_recv + (size_t)v * 2); l[u][1] = ld_line(my_recv + (size_t)v * 2 + 1);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned v = base + u * stride;
            if (on[u]) pg[o[u]] = pull(l[u][0], l[u][1], my_recv + (size_t)v * 2, a.epoch, a, (int)(v / per));
        }
    }
}

// the regions of the packed-gradient buffer that mvn_unpack_grads reads (everything else in it is never written)
int gradient_ranges(const Geo& g, const PackedLayout& P, PeerArgs& a) {
    int n = 0; unsigned at = 0;
    auto add = [&](size_t off, size_t count) {
        a.off4[n] = (unsigned)(off / 4); a.start4[n] = at; at += (unsigned)((count + 3) / 4); ++n;
    };
    MVN_REQUIRE(g.N + 6 <= MAX_RANGES, "gradient exchange: too many layers");
    add(P.win, (size_t)2 * g.A * g.C);
    for (int l = 0; l < g.N; ++l) add(P.layer0 + (size_t)l * P.layer_stride + P.oWz, P.obrs + g.C + g.S - P.oWz);
    add(P.w1p, P.b2 + g.A - P.w1p);
    if (g.video) {
        add(P.wv, P.bv + g.C - P.wv);
        for (int i = 0; i < 3; ++i) add(P.wt[i], P.bt[i] + (size_t)10 * g.C - P.wt[i]);
    }
    a.n_ranges = n; a.start4[n] = at;
    return 0;
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

extern "C" int mvn_peer_layout(const mvn_shape_t* s, size_t* stage_bytes, size_t* recv_bytes) {
    Geo g; MVN_REQUIRE(s && geo_init(g, s) == 0, "mvn_peer_layout: bad shape");
    PackedLayout P; packed_layout(g, P);
    PeerArgs a; int rc = gradient_ranges(g, P, a); if (rc) return rc;
    const size_t lines = align256(((size_t)a.start4[a.n_ranges] + MAX_PEERS) * 32);      // two 16-byte lines per float4
    if (stage_bytes) *stage_bytes = lines;
    if (recv_bytes) *recv_bytes = lines;
    return 0;
}

extern "C" int mvn_peer_alloc(size_t bytes, void** ptr, void* handle64) {
    MVN_REQUIRE(ptr && handle64 && bytes > 0, "mvn_peer_alloc: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    MVN_CUDA(cudaMalloc(ptr, bytes));
    MVN_CUDA(cudaMemset(*ptr, 0, bytes));
    MVN_CUDA(cudaDeviceSynchronize());
    MVN_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, *ptr));
    return 0;
}
extern "C" int mvn_peer_open(const void* handle64, void** ptr) {
    MVN_REQUIRE(ptr && handle64, "mvn_peer_open: null argument");
    cudaIpcMemHandle_t h; memcpy(&h, handle64, sizeof(h));
    MVN_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int mvn_peer_close(void* ptr) { MVN_CUDA(cudaIpcCloseMemHandle(ptr)); return 0; }
extern "C" int mvn_peer_free(void* ptr) { MVN_CUDA(cudaFree(ptr)); return 0; }

extern "C" int mvn_peer_reduce_unpack(const mvn_shape_t* s, void* const* peer_base, int rank, int world, unsigned epoch,
                                      void* packed_grads, float* flat_grads, const int64_t* offsets_dev, float scale, void* stream) {
    Geo g; MVN_REQUIRE(s && geo_init(g, s) == 0, "mvn_peer_reduce_unpack: bad shape");
    MVN_REQUIRE(peer_base && packed_grads && flat_grads && offsets_dev, "mvn_peer_reduce_unpack: null buffer");
    MVN_REQUIRE(world >= 2 && world <= MAX_PEERS && rank >= 0 && rank < world, "mvn_peer_reduce_unpack: world size 2..%d", MAX_PEERS);
    MVN_REQUIRE(epoch != 0, "mvn_peer_reduce_unpack: steps are numbered from 1 (0 marks a line nobody has written)");
    PackedLayout P; packed_layout(g, P);
    PeerArgs a; memset(&a, 0, sizeof(a));
    int rc = gradient_ranges(g, P, a); if (rc) return rc;
    const size_t lines = align256(((size_t)a.start4[a.n_ranges] + MAX_PEERS) * 32);
    for (int r = 0; r < world; ++r) {
        MVN_REQUIRE(peer_base[r], "mvn_peer_reduce_unpack: null peer mapping");
        a.stage[r] = (uint4*)peer_base[r];
        a.recv[r] = (uint4*)((char*)peer_base[r] + lines);
    }
    a.pg = (float*)packed_grads; a.rank = rank; a.world = world; a.epoch = epoch;
    const dim3 grid(mvn_sm_count()), block(PEER_THREADS);
    cudaStream_t st = (cudaStream_t)stream;
    switch (world) {
        case 2: MVN_CUDA(mvn_launch_pdl(peer_allreduce_kernel<2>, grid, block, (size_t)0, st, a)); break;
        case 3: MVN_CUDA(mvn_launch_pdl(peer_allreduce_kernel<3>, grid, block, (size_t)0, st, a)); break;
        case 4: MVN_CUDA(mvn_launch_pdl(peer_allreduce_kernel<4>, grid, block, (size_t)0, st, a)); break;
        case 5: MVN_CUDA(mvn_launch_pdl(peer_allreduce_kernel<5>, grid, block, (size_t)0, st, a)); break;
        case 6: MVN_CUDA(mvn_launch_pdl(peer_allreduce_kernel<6>, grid, block, (size_t)0, st, a)); break;
        case 7: MVN_CUDA(mvn_launch_pdl(peer_allreduce_kernel<7>, grid, block, (size_t)0, st, a)); break;
        default: MVN_CUDA(mvn_launch_pdl(peer_allreduce_kernel<8>, grid, block, (size_t)0, st, a)); break;
    }
    if ((rc = mvn_check_launch("peer_allreduce"))) return rc;
    return mvn_unpack_grads(s, packed_grads, flat_grads, offsets_dev, scale, stream);
}
