// Wide-channel (residual_channels >= 128) tensor-core engine: ONE persistent, warp-specialised tcgen05 GEMM kernel
//
//     D[rows x N] = [A_seg0(t + shift0) | A_seg1(t + shift1) | ...] . W^T          (bf16 x bf16 -> fp32 in TMEM)
//
// whose A operand is a concatenation along K of column ranges of time-major (B, T, cols) bf16 activations, each with its
// own time shift -- a dilation tap of a k=2 dilated conv is "the same tensor, d rows earlier" (movenet/modules.py:36-46),
// out-of-range rows are zero-filled per clip by TMA -- and whose B operand is a weight matrix streamed from L2.  At these
// widths the weights of a layer (2C x 2C bf16 = 512 KB at C = 256) cannot stay in shared memory, so this is the classical
// weight-streaming pipeline:
//
//   warp 0   TMA producer : A k-block [128 rows x 64] + W k-block [256/PAIR rows x 64] per stage, 4-6 stage ring
//   warp 1   MMA issuer   : tcgen05.mma (M = 128 x PAIR, N <= 256, K = 16) x 4 per stage, accumulators in TMEM,
//                            two 256-column accumulators so the epilogue of one N chunk overlaps the MMAs of the next
//   warps 2-9 epilogue    : tcgen05.ld -> the fused element-wise tail of the op (gate, residual + skip, gate derivative,
//                            softmax, leaky-ReLU derivative ...) -> global memory, one thread = one time row
//
// PAIR = 2 runs the kernel as clusters of two CTAs on one TPC with tcgen05.mma.cta_group::2: M = 256 time rows per
// instruction, each CTA stages only HALF of the weight k-block (the pair shares it), which halves the L2 -> SM weight traffic
// that bounds the single-CTA variant.  Every fused layer / head stage of the wide path is an instantiation of this kernel with
// a different epilogue (wide.cu).
#pragma once
#include "tc_common.cuh"

namespace wide {
using namespace tc;

constexpr int BM = 128;                       // time rows per CTA = TMEM lanes
constexpr int BK = 64;                        // K per stage: one 128-byte swizzle row of bf16
constexpr int NCH = 256;                      // accumulator columns per N chunk
constexpr int A_STAGE_BYTES = BM * BK * 2;    // 16 KB
constexpr int N_EPI_WARPS = 8;
constexpr int N_THREADS = 32 * (2 + N_EPI_WARPS);
constexpr int OUT_BUF_BYTES = 32 * 128;                           // one staged store: 32 rows x 128 bytes
constexpr int OUT_STAGE_BYTES = N_EPI_WARPS * 2 * OUT_BUF_BYTES;  // two per epilogue warp: 64 KB

template <int PAIR> struct Cfg {
    static constexpr int B_ROWS = NCH / PAIR;                     // weight rows one CTA stages per k-block
    static constexpr int B_STAGE_BYTES = B_ROWS * BK * 2;
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
    static constexpr int STAGES = PAIR == 1 ? 3 : 5;              // 144 / 160 KB of operand ring
    static constexpr int SMEM = STAGES * STAGE_BYTES + OUT_STAGE_BYTES + 2048;   // + output staging, barriers, alignment slack
};

enum Epi {
    EPI_GATE = 0,        // gated = tanh(f + bf) * sigmoid(g + bg)                      (movenet/modules.py:73-80)
    EPI_RESID_SKIP,      // x' = r + br + x ; skip_sum += s + bs                       (movenet/modules.py:83-91, wavenet.py:181)
    EPI_STORE,           // bf16 store
    EPI_GATE_BWD,        // recomputed f, g + d(gated) -> gated, dz = (df, dg) interleaved
    EPI_ADD_STORE,       // d(x) = acc + d(x')
    EPI_HEAD1,           // a1 = acc + b1 (fp32) ; lrelu(a1) (bf16)                    (movenet/modules.py:139-141)
    EPI_HEAD2,           // z = acc + b2 -> softmax over channels -> (B, A, Tn) fp32   (movenet/wavenet.py:187-191)
    EPI_LRELU_BWD,       // out = acc * lrelu'(aux)  (aux fp32), bf16 store, optional row shift into the T row space
    EPI_DZ,              // dz = d(gated) * (a, b): the gate-derivative factors the forward kept, loaded by TMA into the staging tiles
    EPI_COUNT
};

struct Seg { int map, nkb, shift, c0; };      // source tensor map (0/1), k-blocks, time shift (rows), first column

struct Args {
    int B, rows, tiles_per_clip, n_tiles;     // rows per clip of the A operand's row space; a tile = BM * PAIR rows
    int nseg; Seg seg[3];
    int N, nkb;                               // output columns (multiple of 128), total k-blocks
    int b_row0, b_kb0;                        // first row / first k-block of the weight matrix this GEMM uses
    // epilogue operands (meaning depends on the epilogue)
    const float* bias;                        // [N] fp32 or null
    const void* aux; int ld_aux;              // per-row auxiliary input (bf16 or fp32), row space = the A operand's
    void* out; int ld_out;                    // primary output
    int out_c0;                               // first column of the primary output inside its (wider) tensor
    int out2_c0;                              // ... and of the secondary one (EPI_GATE: the derivative factors; EPI_DZ: where they are read)
    float* csum;                              // EPI_ADD_STORE, N <= 256: per-(CTA, lane quarter) column sums of the stored tile rows
    void* out2; int ld_out2;                  // secondary output
    float* skip;                              // (B, Tout, S) fp32 running skip sum
    int n_resid, S, Tout, RF, skip_init;      // EPI_RESID_SKIP
    int Tn, logits;                           // EPI_HEAD2: columns the caller receives; raw logits instead of softmax
    int out_rows, out_shift, aux_rows, aux_shift;   // EPI_LRELU_BWD: rows per clip of `out` / `aux` and the row shifts into them
    int out_fp32;                                   // EPI_LRELU_BWD: store fp32 instead of bf16
};

// ---- cluster / pair helpers ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_addr(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
template <int PAIR>
__device__ __forceinline__ void tma_a(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    if (PAIR == 1)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
template <int PAIR>
__device__ __forceinline__ void tma_b(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    if (PAIR == 1)
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
template <int PAIR>
__device__ __forceinline__ void umma_p(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    if (PAIR == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of every MMA issued so far -> one arrival on `bar` (PAIR = 2: on the barrier at the same offset in BOTH CTAs)
template <int PAIR>
__device__ __forceinline__ void commit_p(uint32_t bar) {
    if (PAIR == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else
        asm volatile("{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
                     "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
                     ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity) {
    uint32_t done, spins = 0;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 24)) asm volatile("trap;");      // a protocol bug must abort the kernel, not hang the GPU
    } while (!done);
}

// (volatile: under register pressure the compiler otherwise RE-EXECUTES a non-volatile tanh asm at every use of its result
// instead of keeping it in a register -- ncu showed 4 MUFU.TANH per channel instead of 2 in the gate-derivative epilogue)
__device__ __forceinline__ float tanh_v(float x) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_v(0.5f * x), 0.5f); }
__device__ __forceinline__ void st_bf16x8(void* p, const float* v) {
    *(uint4*)p = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}
__device__ __forceinline__ void ld_bf16x8(const void* p, float* v) {
    const uint4 u = *(const uint4*)p;
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

// ---- epilogue prefetch: the per-row auxiliary operand of an epilogue (x for the residual add, d(gated) for the gate
// derivative, d(x') for the data gradient) is loaded into registers BEFORE the thread waits for the accumulator, so its DRAM /
// L2 latency hides under the chunk's MMAs instead of being paid once per 32-column block by the only eight epilogue warps of
// the SM (measured at C = 256: residual + skip GEMM 187 -> see profiles/, gate-derivative GEMM 195 us before)
template <int EPI>
__device__ __forceinline__ void prefetch(const Args& a, uint4* pre, int nc, int n0, int b, int t, bool ok, int half) {
    const size_t row = (size_t)b * a.rows + t;
    if (EPI == EPI_GATE_BWD) {
        const __nv_bfloat16* dg = (const __nv_bfloat16*)a.aux + row * a.ld_aux + (n0 >> 1) + 64 * half;
#pragma unroll
        for (int i = 0; i < 8; ++i) pre[i] = ok ? *(const uint4*)(dg + 8 * i) : make_uint4(0, 0, 0, 0);
    } else if (EPI == EPI_RESID_SKIP || EPI == EPI_ADD_STORE) {
        const int cnt = nc / 64;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int n = n0 + 32 * (half * cnt + u);
            const bool need = ok && u < cnt && a.aux != nullptr && (EPI == EPI_ADD_STORE || n < a.n_resid);
            const __nv_bfloat16* x = (const __nv_bfloat16*)a.aux + row * a.ld_aux + n;
#pragma unroll
            for (int q = 0; q < 4; ++q) pre[4 * u + q] = need ? *(const uint4*)(x + 8 * q) : make_uint4(0, 0, 0, 0);
        }
    }
}
__device__ __forceinline__ void unpack8(const uint4 u, float* v) {
    float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}

// ---- output staging: scattered 16-byte global stores (one row per lane, rows 512+ bytes apart) cap an epilogue at ~1.9 TB/s
// (measured: the residual + skip GEMM spent 130 of its 180 us on them), so the layer epilogues write their results into a
// per-warp shared-memory tile in the 128-byte-swizzled layout and hand it to TMA: full-line, asynchronous stores (or fp32
// add-reductions for the skip sum) that cost the warp one instruction.  Two 4 KB tiles per warp alternate.
struct Stager {
    uint32_t base; int cur, lane;
    uint32_t row, sw;        // this lane's row inside tile 0, and its swizzle term (lane & 7) << 4
    uint32_t lbar[2], lph[2];   // per-tile load barriers (EPI_DZ: tiles are also filled by TMA) and their phases
    __device__ __forceinline__ void init(uint32_t b, int l, uint32_t bar0) {
        base = b; cur = 0; lane = l; row = b + l * 128; sw = (uint32_t)(l & 7) << 4;
        lbar[0] = bar0; lbar[1] = bar0 + 8; lph[0] = lph[1] = 0;
    }
    // TMA load of a [32 rows x 128 bytes] box into tile `buf` (the tile must not be in use by a store: wait_group.read first)
    __device__ __forceinline__ void load(int buf, const CUtensorMap* map, int c0, int c1, int c2) {
        if (lane == 0) {
            mbar_expect_tx_addr(lbar[buf], OUT_BUF_BYTES);
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(base + buf * OUT_BUF_BYTES), "l"(map), "r"(lbar[buf]), "r"(c0), "r"(c1), "r"(c2) : "memory");
        }
    }
    __device__ __forceinline__ void wait_load(int buf) { mbar_wait_addr(lbar[buf], lph[buf]); lph[buf] ^= 1; }
    __device__ __forceinline__ uint4 get(int buf, int k) {
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(row + buf * OUT_BUF_BYTES + (((uint32_t)k << 4) ^ sw)) : "memory");
        return v;
    }
    __device__ __forceinline__ void put_at(int buf, int k, uint4 v) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + buf * OUT_BUF_BYTES + (((uint32_t)k << 4) ^ sw)),
                     "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
    __device__ __forceinline__ void store_buf(int buf, const CUtensorMap* map, int c0, int c1, int c2) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                         ::"l"(map), "r"(base + buf * OUT_BUF_BYTES), "r"(c0), "r"(c1), "r"(c2) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    __device__ __forceinline__ void drain_reads() {      // every tile handed to a TMA store has been read
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
    }
    __device__ __forceinline__ void begin() {          // the tile about to be written was handed to TMA two stores ago
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
    }
    __device__ __forceinline__ void put(int k, uint4 v) {     // 16-byte chunk k (0..7) of this lane's 128-byte row
        const uint32_t addr = row + cur * OUT_BUF_BYTES + (((uint32_t)k << 4) ^ sw);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
    // column sums over the 32 rows of the tile that was flushed last (bf16, 64 columns): lane L owns columns 2L, 2L + 1
    __device__ __forceinline__ void colsum_last(float& s0, float& s1) {
        const uint32_t t0 = base + (cur ^ 1) * OUT_BUF_BYTES + (lane & 3) * 4;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
            uint32_t w;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(t0 + r * 128 + (((lane >> 2) ^ (r & 7)) << 4)) : "memory");
            const float2 v = unpack_bf16(w);
            s0 += v.x; s1 += v.y;
        }
    }
    __device__ __forceinline__ void flush(const CUtensorMap* map, int c0, int c1, int c2, bool add) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
            const uint32_t src = base + cur * OUT_BUF_BYTES;
            if (add) asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];"
                                  ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
            else asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                              ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        cur ^= 1;
    }
};
__device__ __forceinline__ uint4 pack8(const float* v) {
    return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
}

// EPI_DZ: request the factor tiles of this thread-half's first two 32-channel blocks before waiting for the accumulator
__device__ __forceinline__ void dz_prefetch(const Args& a, Stager& sg, const CUtensorMap* mO1, int nc, int n0, int b, int r0, int half) {
    const int cnt = nc / 64, ch = n0 + 32 * half * cnt;
    sg.drain_reads();                               // the previous chunk's stores have read both tiles
    sg.load(0, mO1, a.out2_c0 + 2 * ch, r0, b);
    if (cnt > 1) sg.load(1, mO1, a.out2_c0 + 2 * (ch + 32), r0, b);
}

// layer epilogues (TMA-stored): r0 = first row of this warp's 32 rows inside the clip; mO0 / mO1: tensor maps of the outputs
template <int EPI>
__device__ __forceinline__ void epilogue_tma(const Args& a, const uint4* pre, Stager& sg, const CUtensorMap* mO0, const CUtensorMap* mO1,
                                             uint32_t tm, int nc, int n0, int b, int r0, int half, float* cs) {
    if (EPI == EPI_GATE || EPI == EPI_GATE_BWD) {
        // chunk = [f of 128 channels | g of the same 128 channels]; this thread: channels 64 * half .. + 63 of them
        const int ch0 = (n0 >> 1) + 64 * half;
        uint4 og[8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            uint32_t f[32], g[32];
            tmem_ld32(tm + 64 * half + 32 * u, f);
            tmem_ld32(tm + 128 + 64 * half + 32 * u, g);
            tmem_ld_wait();
            // training forward (EPI_GATE with a second output): the gate's derivative factors are kept for the backward,
            //   a = d gated / df = sigma (1 - tanh^2),  b = d gated / dg = tanh sigma (1 - sigma),   interleaved (a_c, b_c),
            // so that the backward needs neither the recomputed pre-activations (an 8 C^2 GEMM per layer) nor tanh / sigmoid
            const bool keep = EPI == EPI_GATE && a.out2 != nullptr;
            if (EPI == EPI_GATE_BWD || keep) sg.begin();
            if (a.bias) {      // (context-conv biases; null on the audio-only wide path -- one uniform branch, not 64 predicated loads)
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    f[e] = __float_as_uint(__uint_as_float(f[e]) + a.bias[n0 + 64 * half + 32 * u + e]);
                    g[e] = __float_as_uint(__uint_as_float(g[e]) + a.bias[n0 + 128 + 64 * half + 32 * u + e]);
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float d[8], o[8], z0[8], z1[8];
                if (EPI == EPI_GATE_BWD) unpack8(pre[4 * u + q], d);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float th = tanh_v(__uint_as_float(f[8 * q + e])), sgm = sigmoid_fast(__uint_as_float(g[8 * q + e]));
                    o[e] = th * sgm;
                    // d tanh(f) sigma(g) / df = sigma (1 - tanh^2) ; / dg = tanh sigma (1 - sigma)
                    const float dd = EPI == EPI_GATE_BWD ? d[e] : 1.f;
                    const float df = (dd * sgm) * fmaf(-th, th, 1.f), dgv = (dd * o[e]) * (1.f - sgm);
                    if (e < 4) { z0[2 * e] = df; z0[2 * e + 1] = dgv; } else { z1[2 * (e - 4)] = df; z1[2 * (e - 4) + 1] = dgv; }
                }
                if (EPI == EPI_GATE) og[4 * u + q] = pack8(o);       // (the backward reads the gated activations the forward kept)
                if (EPI == EPI_GATE_BWD || keep) { sg.put(2 * q, pack8(z0)); sg.put(2 * q + 1, pack8(z1)); }
            }
            if (EPI == EPI_GATE_BWD) sg.flush(mO1, 2 * ch0 + 64 * u, r0, b, false);      // dz columns interleave (df c, dg c)
            if (keep) sg.flush(mO1, a.out2_c0 + 2 * ch0 + 64 * u, r0, b, false);
        }
        if (EPI == EPI_GATE) {
            sg.begin();
#pragma unroll
            for (int k = 0; k < 8; ++k) sg.put(k, og[k]);
            sg.flush(mO0, a.out_c0 + ch0, r0, b, false);
        }
        return;
    }
    if (EPI == EPI_DZ) {
        // dz[t][2c], dz[t][2c+1] = d(gated)[t][c] * (a_c, b_c).  The factor tile of a 32-channel block ([32 rows x 128 bytes]:
        // 64 interleaved bf16) was requested by dz_prefetch() (blocks 0, 1) or right after the tile's previous store; the
        // products overwrite it in place and the same tile is handed to the TMA store.
        const int cnt = nc / 64;
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
            if (uu >= cnt) break;                  // (warp-uniform)
            const int u = half * cnt + uu, ch = n0 + 32 * u, buf = uu & 1;
            uint32_t v[32];
            tmem_ld32(tm + 32 * u, v);
            tmem_ld_wait();
            sg.wait_load(buf);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint4 ab = sg.get(buf, k);
                const uint32_t w[4] = {ab.x, ab.y, ab.z, ab.w};
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 f2 = unpack_bf16(w[e]);
                    const float d = __uint_as_float(v[4 * k + e]);
                    o[e] = pack_bf16(d * f2.x, d * f2.y);
                }
                sg.put_at(buf, k, make_uint4(o[0], o[1], o[2], o[3]));
            }
            sg.store_buf(buf, mO0, 2 * ch, r0, b);
            if (uu + 2 < cnt) {                    // this tile's next factor block, as soon as the store has read it
                sg.drain_reads();
                sg.load(buf, mO1, a.out2_c0 + 2 * (ch + 64), r0, b);
            }
        }
        return;
    }
    // column-block epilogues: this thread handles columns [half * nc/2, (half + 1) * nc/2) of the chunk, 32 per TMEM load,
    // 64 bf16 (or 32 fp32) columns per staged store
    const int cnt = nc / 64;
#pragma unroll
    for (int uu = 0; uu < 4; ++uu) {
        if (uu >= cnt) break;                  // (warp-uniform)
        const int u = half * cnt + uu;
        const int n = n0 + 32 * u;
        const bool is_skip = EPI == EPI_RESID_SKIP && n >= a.n_resid;      // (never straddles a pair: n_resid % 128 == 0)
        uint32_t v[32];
        tmem_ld32(tm + 32 * u, v);
        tmem_ld_wait();
        if (is_skip) {       // skip_sum (fp32): layer 0 stores, later layers add (TMA reduction, executed by L2: in layer order)
            sg.begin();
#pragma unroll
            for (int q = 0; q < 8; ++q)
                sg.put(q, make_uint4(__float_as_uint(__uint_as_float(v[4 * q]) + a.bias[n + 4 * q]), __float_as_uint(__uint_as_float(v[4 * q + 1]) + a.bias[n + 4 * q + 1]),
                                     __float_as_uint(__uint_as_float(v[4 * q + 2]) + a.bias[n + 4 * q + 2]), __float_as_uint(__uint_as_float(v[4 * q + 3]) + a.bias[n + 4 * q + 3])));
            sg.flush(mO1, n - a.n_resid, r0, b, !a.skip_init);      // (the wide path keeps skip_sum on the T row space, see wide.cu)
            continue;
        }
        if ((uu & 1) == 0) sg.begin();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[8 * q + e]);
            if (EPI == EPI_RESID_SKIP) {
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] += a.bias[n + 8 * q + e];
            }
            if (EPI == EPI_RESID_SKIP || (EPI == EPI_ADD_STORE && a.aux)) {
                float xv[8];
                unpack8(pre[4 * uu + q], xv);
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] += xv[e];
            }
            sg.put(4 * (uu & 1) + q, pack8(o));
        }
        if (uu & 1) {
            sg.flush(mO0, n - 32, r0, b, false);
            // bias gradient of the layer below = column sums of this d(x): taken from the staged tile (rows past the clip are zero)
            if (EPI == EPI_ADD_STORE && a.csum) sg.colsum_last(cs[uu & 2], cs[(uu & 2) + 1]);
        }
    }
}

// ---- epilogues: one thread owns one time row of the tile; `half` selects which half of the chunk's columns it handles --
// tm: TMEM address of this thread's lane at the first column of the accumulator chunk; nc: columns of this chunk (128 / 256);
// n0: first output column of the chunk; b, t: clip and row (row space of the A operand); ok: the row exists.
template <int EPI>
__device__ __forceinline__ void epilogue(const Args& a, const uint4* pre, uint32_t tm, int nc, int n0, int b, int t, bool ok, int half) {
    const size_t row = (size_t)b * a.rows + t;
    (void)pre;
    if (EPI == EPI_HEAD2) {
        // softmax over all nc (= A <= 256) columns of the row: both warps of a lane quarter compute the statistics, each writes
        // its half of the channels; lanes are consecutive time steps, so the channels-first (B, A, Tn) store is coalesced
        const bool live = ok && t < a.Tn;
        float m = -INFINITY, s = 0.f;
        if (!a.logits) {
#pragma unroll 1
            for (int u = 0; u < nc / 32; ++u) {
                uint32_t v[32];
                tmem_ld32(tm + 32 * u, v);
                tmem_ld_wait();
                float m2 = m;
#pragma unroll
                for (int e = 0; e < 32; ++e) m2 = fmaxf(m2, __uint_as_float(v[e]) + a.bias[n0 + 32 * u + e]);
                s *= __expf(m - m2);
#pragma unroll
                for (int e = 0; e < 32; ++e) s += __expf(__uint_as_float(v[e]) + a.bias[n0 + 32 * u + e] - m2);
                m = m2;
            }
        }
        const float inv = a.logits ? 1.f : 1.f / s;
        float* out = (float*)a.out + (size_t)b * a.N * a.Tn + t;
#pragma unroll 1
        for (int u = half * (nc / 64); u < (half + 1) * (nc / 64); ++u) {
            uint32_t v[32];
            tmem_ld32(tm + 32 * u, v);
            tmem_ld_wait();
            if (!live) continue;
#pragma unroll
            for (int e = 0; e < 32; ++e) {
                const float z = __uint_as_float(v[e]) + a.bias[n0 + 32 * u + e];
                out[(size_t)(n0 + 32 * u + e) * a.Tn] = a.logits ? z : __expf(z - m) * inv;
            }
        }
        return;
    }
    // column-block epilogues: this thread handles columns [half * nc/2, (half + 1) * nc/2) of the chunk in blocks of 32
    const int cnt = nc / 64;
#pragma unroll
    for (int uu = 0; uu < 4; ++uu) {
        if (uu >= cnt) break;                  // (warp-uniform)
        const int u = half * cnt + uu;
        uint32_t v[32];
        tmem_ld32(tm + 32 * u, v);
        tmem_ld_wait();
        const int n = n0 + 32 * u;
        if (!ok) continue;
        if (EPI == EPI_HEAD1) {
            if (t >= a.Tn) continue;
            float* a1 = (float*)a.out + ((size_t)b * a.Tn + t) * a.ld_out + n;
            __nv_bfloat16* l1 = (__nv_bfloat16*)a.out2 + row * a.ld_out2 + n;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float o[8], l[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) { o[e] = __uint_as_float(v[8 * q + e]) + a.bias[n + 8 * q + e]; l[e] = mvn_lrelu(o[e]); }
                ((float4*)(a1 + 8 * q))[0] = make_float4(o[0], o[1], o[2], o[3]);
                ((float4*)(a1 + 8 * q))[1] = make_float4(o[4], o[5], o[6], o[7]);
                st_bf16x8(l1 + 8 * q, l);
            }
        } else if (EPI == EPI_LRELU_BWD) {
            if (t >= a.Tn) continue;          // the dropped last column keeps a zero gradient
            const float* pre = (const float*)a.aux + ((size_t)b * a.aux_rows + t + a.aux_shift) * a.ld_aux + n;
            const size_t orow = ((size_t)b * a.out_rows + t + a.out_shift) * a.ld_out + n;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 p0 = ((const float4*)(pre + 8 * q))[0], p1 = ((const float4*)(pre + 8 * q))[1];
                const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
                float o[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(v[8 * q + e]) * mvn_lrelu_grad(pv[e]);
                if (a.out_fp32) {
                    float4* d4 = (float4*)((float*)a.out + orow + 8 * q);
                    d4[0] = make_float4(o[0], o[1], o[2], o[3]); d4[1] = make_float4(o[4], o[5], o[6], o[7]);
                } else st_bf16x8((__nv_bfloat16*)a.out + orow + 8 * q, o);
            }
        }
    }
}

template <int PAIR, int EPI>
__global__ void __launch_bounds__(N_THREADS, 1)
wide_gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapO0,
                 const __grid_constant__ CUtensorMap mapO1, const Args a) {
    using C = Cfg<PAIR>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* out_stage = smem + C::STAGES * C::STAGE_BYTES;
    uint64_t* bars = (uint64_t*)(out_stage + OUT_STAGE_BYTES);
    uint64_t* full = bars;                     // [STAGES] operands landed (the pair's barriers live in the leader CTA)
    uint64_t* empty = bars + C::STAGES;        // [STAGES] operands consumed (one per CTA, signalled by the multicast commit)
    uint64_t* tfull = empty + C::STAGES;       // [2] accumulator chunk complete (one per CTA)
    uint64_t* tempty = tfull + 2;              // [2] accumulator chunk drained by the epilogue warps of every CTA of the pair (leader)
    uint64_t* lbars = tempty + 2;              // [N_EPI_WARPS][2] staging-tile load barriers (EPI_DZ)
    uint32_t* tmem_slot = (uint32_t*)(lbars + 2 * N_EPI_WARPS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR == 1 ? 0u : cluster_ctarank();
    const int cluster = blockIdx.x / PAIR, n_clusters = gridDim.x / PAIR;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, N_EPI_WARPS * PAIR); }
        for (int s = 0; s < 2 * N_EPI_WARPS; ++s) mbar_init(lbars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (PAIR == 1) __syncthreads(); else cluster_sync_all();
    tc_fence_after();
    mvn_griddep_launch();
    mvn_griddep_wait();
    const uint32_t tmem = *tmem_slot;
    const int nchunks = (a.N + NCH - 1) / NCH;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            uint32_t it = 0;
            for (int tile = cluster; tile < a.n_tiles; tile += n_clusters) {
                const int b = tile / a.tiles_per_clip, t0 = ((tile - b * a.tiles_per_clip) * PAIR + (int)rank) * BM;
                for (int j = 0; j < nchunks; ++j) {
                    const int nc = a.N - NCH * j < NCH ? a.N - NCH * j : NCH;
                    int kb = 0;
                    for (int sg = 0; sg < a.nseg; ++sg) {
                        const Seg s = a.seg[sg];
                        const CUtensorMap* mp = s.map ? &mapA1 : &mapA0;
                        for (int i = 0; i < s.nkb; ++i, ++kb, ++it) {
                            const int stage = it % C::STAGES;
                            const uint32_t ph = (it / C::STAGES) & 1;
                            mbar_wait_addr(smem_u32(empty + stage), ph ^ 1);
                            const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES), sb = sa + A_STAGE_BYTES;
                            uint32_t fb = smem_u32(full + stage);
                            if (PAIR == 2) fb = mapa_rank(fb, 0);
                            if (rank == 0) mbar_expect_tx_addr(smem_u32(full + stage), PAIR * C::STAGE_BYTES);
                            tma_a<PAIR>(sa, mp, fb, s.c0 + BK * i, t0 + s.shift, b);
                            tma_b<PAIR>(sb, &mapB, fb, BK * (a.b_kb0 + kb), a.b_row0 + NCH * j + (int)rank * (nc / PAIR));
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer (the leader CTA of a pair) =====
        if (rank == 0) {
            uint32_t it = 0, acc = 0;
            for (int tile = cluster; tile < a.n_tiles; tile += n_clusters) {
                for (int j = 0; j < nchunks; ++j, ++acc) {
                    const int nc = a.N - NCH * j < NCH ? a.N - NCH * j : NCH;
                    const uint32_t buf = acc & 1, idesc = umma_idesc(BM * PAIR, nc);
                    mbar_wait_addr(smem_u32(tempty + buf), ((acc >> 1) & 1) ^ 1);
                    tc_fence_after();
                    for (int kb = 0; kb < a.nkb; ++kb, ++it) {
                        const int stage = it % C::STAGES;
                        mbar_wait_addr(smem_u32(full + stage), (it / C::STAGES) & 1);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
                        const uint64_t da = umma_desc(sa), db = umma_desc(sa + A_STAGE_BYTES);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_p<PAIR>(tmem + buf * NCH, desc_adv(da, k * 32), desc_adv(db, k * 32), idesc, (kb | k) != 0);
                            commit_p<PAIR>(smem_u32(empty + stage));
                            if (kb == a.nkb - 1) commit_p<PAIR>(smem_u32(tfull + buf));
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else {
        // ===== epilogue warps: lane quarter = warp % 4 (the TMEM lanes a warp may read), two warps per quarter split the columns
        const int q = warp & 3, half = (warp - 2) >> 2;
        constexpr bool TMA_OUT = EPI == EPI_GATE || EPI == EPI_GATE_BWD || EPI == EPI_RESID_SKIP || EPI == EPI_STORE || EPI == EPI_ADD_STORE ||
                                 EPI == EPI_DZ;
        Stager sg; sg.init(smem_u32(out_stage + (warp - 2) * 2 * OUT_BUF_BYTES), lane, smem_u32(lbars + 2 * (warp - 2)));
        uint32_t acc = 0;
        float cs[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t te[2];
        for (int s = 0; s < 2; ++s) { te[s] = smem_u32(tempty + s); if (PAIR == 2) te[s] = mapa_rank(te[s], 0); }
        for (int tile = cluster; tile < a.n_tiles; tile += n_clusters) {
            const int b = tile / a.tiles_per_clip, t0 = ((tile - b * a.tiles_per_clip) * PAIR + (int)rank) * BM;
            const int t = t0 + 32 * q + lane;
            const bool ok = t < a.rows;
            for (int j = 0; j < nchunks; ++j, ++acc) {
                const int nc = a.N - NCH * j < NCH ? a.N - NCH * j : NCH;
                const uint32_t buf = acc & 1;
                uint4 pre[16];
                prefetch<EPI>(a, pre, nc, NCH * j, b, t, ok, half);
                if constexpr (EPI == EPI_DZ) dz_prefetch(a, sg, &mapO1, nc, NCH * j, b, t0 + 32 * q, half);
                mbar_wait_addr(smem_u32(tfull + buf), (acc >> 1) & 1);
                tc_fence_after();
                if constexpr (TMA_OUT) epilogue_tma<EPI>(a, pre, sg, &mapO0, &mapO1, tmem + ((uint32_t)(32 * q) << 16) + buf * NCH, nc, NCH * j, b, t0 + 32 * q, half, cs);
                else epilogue<EPI>(a, pre, tmem + ((uint32_t)(32 * q) << 16) + buf * NCH, nc, NCH * j, b, t, ok, half);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) { if (PAIR == 1) mbar_arrive(tempty + buf); else mbar_arrive_remote(te[buf]); }
            }
        }
        if (TMA_OUT && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the staged tiles are read before the CTA exits
        if (EPI == EPI_ADD_STORE && a.csum) {      // one partial row per (CTA, lane quarter); N <= 256: a single chunk per tile
            float* dst = a.csum + ((size_t)blockIdx.x * 4 + q) * a.N + (a.N / 2) * half + 2 * lane;
            dst[0] = cs[0]; dst[1] = cs[1];
            if (a.N > 128) { dst[64] = cs[2]; dst[65] = cs[3]; }
        }
    }
    tc_fence_before();
    if (PAIR == 1) __syncthreads(); else cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        if (PAIR == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

}  // namespace wide
