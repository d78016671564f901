// Cached autoregressive decoding: WaveNet.generate (movenet/wavenet.py:193-239) without the
// per-sample window recompute.  Every layer keeps a ring of its last d inputs ("dilation queue"),
// so one new sample costs one pass over the N layers instead of RF of them.
//
// fp32 (exact) path: one CTA owns CB clips for the whole run and loops over the samples; the
// layer weights are staged in shared memory when they fit (they do for the receptive-field
// config, experiments/04), the queues live in HBM laid out (layer, slot, clip, channel) so a CTA's
// pops and pushes are contiguous across its clips.
#include "common.cuh"
#include "layout.h"

struct DecodeArgs {
    const float* packed; PackedLayout P;
    int N, A, C, S, Kz, video, B, Tctx, ctx_dtype;
    int t_start, n_new, smem_weights;
    float temperature; unsigned seed;
    float* queues; int* last2;           // [B][2] : code[t-2], code[t-1]
    const void* ctx;
    int* out_codes; float* out_logits;
    long long qoff[MVN_MAX_LAYERS];      // element offset of layer l's ring (per clip), before the *B factor
    int dil[MVN_MAX_LAYERS];
    // reference-window mode (stack_size == 1): see the comment above decode_edge_geometry
    int edge, RF;
    int hist[MVN_MAX_LAYERS];            // ring depth of layer l (== dil[l] without the edge chain)
    int age[MVN_MAX_LAYERS];             // the edge chain reads x_l[tau - age[l]]
    int* code_ring;                      // [RF][B] : code at absolute position p lives in slot p % RF
};

// The reference's generate() (movenet/wavenet.py:217-224) evaluates a window of exactly RF samples per step and
// CausalConv1d zero-pads the window's left edge (movenet/modules.py:15-30), so the leftmost column of the window's h0
// lacks its W[:,:,0] x[i-RF-1] term.  Every layer drops d columns on the left (movenet/modules.py:36-46), so exactly ONE
// column per layer -- the leftmost one, at absolute position p_l = tau - sum_{k>l} d_k (tau = i - 1) -- descends from
// that perturbed column.  With stack_size >= 2 it is cut off before the output; with stack_size == 1 the last layer's
// only column IS the leftmost one.  The skip outputs of layers 0..N-2 at tau are unaffected (movenet/modules.py:90-91
// keeps the last column).  Reference-exact decoding therefore = the ordinary queue update, plus the "edge chain"
//     e_0     = W[:,:,1] x[i-RF]
//     e_{l+1} = Wr_l gate(Wz0_l e_l + Wz1_l x_l[p_l] (+ V_l ctx[p_l])) + br_l + x_l[p_l]           l = 0 .. N-2
//     logits  = head( sum_{l<N-1} skip_l[tau] + Ws gate(Wz0 e_{N-1} + Wz1 x_{N-1}[tau] (+ V ctx[tau])) + bs )
// which needs the TRUE layer inputs x_l at age sum_{k>l} d_k: rings of depth max(d_l, age_l) instead of d_l.
static void decode_edge_geometry(const Geo& g, int mode, int* edge, int* hist, int* age) {
    *edge = (mode == MVN_DECODE_REFERENCE && g.St == 1) ? 1 : 0;
    long long above = 0;
    for (int l = g.N - 1; l >= 0; --l) {
        age[l] = (int)above;
        hist[l] = (*edge && above > g.dil[l]) ? (int)above : g.dil[l];
        above += g.dil[l];
    }
}

// out[cb][n] = bias[n] + sum_k W[k][n] * in[cb][k]   for every (cb, n) pair, spread over the CTA
template <int CB>
__device__ __forceinline__ void matvec(const float* __restrict__ W, int ldw, int K, int Nout, const float* in, int ldin,
                                       const float* bias, float* out, int ldout, bool lrelu_in, bool accumulate) {
    for (int idx = threadIdx.x; idx < CB * Nout; idx += blockDim.x) {
        const int cb = idx / Nout, n = idx - cb * Nout;
        float acc = bias ? bias[n] : 0.f;
        const float* x = in + cb * ldin;
        const float* w = W + n;
#pragma unroll 8
        for (int k = 0; k < K; ++k) {
            float xv = x[k];
            if (lrelu_in) xv = mvn_lrelu(xv);
            acc = fmaf(w[(size_t)k * ldw], xv, acc);
        }
        if (accumulate) out[cb * ldout + n] += acc; else out[cb * ldout + n] = acc;
    }
}

template <int CB>
__global__ void __launch_bounds__(256) decode_kernel(const DecodeArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int C = a.C, S = a.S, A = a.A, N = a.N, Kz = a.Kz;
    const int clip0 = blockIdx.x * CB;
    // carve shared memory
    float* h = sm;                       // [CB][C]   current layer input x_l[t]
    float* olds = h + CB * C;            // [N][CB][C] queue pops for this step
    float* zin = olds + (size_t)N * CB * C;   // [CB][Kz] : old | h | ctx
    float* zb = zin + CB * Kz;           // [CB][2C]
    float* gated = zb + CB * 2 * C;      // [CB][C]
    float* rs = gated + CB * C;          // [CB][C+S]
    float* skip = rs + CB * (C + S);     // [CB][S]
    float* a1 = skip + CB * S;           // [CB][A]
    float* zl = a1 + CB * A;             // [CB][A]
    float* ev = zl + CB * A;             // [CB][C]    edge chain e_l (reference-window mode)
    float* xes = ev + (a.edge ? CB * C : 0);            // [N][CB][C] edge taps x_l[tau - age_l]
    float* wsm = xes + (a.edge ? (size_t)N * CB * C : 0);   // staged weights (optional)
    __shared__ int code_prev[CB], code_cur[CB];

    const size_t lsz = (size_t)Kz * 2 * C + 2 * C + (size_t)C * (C + S) + (C + S);   // staged floats per layer
    if (a.smem_weights) {
        for (int l = 0; l < N; ++l) {
            const float* lw = a.packed + a.P.layer0 + (size_t)l * a.P.layer_stride;
            float* dst = wsm + l * lsz;
            for (int i = threadIdx.x; i < Kz * 2 * C; i += blockDim.x) dst[i] = lw[a.P.oWz + i];
            dst += Kz * 2 * C;
            for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) dst[i] = lw[a.P.obz + i];
            dst += 2 * C;
            for (int i = threadIdx.x; i < C * (C + S); i += blockDim.x) dst[i] = lw[a.P.oWrs + i];
            dst += C * (C + S);
            for (int i = threadIdx.x; i < C + S; i += blockDim.x) dst[i] = lw[a.P.obrs + i];
        }
        float* hd = wsm + N * lsz;
        for (int i = threadIdx.x; i < S * A; i += blockDim.x) hd[i] = a.packed[a.P.w1p + i];
        hd += S * A;
        for (int i = threadIdx.x; i < A; i += blockDim.x) hd[i] = a.packed[a.P.b1 + i];
        hd += A;
        for (int i = threadIdx.x; i < A * A; i += blockDim.x) hd[i] = a.packed[a.P.w2p + i];
        hd += A * A;
        for (int i = threadIdx.x; i < A; i += blockDim.x) hd[i] = a.packed[a.P.b2 + i];
    }
    if (threadIdx.x < CB) {
        const int b = clip0 + threadIdx.x;
        code_prev[threadIdx.x] = b < a.B ? a.last2[2 * b] : -1;
        code_cur[threadIdx.x] = b < a.B ? a.last2[2 * b + 1] : -1;
    }
    __syncthreads();

    const float* win = a.packed + a.P.win;
    for (int i = a.t_start; i < a.t_start + a.n_new; ++i) {
        const int tau = i - 1;           // the model consumes x[..tau] and predicts sample i
        // causal conv (movenet/modules.py:15-30) for one-hot inputs: two weight rows
        for (int idx = threadIdx.x; idx < CB * C; idx += blockDim.x) {
            const int cb = idx / C, c = idx - cb * C;
            const int c0 = code_prev[cb], c1 = code_cur[cb];
            float v = 0.f;
            if (c0 >= 0) v += win[(size_t)c0 * C + c];
            if (c1 >= 0) v += win[((size_t)A + c1) * C + c];
            h[idx] = v;
        }
        for (int idx = threadIdx.x; idx < CB * S; idx += blockDim.x) skip[idx] = 0.f;
        // all queue pops of this step: their addresses depend on tau only
        for (int idx = threadIdx.x; idx < N * CB * C; idx += blockDim.x) {
            const int l = idx / (CB * C), r = idx - l * CB * C, cb = r / C, c = r - cb * C;
            const int d = a.dil[l], H = a.hist[l], b = clip0 + cb;
            const float* ring = a.queues + a.qoff[l] * a.B;
            float v = 0.f;
            if (b < a.B && tau - d >= 0) v = ring[((size_t)((tau - d) % H) * a.B + b) * C + c];
            olds[idx] = v;
            if (a.edge && l < N - 1) xes[idx] = b < a.B ? ring[((size_t)((tau - a.age[l]) % H) * a.B + b) * C + c] : 0.f;
        }
        if (a.edge)      // e_0 = W[:,:,1] x[i - RF]: the window's first column without its zero-padded left neighbour
            for (int idx = threadIdx.x; idx < CB * C; idx += blockDim.x) {
                const int cb = idx / C, c = idx - cb * C, b = clip0 + cb;
                ev[idx] = b < a.B ? win[((size_t)A + a.code_ring[(size_t)(i % a.RF) * a.B + b]) * C + c] : 0.f;
            }
        __syncthreads();
        for (int l = 0; l < N; ++l) {
            const float* lw = a.packed + a.P.layer0 + (size_t)l * a.P.layer_stride;
            const float *Wz, *bz, *Wrs, *brs;
            if (a.smem_weights) {
                Wz = wsm + l * lsz; bz = Wz + Kz * 2 * C; Wrs = bz + 2 * C; brs = Wrs + C * (C + S);
            } else { Wz = lw + a.P.oWz; bz = lw + a.P.obz; Wrs = lw + a.P.oWrs; brs = lw + a.P.obrs; }
            const int H = a.hist[l];
            const bool top = l == N - 1;
            if (a.edge && top)     // the last layer's edge tap is its current input
                for (int idx = threadIdx.x; idx < CB * C; idx += blockDim.x) xes[(size_t)l * CB * C + idx] = h[idx];
            // pass 0: the ordinary update at time tau (in reference-window mode the last layer's own outputs are not used);
            // pass 1: the edge column of this layer (reference-window mode only)
            for (int pass = 0; pass < (a.edge ? 2 : 1); ++pass) {
                const bool edge_pass = pass == 1;
                const int tq = edge_pass ? tau - a.age[l] : tau;           // absolute time of the column being evaluated
                // gather [older tap | x_l[tq] | ctx[tq]]; pass 0 pushes x_l[tau] into the ring
                for (int idx = threadIdx.x; idx < CB * C; idx += blockDim.x) {
                    const int cb = idx / C, c = idx - cb * C, b = clip0 + cb;
                    const float hv = edge_pass ? xes[(size_t)l * CB * C + idx] : h[idx];
                    zin[cb * Kz + c] = edge_pass ? ev[idx] : olds[(l * CB + cb) * C + c];
                    zin[cb * Kz + C + c] = hv;
                    if (a.video) zin[cb * Kz + 2 * C + c] = (b < a.B && tq >= 0 && tq < a.Tctx)
                        ? mvn_ld(a.ctx, a.ctx_dtype, ((size_t)b * a.Tctx + tq) * C + c) : 0.f;
                    if (!edge_pass && b < a.B && tau >= 0) a.queues[a.qoff[l] * a.B + ((size_t)(tau % H) * a.B + b) * C + c] = hv;
                }
                if (!edge_pass && a.edge && top) continue;                  // (block-uniform)
                __syncthreads();
                matvec<CB>(Wz, 2 * C, Kz, 2 * C, zin, Kz, bz, zb, 2 * C, false, false);
                __syncthreads();
                for (int idx = threadIdx.x; idx < CB * C; idx += blockDim.x) {
                    const int cb = idx / C, c = idx - cb * C;
                    gated[idx] = tanhf(zb[cb * 2 * C + 2 * c]) * mvn_sigmoid(zb[cb * 2 * C + 2 * c + 1]);
                }
                __syncthreads();
                matvec<CB>(Wrs, C + S, C, C + S, gated, C, brs, rs, C + S, false, false);
                __syncthreads();
                for (int idx = threadIdx.x; idx < CB * (C + S); idx += blockDim.x) {
                    const int cb = idx / (C + S), n = idx - cb * (C + S);
                    if (!edge_pass) { if (n < C) h[cb * C + n] += rs[idx]; else skip[cb * S + n - C] += rs[idx]; }
                    else if (n < C) ev[cb * C + n] = rs[idx] + xes[((size_t)l * CB + cb) * C + n];
                    else if (top) skip[cb * S + n - C] += rs[idx];
                }
                __syncthreads();
            }
        }
        // dense head (movenet/modules.py:133-142)
        const float *W1, *b1, *W2, *b2;
        if (a.smem_weights) { W1 = wsm + N * lsz; b1 = W1 + S * A; W2 = b1 + A; b2 = W2 + A * A; }
        else { W1 = a.packed + a.P.w1p; b1 = a.packed + a.P.b1; W2 = a.packed + a.P.w2p; b2 = a.packed + a.P.b2; }
        matvec<CB>(W1, A, S, A, skip, S, b1, a1, A, true, false);
        __syncthreads();
        matvec<CB>(W2, A, A, A, a1, A, b2, zl, A, true, false);
        __syncthreads();
        // next sample: argmax over channels, ties to the lowest index (torch.argmax); or, for
        // temperature > 0, a draw from softmax(softmax(z) / temperature) (movenet/wavenet.py:227-233)
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int cb = warp; cb < CB; cb += (blockDim.x >> 5)) {
            const int b = clip0 + cb;
            float* zr = zl + cb * A;
            float best = -INFINITY; int arg = 0x7fffffff;
            for (int n = lane; n < A; n += 32) {
                const float v = zr[n];
                if (v > best) { best = v; arg = n; }
                if (a.out_logits && b < a.B) a.out_logits[((size_t)b * a.n_new + (i - a.t_start)) * A + n] = v;
            }
            for (int o = 16; o; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
            }
            if (a.temperature > 0.f) {
                float sum = 0.f;                                   // p = softmax(z)
                for (int n = lane; n < A; n += 32) sum += expf(zr[n] - best);
                for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                const float pmax = 1.f / sum / a.temperature;      // largest p / temperature
                float qsum = 0.f;                                  // q ~ exp(p / temperature - max)
                __syncwarp();
                for (int n = lane; n < A; n += 32) {
                    const float q = expf(expf(zr[n] - best) / sum / a.temperature - pmax);
                    a1[cb * A + n] = q; qsum += q;
                }
                for (int o = 16; o; o >>= 1) qsum += __shfl_xor_sync(0xffffffffu, qsum, o);
                __syncwarp();
                // counter-based uniform in [0,1): one draw per (seed, clip, position)
                unsigned long long x = ((unsigned long long)a.seed << 32) ^ ((unsigned long long)(unsigned)b * 0x9E3779B97F4A7C15ULL) ^ (unsigned long long)(unsigned)i;
                x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 27; x *= 0x94D049BB133111EBULL; x ^= x >> 31;
                const float target = (float)(x >> 40) * (1.f / 16777216.f) * qsum;
                const int chunk = (A + 31) / 32, lo = lane * chunk, hi = min(lo + chunk, A);
                float local = 0.f;
                for (int n = lo; n < hi; ++n) local += a1[cb * A + n];
                float incl = local;
                for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                const float excl = incl - local;
                int pick = -1;
                if (target >= excl && target < incl) {
                    float run = excl; pick = hi - 1;
                    for (int n = lo; n < hi; ++n) { run += a1[cb * A + n]; if (target < run) { pick = n; break; } }
                }
                int chosen = A - 1;                                // rounding fell off the end: last channel
                for (int src = 31; src >= 0; --src) { const int pk = __shfl_sync(0xffffffffu, pick, src); if (pk >= 0) chosen = pk; }
                arg = chosen;
            }
            if (lane == 0) {
                code_prev[cb] = code_cur[cb]; code_cur[cb] = arg;
                if (b < a.B) a.out_codes[(size_t)b * a.n_new + (i - a.t_start)] = arg;
                if (a.edge && b < a.B) a.code_ring[(size_t)(i % a.RF) * a.B + b] = arg;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x < CB) {
        const int b = clip0 + threadIdx.x;
        if (b < a.B) { a.last2[2 * b] = code_prev[threadIdx.x]; a.last2[2 * b + 1] = code_cur[threadIdx.x]; }
    }
}

// ------------------------------------------------------------------------------------------------
// Warp-per-clip variant for narrow models (2C <= 64, e.g. the receptive-field configuration): one warp owns one clip,
// a lane owns one (or two) output channels of every matrix-vector product, the operand vector sits in a few hundred
// bytes of per-warp shared memory (broadcast reads), all weights are staged in shared memory once per CTA.  No
// block-wide barrier inside the sample loop -- only __syncwarp -- so a single clip advances ~2.4x faster than with
// the block-per-clip kernel and an SM interleaves 16 independent clips.  Same fp32 arithmetic as decode_kernel.
#define DW_WARPS 16
template <int C>
__global__ void __launch_bounds__(32 * DW_WARPS, 1) decode_warp_kernel(const DecodeArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int S = a.S, A = a.A, N = a.N, Kz = 2 * C;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * DW_WARPS + warp;
    const size_t lsz = (size_t)Kz * 2 * C + 2 * C + (size_t)C * (C + S) + (C + S);
    // staged weights: per layer Wz | bz | Wrs | brs ; head W1 | b1 | W2 | b2 ; input rows Win[2][A][C]
    float* wsm = sm;
    float* hd = wsm + N * lsz;
    float* win = hd + (size_t)S * A + A + (size_t)A * A + A;
    float* scratch = win + (size_t)2 * A * C;
    const int per_warp = 2 * C + C + 32 + A + 2 * C;    // in | gated | skip (padded) | a1 | edge chain [e_l | x_l[p_l]]
    float* in_s = scratch + (size_t)warp * per_warp;
    float* gated_s = in_s + 2 * C;
    float* skip_s = gated_s + C;
    float* a1_s = skip_s + 32;
    float* ein_s = a1_s + A;
    for (int l = 0; l < N; ++l) {
        const float* lw = a.packed + a.P.layer0 + (size_t)l * a.P.layer_stride;
        float* dst = wsm + l * lsz;
        for (int i = threadIdx.x; i < Kz * 2 * C; i += blockDim.x) dst[i] = lw[a.P.oWz + i];
        dst += Kz * 2 * C;
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) dst[i] = lw[a.P.obz + i];
        dst += 2 * C;
        for (int i = threadIdx.x; i < C * (C + S); i += blockDim.x) dst[i] = lw[a.P.oWrs + i];
        dst += C * (C + S);
        for (int i = threadIdx.x; i < C + S; i += blockDim.x) dst[i] = lw[a.P.obrs + i];
    }
    {
        float* d = hd;
        for (int i = threadIdx.x; i < S * A; i += blockDim.x) d[i] = a.packed[a.P.w1p + i];
        d += S * A;
        for (int i = threadIdx.x; i < A; i += blockDim.x) d[i] = a.packed[a.P.b1 + i];
        d += A;
        for (int i = threadIdx.x; i < A * A; i += blockDim.x) d[i] = a.packed[a.P.w2p + i];
        d += A * A;
        for (int i = threadIdx.x; i < A; i += blockDim.x) d[i] = a.packed[a.P.b2 + i];
        for (int i = threadIdx.x; i < 2 * A * C; i += blockDim.x) win[i] = a.packed[a.P.win + i];
    }
    __syncthreads();
    if (b >= a.B) return;
    int code_prev = a.last2[2 * b], code_cur = a.last2[2 * b + 1];
    const float *W1 = hd, *b1 = W1 + S * A, *W2 = b1 + A, *b2 = W2 + A * A;
    constexpr int NZ = 2 * C / 32;                      // gate pre-activations per lane

    for (int i = a.t_start; i < a.t_start + a.n_new; ++i) {
        const int tau = i - 1;
        for (int c = lane; c < C; c += 32) {
            float v = 0.f;
            if (code_prev >= 0) v += win[(size_t)code_prev * C + c];
            if (code_cur >= 0) v += win[((size_t)A + code_cur) * C + c];
            in_s[C + c] = v;
        }
        skip_s[lane] = 0.f;
        if (a.edge) {     // e_0 = W[:,:,1] x[i - RF] (see decode_edge_geometry)
            const int ce = a.code_ring[(size_t)(i % a.RF) * a.B + b];
            for (int c = lane; c < C; c += 32) ein_s[c] = win[((size_t)A + ce) * C + c];
        }
        for (int l = 0; l < N; ++l) {
            const float* Wz = wsm + l * lsz; const float* bz = Wz + Kz * 2 * C; const float* Wrs = bz + 2 * C; const float* brs = Wrs + C * (C + S);
            const int d = a.dil[l], H = a.hist[l];
            const bool top = l == N - 1;
            __syncwarp();
            for (int c = lane; c < C; c += 32) {         // queue pop / push (reads before the write: the slots may coincide)
                float* ring = a.queues + a.qoff[l] * a.B;
                in_s[c] = tau - d >= 0 ? ring[((size_t)((tau - d) % H) * a.B + b) * C + c] : 0.f;
                if (a.edge) ein_s[C + c] = top ? in_s[C + c] : ring[((size_t)((tau - a.age[l]) % H) * a.B + b) * C + c];
                ring[((size_t)(tau % H) * a.B + b) * C + c] = in_s[C + c];
            }
            // pass 0: the ordinary update at time tau (not needed for the last layer in reference-window mode);
            // pass 1: this layer's edge column
            for (int pass = (a.edge && top) ? 1 : 0; pass < (a.edge ? 2 : 1); ++pass) {
                const float* zin = pass ? ein_s : in_s;
                __syncwarp();
#pragma unroll
                for (int o = 0; o < NZ; ++o) {
                    const int n = 32 * o + lane;
                    float acc0 = bz[n], acc1 = 0.f;
#pragma unroll 8
                    for (int k = 0; k < Kz; k += 4) {
                        const float4 x = *(const float4*)(zin + k);
                        acc0 = fmaf(Wz[(k + 0) * 2 * C + n], x.x, acc0); acc1 = fmaf(Wz[(k + 1) * 2 * C + n], x.y, acc1);
                        acc0 = fmaf(Wz[(k + 2) * 2 * C + n], x.z, acc0); acc1 = fmaf(Wz[(k + 3) * 2 * C + n], x.w, acc1);
                    }
                    const float z = acc0 + acc1;
                    const float partner = __shfl_xor_sync(0xffffffffu, z, 1);    // columns interleave (filter c, gate c)
                    if (!(lane & 1)) gated_s[n >> 1] = tanhf(z) * mvn_sigmoid(partner);
                }
                __syncwarp();
                for (int n = lane; n < C + S; n += 32) {
                    float acc0 = brs[n], acc1 = 0.f;
#pragma unroll 4
                    for (int k = 0; k < C; k += 4) {
                        const float4 x = *(const float4*)(gated_s + k);
                        acc0 = fmaf(Wrs[(k + 0) * (C + S) + n], x.x, acc0); acc1 = fmaf(Wrs[(k + 1) * (C + S) + n], x.y, acc1);
                        acc0 = fmaf(Wrs[(k + 2) * (C + S) + n], x.z, acc0); acc1 = fmaf(Wrs[(k + 3) * (C + S) + n], x.w, acc1);
                    }
                    const float v = acc0 + acc1;
                    if (!pass) { if (n < C) in_s[C + n] += v; else skip_s[n - C] += v; }
                    else if (n < C) ein_s[n] = v + ein_s[C + n];
                    else if (top) skip_s[n - C] += v;
                }
            }
        }
        __syncwarp();
        // dense head
        for (int n = lane; n < A; n += 32) {
            float acc = b1[n];
            for (int s2 = 0; s2 < S; ++s2) acc = fmaf(W1[s2 * A + n], mvn_lrelu(skip_s[s2]), acc);
            a1_s[n] = mvn_lrelu(acc);
        }
        __syncwarp();
        float best = -INFINITY; int arg = 0x7fffffff;
        float zmine[8];                                    // A <= 256: up to 8 logits per lane
        {
            int o = 0;
            for (int n = lane; n < A; n += 32, ++o) {
                float acc0 = b2[n], acc1 = 0.f;
#pragma unroll 8
                for (int k = 0; k < A; k += 4) {
                    const float4 x = *(const float4*)(a1_s + k);
                    acc0 = fmaf(W2[(k + 0) * A + n], x.x, acc0); acc1 = fmaf(W2[(k + 1) * A + n], x.y, acc1);
                    acc0 = fmaf(W2[(k + 2) * A + n], x.z, acc0); acc1 = fmaf(W2[(k + 3) * A + n], x.w, acc1);
                }
                const float z = acc0 + acc1;
                zmine[o] = z;
                if (z > best) { best = z; arg = n; }
                if (a.out_logits) a.out_logits[((size_t)b * a.n_new + (i - a.t_start)) * A + n] = z;
            }
        }
        for (int o = 16; o; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
        }
        if (a.temperature > 0.f) {      // draw from softmax(softmax(z) / temperature) (movenet/wavenet.py:227-231)
            float sum = 0.f;
            { int o = 0; for (int n = lane; n < A; n += 32, ++o) sum += expf(zmine[o] - best); }
            for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float pmax = 1.f / sum / a.temperature;
            float qsum = 0.f;
            __syncwarp();
            { int o = 0; for (int n = lane; n < A; n += 32, ++o) { const float q = expf(expf(zmine[o] - best) / sum / a.temperature - pmax); a1_s[n] = q; qsum += q; } }
            for (int o = 16; o; o >>= 1) qsum += __shfl_xor_sync(0xffffffffu, qsum, o);
            __syncwarp();
            unsigned long long x = ((unsigned long long)a.seed << 32) ^ ((unsigned long long)(unsigned)b * 0x9E3779B97F4A7C15ULL) ^ (unsigned long long)(unsigned)i;
            x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL; x ^= x >> 27; x *= 0x94D049BB133111EBULL; x ^= x >> 31;
            const float target = (float)(x >> 40) * (1.f / 16777216.f) * qsum;
            const int chunk = (A + 31) / 32, lo = lane * chunk, hi = min(lo + chunk, A);
            float local = 0.f;
            for (int n = lo; n < hi; ++n) local += a1_s[n];
            float incl = local;
            for (int o = 1; o < 32; o <<= 1) { const float t2 = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t2; }
            const float excl = incl - local;
            int pick = -1;
            if (target >= excl && target < incl) {
                float run = excl; pick = hi - 1;
                for (int n = lo; n < hi; ++n) { run += a1_s[n]; if (target < run) { pick = n; break; } }
            }
            int chosen = A - 1;
            for (int src = 31; src >= 0; --src) { const int pk = __shfl_sync(0xffffffffu, pick, src); if (pk >= 0) chosen = pk; }
            arg = chosen;
        }
        code_prev = code_cur; code_cur = arg;
        if (lane == 0) {
            a.out_codes[(size_t)b * a.n_new + (i - a.t_start)] = arg;
            if (a.edge) a.code_ring[(size_t)(i % a.RF) * a.B + b] = arg;
        }
    }
    if (lane == 0) { a.last2[2 * b] = code_prev; a.last2[2 * b + 1] = code_cur; }
}

template <int C>
static int launch_decode_warp(DecodeArgs& a, const Geo& g, cudaStream_t st, bool* used) {
    const size_t lsz = (size_t)2 * C * 2 * C + 2 * C + (size_t)C * (C + g.S) + (C + g.S);
    const size_t w_floats = g.N * lsz + (size_t)g.S * g.A + g.A + (size_t)g.A * g.A + g.A + (size_t)2 * g.A * C;
    const size_t per_warp = 2 * C + C + 32 + g.A + 2 * C;
    const size_t smem = (w_floats + DW_WARPS * per_warp) * 4;
    *used = false;
    if (smem > 220 * 1024 || g.S > 32 || g.A > 256 || g.A % 4) return 0;
    *used = true;
    MVN_CUDA(cudaFuncSetAttribute(decode_warp_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    decode_warp_kernel<C><<<mvn_cdiv(g.B, DW_WARPS), 32 * DW_WARPS, smem, st>>>(a);
    return mvn_check_launch("decode_warp_steps");
}

// Fill the rings from the layer inputs of a forward pass over the T-column prompt.  The first decode
// step re-evaluates time T-1 (the last prompt sample) itself, so the rings must hold the d inputs
// BEFORE it: ring_l[tau % d] = x_l[tau] for tau in [T-1-d, T-1).
// (d = the ring depth; Tp = prompt length, T = row stride of x in time steps)
__global__ void decode_prefill_kernel(const void* __restrict__ x, int adt, int B, int T, int Tp, int C, int d,
                                      float* __restrict__ ring) {
    const long long n = (long long)B * d * C;
    const int Tend = Tp - 1;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C); const long long r = i / C; const int b = (int)(r % B); const int slot = (int)(r / B);
        int tau = (Tend / d) * d + slot; if (tau >= Tend) tau -= d;     // the time in [Tend-d, Tend) living in this slot
        ring[i] = tau >= 0 ? mvn_ld(x, adt, ((size_t)b * T + tau) * C + c) : 0.f;
    }
}

__global__ void decode_last2_kernel(const int* __restrict__ codes, int B, int T, int Tp, int* __restrict__ last2) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    last2[2 * b] = Tp >= 2 ? codes[(size_t)b * T + Tp - 2] : -1;
    last2[2 * b + 1] = Tp >= 1 ? codes[(size_t)b * T + Tp - 1] : -1;
}
// code_ring[p % RF][b] = codes[b][p] for the last RF prompt positions
__global__ void decode_code_ring_kernel(const int* __restrict__ codes, int B, int T, int Tp, int RF, int* __restrict__ code_ring) {
    const long long n = (long long)B * RF;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i % B), p = Tp - RF + (int)(i / B);
        code_ring[(size_t)(p % RF) * B + b] = codes[(size_t)b * T + p];
    }
}

static size_t queue_floats(const Geo& g, const int* hist, long long* qoff) {
    long long o = 0;
    for (int l = 0; l < g.N; ++l) { if (qoff) qoff[l] = o; o += (long long)hist[l] * g.C; }
    return (size_t)o;
}

extern "C" size_t mvn_decode_state_bytes(const mvn_shape_t* s, int mode) {
    Geo g; if (geo_init(g, s)) return 0;
    int edge, hist[MVN_MAX_LAYERS], age[MVN_MAX_LAYERS];
    decode_edge_geometry(g, mode, &edge, hist, age);
    return al256(queue_floats(g, hist, nullptr) * (size_t)g.B * 4) + al256((size_t)g.B * 2 * 4) +
           (edge ? al256((size_t)g.RF * g.B * 4) : 0);
}

extern "C" int mvn_decode_prefill(const mvn_shape_t* s, const void* acts, void* state, int mode, int prompt_frames,
                                  void* stream) {
    Geo g; MVN_REQUIRE(s && geo_init(g, s) == 0, "mvn_decode_prefill: bad shape");
    MVN_REQUIRE(acts && state, "mvn_decode_prefill: null buffer");
    const int Tp = prompt_frames > 0 ? prompt_frames : g.T;
    MVN_REQUIRE(Tp <= g.T, "mvn_decode_prefill: prompt_frames (%d) exceeds the forward pass's frames (%d)", Tp, g.T);
    int edge, hist[MVN_MAX_LAYERS], age[MVN_MAX_LAYERS];
    decode_edge_geometry(g, mode, &edge, hist, age);
    MVN_REQUIRE(!edge || Tp >= g.RF, "mvn_decode_prefill: the reference-window mode needs a prompt of at least receptive_fields (%d) columns", g.RF);
    ActsLayout AL; acts_layout(g, AL);
    long long qoff[MVN_MAX_LAYERS];
    const size_t qf = queue_floats(g, hist, qoff);
    float* queues = (float*)state;
    int* last2 = (int*)((char*)state + al256(qf * (size_t)g.B * 4));
    cudaStream_t st = (cudaStream_t)stream;
    for (int l = 0; l < g.N; ++l) {
        const void* x = (const char*)acts + AL.x0 + (size_t)l * AL.x_stride;
        const long long n = (long long)g.B * hist[l] * g.C;
        decode_prefill_kernel<<<mvn_cdiv(n, 256) < 1184 ? mvn_cdiv(n, 256) : 1184, 256, 0, st>>>(
            x, g.adt, g.B, g.T, Tp, g.C, hist[l], queues + qoff[l] * g.B);
    }
    const int* codes = (const int*)((const char*)acts + AL.codes);
    decode_last2_kernel<<<mvn_cdiv(g.B, 128), 128, 0, st>>>(codes, g.B, g.T, Tp, last2);
    if (edge) {
        int* code_ring = last2 + al256((size_t)g.B * 2 * 4) / 4;
        const long long n = (long long)g.B * g.RF;
        decode_code_ring_kernel<<<mvn_cdiv(n, 256) < 1184 ? mvn_cdiv(n, 256) : 1184, 256, 0, st>>>(codes, g.B, g.T, Tp, g.RF, code_ring);
    }
    return mvn_check_launch("decode_prefill");
}

template <int CB>
static int launch_decode(DecodeArgs& a, const Geo& g, cudaStream_t st) {
    const size_t act_floats = (size_t)CB * g.C + (size_t)g.N * CB * g.C + (size_t)CB * g.Kz + (size_t)CB * 2 * g.C +
                              (size_t)CB * g.C + (size_t)CB * (g.C + g.S) + (size_t)CB * g.S + 2 * (size_t)CB * g.A +
                              (a.edge ? (size_t)CB * g.C + (size_t)g.N * CB * g.C : 0);
    const size_t lsz = (size_t)g.Kz * 2 * g.C + 2 * g.C + (size_t)g.C * (g.C + g.S) + (g.C + g.S);
    const size_t w_floats = g.N * lsz + (size_t)g.S * g.A + g.A + (size_t)g.A * g.A + g.A;
    const size_t cap = 220 * 1024;
    MVN_REQUIRE(act_floats * 4 <= cap, "decode: model too large for the per-CTA activation staging");
    a.smem_weights = (act_floats + w_floats) * 4 <= cap;
    const size_t smem = (act_floats + (a.smem_weights ? w_floats : 0)) * 4;
    MVN_CUDA(cudaFuncSetAttribute(decode_kernel<CB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap));
    decode_kernel<CB><<<mvn_cdiv(g.B, CB), 256, smem, st>>>(a);
    return mvn_check_launch("decode_steps");
}

extern "C" int mvn_decode_steps(const mvn_shape_t* s, const void* packed, void* state, const void* ctx, int mode,
                                int t_start, int n_new, int* out_codes, float* out_logits, float temperature,
                                unsigned seed, void* stream) {
    Geo g; MVN_REQUIRE(s && geo_init(g, s) == 0, "mvn_decode_steps: bad shape");
    MVN_REQUIRE(packed && state && out_codes && n_new >= 0 && t_start >= 1, "mvn_decode_steps: bad arguments");
    MVN_REQUIRE(!g.video || ctx, "mvn_decode_steps: shape says video but ctx is null");
    if (n_new == 0) return 0;
    DecodeArgs args;
    memset(&args, 0, sizeof(args));
    args.packed = (const float*)packed; packed_layout(g, args.P);
    args.N = g.N; args.A = g.A; args.C = g.C; args.S = g.S; args.Kz = g.Kz; args.video = g.video; args.B = g.B;
    args.Tctx = g.T; args.ctx_dtype = g.adt; args.t_start = t_start; args.n_new = n_new;
    args.temperature = temperature; args.seed = seed;
    decode_edge_geometry(g, mode, &args.edge, args.hist, args.age);
    args.RF = g.RF;
    MVN_REQUIRE(!args.edge || t_start >= g.RF, "mvn_decode_steps: the reference-window mode starts at t_start >= receptive_fields");
    const size_t qf = queue_floats(g, args.hist, args.qoff);
    args.queues = (float*)state; args.last2 = (int*)((char*)state + al256(qf * (size_t)g.B * 4));
    args.code_ring = args.last2 + al256((size_t)g.B * 2 * 4) / 4;
    args.ctx = ctx; args.out_codes = out_codes; args.out_logits = out_logits;
    for (int l = 0; l < g.N; ++l) args.dil[l] = g.dil[l];
    cudaStream_t st = (cudaStream_t)stream;
    if (!g.video && (g.C == 16 || g.C == 32)) {           // narrow model: one warp per clip, no block-wide barriers
        bool used = false;
        int rc = g.C == 16 ? launch_decode_warp<16>(args, g, st, &used) : launch_decode_warp<32>(args, g, st, &used);
        if (used) return rc;
    }
    // clips per CTA: keep every SM busy first, then amortise weight reads over more clips
    const size_t per_clip = (size_t)g.N * g.C * 4 * (args.edge ? 2 : 1);
    if (g.B >= 148 * 8 && per_clip * 8 <= 64 * 1024) return launch_decode<8>(args, g, st);
    if (g.B >= 148 * 4 && per_clip * 4 <= 64 * 1024) return launch_decode<4>(args, g, st);
    if (g.B >= 148 * 2) return launch_decode<2>(args, g, st);
    return launch_decode<1>(args, g, st);
}
