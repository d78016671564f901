// Host-side geometry: where everything lives inside the caller-owned buffers.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include "../../include/movenet_b200.h"

#define MVN_MAX_LAYERS 256
#define MVN_DET_WS_FLOATS (8u << 20)     // 32 MB
// tensor-core weight image of one layer: 3 Wz chunks + [Wr|Ws] (<=128 rows) of 16 KB each, + 1 KB of biases
#define MVN_TC_IMG_BYTES (4 * 16384 + 1024)

struct Geo {
    int L, St, A, C, S, Cin, B, T, video, adt, remove_last, logits, no_grad;
    int Cl;       // the model's residual_channels; C is the PHYSICAL channel count of every internal buffer: in the
                  // tensor-core (bf16) mode narrower models are zero-padded to 64 channels (the pad stays exactly zero
                  // through every layer: tanh(0)*sigmoid(0) = 0), so one set of C = 64 kernels serves them all
    int N;        // layers
    int RF;       // receptive_fields (movenet/wavenet.py:125-134)
    int Tout;     // T - RF + 1      (movenet/wavenet.py:136-147)
    int Tn;       // Tout - remove_last : columns the caller receives
    int Kz;       // 2C (+C with video): contraction length of the gate GEMM
    int es;       // bytes per activation element
    int dil[MVN_MAX_LAYERS];
};

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

static inline int geo_init(Geo& g, const mvn_shape_t* s) {
    g.L = s->layer_size; g.St = s->stack_size; g.A = s->input_channels; g.Cl = s->residual_channels;
    g.C = (s->act_dtype == MVN_DTYPE_BF16 && g.Cl < 64 && s->skip_channels % 8 == 0 && s->skip_channels >= 8 &&
           s->skip_channels <= (s->has_video ? 32 : 64)) ? 64 : g.Cl;
    g.S = s->skip_channels; g.Cin = s->context_in_channels; g.B = s->batch; g.T = s->frames;
    g.video = s->has_video; g.adt = s->act_dtype; g.remove_last = s->remove_last; g.logits = s->output_logits;
    g.no_grad = s->no_grad;
    g.N = g.L * g.St;
    if (g.L < 1 || g.St < 1 || g.N > MVN_MAX_LAYERS || g.L > 24) return -1;
    long long rf = g.St;
    for (int st = 0; st < g.St; ++st)
        for (int x = 0; x < g.L; ++x) { g.dil[st * g.L + x] = 1 << x; rf += (1LL << x); }
    if (rf > 0x7fffffffLL) return -1;
    g.RF = (int)rf;
    g.Tout = g.T - g.RF + 1;
    g.Tn = g.Tout - (g.remove_last ? 1 : 0);
    g.Kz = g.video ? 3 * g.C : 2 * g.C;
    g.es = g.adt == MVN_DTYPE_BF16 ? 2 : 4;
    return 0;
}

// Wide-channel tensor-core path (wide.cu, wide_gemm.cuh): weight-streaming tcgen05 GEMMs for residual_channels >= 128
// (the widened scale-up shape, BASELINE configs[3]); audio-only, every channel count a multiple of 128, A <= 256 (one
// accumulator chunk holds a whole softmax row)
static inline int wide_ok(const Geo& g) {
    return g.adt == MVN_DTYPE_BF16 && !g.video && g.C >= 128 && g.C % 128 == 0 && g.C <= 1024 && g.S >= 128 && g.S % 128 == 0 &&
           g.S <= 1024 && g.A >= 128 && g.A % 128 == 0 && g.A <= 256;
}

// The wide engine's head kernels alone (DenseConv + softmax and their backward) also serve shapes whose residual stack runs
// on the fused C <= 64 kernels but whose head the smem-resident head kernels do not cover: the reference's own test
// architecture (A = 256, C = 64, S = 64; tests/test_model.py:42-48)
static inline int wide_head_ok(const Geo& g) {
    return g.adt == MVN_DTYPE_BF16 && g.A >= 128 && g.A % 128 == 0 && g.A <= 256 && g.S >= 64 && g.S % 64 == 0 && g.S <= 1024;
}

// ---- packed weights (fp32 elements) -------------------------------------------------------------
struct PackedLayout {
    size_t win;                 // [2][A][C]      Win[tap][a][c] = causal_conv.conv.weight[c][a][tap]
    size_t layer0, layer_stride;
    // inside one layer
    size_t oWz;                 // [Kz][2C]   rows: tap0 C | tap1 C | ctx C ; cols interleaved (filter c, gate c)
    size_t obz;                 // [2C]       context conv biases (0 without video)
    size_t oWrs;                // [C][C+S]   cols: residual C | skip S
    size_t obrs;                // [C+S]
    size_t oWzT;                // [2C][Kz]   transpose of Wz (backward data)
    size_t oWrsT;               // [C+S][C]   transpose of Wrs
    size_t oTc;                 // tensor-core shared-memory image (bf16, swizzled; see layer_tc.cu), C == 64 only
    size_t w1p, b1, w2p, b2;    // head: [S][A], [A], [A][A], [A]
    size_t w1pT, w2pT;          // [A][S], [A][A]
    size_t tc_head;             // tensor-core image of conv2.weight (bf16, A/64 chunks of [A][64], 128B swizzle), A == 64 or 128
    size_t wv, bv;              // video conv: [4096*Cin][C], [C]
    size_t wt[3], bt[3], wtT[3];// transposed convs: [C][10C], [10C] (bias tiled), [10C][C]
    size_t tc_up;               // tensor-core image of the last upsampler level (bf16 [640][64] + bias), video && C == 64
    size_t tc_up01[2];          // ... and of the first two levels
    // wide path: bf16 K-major matrices [rows][K] streamed by TMA (offsets in fp32 elements; per layer relative to the layer base)
    size_t wWz;                 // [2C][2C]   rows in chunks of 256 = (filter | gate) of 128 channels ; K = tap0 C | tap1 C
    size_t wWrs;                // [C][2C]    [Wr | I]: x' = [gated | x] . this^T + br -- the residual add rides through the tensor core (exact)
    size_t wWrsT;               // [C][C+S]   d(gated) = [d(x') | d(skip)] . this^T
    size_t wWzT;                // [C][4C]    d(x) = [dz(t) | dz(t+d)] . this^T ; dz columns interleaved (df c, dg c)
    size_t wH1, wH2, wH2T, wH1T;// head: [A][S], [A][A], [A][A] (transposed), [S][A]
    size_t wWsAll;              // [S][N C]   every layer's skip 1x1 conv side by side: skip_sum = [gated_0 | gated_1 | ...] . this^T
    size_t wbsum;               // [S] fp32   sum over layers of the skip biases
    size_t total;               // elements
};

static inline void packed_layout(const Geo& g, PackedLayout& p) {
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += (n + 63) & ~(size_t)63; return r; };
    const size_t A = g.A, C = g.C, S = g.S, Kz = g.Kz;
    p.win = take(2 * A * C);
    size_t l0 = o;
    p.oWz = take(Kz * 2 * C) - l0;
    p.obz = take(2 * C) - l0;
    p.oWrs = take(C * (C + S)) - l0;
    p.obrs = take(C + S) - l0;
    p.oWzT = take(2 * C * Kz) - l0;
    p.oWrsT = take((C + S) * C) - l0;
    p.oTc = take(g.C == 64 ? MVN_TC_IMG_BYTES / 4 : 0) - l0;
    const bool wide = wide_ok(g);
    p.wWz = take(wide ? 2 * C * 2 * C / 2 : 0) - l0;
    p.wWrs = take(wide ? C * 2 * C / 2 : 0) - l0;
    p.wWrsT = take(wide ? C * (C + S) / 2 : 0) - l0;
    p.wWzT = take(wide ? C * 4 * C / 2 : 0) - l0;
    p.layer0 = l0;
    p.layer_stride = o - l0;
    o = l0 + p.layer_stride * g.N;
    p.w1p = take(S * A); p.b1 = take(A); p.w2p = take(A * A); p.b2 = take(A);
    p.w1pT = take(A * S); p.w2pT = take(A * A);
    p.tc_head = take((A == 64 || A == 128) ? A * A / 2 : 0);
    const bool whead = wide || wide_head_ok(g);
    p.wH1 = take(whead ? A * S / 2 : 0); p.wH2 = take(whead ? A * A / 2 : 0);
    p.wH2T = take(whead ? A * A / 2 : 0); p.wH1T = take(whead ? S * A / 2 : 0);
    p.wWsAll = take(wide ? S * (size_t)g.N * C / 2 : 0); p.wbsum = take(wide ? S : 0);
    if (g.video) {
        p.wv = take((size_t)4096 * g.Cin * C); p.bv = take(C);
        for (int i = 0; i < 3; ++i) { p.wt[i] = take(C * 10 * C); p.bt[i] = take(10 * C); p.wtT[i] = take(10 * C * C); }
        p.tc_up = take(C == 64 ? (5 * 16384 + 3072) / 4 : 0);
        for (int i = 0; i < 2; ++i) p.tc_up01[i] = take(C == 64 ? (5 * 16384 + 3072) / 4 : 0);
    } else {
        p.tc_up = 0; p.tc_up01[0] = p.tc_up01[1] = 0;
        p.wv = p.bv = 0;
        for (int i = 0; i < 3; ++i) p.wt[i] = p.bt[i] = p.wtT[i] = 0;
    }
    p.total = o;
}

// ---- activations kept from forward to backward (byte offsets) -----------------------------------
struct ActsLayout {
    size_t codes;     // int32 [B*T]   argmax over channels of every audio column
    size_t dense;     // uint8 [B*T]   1 where the column is not an exact one-hot
    size_t x0;        // layer inputs x_0..x_{N-1}, each (B,T,C) act dtype ; x_0 = causal conv output
    size_t x_stride;
    size_t ctx;       // (B,T,C) act dtype (video only)
    size_t skip;      // (B,Tout,S) fp32 ; wide path: (B,T,S), row t = time t (TMA stores cannot start at a negative row)
    size_t a1;        // (B,Tn,A) fp32 : dense_conv.conv1 output (pre-activation)
    size_t enc, u1, u2; // video: (B,160,C) (B,1600,C) (B,16000,C) fp32
    size_t w_gated;   // wide path: (B,T,N C) bf16, the gated activations of every layer side by side (layer l: columns l C ..)
    size_t w_gab;     // wide path, training: (B,T,N 2C) bf16, the gate's derivative factors (a_c, b_c) interleaved, per layer:
                      //   a = sigma(g) (1 - tanh(f)^2), b = tanh(f) sigma(g) (1 - sigma(g))  ->  dz = d(gated) * (a, b)
    size_t total;
};

static inline void acts_layout(const Geo& g, ActsLayout& a) {
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += al256(n); return r; };
    const size_t BT = (size_t)g.B * g.T;
    a.codes = take(BT * 4);
    a.dense = take(BT);
    a.x_stride = al256(BT * g.C * g.es);
    a.x0 = take(a.x_stride * g.N);
    a.ctx = g.video ? take(BT * g.C * g.es) : 0;
    a.skip = take((size_t)g.B * (wide_ok(g) ? g.T : (g.Tout > 0 ? g.Tout : 0)) * g.S * 4);
    a.a1 = take((size_t)g.B * (g.Tout > 0 ? g.Tout : 0) * g.A * 4);
    if (g.video) {
        a.enc = take((size_t)g.B * 160 * g.C * 4);
        a.u1 = take((size_t)g.B * 1600 * g.C * 4);
        a.u2 = take((size_t)g.B * 16000 * g.C * 4);
    } else a.enc = a.u1 = a.u2 = 0;
    a.w_gated = take(wide_ok(g) ? BT * g.N * g.C * 2 : 0);
    a.w_gab = take(wide_ok(g) && !g.no_grad ? BT * g.N * 2 * g.C * 2 : 0);
    a.total = o;
}

// ---- reusable workspace (byte offsets) ----------------------------------------------------------
#define MVN_TC_PARTIAL_SLOT_BYTES ((size_t)2 * 148 * (128 * 256 + 256) * 4)
#define MVN_TC_PARTIAL_SLOTS 6
struct ScratchLayout {
    size_t gated;     // (B,T,C) act dtype
    size_t z;         // (B,Tn,A) fp32 : head logits, time-major ; backward: d(logits)
    size_t da1;       // (B,Tn,A) fp32
    size_t dskip;     // (B,Tout,S) fp32
    size_t dgated;    // (B,T,C) act dtype
    size_t dz;        // (B,T,2C) act dtype
    size_t dxa, dxb;  // (B,T,C) act dtype, ping-pong
    size_t dctx;      // (B,T,C) fp32
    size_t du2, du1, denc;
    size_t tc_partial; // per-CTA partial weight gradients of the head / input / upsampler / video tensor-core kernels: MVN_TC_PARTIAL_SLOTS
                       // slots of MVN_TC_PARTIAL_SLOT_BYTES (one per producer: their reductions are deferred to side streams)
    size_t tc_layer_partial; // ... and of the layer backward kernel, one slot per layer (reduced together at the end)
    size_t det_ws;    // fp32 partial products of the exact-mode split reductions (added in a fixed order: no atomics)
    // wide path (wide.cu)
    size_t w_l0;      // (B,Tout,S) bf16 : lrelu(skip_sum), the head's first A operand
    size_t w_ds16;    // (B,T,S) bf16    : d(skip) on the T row space (zero outside the last Tn rows of a clip)
    size_t w_oh16;    // (B,T,A) bf16    : the audio (one-hot) as a GEMM operand of the input conv's weight gradient
    size_t w_colsum;  // fp32 partial column sums (bias gradients)
    size_t w_wgpart;  // fp32 partial weight-gradient blocks, one [256][512] per CTA pair (wide_wgrad.cuh)
    size_t total;
};

static inline void scratch_layout(const Geo& g, ScratchLayout& w) {
    size_t o = 0;
    auto take = [&](size_t n) { size_t r = o; o += al256(n); return r; };
    const size_t BT = (size_t)g.B * g.T;
    const size_t BTo = (size_t)g.B * (g.Tout > 0 ? g.Tout : 0);
    w.gated = take(BT * g.C * g.es);
    w.z = take(BTo * g.A * 4);
    w.da1 = take(BTo * g.A * 4);
    w.dskip = take(BTo * g.S * 4);
    w.dgated = take(BT * g.C * g.es);
    w.dz = take(BT * 2 * g.C * g.es);
    w.dxa = take(BT * g.C * g.es);
    w.dxb = take(BT * g.C * g.es);
    if (g.video) {
        w.dctx = take(BT * g.C * 4);
        w.du2 = take((size_t)g.B * 16000 * g.C * 4);
        w.du1 = take((size_t)g.B * 1600 * g.C * 4);
        w.denc = take((size_t)g.B * 160 * g.C * 4);
    } else w.dctx = w.du2 = w.du1 = w.denc = 0;
    // head | input | upsampler levels 0..2 | video encoder ; then one slot per layer
    w.tc_partial = take(g.adt == MVN_DTYPE_BF16 && (g.C == 64 || g.A == 64 || g.A == 128) ? (size_t)MVN_TC_PARTIAL_SLOTS * MVN_TC_PARTIAL_SLOT_BYTES : 0);
    w.tc_layer_partial = take(g.adt == MVN_DTYPE_BF16 && g.C == 64 ? (size_t)g.N * 148 * (128 * 256 + 256) * 4 : 0);
    w.det_ws = take((size_t)MVN_DET_WS_FLOATS * 4);
    const bool wide = wide_ok(g);
    w.w_l0 = take(wide || wide_head_ok(g) ? BTo * g.S * 2 : 0);
    w.w_ds16 = take(wide ? BT * g.S * 2 : 0);
    w.w_oh16 = take(wide ? BT * g.A * 2 : 0);
    w.w_colsum = take(wide || wide_head_ok(g) ? (size_t)1024 * 1024 * 4 : 0);
    w.w_wgpart = take(wide ? (size_t)80 * 256 * 512 * 4 : 0);
    w.total = o;
}
