// Re-layout of the reference's parameters (state_dict tensors) into the packed
// form the kernels consume, and the reverse map for gradients.
#include "common.cuh"
#include "layout.h"
#include "layer_tc.h"

// one block row per layer: blockIdx.y = layer.  ptrs = device table in state_dict order.
// C: physical channels of the packed layout, Cl <= C: channels of the reference tensors (the rest is zero padding)
__global__ void pack_layer_kernel(const float* const* __restrict__ ptrs, float* __restrict__ packed,
                                  PackedLayout P, int C, int Cl, int S, int Kz, int video, int biases_only) {
    MVN_PDL_PROLOGUE();
    const int l = blockIdx.y;
    const float* const* lp = ptrs + MVN_PARAM_LAYER(l, 0);
    const float *wf = lp[0], *wg = lp[1], *vf = lp[2], *bvf = lp[3], *vg = lp[4], *bvg = lp[5], *wr = lp[6],
                *br = lp[7], *ws = lp[8], *bs = lp[9];
    float* base = packed + P.layer0 + (size_t)l * P.layer_stride;
    const int nWz = Kz * 2 * C, nbz = 2 * C, nWrs = C * (C + S), nbrs = C + S;
    const int total = nWz + nbz + nWrs + nbrs;
    // (the wide tensor-core path reads its own bf16 matrices, wide.cu: only the bias vectors of this layout are used there)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int j = i;
        if (biases_only && (j < nWz || (j >= nWz + nbz && j < nWz + nbz + nWrs))) continue;
        if (j < nWz) {
            const int k = j / (2 * C), n = j % (2 * C), c = n >> 1, gate = n & 1;
            float v = 0.f;
            const int kk = k % C;            // input channel inside its tap / context block
            if (c < Cl && kk < Cl) {
                if (k < C) v = (gate ? wg : wf)[((size_t)c * Cl + kk) * 2 + 0];
                else if (k < 2 * C) v = (gate ? wg : wf)[((size_t)c * Cl + kk) * 2 + 1];
                else v = (gate ? vg : vf)[(size_t)c * Cl + kk];
            }
            base[P.oWz + j] = v;
            base[P.oWzT + (size_t)n * Kz + k] = v;
            continue;
        }
        j -= nWz;
        if (j < nbz) { base[P.obz + j] = (video && (j >> 1) < Cl) ? ((j & 1) ? bvg : bvf)[j >> 1] : 0.f; continue; }
        j -= nbz;
        if (j < nWrs) {
            const int k = j / (C + S), n = j % (C + S);
            float v = 0.f;
            if (k < Cl) { if (n < C) { if (n < Cl) v = wr[(size_t)n * Cl + k]; } else v = ws[(size_t)(n - C) * Cl + k]; }
            base[P.oWrs + j] = v;
            base[P.oWrsT + (size_t)n * C + k] = v;
            continue;
        }
        j -= nWrs;
        base[P.obrs + j] = j < C ? (j < Cl ? br[j] : 0.f) : bs[j - C];
    }
}

__global__ void unpack_layer_kernel(float* __restrict__ flat, const long long* __restrict__ offs,
                                    const float* __restrict__ packed, PackedLayout P, int C, int Cl, int S, int Kz, int video, float scale) {
    MVN_PDL_PROLOGUE();
    const int l = blockIdx.y;
    float* lp[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) { const long long o = offs[MVN_PARAM_LAYER(l, j)]; lp[j] = o < 0 ? nullptr : flat + o; }
    const float* base = packed + P.layer0 + (size_t)l * P.layer_stride;
    const int nW = Cl * Cl * 2, nV = Cl * Cl;
    // segments (reference shapes, Cl channels): wf, wg | vf, vg | bvf, bvg | wr | br | ws | bs
    const int total = 2 * nW + 2 * nV + 2 * Cl + nV + Cl + S * Cl + S;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int j = i;
        if (j < 2 * nW) {
            const int gate = j >= nW; if (gate) j -= nW;
            const int tap = j & 1, k = (j >> 1) % Cl, c = (j >> 1) / Cl;
            if (lp[gate]) lp[gate][j] = scale * base[P.oWz + (size_t)(tap * C + k) * 2 * C + 2 * c + gate];
            continue;
        }
        j -= 2 * nW;
        if (j < 2 * nV) {
            const int gate = j >= nV; if (gate) j -= nV;
            const int k = j % Cl, c = j / Cl;
            float* dst = lp[gate ? 4 : 2];
            if (dst) dst[j] = video ? scale * base[P.oWz + (size_t)(2 * C + k) * 2 * C + 2 * c + gate] : 0.f;
            continue;
        }
        j -= 2 * nV;
        if (j < 2 * Cl) {
            const int gate = j >= Cl; if (gate) j -= Cl;
            float* dst = lp[gate ? 5 : 3];
            if (dst) dst[j] = video ? scale * base[P.obz + 2 * j + gate] : 0.f;
            continue;
        }
        j -= 2 * Cl;
        if (j < nV) { const int k = j % Cl, n = j / Cl; if (lp[6]) lp[6][j] = scale * base[P.oWrs + (size_t)k * (C + S) + n]; continue; }
        j -= nV;
        if (j < Cl) { if (lp[7]) lp[7][j] = scale * base[P.obrs + j]; continue; }
        j -= Cl;
        if (j < S * Cl) { const int k = j % Cl, s = j / Cl; if (lp[8]) lp[8][j] = scale * base[P.oWrs + (size_t)k * (C + S) + C + s]; continue; }
        j -= S * Cl;
        if (lp[9]) lp[9][j] = scale * base[P.obrs + C + j];
    }
}

// input conv, head, video: blockIdx.y selects the group
__global__ void pack_misc_kernel(const float* const* __restrict__ ptrs, float* __restrict__ packed, PackedLayout P,
                                 int A, int C, int Cl, int S, int Cin, int N, int video) {
    MVN_PDL_PROLOGUE();
    const int grp = blockIdx.y;
    const int stride = gridDim.x * blockDim.x, i0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (grp == 0) {            // Win[tap][a][c] = w[c][a][tap]
        const float* w = ptrs[MVN_PARAM_CAUSAL_W];
        for (int i = i0; i < 2 * A * C; i += stride) {
            const int c = i % C, a = (i / C) % A, tap = i / (A * C);
            packed[P.win + i] = c < Cl ? w[((size_t)c * A + a) * 2 + tap] : 0.f;
        }
    } else if (grp == 1) {     // head
        const float *w1 = ptrs[MVN_PARAM_DENSE(N, 0)], *b1 = ptrs[MVN_PARAM_DENSE(N, 1)],
                    *w2 = ptrs[MVN_PARAM_DENSE(N, 2)], *b2 = ptrs[MVN_PARAM_DENSE(N, 3)];
        for (int i = i0; i < S * A; i += stride) {        // w1p[s][a] = w1[a][s]
            const int a = i % A, s = i / A;
            const float v = w1[(size_t)a * S + s];
            packed[P.w1p + i] = v; packed[P.w1pT + (size_t)a * S + s] = v;
        }
        for (int i = i0; i < A * A; i += stride) {        // w2p[k][n] = w2[n][k]
            const int n = i % A, k = i / A;
            const float v = w2[(size_t)n * A + k];
            packed[P.w2p + i] = v; packed[P.w2pT + (size_t)n * A + k] = v;
        }
        for (int i = i0; i < A; i += stride) { packed[P.b1 + i] = b1[i]; packed[P.b2 + i] = b2[i]; }
    } else if (video && grp == 2) {   // video conv: wv[(hw)*Cin + ci][c] = w[c][ci][0][h][w]
        const float *w = ptrs[MVN_PARAM_VIDEO_CONV_W], *b = ptrs[MVN_PARAM_VIDEO_CONV_B];
        const int K = 4096 * Cin;
        for (int i = i0; i < K * C; i += stride) {        // hw fastest: the reads of the reference tensor are coalesced
            const int hw = i & 4095, r = i >> 12, ci = r % Cin, c = r / Cin;
            packed[P.wv + ((size_t)hw * Cin + ci) * C + c] = c < Cl ? w[((size_t)c * Cin + ci) * 4096 + hw] : 0.f;
        }
        for (int i = i0; i < C; i += stride) packed[P.bv + i] = i < Cl ? b[i] : 0.f;
    } else if (video && grp >= 3 && grp < 6) {   // ConvTranspose1d: wt[ci][j*C + co] = w[ci][co][j]
        const int lv = grp - 3;
        const float *w = ptrs[MVN_PARAM_VT_W(lv)], *b = ptrs[MVN_PARAM_VT_B(lv)];
        for (int i = i0; i < C * 10 * C; i += stride) {
            const int n = i % (10 * C), ci = i / (10 * C), j = n / C, co = n % C;
            const float v = (ci < Cl && co < Cl) ? w[((size_t)ci * Cl + co) * 10 + j] : 0.f;
            packed[P.wt[lv] + i] = v;
            packed[P.wtT[lv] + (size_t)n * C + ci] = v;
        }
        for (int i = i0; i < 10 * C; i += stride) packed[P.bt[lv] + i] = (i % C) < Cl ? b[i % C] : 0.f;
    }
}

__global__ void unpack_misc_kernel(float* __restrict__ flat, const long long* __restrict__ offs,
                                   const float* __restrict__ packed, PackedLayout P,
                                   int A, int C, int Cl, int S, int Cin, int N, int video, float scale) {
    MVN_PDL_PROLOGUE();
    const int grp = blockIdx.y;
    auto gp = [&](int i) -> float* { const long long o = offs[i]; return o < 0 ? nullptr : flat + o; };
    const int stride = gridDim.x * blockDim.x, i0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (grp == 0) {
        float* w = gp(MVN_PARAM_CAUSAL_W);
        if (w) for (int i = i0; i < 2 * A * C; i += stride) {
            const int c = i % C, a = (i / C) % A, tap = i / (A * C);
            if (c < Cl) w[((size_t)c * A + a) * 2 + tap] = scale * packed[P.win + i];
        }
    } else if (grp == 1) {
        float *w1 = gp(MVN_PARAM_DENSE(N, 0)), *b1 = gp(MVN_PARAM_DENSE(N, 1)),
              *w2 = gp(MVN_PARAM_DENSE(N, 2)), *b2 = gp(MVN_PARAM_DENSE(N, 3));
        if (w1) for (int i = i0; i < S * A; i += stride) { const int a = i % A, s = i / A; w1[(size_t)a * S + s] = scale * packed[P.w1p + i]; }
        if (w2) for (int i = i0; i < A * A; i += stride) { const int n = i % A, k = i / A; w2[(size_t)n * A + k] = scale * packed[P.w2p + i]; }
        for (int i = i0; i < A; i += stride) { if (b1) b1[i] = scale * packed[P.b1 + i]; if (b2) b2[i] = scale * packed[P.b2 + i]; }
    } else if (grp == 2) {
        float *w = gp(MVN_PARAM_VIDEO_CONV_W), *b = gp(MVN_PARAM_VIDEO_CONV_B);
        const int K = 4096 * Cin;
        if (w) for (int i = i0; i < K * C; i += stride) {    // hw fastest: the writes of the reference-shaped gradient are coalesced
            const int hw = i & 4095, r = i >> 12, ci = r % Cin, c = r / Cin;
            if (c < Cl) w[((size_t)c * Cin + ci) * 4096 + hw] = video ? scale * packed[P.wv + ((size_t)hw * Cin + ci) * C + c] : 0.f;
        }
        if (b) for (int i = i0; i < Cl; i += stride) b[i] = video ? scale * packed[P.bv + i] : 0.f;
    } else if (grp >= 3 && grp < 6) {
        const int lv = grp - 3;
        float *w = gp(MVN_PARAM_VT_W(lv)), *b = gp(MVN_PARAM_VT_B(lv));
        if (w) for (int i = i0; i < C * 10 * C; i += stride) {
            const int n = i % (10 * C), ci = i / (10 * C), j = n / C, co = n % C;
            if (ci < Cl && co < Cl) w[((size_t)ci * Cl + co) * 10 + j] = video ? scale * packed[P.wt[lv] + i] : 0.f;
        }
        if (b) for (int i = i0; i < Cl; i += stride) {
            float acc = 0.f;
            if (video) for (int j = 0; j < 10; ++j) acc += scale * packed[P.bt[lv] + j * C + i];
            b[i] = acc;
        }
    }
}

extern "C" int mvn_pack_weights(const mvn_shape_t* s, const void* const* param_ptrs_dev, void* packed, void* stream) {
    Geo g; MVN_REQUIRE(geo_init(g, s) == 0, "mvn_pack_weights: bad shape");
    PackedLayout P; packed_layout(g, P);
    cudaStream_t st = (cudaStream_t)stream;
    const float* const* ptrs = (const float* const*)param_ptrs_dev;
    // Seven small latency-bound kernels: the independent ones run side by side -- the per-layer re-layout and the tensor-core
    // layer images (both read only the parameters) on two side streams, the chain pack_misc -> head image -> upsampler images
    // (each reads what the one before wrote) on the caller's stream; the caller's stream then waits for the side streams.
    cudaStream_t s0 = mvn_side_stream(0), s1 = mvn_side_stream(1);
    if (!s0 || !s1) s0 = s1 = st;
    int rc;
    if ((rc = mvn_stream_after(s0, st)) || (rc = mvn_stream_after(s1, st))) return rc;
    dim3 gl(8, g.N);
    MVN_CUDA(mvn_launch_pdl(pack_layer_kernel, dim3(gl), dim3(256), (size_t)(0), s0, ptrs, (float*)packed, P, g.C, g.Cl, g.S, g.Kz, g.video, (int)mvn_wide_supported(g)));
    dim3 gm(128, 6);
    MVN_CUDA(mvn_launch_pdl(pack_misc_kernel, dim3(gm), dim3(256), (size_t)(0), st, ptrs, (float*)packed, P, g.A, g.C, g.Cl, g.S, g.Cin, g.N, g.video));
    if ((rc = mvn_check_launch("pack_weights"))) return rc;
    if (g.adt == MVN_DTYPE_BF16 && mvn_tc_layer_supported(g.C, g.S, g.video))
        if ((rc = mvn_tc_pack(ptrs, (float*)packed, P, g, s1))) return rc;
    if (mvn_wide_head_supported(g) && (rc = mvn_wide_pack(ptrs, (float*)packed, P, g, s1))) return rc;
    if (g.adt == MVN_DTYPE_BF16 && mvn_tc_head_supported(g.A, g.S))   // conv2.weight is (A, A, 1): already [n][k]
        if ((rc = mvn_tc_head_pack((const float*)packed + P.w2pT, (float*)packed, P, g.A, st))) return rc;
    if (g.adt == MVN_DTYPE_BF16 && g.video && mvn_tc_upsample_supported(g.C))
    {
        if ((rc = mvn_tc_upsample_pack((const float*)packed + P.wt[2], (const float*)packed + P.bt[2], (float*)packed + P.tc_up, st))) return rc;
        for (int i = 0; i < 2; ++i)
            if ((rc = mvn_tc_upsample_pack((const float*)packed + P.wt[i], (const float*)packed + P.bt[i], (float*)packed + P.tc_up01[i], st))) return rc;
    }
    if ((rc = mvn_stream_after(st, s0)) || (rc = mvn_stream_after(st, s1))) return rc;
    return 0;
}

extern "C" int mvn_unpack_grads(const mvn_shape_t* s, const void* packed_grads, float* flat_grads,
                                const int64_t* offsets_dev, float scale, void* stream) {
    Geo g; MVN_REQUIRE(geo_init(g, s) == 0, "mvn_unpack_grads: bad shape");
    PackedLayout P; packed_layout(g, P);
    cudaStream_t st = (cudaStream_t)stream;
    MVN_REQUIRE(packed_grads && flat_grads && offsets_dev, "mvn_unpack_grads: null buffer");
    const long long* offs = (const long long*)offsets_dev;
    dim3 gl(8, g.N);
    MVN_CUDA(mvn_launch_pdl(unpack_layer_kernel, dim3(gl), dim3(256), (size_t)(0), st, flat_grads, offs, (const float*)packed_grads, P, g.C, g.Cl, g.S, g.Kz, g.video, scale));
    dim3 gm(128, 6);
    MVN_CUDA(mvn_launch_pdl(unpack_misc_kernel, dim3(gm), dim3(256), (size_t)(0), st, flat_grads, offs, (const float*)packed_grads, P, g.A, g.C, g.Cl, g.S, g.Cin, g.N, g.video, scale));
    return mvn_check_launch("unpack_grads");
}
