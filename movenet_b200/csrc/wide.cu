// Wide-channel tensor-core path (residual_channels >= 128: the widened scale-up shape, BASELINE configs[3] / SURVEY "03w").
//
// Every GEMM of the layer and of the head runs on the weight-streaming tcgen05 engine of wide_gemm.cuh with its element-wise
// tail fused into the epilogue:
//   forward, per layer   F1  [x(t-d) | x(t)] . Wz^T -> gate                       -> gated            (modules.py:73-80)
//                        F2  gated . [Wr | Ws]^T    -> + br + x(t), skip_sum +=   -> x', skip_sum     (modules.py:83-91)
//   head                 H1  lrelu(skip_sum) . W1^T + b1                          -> a1, lrelu(a1)    (modules.py:139-141)
//                        H2  lrelu(a1) . W2^T + b2 -> softmax -> (B, A, Tn)       -> probabilities    (wavenet.py:187-191)
//   backward, head       HB1 dz . W2 * lrelu'(a1) -> d(a1) ; HB2 d(a1) . W1 * lrelu'(skip_sum) -> d(skip) (bf16, T row space)
//   backward, per layer  B1a [d(x') | d(skip)] . [Wr | Ws] -> d(gated)
//                        B1b recompute [x(t-d) | x(t)] . Wz^T, gate derivative    -> gated, dz = (df, dg)
//                        B2  [dz(t) | dz(t+d)] . Wz -> + d(x')                    -> d(x)
// The weight gradients are PLAIN GEMMs with K = time (X^T . dz, gated^T . [d(x') | d(skip)], ...): they go to cuBLAS
// (bf16 operands, fp32 accumulate and output), loaded at run time from the process (no link-time dependency); bias gradients
// are deterministic two-stage column sums.  Results land in the packed-gradient layout of layout.h, so mvn_unpack_grads and
// everything above the C ABI is unchanged.
#include <dlfcn.h>
#include <cstdlib>
#include <mutex>
#include "wide_gemm.cuh"
#include "wide_wgrad.cuh"
#include "layer_tc.h"

using namespace wide;

namespace {

// ------------------------------------------------------------------------------------------------ tensor maps (cached)
struct WMapSlot { const void* ptr; unsigned long long d0, d1, d2; unsigned box0, box1; bool used; CUtensorMap map; };
constexpr int WMAP_SLOTS = 1024;
WMapSlot g_wmaps[WMAP_SLOTS];
std::mutex g_wmaps_mu;

EncodeTiledFn w_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

// tensor [d2][d1][d0] (d0 contiguous), box {box0, box1, 1} with box0 elements = 128 bytes (64 bf16 or 32 fp32), 128-byte
// swizzle; d2 == 0: a 2-D matrix [d1][d0]
int w_map(CUtensorMap* map, const void* ptr, unsigned long long d0, unsigned long long d1, unsigned long long d2, unsigned box1,
          unsigned box0 = 64) {
    const unsigned es = box0 == 64 ? 2 : 4;
    const size_t h = ((size_t)(uintptr_t)ptr >> 8) * 0x9E3779B97F4A7C15ull ^ (d0 * 31 + d1 * 131 + d2 * 1313 + box1 * 7 + box0);
    const int i0 = (int)((h >> 17) % WMAP_SLOTS);
    {
        std::lock_guard<std::mutex> lk(g_wmaps_mu);
        for (int p = 0; p < 8; ++p) {
            const WMapSlot& s = g_wmaps[(i0 + p) % WMAP_SLOTS];
            if (s.used && s.ptr == ptr && s.d0 == d0 && s.d1 == d1 && s.d2 == d2 && s.box1 == box1 && s.box0 == box0) { *map = s.map; return 0; }
        }
    }
    EncodeTiledFn fn = w_encode_fn();
    MVN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
    MVN_REQUIRE((((uintptr_t)ptr) & 15) == 0 && (d0 * es) % 16 == 0, "wide path: operands must be 16-byte aligned");
    const int rank = d2 ? 3 : 2;
    cuuint64_t dims[3] = {d0, d1, d2 ? d2 : 1};
    cuuint64_t strides[2] = {d0 * es, d0 * d1 * es};
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MVN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    std::lock_guard<std::mutex> lk(g_wmaps_mu);
    int victim = i0;
    for (int p = 0; p < 8; ++p) if (!g_wmaps[(i0 + p) % WMAP_SLOTS].used) { victim = (i0 + p) % WMAP_SLOTS; break; }
    WMapSlot& s = g_wmaps[victim];
    s.ptr = ptr; s.d0 = d0; s.d1 = d1; s.d2 = d2; s.box1 = box1; s.box0 = box0; s.map = *map; s.used = true;
    return 0;
}

// ------------------------------------------------------------------------------------------------ engine launcher
// MOVENET_B200_WIDE_PAIR=1: single-CTA tiles (M = 128) instead of CTA pairs (read per call: the tests switch it)
int pair_mode() {
    const char* e = getenv("MOVENET_B200_WIDE_PAIR");
    return (e && atoi(e) == 1) ? 1 : 2;
}

struct Operand { const void* ptr; int cols; };       // time-major bf16 (B, rows, cols)
struct Output { const void* ptr; int cols, rows, fp32; };   // TMA-stored result: time-major (B, rows, cols) bf16 or fp32

template <int PAIR, int EPI>
int launch_t(const CUtensorMap& mA0, const CUtensorMap& mA1, const CUtensorMap& mB, const CUtensorMap& mO0, const CUtensorMap& mO1,
             const Args& a, cudaStream_t st) {
    using C = Cfg<PAIR>;
    static MvnSmemAttr attr;
    MVN_CUDA(mvn_ensure_smem(wide_gemm_kernel<PAIR, EPI>, C::SMEM, attr));
    int clusters = mvn_sm_count() / PAIR;
    if (clusters > a.n_tiles) clusters = a.n_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * PAIR); cfg.blockDim = dim3(N_THREADS); cfg.dynamicSmemBytes = C::SMEM; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = PAIR; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = PAIR == 2 ? 2 : 1;
    MVN_CUDA(cudaLaunchKernelEx(&cfg, wide_gemm_kernel<PAIR, EPI>, mA0, mA1, mB, mO0, mO1, a));
    static const char* names[EPI_COUNT] = {"wide_gemm<gate>", "wide_gemm<resid_skip>", "wide_gemm<store>", "wide_gemm<gate_bwd>",
                                           "wide_gemm<add_store>", "wide_gemm<head1>", "wide_gemm<head2>", "wide_gemm<lrelu_bwd>", "wide_gemm<dz>"};
    return mvn_check_launch(names[EPI]);
}

template <int EPI>
int launch(const Operand& A0, const Operand& A1, const void* W, int w_rows, int w_k, Args& a, int B, int rows, cudaStream_t st,
           Output O0 = Output{nullptr, 0, 0, 0}, Output O1 = Output{nullptr, 0, 0, 0}) {
    const int pair = pair_mode();
    a.B = B; a.rows = rows;
    a.tiles_per_clip = (rows + BM * pair - 1) / (BM * pair);
    a.n_tiles = a.tiles_per_clip * B;
    a.nkb = 0;
    for (int i = 0; i < a.nseg; ++i) a.nkb += a.seg[i].nkb;
    MVN_REQUIRE(a.N % 64 == 0 && a.N > 0 && a.nkb > 0, "wide path: bad GEMM shape");
    CUtensorMap mA0, mA1, mB;
    int rc;
    if ((rc = w_map(&mA0, A0.ptr, A0.cols, rows, B, BM))) return rc;
    const Operand& A1r = A1.ptr ? A1 : A0;
    if ((rc = w_map(&mA1, A1r.ptr, A1r.cols, rows, B, BM))) return rc;
    if ((rc = w_map(&mB, W, w_k, w_rows, 0, NCH / pair))) return rc;
    CUtensorMap mO0 = mA0, mO1 = mA0;                 // (unused by the epilogues that store directly)
    if (O0.ptr && (rc = w_map(&mO0, O0.ptr, O0.cols, O0.rows, B, 32, O0.fp32 ? 32 : 64))) return rc;
    if (O1.ptr && (rc = w_map(&mO1, O1.ptr, O1.cols, O1.rows, B, 32, O1.fp32 ? 32 : 64))) return rc;
    return pair == 2 ? launch_t<2, EPI>(mA0, mA1, mB, mO0, mO1, a, st) : launch_t<1, EPI>(mA0, mA1, mB, mO0, mO1, a, st);
}

Args new_args() { Args a; memset(&a, 0, sizeof(a)); return a; }
void seg(Args& a, int map, int cols, int shift, int c0 = 0) { a.seg[a.nseg++] = Seg{map, cols / BK, shift, c0}; }

// ------------------------------------------------------------------------------------------------ small kernels
// bf16 K-major weight matrices from the reference tensors (ptrs in state_dict order); blockIdx.y = layer, or N for the head
__global__ void wide_pack_kernel(const float* const* __restrict__ ptrs, float* __restrict__ packed, PackedLayout P, int C, int S,
                                 int A, int N, int layers) {
    MVN_PDL_PROLOGUE();
    const int l = blockIdx.y;
    const int i0 = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    if (l < N) {
        if (!layers) return;          // head-only use (the residual stack runs on the fused C <= 64 kernels)
        const float* const* lp = ptrs + MVN_PARAM_LAYER(l, 0);
        const float *wf = lp[0], *wg = lp[1], *wr = lp[6], *ws = lp[8];
        float* base = packed + P.layer0 + (size_t)l * P.layer_stride;
        __nv_bfloat16* Wz = (__nv_bfloat16*)(base + P.wWz);
        __nv_bfloat16* Wrs = (__nv_bfloat16*)(base + P.wWrs);
        __nv_bfloat16* WrsT = (__nv_bfloat16*)(base + P.wWrsT);
        __nv_bfloat16* WzT = (__nv_bfloat16*)(base + P.wWzT);
        const int K = 2 * C;
        for (int i = i0; i < 2 * C * K; i += stride) {        // Wz: row = 256 j + n ; n < 128 filter, else gate, channel 128 j + n % 128
            const int row = i / K, k = i - row * K, n = row & 255, ch = 128 * (row >> 8) + (n & 127), gate = n >> 7;
            const int tap = k / C, cin = k - tap * C;
            Wz[i] = __float2bfloat16((gate ? wg : wf)[((size_t)ch * C + cin) * 2 + tap]);
        }
        for (int i = i0; i < C * 2 * C; i += stride) {        // [Wr | I][n][k]: k < C the residual 1x1 conv, k >= C the identity
            const int n = i / (2 * C), k = i - n * 2 * C;
            Wrs[i] = __float2bfloat16(k < C ? wr[(size_t)n * C + k] : (k - C == n ? 1.f : 0.f));
        }
        for (int i = i0; i < (C + S) * C; i += stride) {      // WrsT[k][n] = [Wr | Ws][n][k]
            const int n = i / C, k = i - n * C;
            const float v = n < C ? wr[(size_t)n * C + k] : ws[(size_t)(n - C) * C + k];
            WrsT[(size_t)k * (C + S) + n] = __float2bfloat16(v);
        }
        for (int i = i0; i < C * 4 * C; i += stride) {        // WzT[c][k]: k = seg * 2C + 2 o + gate ; seg 0 <-> tap 1 (dz(t)), seg 1 <-> tap 0 (dz(t+d))
            const int c = i / (4 * C), k = i - c * 4 * C, sgm = k / (2 * C), kk = k - sgm * 2 * C, o = kk >> 1, gate = kk & 1;
            WzT[i] = __float2bfloat16((gate ? wg : wf)[((size_t)o * C + c) * 2 + (sgm == 0 ? 1 : 0)]);
        }
    } else {
        const float *w1 = ptrs[MVN_PARAM_DENSE(N, 0)], *w2 = ptrs[MVN_PARAM_DENSE(N, 2)];
        __nv_bfloat16* H1 = (__nv_bfloat16*)(packed + P.wH1);
        __nv_bfloat16* H2 = (__nv_bfloat16*)(packed + P.wH2);
        __nv_bfloat16* H2T = (__nv_bfloat16*)(packed + P.wH2T);
        __nv_bfloat16* H1T = (__nv_bfloat16*)(packed + P.wH1T);
        for (int i = i0; i < A * S; i += stride) {            // conv1.weight (A, S, 1)
            const int a = i / S, s = i - a * S;
            H1[i] = __float2bfloat16(w1[i]);
            H1T[(size_t)s * A + a] = __float2bfloat16(w1[i]);
        }
        for (int i = i0; i < A * A; i += stride) {            // conv2.weight (A, A, 1): [n][k]
            const int n = i / A, k = i - n * A;
            H2[i] = __float2bfloat16(w2[i]);
            H2T[(size_t)k * A + n] = __float2bfloat16(w2[i]);
        }
        __nv_bfloat16* WsAll = (__nv_bfloat16*)(packed + P.wWsAll);
        const int KA = layers ? N * C : 0;
        for (int i = i0; i < S * KA; i += stride) {           // WsAll[s][l C + c] = conv_skip_l.weight[s][c]
            const int sidx = i / KA, k = i - sidx * KA, l2 = k / C, c = k - l2 * C;
            WsAll[i] = __float2bfloat16(ptrs[MVN_PARAM_LAYER(l2, 8)][(size_t)sidx * C + c]);
        }
        for (int i = i0; layers && i < S; i += stride) {      // fixed order: deterministic
            float acc = 0.f;
            for (int l2 = 0; l2 < N; ++l2) acc += ptrs[MVN_PARAM_LAYER(l2, 9)][i];
            packed[P.wbsum + i] = acc;
        }
    }
}

// dst (B, rows_dst, cols) bf16: row t = lrelu(src row t + shift) for t < rows_valid, zero beyond; src is (B, src_rows, cols) fp32
__global__ void lrelu16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int rows_valid, int rows_dst, int cols,
                               int src_rows, int shift) {
    MVN_PDL_PROLOGUE();
    const long long n8 = (long long)B * rows_dst * cols / 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i * 8, r = e / cols; const int c = (int)(e - r * cols);
        const int b = (int)(r / rows_dst), t = (int)(r - (long long)b * rows_dst);
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (t < rows_valid) {
            const float4* s = (const float4*)(src + ((size_t)b * src_rows + t + shift) * cols + c);
            const float4 a = s[0], q = s[1];
            v[0] = mvn_lrelu(a.x); v[1] = mvn_lrelu(a.y); v[2] = mvn_lrelu(a.z); v[3] = mvn_lrelu(a.w);
            v[4] = mvn_lrelu(q.x); v[5] = mvn_lrelu(q.y); v[6] = mvn_lrelu(q.z); v[7] = mvn_lrelu(q.w);
        }
        st_bf16x8(dst + e, v);
    }
}

// audio as a bf16 GEMM operand (B, T, A): exact for one-hot columns (the codes), rounded values for dense columns
__global__ void onehot16_kernel(const float* __restrict__ audio, const int* __restrict__ codes, const unsigned char* __restrict__ dense,
                                __nv_bfloat16* __restrict__ oh, int B, int T, int A) {
    MVN_PDL_PROLOGUE();
    const long long n8 = (long long)B * T * A / 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const long long e = i * 8, r = e / A; const int a0 = (int)(e - r * A);
        const int b = (int)(r / T), t = (int)(r - (long long)b * T);
        float v[8];
        if (!dense[r]) { const int c = codes[r];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (a0 + j == c) ? 1.f : 0.f;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = audio[((size_t)b * A + a0 + j) * T + t];
        }
        st_bf16x8(oh + e, v);
    }
}

// d(logits) (B, Tout, A) bf16 time-major (rows >= Tn zero) from the channels-first fp32 tensors:
//   logits mode          dz = d(out)
//   d(out) given         dz = p * (dp - <dp, p>)                                   (softmax backward, wavenet.py:191)
//   fused loss           dp = gs * (softmax_c(p) - onehot(target)), gs = d(loss) / (B Tn), then as above (loss.py)
__global__ void __launch_bounds__(256) head_dz16_kernel(const float* __restrict__ probs, const float* __restrict__ dout,
                                                        const long long* __restrict__ target, const float* __restrict__ gloss,
                                                        __nv_bfloat16* __restrict__ dz, int A, int Tn, int Tout, int logits, float inv_count) {
    extern __shared__ float tile[];        // p[32][A+1] (, dp[32][A+1] when d(out) is given)
    const int b = blockIdx.y, j0 = blockIdx.x * 32, ld = A + 1;
    float* tp = tile; float* td = tile + 32 * ld;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int j = j0 + lane;
    if (j < Tn) {
#pragma unroll 4
        for (int a = warp; a < A; a += nw) {
            const size_t o = ((size_t)b * A + a) * Tn + j;
            if (dout) td[lane * ld + a] = dout[o];
            if (!logits) tp[lane * ld + a] = probs[o];
        }
    }
    __syncthreads();
    const int per = A / 32;                // channels per lane: lane, lane + 32, ... (A % 32 == 0, A <= 256; conflict-free tile reads)
    for (int r = warp; r < 32; r += nw) {
        const int jj = j0 + r;
        if (jj >= Tout) continue;
        __nv_bfloat16* dr = dz + ((size_t)b * Tout + jj) * A;
        if (jj >= Tn) { for (int e = 0; e < per; ++e) dr[lane + 32 * e] = __float2bfloat16(0.f); continue; }
        if (logits) { for (int e = 0; e < per; ++e) dr[lane + 32 * e] = __float2bfloat16(td[r * ld + lane + 32 * e]); continue; }
        float p[8], dp[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) p[e] = e < per ? tp[r * ld + lane + 32 * e] : 0.f;
        if (dout) {
#pragma unroll
            for (int e = 0; e < 8; ++e) dp[e] = e < per ? td[r * ld + lane + 32 * e] : 0.f;
        } else {
            float m = -INFINITY;
#pragma unroll
            for (int e = 0; e < 8; ++e) if (e < per) m = fmaxf(m, p[e]);
            for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float z = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) { dp[e] = e < per ? expf(p[e] - m) : 0.f; z += dp[e]; }
            for (int o = 16; o; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
            const float gs = gloss[0] * inv_count, inv = gs / z;
            const int tg = (int)target[(size_t)b * Tn + jj];
#pragma unroll
            for (int e = 0; e < 8; ++e) dp[e] = dp[e] * inv - (lane + 32 * e == tg ? gs : 0.f);
        }
        float dot = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) dot = fmaf(dp[e], p[e], dot);
        for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
#pragma unroll
        for (int e = 0; e < 8; ++e) if (e < per) dr[lane + 32 * e] = __float2bfloat16(p[e] * (dp[e] - dot));
    }
}

// deterministic column sums of a bf16 matrix [rows][cols] (cols % 2 == 0, cols <= 1024): stage 1 writes one partial row per
// block, stage 2 adds the partials of a column in a fixed order
#define CS_BLOCKS 296
__global__ void __launch_bounds__(512) colsum1_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int cols, float* __restrict__ partial) {
    MVN_PDL_PROLOGUE();
    // 16-byte loads: cols / 8 threads cover a row, blockDim / (cols / 8) rows are in flight per step, two steps unrolled
    const int c8 = cols / 8, tx = threadIdx.x % c8, ty = threadIdx.x / c8, ny = blockDim.x / c8;
    const long long per = (rows + gridDim.x - 1) / gridDim.x, r0 = blockIdx.x * per, r1 = r0 + per < rows ? r0 + per : rows;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (ty < ny) {
        long long r = r0 + ty;
        for (; r + ny < r1; r += 2 * ny) {
            float v0[8], v1[8];
            ld_bf16x8(x + r * cols + 8 * tx, v0);
            ld_bf16x8(x + (r + ny) * cols + 8 * tx, v1);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += v0[e] + v1[e];
        }
        if (r < r1) {
            float v0[8];
            ld_bf16x8(x + r * cols + 8 * tx, v0);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += v0[e];
        }
    }
    __shared__ float red[8][512];
#pragma unroll
    for (int e = 0; e < 8; ++e) red[e][threadIdx.x] = acc[e];
    __syncthreads();
    if (ty == 0) {
        for (int y = 1; y < ny; ++y)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += red[e][y * c8 + tx];
#pragma unroll
        for (int e = 0; e < 8; ++e) partial[(size_t)blockIdx.x * cols + 8 * tx + e] = acc[e];
    }
}
// blockDim = (32, RED_SPLIT): 32 columns per block, the partial rows of a column dealt to RED_SPLIT threads and combined in a
// fixed order (tc::column_sum)
__global__ void colsum2_kernel(const float* __restrict__ partial, int nblk, int cols, float* __restrict__ out) {
    MVN_PDL_PROLOGUE();
    const int c = blockIdx.x * 32 + threadIdx.x;
    const float tot = column_sum(partial, nblk, (size_t)cols, (size_t)c, c < cols);
    if (threadIdx.y == 0 && c < cols) out[c] = tot;
}

int colsum(const void* x, long long rows, int cols, float* partial, float* out, cudaStream_t st) {
    MVN_REQUIRE(cols % 8 == 0 && cols <= 1024, "wide path: column sum width");
    MVN_CUDA(mvn_launch_pdl(colsum1_kernel, dim3(CS_BLOCKS), dim3(512), (size_t)0, st, (const __nv_bfloat16*)x, rows, cols, partial));
    int rc = mvn_check_launch("colsum1"); if (rc) return rc;
    MVN_CUDA(mvn_launch_pdl(colsum2_kernel, dim3(mvn_cdiv(cols, 32)), dim3(32, RED_SPLIT), (size_t)0, st, (const float*)partial, (int)CS_BLOCKS, cols, out));
    return mvn_check_launch("colsum2");
}

// ------------------------------------------------------------------------------------------------ cuBLAS (plain GEMMs only)
typedef int (*CreateFn)(void**);
typedef int (*SetStreamFn)(void*, cudaStream_t);
typedef int (*SetWorkspaceFn)(void*, void*, size_t);
typedef int (*GemmExFn)(void*, int, int, int, int, int, const void*, const void*, int, int, const void*, int, int, const void*, void*, int,
                        int, int, int);
struct Blas {
    void* handle[MVN_MAX_DEVICES] = {};
    SetStreamFn set_stream = nullptr; GemmExFn gemm = nullptr; CreateFn create = nullptr;
    std::mutex mu; bool tried = false;
};
Blas g_blas;

int blas_handle(void** out) {
    std::lock_guard<std::mutex> lk(g_blas.mu);
    if (!g_blas.tried) {
        g_blas.tried = true;
        void* lib = dlopen("libcublas.so.12", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libcublas.so", RTLD_NOW | RTLD_GLOBAL);
        if (lib) {
            g_blas.create = (CreateFn)dlsym(lib, "cublasCreate_v2");
            g_blas.set_stream = (SetStreamFn)dlsym(lib, "cublasSetStream_v2");
            g_blas.gemm = (GemmExFn)dlsym(lib, "cublasGemmEx");
        }
    }
    MVN_REQUIRE(g_blas.create && g_blas.set_stream && g_blas.gemm, "wide path: cuBLAS (libcublas.so.12) is not loadable in this process");
    int dev = 0;
    MVN_CUDA(cudaGetDevice(&dev));
    MVN_REQUIRE(dev >= 0 && dev < MVN_MAX_DEVICES, "wide path: device index");
    if (!g_blas.handle[dev]) MVN_REQUIRE(g_blas.create(&g_blas.handle[dev]) == 0, "cublasCreate failed");
    *out = g_blas.handle[dev];
    return 0;
}

// row-major C[m x n] (ld ldc, fp32) (+)= A^T . B with A row-major [k x m] (ld lda), B row-major [k x n] (ld ldb), both bf16:
// the weight-gradient shape (K = time).  In cuBLAS's column-major terms: C^T[n x m] = B^T-view(N) x A-view(T).
int gemm_tn(const void* A, int lda, const void* Bm, int ldb, float* Cm, int ldc, int m, int n, long long k, bool accumulate,
            cudaStream_t st) {
    if (k <= 0) return 0;
    void* h; int rc = blas_handle(&h); if (rc) return rc;
    MVN_REQUIRE(g_blas.set_stream(h, st) == 0, "cublasSetStream failed");
    const float alpha = 1.f, beta = accumulate ? 1.f : 0.f;
    // CUBLAS_OP_N = 0, CUBLAS_OP_T = 1 ; CUDA_R_16BF = 14, CUDA_R_32F = 0 ; CUBLAS_COMPUTE_32F = 68 ; CUBLAS_GEMM_DEFAULT = -1
    const int s = g_blas.gemm(h, 0, 1, n, m, (int)k, &alpha, Bm, 14, ldb, A, 14, lda, &beta, Cm, 0, ldc, 68, -1);
    MVN_REQUIRE(s == 0, "cublasGemmEx failed (%d) for m=%d n=%d k=%lld", s, m, n, k);
    return 0;
}

// ---- tcgen05 weight gradients (wide_wgrad.cuh) -------------------------------------------------------------------------
// MOVENET_B200_WIDE_WGRAD=cublas sends them to cuBLAS instead (also the route for channel counts that are not multiples of 256)
bool wgrad_tc_ok(const Geo& g) {
    const char* e = getenv("MOVENET_B200_WIDE_WGRAD");
    return !(e && e[0] == 'c') && g.C % 256 == 0 && g.S % 256 == 0 && mvn_sm_count() >= 2;
}

struct WgTensors { const void* ptr[5]; int cols[5]; };     // time-major bf16 (B, T, cols)

int wgrad_launch(WgArgs& a, const WgTensors& t, const Geo& g, float* partial, cudaStream_t st) {
    const int pairs = mvn_sm_count() / 2;
    MVN_REQUIRE(a.n_jobs >= 1 && a.n_jobs <= WG_MAX_JOBS && a.n_jobs <= pairs && pairs <= 80, "wide path: weight-gradient job count");
    a.B = g.B; a.T = g.T; a.kb_per_clip = (g.T + 63) / 64; a.n_splits = pairs / a.n_jobs; a.partial = partial;
    CUtensorMap m[5];
    int rc;
    for (int i = 0; i < 5; ++i)
        if ((rc = w_map(&m[i], t.ptr[i] ? t.ptr[i] : t.ptr[0], t.ptr[i] ? t.cols[i] : t.cols[0], g.T, g.B, 64))) return rc;
    static MvnSmemAttr attr;
    MVN_CUDA(mvn_ensure_smem(wide_wgrad_kernel, WG_SMEM, attr));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * a.n_jobs * a.n_splits); cfg.blockDim = dim3(N_THREADS); cfg.dynamicSmemBytes = WG_SMEM; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = 2; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    MVN_CUDA(cudaLaunchKernelEx(&cfg, wide_wgrad_kernel, m[0], m[1], m[2], m[3], m[4], a));
    if ((rc = mvn_check_launch("wide_wgrad"))) return rc;
    MVN_REQUIRE(!a.cs_partial || a.cs_cols <= 256, "wide path: passenger column sum width");
    MVN_CUDA(mvn_launch_pdl(wide_wgrad_reduce_kernel, dim3(128, a.n_jobs + (a.cs_partial ? 1 : 0)), dim3(256), (size_t)0, st, a));
    return mvn_check_launch("wide_wgrad_reduce");
}

}  // namespace

// ================================================================================================ public (library-internal) API
int mvn_wide_supported(const Geo& g) { return wide_ok(g); }

int mvn_wide_head_supported(const Geo& g) { return wide_ok(g) || wide_head_ok(g); }

int mvn_wide_pack(const float* const* param_ptrs_dev, float* packed, const PackedLayout& P, const Geo& g, cudaStream_t st) {
    dim3 grid(64, g.N + 1);
    MVN_CUDA(mvn_launch_pdl(wide_pack_kernel, dim3(grid), dim3(256), (size_t)0, st, param_ptrs_dev, packed, P, g.C, g.S, g.A, g.N, (int)wide_ok(g)));
    return mvn_check_launch("wide_pack");
}

// GatedResidualConv1d.forward (movenet/modules.py:67-93): F1 (gate) + F2 (residual 1x1 conv + x).  The gated activations of
// every layer are kept side by side in one (B, T, N C) tensor: the skip 1x1 convs of ALL layers then run as ONE GEMM with
// K = N C whose fp32 result is written once (mvn_wide_skip_fwd) instead of a 2 KB-per-row read-modify-write of skip_sum per
// layer (which made the per-layer residual + skip GEMM HBM-bound), and the backward reads them for the 1x1 convs' weight
// gradients instead of re-writing them.  The last layer's residual output is discarded (modules.py:125-130): no F2.
int mvn_wide_layer_fwd(const void* x_in, void* x_out, void* gated_all, void* gab_all, const float* packed, const PackedLayout& P,
                       const Geo& g, int l, cudaStream_t st) {
    const int C = g.C, d = g.dil[l], NC = g.N * C;
    const float* lw = packed + P.layer0 + (size_t)l * P.layer_stride;
    int rc;
    {   // gab_all != null (training): the gate's derivative factors are kept next to the gated activations
        Args a = new_args();
        seg(a, 0, C, -d); seg(a, 0, C, 0);
        a.N = 2 * C; a.out = gated_all; a.out_c0 = l * C; a.out2 = gab_all; a.out2_c0 = l * 2 * C;
        if ((rc = launch<EPI_GATE>(Operand{x_in, C}, Operand{nullptr, 0}, lw + P.wWz, 2 * C, 2 * C, a, g.B, g.T, st, Output{gated_all, NC, g.T, 0},
                                   Output{gab_all ? gab_all : gated_all, gab_all ? 2 * NC : NC, g.T, 0}))) return rc;
    }
    if (!x_out) return 0;
    // x' = Wr gated + br + x as ONE product [gated | x] . [Wr | I]^T: the residual add costs C^2 more MACs on the tensor pipe
    // (bf16 x times 1.0, accumulated in fp32: exact) instead of a per-row global load in the epilogue, which with only
    // K = C of MMA work per tile to hide it made this GEMM 74 us instead of 50 (measured)
    Args a = new_args();
    seg(a, 0, C, 0, l * C); seg(a, 1, C, 0);
    a.N = C; a.bias = lw + P.obrs; a.n_resid = C; a.aux = nullptr; a.out = x_out; a.ld_out = C;
    return launch<EPI_RESID_SKIP>(Operand{gated_all, NC}, Operand{x_in, C}, lw + P.wWrs, C, 2 * C, a, g.B, g.T, st,
                                  Output{x_out, C, g.T, 0}, Output{x_out, C, g.T, 0});
}

// skip_sum = sum_l (Ws_l gated_l + bs_l) (movenet/modules.py:90-91, wavenet.py:181) as one GEMM over the layer-concatenated
// gated activations; fp32, on the T row space (row t = time t; the head reads rows >= RF - 1)
int mvn_wide_skip_fwd(const void* gated_all, float* skip, const float* packed, const PackedLayout& P, const Geo& g, cudaStream_t st) {
    const int NC = g.N * g.C;
    Args a = new_args();
    seg(a, 0, NC, 0);
    a.N = g.S; a.bias = packed + P.wbsum; a.n_resid = 0; a.skip_init = 1; a.S = g.S; a.RF = g.RF; a.Tout = g.Tout;
    return launch<EPI_RESID_SKIP>(Operand{gated_all, NC}, Operand{nullptr, 0}, packed + P.wWsAll, g.S, NC, a, g.B, g.T, st,
                                  Output{gated_all, NC, g.T, 0}, Output{skip, g.S, g.T, 1});
}

// DenseConv + drop-last + softmax (movenet/modules.py:133-142, wavenet.py:183-191)
// skip_on_T: skip_sum lives on the T row space (the wide layer path: row j of the head is row j + RF - 1 of it); otherwise it
// is the (B, Tout, S) tensor of the fused C <= 64 layer kernels
int mvn_wide_head_fwd(const float* packed, const PackedLayout& P, const Geo& g, const float* skip, int skip_on_T, float* a1, float* out,
                      void* l0, void* l1, cudaStream_t st) {
    int rc;
    const int srows = skip_on_T ? g.T : g.Tout, sshift = skip_on_T ? g.RF - 1 : 0;
    MVN_CUDA(mvn_launch_pdl(lrelu16_kernel, dim3(2 * mvn_sm_count()), dim3(256), (size_t)0, st, skip, (__nv_bfloat16*)l0, g.B, g.Tout, g.Tout, g.S, srows, sshift));
    if ((rc = mvn_check_launch("lrelu16"))) return rc;
    {
        Args a = new_args();
        seg(a, 0, g.S, 0);
        a.N = g.A; a.bias = packed + P.b1; a.out = a1; a.ld_out = g.A; a.out2 = l1; a.ld_out2 = g.A; a.Tn = g.Tn;
        if ((rc = launch<EPI_HEAD1>(Operand{l0, g.S}, Operand{nullptr, 0}, packed + P.wH1, g.A, g.S, a, g.B, g.Tout, st))) return rc;
    }
    Args a = new_args();
    seg(a, 0, g.A, 0);
    a.N = g.A; a.bias = packed + P.b2; a.out = out; a.Tn = g.Tn; a.logits = g.logits;
    return launch<EPI_HEAD2>(Operand{l1, g.A}, Operand{nullptr, 0}, packed + P.wH2, g.A, g.A, a, g.B, g.Tout, st);
}

// head backward: d(skip) (bf16, T row space) and the head's weight / bias gradients
// ds16 != null: d(skip) as bf16 on the T row space (what the wide layer kernels read); else dskip32: fp32 (B, Tout, S) (what the
// fused C <= 64 layer kernels read)
int mvn_wide_head_bwd(const float* packed, const PackedLayout& P, const Geo& g, const float* skip, int skip_on_T, const float* a1,
                      const float* probs, const float* dout, const long long* target, const float* grad_loss, void* dzh, void* da1,
                      void* l0, void* l1, void* ds16, float* dskip32, float* colsum_ws, float* pg, cudaStream_t st) {
    int rc;
    const long long rows = (long long)g.B * g.Tout;
    const int srows = skip_on_T ? g.T : g.Tout, sshift = skip_on_T ? g.RF - 1 : 0;
    if (ds16) MVN_CUDA(cudaMemsetAsync(ds16, 0, (size_t)g.B * g.T * g.S * 2, st));
    else MVN_CUDA(cudaMemsetAsync(dskip32, 0, (size_t)g.B * g.Tout * g.S * 4, st));
    if (g.Tn <= 0) return 0;
    {
        dim3 grid(mvn_cdiv(g.Tout, 32), g.B);
        const size_t smem = (size_t)(dout ? 2 : 1) * 32 * (g.A + 1) * 4;      // (the d(out) tile only when d(out) is given)
        static MvnSmemAttr attr;
        MVN_CUDA(mvn_ensure_smem(head_dz16_kernel, (int)((size_t)2 * 32 * (g.A + 1) * 4), attr));
        head_dz16_kernel<<<grid, 256, smem, st>>>(probs, dout, target, grad_loss, (__nv_bfloat16*)dzh, g.A, g.Tn, g.Tout, g.logits,
                                                  1.f / ((float)g.B * (float)g.Tn));
        if ((rc = mvn_check_launch("head_dz16"))) return rc;
    }
    MVN_CUDA(mvn_launch_pdl(lrelu16_kernel, dim3(2 * mvn_sm_count()), dim3(256), (size_t)0, st, a1, (__nv_bfloat16*)l1, g.B, g.Tn, g.Tout, g.A, g.Tn, 0));
    if ((rc = mvn_check_launch("lrelu16"))) return rc;
    MVN_CUDA(mvn_launch_pdl(lrelu16_kernel, dim3(2 * mvn_sm_count()), dim3(256), (size_t)0, st, skip, (__nv_bfloat16*)l0, g.B, g.Tout, g.Tout, g.S, srows, sshift));
    if ((rc = mvn_check_launch("lrelu16"))) return rc;
    // conv2: dW2p[k][n] = sum_t lrelu(a1)[t][k] dz[t][n] ; db2 = sum_t dz
    if ((rc = gemm_tn(l1, g.A, dzh, g.A, pg + P.w2p, g.A, g.A, g.A, rows, false, st))) return rc;
    if ((rc = colsum(dzh, rows, g.A, colsum_ws, pg + P.b2, st))) return rc;
    {   // d(a1) = (dz . W2) * lrelu'(a1)
        Args a = new_args();
        seg(a, 0, g.A, 0);
        a.N = g.A; a.aux = a1; a.ld_aux = g.A; a.aux_rows = g.Tn; a.out = da1; a.ld_out = g.A; a.out_rows = g.Tout; a.out_shift = 0; a.Tn = g.Tn;
        MVN_CUDA(cudaMemsetAsync(da1, 0, (size_t)rows * g.A * 2, st));          // the rows >= Tn stay zero
        if ((rc = launch<EPI_LRELU_BWD>(Operand{dzh, g.A}, Operand{nullptr, 0}, packed + P.wH2T, g.A, g.A, a, g.B, g.Tout, st))) return rc;
    }
    // conv1: dW1p[s][a] = sum_t lrelu(skip)[t][s] d(a1)[t][a] ; db1 = sum_t d(a1)
    if ((rc = gemm_tn(l0, g.S, da1, g.A, pg + P.w1p, g.A, g.S, g.A, rows, false, st))) return rc;
    if ((rc = colsum(da1, rows, g.A, colsum_ws, pg + P.b1, st))) return rc;
    // d(skip_sum) = (d(a1) . W1) * lrelu'(skip_sum), written at row t = j + RF - 1 of the (B, T, S) gradient every layer reads
    Args a = new_args();
    seg(a, 0, g.A, 0);
    a.N = g.S; a.aux = skip; a.ld_aux = g.S; a.aux_rows = srows; a.aux_shift = sshift; a.ld_out = g.S; a.Tn = g.Tn;
    if (ds16) { a.out = ds16; a.out_rows = g.T; a.out_shift = g.RF - 1; }
    else { a.out = dskip32; a.out_rows = g.Tout; a.out_shift = 0; a.out_fp32 = 1; }
    return launch<EPI_LRELU_BWD>(Operand{da1, g.A}, Operand{nullptr, 0}, packed + P.wH1T, g.S, g.A, a, g.B, g.Tout, st);
}

// d(skip) bias gradient: the same column sum for every layer (d skip_l = d skip_sum on the last Tout rows, modules.py:90-91)
int mvn_wide_skip_bias_grad(const void* ds16, const Geo& g, float* colsum_ws, float* out_S, cudaStream_t st) {
    return colsum(ds16, (long long)g.B * g.T, g.S, colsum_ws, out_S, st);
}

// backward of one layer: dx_next = d(x_{l+1}) (null for the last layer, whose residual output is discarded), dx_cur = d(x_l)
int mvn_wide_layer_bwd(const void* x_in, const void* dx_next, void* dx_cur, const void* ds16, const void* gab_all, const void* gated_all, void* dz,
                       const float* dbs, const float* packed, float* pg, float* colsum_ws, float* wgpart, const PackedLayout& P, const Geo& g,
                       int l, cudaStream_t st) {
    const int C = g.C, S = g.S, d = g.dil[l], NC = g.N * C;
    const __nv_bfloat16* gated = (const __nv_bfloat16*)gated_all + (size_t)l * C;       // layer l's columns, row stride N C
    const float* lw = packed + P.layer0 + (size_t)l * P.layer_stride;
    float* lg = pg + P.layer0 + (size_t)l * P.layer_stride;
    const long long rows = (long long)g.B * g.T;
    int rc;
    {   // dz = (Wr^T d(x') + Ws^T d(skip)) * (a, b): the d(gated) GEMM with the gate derivative in its epilogue -- the factors
        // (a, b) were kept by the forward's gate GEMM, so the backward runs no recompute GEMM and no tanh / sigmoid at all
        Args a = new_args();
        if (dx_next) seg(a, 0, C, 0);
        seg(a, 1, S, 0);
        a.N = C; a.b_kb0 = dx_next ? 0 : C / BK; a.out = dz; a.ld_out = 2 * C; a.out2_c0 = l * 2 * C;
        if ((rc = launch<EPI_DZ>(Operand{dx_next ? dx_next : ds16, dx_next ? C : S}, Operand{ds16, S}, lw + P.wWrsT, C, C + S, a, g.B, g.T, st,
                                 Output{dz, 2 * C, g.T, 0}, Output{gab_all, 2 * NC, g.T, 0}))) return rc;
    }
    // the residual-bias gradient of layer l - 1 is the column sum of this layer's d(x): taken inside the d(x) GEMM's epilogue
    // from the staged output tiles when a tile holds every column (C <= 256), else by a separate pass over d(x_{l+1})
    const bool fused_colsum = C <= NCH;
    int cs_rows = 0; float* cs_out = nullptr;
    {   // d(x_l)[t] = d(x_{l+1})[t] + W1^T dz[t] + W0^T dz[t + d]
        Args a = new_args();
        seg(a, 0, 2 * C, 0); seg(a, 0, 2 * C, d);
        a.N = C; a.aux = dx_next; a.ld_aux = C; a.out = dx_cur; a.ld_out = C;
        a.csum = (fused_colsum && l > 0) ? colsum_ws : nullptr;
        if ((rc = launch<EPI_ADD_STORE>(Operand{dz, 2 * C}, Operand{nullptr, 0}, lw + P.wWzT, C, 4 * C, a, g.B, g.T, st, Output{dx_cur, C, g.T, 0}))) return rc;
        if (a.csum) {      // per-CTA, per-staging-warp partial rows; reduced by the weight-gradient reduce kernel below (or here)
            const int pair = pair_mode();
            int clusters = mvn_sm_count() / pair; if (clusters > a.n_tiles) clusters = a.n_tiles;
            cs_rows = clusters * pair * 4;
            cs_out = pg + P.layer0 + (size_t)(l - 1) * P.layer_stride + P.obrs;
            if (!wgrad_tc_ok(g)) {
                MVN_CUDA(mvn_launch_pdl(colsum2_kernel, dim3(mvn_cdiv(C, 32)), dim3(32, RED_SPLIT), (size_t)0, st, (const float*)colsum_ws, cs_rows, C, cs_out));
                if ((rc = mvn_check_launch("colsum2"))) return rc;
            }
        }
    }
    // weight gradients (K = time): packed layout oWz[k = tap C + c_in][2 c_out + gate], oWrs[k = c][n]
    if (wgrad_tc_ok(g)) {
        WgArgs w; memset(&w, 0, sizeof(w));
        if (cs_out) { w.cs_partial = colsum_ws; w.cs_out = cs_out; w.cs_rows = cs_rows; w.cs_cols = C; }
        WgTensors t = {{x_in, gated_all, dz, dx_next, ds16}, {C, NC, 2 * C, C, S}};
        for (int tap = 0; tap < 2; ++tap)               // dWz[tap C + c_in][n] = sum_t x[t - (1 - tap) d][c_in] dz[t][n]
            for (int mt = 0; mt < C / 256; ++mt)
                for (int n0 = 0; n0 < 2 * C; n0 += 512) {
                    WgJob& j = w.job[w.n_jobs++];
                    j.a_map = 0; j.a_c0 = 256 * mt; j.a_shift = tap == 0 ? -d : 0;
                    j.n = 2 * C - n0 < 512 ? 2 * C - n0 : 512;
                    j.b_map[0] = j.b_map[1] = 2; j.b_c0[0] = n0; j.b_c0[1] = 0; j.b_n0 = j.n;
                    j.dst = lg + P.oWz + (size_t)(tap * C + 256 * mt) * 2 * C + n0; j.ld = 2 * C;
                }
        const int nx = dx_next ? C : 0, ntot = nx + S;  // d[Wr | Ws][c][n] = sum_t gated[t][c] [d(x') | d(skip)][t][n]
        for (int mt = 0; mt < C / 256; ++mt)
            for (int n0 = 0; n0 < ntot; n0 += 512) {
                WgJob& j = w.job[w.n_jobs++];
                j.a_map = 1; j.a_c0 = l * C + 256 * mt; j.a_shift = 0;
                j.n = ntot - n0 < 512 ? ntot - n0 : 512;
                j.b_map[0] = 3; j.b_map[1] = 4;
                j.b_n0 = nx - n0 < 0 ? 0 : (nx - n0 > j.n ? j.n : nx - n0);
                j.b_c0[0] = n0; j.b_c0[1] = n0 > nx ? n0 - nx : 0;
                j.dst = lg + P.oWrs + (size_t)(256 * mt) * (C + S) + (C - nx) + n0; j.ld = C + S;
            }
        if ((rc = wgrad_launch(w, t, g, wgpart, st))) return rc;
        if (dx_next && !fused_colsum && (rc = colsum(dx_next, rows, C, colsum_ws, lg + P.obrs, st))) return rc;
        MVN_CUDA(cudaMemcpyAsync(lg + P.obrs + C, dbs, (size_t)S * 4, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    const __nv_bfloat16* x = (const __nv_bfloat16*)x_in; const __nv_bfloat16* dzp = (const __nv_bfloat16*)dz;
    for (int b = 0; b < g.B; ++b)       // tap 0 pairs x[t - d] with dz[t]: per clip
        if ((rc = gemm_tn(x + (size_t)b * g.T * C, C, dzp + ((size_t)b * g.T + d) * 2 * C, 2 * C, lg + P.oWz, 2 * C, C, 2 * C, g.T - d, b > 0, st))) return rc;
    if (g.T - d <= 0) MVN_CUDA(cudaMemsetAsync(lg + P.oWz, 0, (size_t)C * 2 * C * 4, st));
    if ((rc = gemm_tn(x, C, dzp, 2 * C, lg + P.oWz + (size_t)C * 2 * C, 2 * C, C, 2 * C, rows, false, st))) return rc;
    if (dx_next) {
        if ((rc = gemm_tn(gated, NC, dx_next, C, lg + P.oWrs, C + S, C, C, rows, false, st))) return rc;
        if (!fused_colsum && (rc = colsum(dx_next, rows, C, colsum_ws, lg + P.obrs, st))) return rc;
    }
    if ((rc = gemm_tn(gated, NC, ds16, S, lg + P.oWrs + C, C + S, C, S, rows, false, st))) return rc;
    MVN_CUDA(cudaMemcpyAsync(lg + P.obrs + C, dbs, (size_t)S * 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// causal input conv weight gradient: dWin[tap][a][c] = sum_t x[a][t - 1 + tap] d(h0)[t][c]   (movenet/modules.py:15-30)
int mvn_wide_input_bwd(const float* audio, const int* codes, const unsigned char* dense, const void* dh0, void* oh16, float* pg,
                       const PackedLayout& P, const Geo& g, cudaStream_t st) {
    int rc;
    MVN_CUDA(mvn_launch_pdl(onehot16_kernel, dim3(4 * mvn_sm_count()), dim3(256), (size_t)0, st, audio, codes, dense, (__nv_bfloat16*)oh16, g.B, g.T, g.A));
    if ((rc = mvn_check_launch("onehot16"))) return rc;
    const __nv_bfloat16* oh = (const __nv_bfloat16*)oh16; const __nv_bfloat16* dh = (const __nv_bfloat16*)dh0;
    float* dwin = pg + P.win;
    for (int b = 0; b < g.B; ++b)
        if ((rc = gemm_tn(oh + (size_t)b * g.T * g.A, g.A, dh + ((size_t)b * g.T + 1) * g.C, g.C, dwin, g.C, g.A, g.C, g.T - 1, b > 0, st))) return rc;
    if (g.T - 1 <= 0) MVN_CUDA(cudaMemsetAsync(dwin, 0, (size_t)g.A * g.C * 4, st));
    return gemm_tn(oh, g.A, dh, g.C, dwin + (size_t)g.A * g.C, g.C, g.A, g.C, (long long)g.B * g.T, false, st);
}
