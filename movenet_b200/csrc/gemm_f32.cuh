// fp32 (CUDA-core FFMA) building blocks of the exact-mode path.
//
//   row_gemm : out[r][n] = epilogue( sum_i sum_k act_i(src_i[row_i(r)][k]) * W_i[k][n] + bias[n] )
//              rows are (clip b, time t) pairs; every operand has its own
//              (T, shift) row map so dilation taps are "the same tensor, t-d"
//              with zeros outside [0,T) -- the reference shortens its tensors
//              instead (movenet/modules.py:36-46); both agree on the columns
//              the reference keeps.
//   tn_gemm  : dW_i[k][n] += sum_r src_i[row_i(r)][k] * q[row_q(r)][n]   (weight gradients)
//
// These are deliberately plain tiled SGEMMs: they are the exact-arithmetic
// mode (token-exact argmax, gradient checks) and the stand-in for ops that do
// not have a tcgen05 kernel yet.  The tensor-core path lives in layer_tc.cu.
#pragma once
#include "common.cuh"

#define MVN_MAX_SRC 4

enum {
    EPI_STORE = 0,        // out = v
    EPI_ACCUM = 1,        // out += v
    EPI_GATE = 2,         // (f,g) column pairs -> out[n/2] = tanh(f)*sigmoid(g)
    EPI_GATE_BWD = 3,     // out[n],out[n+1] = dz_f,dz_g from aux=dgated ; out2[n/2] = gated
    EPI_RESID_SKIP = 4,   // n<split: out = v + aux ; n>=split: out2[n-split] += v
    EPI_ADD_AUX = 5,      // out = v + aux
    EPI_MUL_LRELU_GRAD = 6 // out = v * lrelu'(aux)
};

struct GemmSrc {
    const void* ptr;
    const float* W;   // W[k * ldw + n]
    int dtype, ld, K, T, shift, pre, ldw;
};

struct RowGemmArgs {
    long long rows;
    int Trow, N, nsrc;
    GemmSrc src[MVN_MAX_SRC];
    const float* bias;
    int epi;
    void* out;  int out_dtype, ldo, out_T, out_shift;
    void* out2; int out2_dtype, ldo2, out2_T, out2_shift;
    const void* aux; int aux_dtype, lda, aux_T, aux_shift;
    int split;
    int allow_ksplit; // caller opt-in: few rows, long contraction (needs `ws`)
    int ksplit;      // > 1: the K chunks are dealt round-robin to gridDim.z CTAs, each writes its partial [rows][N] into `ws`;
                     // a second kernel adds the partials in a fixed order (deterministic: no fp32 atomics) (EPI_STORE only)
    float* ws; size_t ws_floats;
};

struct TnSrc {
    const void* ptr;
    float* out;       // out[k * ldo + n]
    int dtype, ld, K, T, shift, pre, ldo;
};

struct TnGemmArgs {
    long long rows, rows_per_cta;
    int Trow, N, nsrc;
    TnSrc src[MVN_MAX_SRC];
    const void* q; int q_dtype, ldq, q_T, q_shift;
    float* dbias;
    // row slices (gridDim.z) write partial products into ws[z][ktot + 1][N] (the last row: the bias sums); a second kernel adds
    // them to the outputs in slice order -- deterministic, no fp32 atomics
    float* ws; size_t ws_floats; int ktot;
};

__device__ __forceinline__ long long mvn_map_row(long long b, int t, int T, int shift) {
    int ts = t + shift;
    return (ts >= 0 && ts < T) ? b * (long long)T + ts : -1;
}

// load 8 consecutive elements of one operand row into v (zero filled)
__device__ __forceinline__ void mvn_load8(const void* ptr, int dtype, long long row, int ld, int k, int K,
                                          int pre, float* v) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (row < 0 || k >= K) return;
    long long base = row * (long long)ld + k;
    if (k + 8 <= K) {
        if (dtype == MVN_F32) {
            const float* p = (const float*)ptr + base;
            if ((((uintptr_t)p) & 15) == 0) {
                float4 a = *(const float4*)p, b = *(const float4*)(p + 4);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = p[j];
            }
        } else {
            const __nv_bfloat16* p = (const __nv_bfloat16*)ptr + base;
            if ((((uintptr_t)p) & 15) == 0) {
                uint4 raw = *(const uint4*)p;
                const __nv_bfloat162* h = (const __nv_bfloat162*)&raw;
#pragma unroll
                for (int j = 0; j < 4; ++j) { float2 f = __bfloat1622float2(h[j]); v[2 * j] = f.x; v[2 * j + 1] = f.y; }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = __bfloat162float(p[j]);
            }
        }
    } else {
        for (int j = 0; j < 8 && k + j < K; ++j) v[j] = mvn_ld(ptr, dtype, base + j);
    }
    if (pre) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = mvn_lrelu(v[j]);
    }
}

__device__ __forceinline__ void mvn_load4w(const float* W, int ldw, int k, int K, int n, int N, float* w) {
    w[0] = w[1] = w[2] = w[3] = 0.f;
    if (k >= K || n >= N) return;
    const float* p = W + (long long)k * ldw + n;
    if (n + 4 <= N && (((uintptr_t)p) & 15) == 0) {
        float4 a = *(const float4*)p; w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    } else {
        for (int j = 0; j < 4 && n + j < N; ++j) w[j] = p[j];
    }
}

#define RG_BM 128
#define RG_BN 64
#define RG_BK 16

__global__ void __launch_bounds__(256) row_gemm_kernel(const RowGemmArgs a) {
    MVN_PDL_PROLOGUE();
    __shared__ __align__(16) float As[RG_BK][RG_BM + 4];
    __shared__ __align__(16) float Ws[RG_BK][RG_BN];
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const long long r0 = (long long)blockIdx.x * RG_BM;
    const int n0 = blockIdx.y * RG_BN;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int lrow = tid >> 1, kh = (tid & 1) * 8;
    const long long lr = r0 + lrow;
    const long long lb = lr / a.Trow;
    const int lt = (int)(lr - lb * a.Trow);
    const int wk = tid >> 4, wn = (tid & 15) * 4;
    int chunk = 0;

    for (int s = 0; s < a.nsrc; ++s) {
        const GemmSrc& S = a.src[s];
        const long long srow = (lr < a.rows) ? mvn_map_row(lb, lt, S.T, S.shift) : -1;
        for (int k0 = 0; k0 < S.K; k0 += RG_BK) {
            if (a.ksplit > 1 && (chunk++ % a.ksplit) != (int)blockIdx.z) continue;
            float v[8], w[4];
            mvn_load8(S.ptr, S.dtype, srow, S.ld, k0 + kh, S.K, S.pre, v);
            mvn_load4w(S.W, S.ldw, k0 + wk, S.K, n0 + wn, a.N, w);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 8; ++j) As[kh + j][lrow] = v[j];
            *(float4*)&Ws[wk][wn] = make_float4(w[0], w[1], w[2], w[3]);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < RG_BK; ++k) {
                float4 a0 = *(const float4*)&As[k][ty * 8], a1 = *(const float4*)&As[k][ty * 8 + 4];
                float4 b = *(const float4*)&Ws[k][tx * 4];
                float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
        }
    }

    const int nb = n0 + tx * 4;
    float bias[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.bias && (a.ksplit <= 1 || blockIdx.z == 0)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (nb + j < a.N) bias[j] = a.bias[nb + j];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long r = r0 + ty * 8 + i;
        if (r >= a.rows) continue;
        const long long b = r / a.Trow;
        const int t = (int)(r - b * a.Trow);
        const long long orow = mvn_map_row(b, t, a.out_T, a.out_shift);
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
        switch (a.epi) {
        case EPI_STORE:
            if (orow >= 0) {
                if (a.ksplit > 1) { for (int j = 0; j < 4; ++j) if (nb + j < a.N) a.ws[((size_t)blockIdx.z * a.rows + orow) * a.N + nb + j] = v[j]; }
                else for (int j = 0; j < 4; ++j) if (nb + j < a.N) mvn_st(a.out, a.out_dtype, orow * a.ldo + nb + j, v[j]);
            }
            break;
        case EPI_ACCUM:
            if (orow >= 0)
                for (int j = 0; j < 4; ++j) if (nb + j < a.N) {
                    long long o = orow * a.ldo + nb + j;
                    mvn_st(a.out, a.out_dtype, o, mvn_ld(a.out, a.out_dtype, o) + v[j]);
                }
            break;
        case EPI_GATE:
            if (orow >= 0)
                for (int j = 0; j < 4; j += 2) if (nb + j + 1 < a.N)
                    mvn_st(a.out, a.out_dtype, orow * a.ldo + ((nb + j) >> 1), tanhf(v[j]) * mvn_sigmoid(v[j + 1]));
            break;
        case EPI_GATE_BWD: {
            const long long arow = mvn_map_row(b, t, a.aux_T, a.aux_shift);
            const long long o2 = mvn_map_row(b, t, a.out2_T, a.out2_shift);
            for (int j = 0; j < 4; j += 2) if (nb + j + 1 < a.N) {
                const int c = (nb + j) >> 1;
                const float th = tanhf(v[j]), sg = mvn_sigmoid(v[j + 1]);
                const float dg = arow >= 0 ? mvn_ld(a.aux, a.aux_dtype, arow * a.lda + c) : 0.f;
                if (orow >= 0) {
                    mvn_st(a.out, a.out_dtype, orow * a.ldo + nb + j, dg * sg * (1.f - th * th));
                    mvn_st(a.out, a.out_dtype, orow * a.ldo + nb + j + 1, dg * th * sg * (1.f - sg));
                }
                if (a.out2 && o2 >= 0) mvn_st(a.out2, a.out2_dtype, o2 * a.ldo2 + c, th * sg);
            }
        } break;
        case EPI_RESID_SKIP: {
            const long long arow = mvn_map_row(b, t, a.aux_T, a.aux_shift);
            const long long o2 = mvn_map_row(b, t, a.out2_T, a.out2_shift);
            for (int j = 0; j < 4; ++j) {
                const int n = nb + j;
                if (n >= a.N) break;
                if (n < a.split) {
                    if (a.out && orow >= 0) {
                        const float x = arow >= 0 ? mvn_ld(a.aux, a.aux_dtype, arow * a.lda + n) : 0.f;
                        mvn_st(a.out, a.out_dtype, orow * a.ldo + n, v[j] + x);
                    }
                } else if (o2 >= 0) {
                    long long o = o2 * a.ldo2 + (n - a.split);
                    mvn_st(a.out2, a.out2_dtype, o, mvn_ld(a.out2, a.out2_dtype, o) + v[j]);
                }
            }
        } break;
        case EPI_ADD_AUX: {
            const long long arow = a.aux ? mvn_map_row(b, t, a.aux_T, a.aux_shift) : -1;
            if (orow >= 0)
                for (int j = 0; j < 4; ++j) if (nb + j < a.N) {
                    const float x = arow >= 0 ? mvn_ld(a.aux, a.aux_dtype, arow * a.lda + nb + j) : 0.f;
                    mvn_st(a.out, a.out_dtype, orow * a.ldo + nb + j, v[j] + x);
                }
        } break;
        case EPI_MUL_LRELU_GRAD: {
            const long long arow = mvn_map_row(b, t, a.aux_T, a.aux_shift);
            if (orow >= 0)
                for (int j = 0; j < 4; ++j) if (nb + j < a.N) {
                    const float pre = arow >= 0 ? mvn_ld(a.aux, a.aux_dtype, arow * a.lda + nb + j) : 0.f;
                    mvn_st(a.out, a.out_dtype, orow * a.ldo + nb + j, v[j] * mvn_lrelu_grad(pre));
                }
        } break;
        }
    }
}

// out[r][n] = sum_z ws[z][r][n], z ascending
__global__ void ksplit_reduce_kernel(const float* __restrict__ ws, int ks, long long rows, int N, float* __restrict__ out, int ldo) {
    MVN_PDL_PROLOGUE();
    const long long n_el = rows * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += (long long)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int z = 0; z < ks; ++z) acc += ws[(size_t)z * n_el + i];
        out[(i / N) * ldo + i % N] = acc;
    }
}

static inline int mvn_row_gemm(RowGemmArgs a, cudaStream_t st) {
    if (a.rows <= 0 || a.N <= 0) return 0;
    dim3 grid(mvn_cdiv(a.rows, RG_BM), mvn_cdiv(a.N, RG_BN));
    // few rows but a long contraction (the small video-upsampler GEMMs): split K so the grid fills the chip
    int ktot = 0;
    for (int s = 0; s < a.nsrc; ++s) ktot += a.src[s].K;
    a.ksplit = 1;
    if (a.allow_ksplit && a.ws && a.epi == EPI_STORE && a.out_dtype == MVN_F32 && a.out_shift == 0 && a.out_T == a.Trow && grid.x * grid.y < 148 && ktot >= 256) {
        int ks = (2 * 148) / (grid.x * grid.y);
        if (ks > ktot / (2 * RG_BK)) ks = ktot / (2 * RG_BK);
        while (ks > 1 && (size_t)ks * a.rows * a.N > a.ws_floats) --ks;
        if (ks > 1) { a.ksplit = ks; grid.z = ks; }
    }
    MVN_CUDA(mvn_launch_pdl(row_gemm_kernel, dim3(grid), dim3(256), (size_t)(0), st, a));
    int rc = mvn_check_launch("row_gemm");
    if (rc || a.ksplit <= 1) return rc;
    const long long n_el = a.rows * a.N;
    MVN_CUDA(mvn_launch_pdl(ksplit_reduce_kernel, dim3(mvn_cdiv(n_el, 256) < 592 ? mvn_cdiv(n_el, 256) : 592), dim3(256), (size_t)(0), st,
                            (const float*)a.ws, a.ksplit, a.rows, a.N, (float*)a.out, a.ldo));
    return mvn_check_launch("ksplit_reduce");
}

#define TN_BK 64
#define TN_BN 64
#define TN_BR 16

__global__ void __launch_bounds__(256) tn_gemm_kernel(const TnGemmArgs a) {
    MVN_PDL_PROLOGUE();
    __shared__ __align__(16) float Ps[TN_BR][TN_BK];
    __shared__ __align__(16) float Qs[TN_BR][TN_BN];
    // which (source, k tile) is this block?
    int s = 0, kt = blockIdx.x;
    while (s < a.nsrc && kt >= (a.src[s].K + TN_BK - 1) / TN_BK) { kt -= (a.src[s].K + TN_BK - 1) / TN_BK; ++s; }
    if (s >= a.nsrc) return;
    const TnSrc& S = a.src[s];
    const int k0 = kt * TN_BK, n0 = blockIdx.y * TN_BN;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int lrow = tid >> 4, lcol = (tid & 15) * 4;
    const long long rbeg = (long long)blockIdx.z * a.rows_per_cta;
    long long rend = rbeg + a.rows_per_cta;
    if (rend > a.rows) rend = a.rows;
    float acc[4][4], qsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const bool do_bias = (a.dbias != nullptr) && blockIdx.x == 0 && ty == 0;

    for (long long rc = rbeg; rc < rend; rc += TN_BR) {
        const long long r = rc + lrow;
        float p[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
        if (r < rend) {
            const long long b = r / a.Trow;
            const int t = (int)(r - b * a.Trow);
            const long long prow = mvn_map_row(b, t, S.T, S.shift);
            const long long qrow = mvn_map_row(b, t, a.q_T, a.q_shift);
            if (prow >= 0)
                for (int j = 0; j < 4; ++j) if (k0 + lcol + j < S.K) {
                    float v = mvn_ld(S.ptr, S.dtype, prow * S.ld + k0 + lcol + j);
                    p[j] = S.pre ? mvn_lrelu(v) : v;
                }
            if (qrow >= 0)
                for (int j = 0; j < 4; ++j) if (n0 + lcol + j < a.N)
                    q[j] = mvn_ld(a.q, a.q_dtype, qrow * a.ldq + n0 + lcol + j);
        }
        __syncthreads();
        *(float4*)&Ps[lrow][lcol] = make_float4(p[0], p[1], p[2], p[3]);
        *(float4*)&Qs[lrow][lcol] = make_float4(q[0], q[1], q[2], q[3]);
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < TN_BR; ++rr) {
            float4 pv = *(const float4*)&Ps[rr][ty * 4];
            float4 qv = *(const float4*)&Qs[rr][tx * 4];
            float pa[4] = {pv.x, pv.y, pv.z, pv.w}, qa[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(pa[i], qa[j], acc[i][j]);
            if (do_bias) {
#pragma unroll
                for (int j = 0; j < 4; ++j) qsum[j] += qa[j];
            }
        }
    }
    int koff = 0;
    for (int i = 0; i < s; ++i) koff += a.src[i].K;
    float* part = a.ws + (size_t)blockIdx.z * (a.ktot + 1) * a.N;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = k0 + ty * 4 + i;
        if (k >= S.K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < a.N) part[(size_t)(koff + k) * a.N + n] = acc[i][j];
        }
    }
    if (do_bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < a.N) part[(size_t)a.ktot * a.N + n] = qsum[j];
        }
    }
}

// out_s[k][n] += sum_z ws[z][koff_s + k][n] ; dbias[n] += sum_z ws[z][ktot][n]   (z ascending: a fixed order)
__global__ void tn_reduce_kernel(const TnGemmArgs a, int nz) {
    MVN_PDL_PROLOGUE();
    const long long n_el = (long long)(a.ktot + 1) * a.N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_el; i += (long long)gridDim.x * blockDim.x) {
        const int kk = (int)(i / a.N), n = (int)(i - (long long)kk * a.N);
        float* dst;
        if (kk == a.ktot) { if (!a.dbias) continue; dst = a.dbias + n; }
        else {
            int s = 0, k = kk;
            while (k >= a.src[s].K) { k -= a.src[s].K; ++s; }
            dst = a.src[s].out + (long long)k * a.src[s].ldo + n;
        }
        float acc = 0.f;
        for (int z = 0; z < nz; ++z) acc += a.ws[(size_t)z * n_el + i];
        *dst += acc;
    }
}

static inline int mvn_tn_gemm(TnGemmArgs a, cudaStream_t st) {
    if (a.rows <= 0 || a.N <= 0) return 0;
    int ktiles = 0;
    for (int s = 0; s < a.nsrc; ++s) ktiles += mvn_cdiv(a.src[s].K, TN_BK);
    const int ntiles = mvn_cdiv(a.N, TN_BN);
    // aim for ~4 waves of 148 SMs x 2 resident CTAs; at least 256 rows per CTA
    long long want = (148LL * 8) / ((long long)ktiles * ntiles);
    if (want < 1) want = 1;
    long long rpc = (a.rows + want - 1) / want;
    if (rpc < 64) rpc = 64;
    rpc = ((rpc + TN_BR - 1) / TN_BR) * TN_BR;
    a.ktot = 0;
    for (int s = 0; s < a.nsrc; ++s) a.ktot += a.src[s].K;
    if (!a.ws) { mvn_set_error("tn_gemm: no workspace for the partial products"); return -1; }
    const size_t per_slice = (size_t)(a.ktot + 1) * a.N;
    while (mvn_cdiv(a.rows, rpc) > 1 && (size_t)mvn_cdiv(a.rows, rpc) * per_slice > a.ws_floats) rpc *= 2;   // fewer, longer row slices
    if ((size_t)mvn_cdiv(a.rows, rpc) * per_slice > a.ws_floats) { mvn_set_error("tn_gemm: workspace too small"); return -1; }
    a.rows_per_cta = rpc;
    dim3 grid(ktiles, ntiles, mvn_cdiv(a.rows, rpc));
    MVN_CUDA(mvn_launch_pdl(tn_gemm_kernel, dim3(grid), dim3(256), (size_t)(0), st, a));
    int rc = mvn_check_launch("tn_gemm");
    if (rc) return rc;
    MVN_CUDA(mvn_launch_pdl(tn_reduce_kernel, dim3(mvn_cdiv((long long)per_slice, 256) < 592 ? mvn_cdiv((long long)per_slice, 256) : 592), dim3(256),
                            (size_t)(0), st, a, (int)grid.z));
    return mvn_check_launch("tn_reduce");
}
