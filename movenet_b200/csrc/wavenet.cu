// WaveNet forward / backward orchestration and the non-GEMM kernels
// (one-hot detection, input gather, softmax <-> channels-first transposes).
#include <stdlib.h>
#include "common.cuh"
#include "gemm_f32.cuh"
#include "layout.h"
#include "layer_tc.h"

// ------------------------------------------------------------------------------------------------
// audio (B,A,T) fp32 -> codes[b][t] = argmax_a (first max, like torch.argmax), dense[b][t] = 1 when
// the column is not an exact one-hot (then the input conv takes the dense path).
// The same codes are the training targets: target[t] = codes[t + RF]
// (movenet/pytorch_lightning_trainer.py:64).
// four columns per thread through 16-byte loads (T % 4 == 0, 16-byte aligned rows)
__global__ void codes4_kernel(const float* __restrict__ audio, int A, int T, int* __restrict__ codes,
                              unsigned char* __restrict__ dense) {
    MVN_PDL_PROLOGUE();
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) * 4, b = blockIdx.y;
    if (t >= T) return;
    const float* p = audio + (size_t)b * A * T + t;
    float best[4]; int arg[4] = {0, 0, 0, 0}, ones[4], other[4];
    {
        const float4 v4 = *(const float4*)p;
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { best[j] = v[j]; ones[j] = (v[j] == 1.f); other[j] = (v[j] != 0.f && v[j] != 1.f); }
    }
#pragma unroll 4
    for (int a = 1; a < A; ++a) {
        const float4 v4 = *(const float4*)(p + (size_t)a * T);
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (v[j] > best[j] || (v[j] != v[j] && best[j] == best[j])) { best[j] = v[j]; arg[j] = a; }   // NaN wins like torch
            ones[j] += (v[j] == 1.f); other[j] += (v[j] != 0.f && v[j] != 1.f);
        }
    }
    *(int4*)(codes + (size_t)b * T + t) = make_int4(arg[0], arg[1], arg[2], arg[3]);
    uchar4 d;
    d.x = (ones[0] == 1 && other[0] == 0) ? 0 : 1; d.y = (ones[1] == 1 && other[1] == 0) ? 0 : 1;
    d.z = (ones[2] == 1 && other[2] == 0) ? 0 : 1; d.w = (ones[3] == 1 && other[3] == 0) ? 0 : 1;
    *(uchar4*)(dense + (size_t)b * T + t) = d;
}

__global__ void codes_kernel(const float* __restrict__ audio, int A, int T, int* __restrict__ codes,
                             unsigned char* __restrict__ dense) {
    MVN_PDL_PROLOGUE();
    const int t = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (t >= T) return;
    const float* p = audio + (size_t)b * A * T + t;
    float best = p[0]; int arg = 0, ones = (best == 1.f), other = (best != 0.f && best != 1.f);
    for (int a = 1; a < A; ++a) {
        const float v = p[(size_t)a * T];
        if (v > best || (v != v && best == best)) { best = v; arg = a; }   // NaN wins like torch
        ones += (v == 1.f); other += (v != 0.f && v != 1.f);
    }
    codes[(size_t)b * T + t] = arg;
    dense[(size_t)b * T + t] = (ones == 1 && other == 0) ? 0 : 1;
}

// CausalConv1d (movenet/modules.py:15-30): h0[t] = W[:,:,0] x[t-1] + W[:,:,1] x[t], x[-1] = 0.
// One-hot columns are a gather of two rows of Win[tap][a][:]; anything else takes the dense sum.
__global__ void input_fwd_kernel(const float* __restrict__ audio, const int* __restrict__ codes,
                                 const unsigned char* __restrict__ dense, const float* __restrict__ win,
                                 void* __restrict__ h0, int adt, int A, int C, int T, long long rows) {
    MVN_PDL_PROLOGUE();
    const int cg = (C + 7) / 8;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= rows * cg) return;
    const long long row = idx / cg;
    const int c0 = (int)(idx % cg) * 8;
    const long long b = row / T; const int t = (int)(row % T);
    const bool vec = (C % 8 == 0);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int tap = 0; tap < 2; ++tap) {
        const int ts = t - 1 + tap;
        if (ts < 0) continue;
        const long long r = b * T + ts;
        const float* wt = win + (size_t)tap * A * C;
        if (!dense[r]) {
            const float* wr = wt + (size_t)codes[r] * C + c0;
            if (vec) {
                const float4 w0 = ((const float4*)wr)[0], w1 = ((const float4*)wr)[1];
                acc[0] += w0.x; acc[1] += w0.y; acc[2] += w0.z; acc[3] += w0.w;
                acc[4] += w1.x; acc[5] += w1.y; acc[6] += w1.z; acc[7] += w1.w;
            } else {
                for (int j = 0; j < 8 && c0 + j < C; ++j) acc[j] += wr[j];
            }
        } else {
            for (int a = 0; a < A; ++a) {
                const float x = audio[((size_t)b * A + a) * T + ts];
                if (x != 0.f) for (int j = 0; j < 8 && c0 + j < C; ++j) acc[j] = fmaf(x, wt[(size_t)a * C + c0 + j], acc[j]);
            }
        }
    }
    if (vec && adt == MVN_BF16) {
        __nv_bfloat162 o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
        *(uint4*)((__nv_bfloat16*)h0 + row * C + c0) = *(const uint4*)o;
    } else if (vec) {
        float4* d = (float4*)((float*)h0 + row * C + c0);
        d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]); d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
        for (int j = 0; j < 8 && c0 + j < C; ++j) mvn_st(h0, adt, row * C + c0 + j, acc[j]);
    }
}

// The same for C % 8 == 0 with the two tap tables Win[tap][a][:] (2 A C floats) staged in shared memory: persistent
// CTAs, 8 channels per thread, C/8 threads per row, so a row is one coalesced 128- or 256-byte store and the gather never
// leaves the SM.  Columns that are not one-hot take the dense sum from the same shared-memory table.
__global__ void __launch_bounds__(256) input_fwd_smem_kernel(const float* __restrict__ audio, const int* __restrict__ codes,
                                                            const unsigned char* __restrict__ dense, const float* __restrict__ win,
                                                            void* __restrict__ h0, int adt, int A, int C, int T, int rows) {
    MVN_PDL_PROLOGUE();
    extern __shared__ float swin[];
    constexpr int CHUNK = 256;                       // rows per step of a CTA: their codes are fetched together (one global latency)
    __shared__ int scode[CHUNK + 1];                 // entry i = row (chunk start - 1 + i); negative: not one-hot (dense path)
    for (int i = threadIdx.x; i < 2 * A * C / 4; i += blockDim.x) ((float4*)swin)[i] = ((const float4*)win)[i];
    const int cg = C / 8, rows_per_pass = blockDim.x / cg;
    const int sub = threadIdx.x / cg, c0 = (threadIdx.x % cg) * 8;
    for (int chunk0 = blockIdx.x * CHUNK; chunk0 < rows; chunk0 += gridDim.x * CHUNK) {
        __syncthreads();                              // the table is in place / the previous chunk's codes are no longer read
        for (int i = threadIdx.x; i < CHUNK + 1; i += blockDim.x) {
            const int rr = chunk0 - 1 + i;
            scode[i] = (rr >= 0 && rr < rows) ? (dense[rr] ? -1 : codes[rr]) : 0;
        }
        __syncthreads();
        for (int local = sub; local < CHUNK && chunk0 + local < rows; local += rows_per_pass) {
            const int row = chunk0 + local, b = row / T, t = row - b * T;
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int tap = 0; tap < 2; ++tap) {
                const int ts = t - 1 + tap;
                if (ts < 0) continue;                 // x[-1] = 0: the first column of a clip has no tap 0
                const float* wt = swin + tap * A * C;
                const int code = scode[local + tap];
                if (code >= 0) {
                    const float* wr = wt + code * C + c0;
                    const float4 w0 = ((const float4*)wr)[0], w1 = ((const float4*)wr)[1];
                    acc[0] += w0.x; acc[1] += w0.y; acc[2] += w0.z; acc[3] += w0.w;
                    acc[4] += w1.x; acc[5] += w1.y; acc[6] += w1.z; acc[7] += w1.w;
                } else {
                    for (int ch = 0; ch < A; ++ch) {
                        const float x = audio[((size_t)b * A + ch) * T + ts];
                        if (x != 0.f) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) acc[j] = fmaf(x, wt[ch * C + c0 + j], acc[j]);
                        }
                    }
                }
            }
            if (adt == MVN_BF16) {
                __nv_bfloat162 o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
                *(uint4*)((__nv_bfloat16*)h0 + (size_t)row * C + c0) = *(const uint4*)o;
            } else {
                float4* d = (float4*)((float*)h0 + (size_t)row * C + c0);
                d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]); d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
            }
        }
    }
}

// Conv3d with a (1,64,64) kernel over 160-frame clips: [B*160 rows] x [4096*Cin] . [4096*Cin x C].  Few rows, long K:
// split K over CTAs (32 rows x all channels x one K slice each); every slice writes its partial product, a second kernel adds
// the slices in a fixed order on top of the bias (deterministic: no fp32 atomics).
#define VC_ROWS 32
#define VC_K 256
__global__ void __launch_bounds__(256) video_conv_kernel(const float* __restrict__ video, const float* __restrict__ wv,
                                                         float* __restrict__ part, int rows, int K, int C) {
    MVN_PDL_PROLOGUE();
    __shared__ float xs[VC_ROWS][VC_K];
    const int r0 = blockIdx.x * VC_ROWS, k0 = blockIdx.y * VC_K;
    for (int i = threadIdx.x; i < VC_ROWS * VC_K; i += blockDim.x) {
        const int rr = i / VC_K, kk = i % VC_K;
        xs[rr][kk] = (r0 + rr < rows && k0 + kk < K) ? video[(size_t)(r0 + rr) * K + k0 + kk] : 0.f;
    }
    __syncthreads();
    const int cpg = blockDim.x / 4;              // 4 thread groups, 8 rows each
    const int grp = threadIdx.x / cpg, cl = threadIdx.x % cpg;
    for (int c = cl; c < C; c += cpg) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const float* w = wv + (size_t)k0 * C + c;
        const int kmax = min(VC_K, K - k0);
#pragma unroll 4
        for (int kk = 0; kk < kmax; ++kk) {
            const float wvv = w[(size_t)kk * C];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(xs[8 * grp + i][kk], wvv, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (r0 + 8 * grp + i < rows) part[((size_t)blockIdx.y * rows + r0 + 8 * grp + i) * C + c] = acc[i];
    }
}
__global__ void video_conv_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bias, float* __restrict__ enc,
                                         int rows, int C, int nk) {
    MVN_PDL_PROLOGUE();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * C) return;
    float acc = bias[i % C];
    for (int k = 0; k < nk; ++k) acc += part[(size_t)k * rows * C + i];
    enc[i] = acc;
}

// causal-conv weight gradient, deterministic: block (tap * A + a, z) owns class a of tap `tap` over slice z of the rows.  It
// scans its rows 256 at a time (coalesced loads of the codes), compacts the rows whose input sample is class a (one-hot
// columns) or has x[a][t] != 0 (dense columns) IN ROW ORDER with a ballot / prefix pass, then adds their d(h0) rows, thread =
// channel.  The z slices are added in order by input_bwd_reduce_kernel.  d(h0)[t] = dh0[t] (+ dh0b[t + shift_b] for a (P, U)
// gradient pair).
__global__ void __launch_bounds__(256) input_bwd_det_kernel(const float* __restrict__ audio, const int* __restrict__ codes,
                                     const unsigned char* __restrict__ dense, const void* __restrict__ dh0,
                                     const void* __restrict__ dh0b, int shift_b, int adt, float* __restrict__ part,
                                     int A, int C, int T, long long rows, int rows_per_z) {
    MVN_PDL_PROLOGUE();
    __shared__ int s_row[256];
    __shared__ float s_x[256];
    __shared__ int s_wcnt[8], s_total;
    const int tap = blockIdx.x / A, a = blockIdx.x - tap * A, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long rbeg = (long long)blockIdx.y * rows_per_z;
    long long rend = rbeg + rows_per_z; if (rend > rows) rend = rows;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};          // channels tid, tid + 256, ... (C <= 1024)
    for (long long chunk = rbeg; chunk < rend; chunk += 256) {
        const long long row = chunk + tid;
        float x = 0.f;
        if (row < rend) {
            const long long b = row / T; const int t = (int)(row - b * T), ts = t - 1 + tap;
            if (ts >= 0) {
                const long long r = b * T + ts;
                if (!dense[r]) x = codes[r] == a ? 1.f : 0.f;
                else x = audio[((size_t)b * A + a) * T + ts];
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, x != 0.f);
        if (lane == 0) s_wcnt[warp] = __popc(m);
        __syncthreads();
        int base = 0;
        for (int w = 0; w < warp; ++w) base += s_wcnt[w];
        if (x != 0.f) { const int pos = base + __popc(m & ((1u << lane) - 1)); s_row[pos] = (int)(row - chunk); s_x[pos] = x; }
        if (tid == 0) { int tot = 0; for (int w = 0; w < 8; ++w) tot += s_wcnt[w]; s_total = tot; }
        __syncthreads();
        const int n = s_total;
        for (int i = 0; i < n; ++i) {
            const long long r2 = chunk + s_row[i];
            const int t2 = (int)(r2 % T);
            const float xv = s_x[i];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = tid + 256 * j;
                if (c < C) {
                    float g = mvn_ld(dh0, adt, r2 * C + c);
                    if (dh0b && t2 + shift_b < T) g += mvn_ld(dh0b, adt, (r2 + shift_b) * C + c);
                    acc[j] = fmaf(xv, g, acc[j]);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = tid + 256 * j;
        if (c < C) part[((size_t)blockIdx.y * 2 * A + blockIdx.x) * C + c] = acc[j];
    }
}
// The same gradient when the whole [2][A][C] table fits in shared memory (every shape but the widest): block z owns a slice
// of the rows and walks it IN ORDER, thread = channel, adding each d(h0) row into the table entry of the row's class (both
// taps) -- no scan per class, no atomics, a fixed order.  part[z] = the block's table.
__global__ void input_bwd_table_kernel(const float* __restrict__ audio, const int* __restrict__ codes,
                                       const unsigned char* __restrict__ dense, const void* __restrict__ dh0,
                                       const void* __restrict__ dh0b, int shift_b, int adt, float* __restrict__ part,
                                       int A, int C, int T, long long rows, int rows_per_z) {
    MVN_PDL_PROLOGUE();
    extern __shared__ float tab[];                 // [2][A][C]
    const int n = 2 * A * C, c = threadIdx.x;
    for (int i = c; i < n; i += blockDim.x) tab[i] = 0.f;
    __syncthreads();
    // (32-bit row arithmetic: B * T < 2^31 is checked by the callers; 64-bit divisions per row were most of this kernel's time)
    const int rbeg = blockIdx.x * rows_per_z;
    int rend = rbeg + rows_per_z; if (rend > (int)rows) rend = (int)rows;
    // (each thread touches only column c of the table: no hazards between threads, program order within a thread).  32 rows per
    // round: their gradients are fetched together (32 loads in flight per thread), their classes staged in shared memory, then
    // applied one after the other
    __shared__ int s_code[33];                     // class of row chunk - 1 + i (tap 0 reads i, tap 1 reads i + 1); -1: dense, -2: none
    int tc = rbeg < rend ? rbeg % T : 0;           // time index of the chunk's first row
    for (int chunk = rbeg; chunk < rend; chunk += 32, tc = (tc + 32) % T) {
        __syncthreads();
        for (int i = c; i < 33; i += blockDim.x) {
            const int rr = chunk - 1 + i;            // the input sample one before / at each row of the chunk
            int code = -2;
            if (rr >= 0 && rr < rend) code = dense[rr] ? -1 : codes[rr];
            s_code[i] = code;
        }
        float gbuf[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int row = chunk + i;
            float g = 0.f;
            if (c < C && row < rend) {
                g = mvn_ld(dh0, adt, (long long)row * C + c);
                int t = tc + i; if (t >= T) t -= T;
                if (dh0b && t + shift_b < T) g += mvn_ld(dh0b, adt, (long long)(row + shift_b) * C + c);
            }
            gbuf[i] = g;
        }
        __syncthreads();
        if (c < C) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int row = chunk + i;
                if (row >= rend) break;
                int t = tc + i; if (t >= T) t -= T;
                const float g = gbuf[i];
#pragma unroll
                for (int tap = 0; tap < 2; ++tap) {
                    const int ts = t - 1 + tap;
                    if (ts < 0) continue;                  // x[-1] = 0: the first column of a clip has no tap 0
                    const int code = s_code[i + tap];
                    if (code >= 0) tab[(tap * A + code) * C + c] += g;
                    else if (code == -1) {
                        const int b = row / T;
                        for (int a = 0; a < A; ++a) {
                            const float x = audio[((size_t)b * A + a) * T + ts];
                            if (x != 0.f) tab[(tap * A + a) * C + c] = fmaf(x, g, tab[(tap * A + a) * C + c]);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int i = c; i < n; i += blockDim.x) part[(size_t)blockIdx.x * n + i] = tab[i];
}
__global__ void input_bwd_reduce_kernel(const float* __restrict__ part, float* __restrict__ dwin, int n, int nz) {
    MVN_PDL_PROLOGUE();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = 0.f;
    int z = 0;
    for (; z + 8 <= nz; z += 8) {               // eight loads in flight, added in slice order
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = part[(size_t)(z + e) * n + i];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc += v[e];
    }
    for (; z < nz; ++z) acc += part[(size_t)z * n + i];
    dwin[i] += acc;
}

// z (B,Tn,A) time-major fp32 -> out (B,A,Tn) channels-first: softmax over A (movenet/wavenet.py:191)
// or a plain transpose when the caller asked for logits.
__global__ void softmax_nct_fwd_kernel(const float* __restrict__ z, float* __restrict__ out, int A, int Tn,
                                       int logits) {
    extern __shared__ float tile[];      // [32][A+1]
    const int b = blockIdx.y, j0 = blockIdx.x * 32, ld = A + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int r = warp; r < 32; r += nw) {
        const int j = j0 + r;
        if (j >= Tn) continue;
        const float* zr = z + ((size_t)b * Tn + j) * A;
        float m = -INFINITY;
        for (int a = lane; a < A; a += 32) { const float v = zr[a]; tile[r * ld + a] = v; m = fmaxf(m, v); }
        if (!logits) {
            for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            float sum = 0.f;
            for (int a = lane; a < A; a += 32) { const float e = expf(tile[r * ld + a] - m); tile[r * ld + a] = e; sum += e; }
            for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float inv = 1.f / sum;
            for (int a = lane; a < A; a += 32) tile[r * ld + a] *= inv;
        }
    }
    __syncthreads();
    const int j = j0 + lane;
    if (j < Tn)
        for (int a = warp; a < A; a += nw) out[((size_t)b * A + a) * Tn + j] = tile[lane * ld + a];
}

// d(logits) time-major from d(out) channels-first: dz = p * (dp - sum_a dp*p)  (or transpose for logits)
__global__ void softmax_nct_bwd_kernel(const float* __restrict__ probs, const float* __restrict__ dout,
                                       float* __restrict__ dz, int A, int Tn, int logits) {
    extern __shared__ float tile[];      // p[32][A+1], dp[32][A+1]
    const int b = blockIdx.y, j0 = blockIdx.x * 32, ld = A + 1;
    float* tp = tile; float* td = tile + 32 * ld;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int j = j0 + lane;
    if (j < Tn)
        for (int a = warp; a < A; a += nw) {
            const size_t o = ((size_t)b * A + a) * Tn + j;
            td[lane * ld + a] = dout[o];
            if (!logits) tp[lane * ld + a] = probs[o];
        }
    __syncthreads();
    for (int r = warp; r < 32; r += nw) {
        const int jj = j0 + r;
        if (jj >= Tn) continue;
        float* dr = dz + ((size_t)b * Tn + jj) * A;
        if (logits) { for (int a = lane; a < A; a += 32) dr[a] = td[r * ld + a]; continue; }
        float dot = 0.f;
        for (int a = lane; a < A; a += 32) dot = fmaf(td[r * ld + a], tp[r * ld + a], dot);
        for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        for (int a = lane; a < A; a += 32) dr[a] = tp[r * ld + a] * (td[r * ld + a] - dot);
    }
}

__global__ void to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2bfloat16(src[i]);
}

__global__ void to_f32_kernel(const void* __restrict__ src, int dtype, float* __restrict__ dst, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = mvn_ld(src, dtype, i);
}

// ------------------------------------------------------------------------------------------------
static GemmSrc make_src(const void* ptr, int dtype, int ld, int K, int T, int shift, int pre, const float* W, int ldw) {
    GemmSrc s; s.ptr = ptr; s.W = W; s.dtype = dtype; s.ld = ld; s.K = K; s.T = T; s.shift = shift; s.pre = pre; s.ldw = ldw;
    return s;
}
static void set_out(RowGemmArgs& a, void* out, int dtype, int ld, int T, int shift) {
    a.out = out; a.out_dtype = dtype; a.ldo = ld; a.out_T = T; a.out_shift = shift;
}
static void set_out2(RowGemmArgs& a, void* out, int dtype, int ld, int T, int shift) {
    a.out2 = out; a.out2_dtype = dtype; a.ldo2 = ld; a.out2_T = T; a.out2_shift = shift;
}
static void set_aux(RowGemmArgs& a, const void* aux, int dtype, int ld, int T, int shift) {
    a.aux = aux; a.aux_dtype = dtype; a.lda = ld; a.aux_T = T; a.aux_shift = shift;
}
static RowGemmArgs new_args(long long rows, int Trow, int N, int epi, const float* bias) {
    RowGemmArgs a; memset(&a, 0, sizeof(a));
    a.rows = rows; a.Trow = Trow; a.N = N; a.epi = epi; a.bias = bias;
    return a;
}
static TnSrc make_tn(const void* ptr, int dtype, int ld, int K, int T, int shift, int pre, float* out, int ldo) {
    TnSrc s; s.ptr = ptr; s.out = out; s.dtype = dtype; s.ld = ld; s.K = K; s.T = T; s.shift = shift; s.pre = pre; s.ldo = ldo;
    return s;
}

struct Ctx {
    Geo g; PackedLayout P; ActsLayout AL; ScratchLayout SL;
    const float* packed; char* acts; char* scratch; cudaStream_t st;
    // backward, tensor-core path: the small fixed-order reductions of the per-CTA partials are deferred (launched on side streams
    // at two points of the pass instead of between the persistent kernels); which ones are pending
    mutable int defer = 0, pend_head = 0, pend_input = 0, pend_up = 0, pend_video = 0;
    float* tc_part(int slot) const { return (float*)(scratch + SL.tc_partial + (size_t)slot * MVN_TC_PARTIAL_SLOT_BYTES); }
    const float* lw(int l) const { return packed + P.layer0 + (size_t)l * P.layer_stride; }
    void* x(int l) const { return acts + AL.x0 + (size_t)l * AL.x_stride; }
    float* det_ws() const { return scratch ? (float*)(scratch + SL.det_ws) : nullptr; }
};
// every split reduction of the exact-mode GEMMs goes through the workspace (partials added in a fixed order: deterministic)
static int tn_gemm(const Ctx& c, TnGemmArgs& t) { t.ws = c.det_ws(); t.ws_floats = MVN_DET_WS_FLOATS; return mvn_tn_gemm(t, c.st); }
static int row_gemm(const Ctx& c, RowGemmArgs& a) { a.ws = c.det_ws(); a.ws_floats = MVN_DET_WS_FLOATS; return mvn_row_gemm(a, c.st); }

static int ctx_init(Ctx& c, const mvn_shape_t* s, const void* packed, const void* acts, const void* scratch, void* stream,
                    const char* who) {
    MVN_REQUIRE(s != nullptr, "%s: null shape", who);
    MVN_REQUIRE(geo_init(c.g, s) == 0, "%s: bad layer_size/stack_size", who);
    MVN_REQUIRE(c.g.A >= 1 && c.g.C >= 1 && c.g.S >= 1 && c.g.B >= 1, "%s: bad channel/batch sizes", who);
    MVN_REQUIRE(c.g.Tout >= 1, "%s: input time steps (%d) must be larger than the receptive fields (%d)", who, c.g.T, c.g.RF);
    MVN_REQUIRE(!c.g.video || c.g.T == MVN_MAX_AUDIO_FRAMES, "%s: video conditioning needs %d frames", who, MVN_MAX_AUDIO_FRAMES);
    MVN_REQUIRE(c.g.adt == MVN_DTYPE_F32 || c.g.adt == MVN_DTYPE_BF16, "%s: bad act_dtype", who);
    MVN_REQUIRE((long long)c.g.B * c.g.T < 0x7fffffffLL, "%s: batch*frames overflows", who);
    packed_layout(c.g, c.P); acts_layout(c.g, c.AL); scratch_layout(c.g, c.SL);
    c.packed = (const float*)packed; c.acts = (char*)acts; c.scratch = (char*)scratch; c.st = (cudaStream_t)stream;
    return 0;
}

extern "C" size_t mvn_packed_bytes(const mvn_shape_t* s) {
    Geo g; if (geo_init(g, s)) return 0; PackedLayout P; packed_layout(g, P); return P.total * 4;
}
extern "C" size_t mvn_acts_bytes(const mvn_shape_t* s) {
    Geo g; if (geo_init(g, s)) return 0; ActsLayout a; acts_layout(g, a); return a.total;
}
extern "C" size_t mvn_scratch_bytes(const mvn_shape_t* s) {
    Geo g; if (geo_init(g, s)) return 0; ScratchLayout w; scratch_layout(g, w); return w.total;
}
extern "C" size_t mvn_acts_offset(const mvn_shape_t* s, int which, int layer) {
    Geo g; if (!s || geo_init(g, s)) return 0; ActsLayout a; acts_layout(g, a);
    if (which == 0) return a.x0 + (size_t)(layer >= 0 && layer < g.N ? layer : 0) * a.x_stride;
    if (which == 2) return a.ctx;
    if (which == 3) return a.w_gab;
    if (which == 4) return a.w_gated;
    return 0;
}
// which kernel family serves this shape: 0 CUDA-core fp32 / generic, 1 tcgen05 fused layer kernels (C = 64 physical),
// 2 wide-channel weight-streaming tcgen05 GEMMs (C >= 128)
extern "C" int mvn_kernel_path(const mvn_shape_t* s) {
    Geo g; if (!s || geo_init(g, s)) return -1;
    if (mvn_wide_supported(g)) return 2;
    return g.adt == MVN_DTYPE_BF16 && mvn_tc_layer_supported(g.C, g.S, g.video) ? 1 : 0;
}
extern "C" int mvn_receptive_fields(int layer_size, int stack_size) {
    mvn_shape_t s; memset(&s, 0, sizeof(s)); s.layer_size = layer_size; s.stack_size = stack_size;
    Geo g; if (geo_init(g, &s)) return -1; return g.RF;
}
extern "C" int mvn_output_size(int layer_size, int stack_size, int frames) {
    const int rf = mvn_receptive_fields(layer_size, stack_size);
    return rf < 0 ? -1 : frames - rf + 1;
}

extern "C" int mvn_onehot_to_codes(const float* audio, int B, int A, int T, int* codes, unsigned char* dense, void* stream) {
    MVN_REQUIRE(audio && codes && dense && B > 0 && A > 0 && T > 0, "mvn_onehot_to_codes: bad arguments");
    if (T % 4 == 0 && (((uintptr_t)audio | (uintptr_t)codes | (uintptr_t)dense) & 15) == 0) {
        dim3 grid4(mvn_cdiv(T / 4, 128), B);
        MVN_CUDA(mvn_launch_pdl(codes4_kernel, grid4, dim3(128), (size_t)(0), (cudaStream_t)stream, audio, A, T, codes, dense));
        return mvn_check_launch("onehot_to_codes4");
    }
    dim3 grid(mvn_cdiv(T, 256), B);
    MVN_CUDA(mvn_launch_pdl(codes_kernel, dim3(grid), dim3(256), (size_t)(0), (cudaStream_t)stream, audio, A, T, codes, dense));
    return mvn_check_launch("onehot_to_codes");
}

// int64 class indices (B,T) -> the int32 codes / "not one-hot" flags the kernels consume
__global__ void codes_from_int64_kernel(const long long* __restrict__ src, int A, long long n, int* __restrict__ codes,
                                        unsigned char* __restrict__ dense) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long v = src[i];
        codes[i] = (v >= 0 && v < A) ? (int)v : 0;
        dense[i] = 0;
    }
}

extern "C" int mvn_codes_input(const mvn_shape_t* s, const int64_t* codes, void* acts, void* stream) {
    Ctx c; int rc = ctx_init(c, s, nullptr, acts, nullptr, stream, "mvn_codes_input"); if (rc) return rc;
    MVN_REQUIRE(codes && acts, "mvn_codes_input: null buffer");
    const long long n = (long long)c.g.B * c.g.T;
    codes_from_int64_kernel<<<mvn_cdiv(n, 256) < 2368 ? mvn_cdiv(n, 256) : 2368, 256, 0, c.st>>>(
        (const long long*)codes, c.g.A, n, (int*)(c.acts + c.AL.codes), (unsigned char*)(c.acts + c.AL.dense));
    return mvn_check_launch("codes_input");
}

static int input_fwd(const Ctx& c, const float* audio) {
    const Geo& g = c.g;
    int* codes = (int*)(c.acts + c.AL.codes); unsigned char* dense = (unsigned char*)(c.acts + c.AL.dense);
    int rc = 0;
    if (audio) rc = mvn_onehot_to_codes(audio, g.B, g.A, g.T, codes, dense, c.st);   // null: mvn_codes_input filled them
    if (rc) return rc;
    const long long rows = (long long)g.B * g.T, n = rows * ((g.C + 7) / 8);
    const size_t table = (size_t)2 * g.A * g.C * 4;
    if (g.C % 8 == 0 && 256 % (g.C / 8) == 0 && table <= 96 * 1024) {
        static MvnSmemAttr attr;
        MVN_CUDA(mvn_ensure_smem(input_fwd_smem_kernel, (int)table, attr));
        const int per_sm = table <= 48 * 1024 ? 4 : 2;
        MVN_CUDA(mvn_launch_pdl(input_fwd_smem_kernel, dim3(148 * per_sm), dim3(256), table, c.st, audio, (const int*)codes,
                                (const unsigned char*)dense, c.packed + c.P.win, c.x(0), g.adt, g.A, g.C, g.T, (int)rows));
        return mvn_check_launch("input_fwd_smem");
    }
    MVN_CUDA(mvn_launch_pdl(input_fwd_kernel, dim3(mvn_cdiv(n, 256)), dim3(256), (size_t)(0), c.st, audio, codes, dense, c.packed + c.P.win, c.x(0), g.adt, g.A, g.C, g.T, rows));
    return mvn_check_launch("input_fwd");
}

static int video_fwd(const Ctx& c, const float* video) {
    const Geo& g = c.g; const int C = g.C;
    float* enc = (float*)(c.acts + c.AL.enc); float* u1 = (float*)(c.acts + c.AL.u1); float* u2 = (float*)(c.acts + c.AL.u2);
    void* ctx = c.acts + c.AL.ctx;
    int rc;
    const bool tc3 = g.adt == MVN_DTYPE_BF16 && mvn_tc_upsample_supported(C);   // all three upsampler levels on tensor cores: u1, u2 kept in bf16
    void* enc16 = (char*)u1 + (size_t)g.B * 1600 * C * 2;       // bf16 copy of the encoder output in the (free) second half of the u1 slot
    {   // Conv3d with a (1,64,64) kernel = one 4096*Cin -> C linear map per frame (movenet/wavenet.py:94-98,152)
        const int rows = g.B * 160, K = 4096 * g.Cin;
        // the K-slice partials live in the (not yet written) second upsampler level's slot of the activation buffer
        float* part = u2;
        if (tc3 && mvn_tc_video_supported(C, K)) {       // tcgen05 (video_tc.cu); writes the bf16 copy too
            MVN_REQUIRE(mvn_tc_video_partial_floats(rows, K) <= (size_t)g.B * 16000 * C, "video encoder: context_in_channels too large for the split-K workspace");
            if ((rc = mvn_tc_video_fwd(video, c.packed + c.P.wv, c.packed + c.P.bv, part, enc, enc16, rows, K, c.st))) return rc;
        } else {
            dim3 grid(mvn_cdiv(rows, VC_ROWS), mvn_cdiv(K, VC_K));
            MVN_REQUIRE((size_t)grid.y * rows * C <= (size_t)g.B * 16000 * C, "video encoder: context_in_channels too large for the split-K workspace");
            MVN_CUDA(mvn_launch_pdl(video_conv_kernel, dim3(grid), dim3(256), (size_t)(0), c.st, video, c.packed + c.P.wv, part, rows, K, C));
            if ((rc = mvn_check_launch("video_conv"))) return rc;
            MVN_CUDA(mvn_launch_pdl(video_conv_reduce_kernel, dim3(mvn_cdiv((long long)rows * C, 256)), dim3(256), (size_t)(0), c.st,
                                    (const float*)part, c.packed + c.P.bv, enc, rows, C, (int)grid.y));
            if ((rc = mvn_check_launch("video_conv_reduce"))) return rc;
            if (tc3) {
                to_bf16_kernel<<<mvn_cdiv((long long)g.B * 160 * C, 256), 256, 0, c.st>>>(enc, (__nv_bfloat16*)enc16, (long long)g.B * 160 * C);
                if ((rc = mvn_check_launch("to_bf16"))) return rc;
            }
        }
    }
    // ConvTranspose1d(k=10, stride=10): out[10 i + j] = W[:, :, j]^T in[i] + b -- a [rows x C] x [C x 10C] GEMM whose
    // row-major output IS the time-major upsampled signal (movenet/wavenet.py:102-118,154)
    const void* in[3] = {enc, u1, u2}; void* out[3] = {u1, u2, ctx};
    const int len[3] = {160, 1600, 16000};
    if (tc3) {
        if ((rc = mvn_tc_upsample_fwd(c.packed + c.P.tc_up01[0], enc16, u1, g.B * len[0], c.st))) return rc;
        if ((rc = mvn_tc_upsample_fwd(c.packed + c.P.tc_up01[1], u1, u2, g.B * len[1], c.st))) return rc;
        return mvn_tc_upsample_fwd(c.packed + c.P.tc_up, u2, ctx, g.B * len[2], c.st);
    }
    for (int i = 0; i < 3; ++i) {
        const int rows = g.B * len[i];
        RowGemmArgs a = new_args(rows, rows, 10 * C, EPI_STORE, c.packed + c.P.bt[i]);
        a.nsrc = 1; a.src[0] = make_src(in[i], MVN_F32, C, C, rows, 0, 0, c.packed + c.P.wt[i], 10 * C);
        set_out(a, out[i], i == 2 ? g.adt : MVN_F32, 10 * C, rows, 0);
        a.allow_ksplit = 1;
        if ((rc = row_gemm(c, a))) return rc;
    }
    return 0;
}

// GatedResidualConv1d.forward (movenet/modules.py:67-93)
static int layer_fwd(const Ctx& c, int l) {
    const Geo& g = c.g; const int C = g.C, S = g.S, d = g.dil[l];
    const long long rows = (long long)g.B * g.T;
    const float* lw = c.lw(l);
    const bool last = (l == g.N - 1);
    float* skip = (float*)(c.acts + c.AL.skip);
    void* ctx = g.video ? c.acts + c.AL.ctx : nullptr;
    if (mvn_wide_supported(g))
        return mvn_wide_layer_fwd(c.x(l), last ? nullptr : c.x(l + 1), c.acts + c.AL.w_gated, g.no_grad ? nullptr : c.acts + c.AL.w_gab,
                                  c.packed, c.P, g, l, c.st);
    if (g.adt == MVN_DTYPE_BF16 && mvn_tc_layer_supported(g.C, g.S, g.video)) {
        return mvn_tc_layer_fwd(c.x(l), ctx, last ? nullptr : c.x(l + 1), skip, lw, c.P, g, l, c.st);
    }
    void* gated = c.scratch + c.SL.gated;
    int rc;
    RowGemmArgs a = new_args(rows, g.T, 2 * C, EPI_GATE, lw + c.P.obz);
    a.nsrc = 2;
    a.src[0] = make_src(c.x(l), g.adt, C, C, g.T, -d, 0, lw + c.P.oWz, 2 * C);
    a.src[1] = make_src(c.x(l), g.adt, C, C, g.T, 0, 0, lw + c.P.oWz + (size_t)C * 2 * C, 2 * C);
    if (g.video) { a.nsrc = 3; a.src[2] = make_src(ctx, g.adt, C, C, g.T, 0, 0, lw + c.P.oWz + (size_t)2 * C * 2 * C, 2 * C); }
    set_out(a, gated, g.adt, C, g.T, 0);
    if ((rc = row_gemm(c, a))) return rc;

    // residual + skip 1x1 convs; the last layer's residual is discarded (movenet/modules.py:125-130)
    RowGemmArgs r = new_args(rows, g.T, last ? S : C + S, EPI_RESID_SKIP, lw + c.P.obrs + (last ? C : 0));
    r.nsrc = 1; r.src[0] = make_src(gated, g.adt, C, C, g.T, 0, 0, lw + c.P.oWrs + (last ? C : 0), C + S);
    r.split = last ? 0 : C;
    if (!last) { set_out(r, c.x(l + 1), g.adt, C, g.T, 0); set_aux(r, c.x(l), g.adt, C, g.T, 0); }
    set_out2(r, skip, MVN_F32, S, g.Tout, -(g.RF - 1));
    return row_gemm(c, r);
}

// DenseConv (movenet/modules.py:133-142) + drop-last + softmax (movenet/wavenet.py:183-191)
static int head_fwd(const Ctx& c, float* out) {
    const Geo& g = c.g;
    if (g.Tn <= 0) return 0;
    const long long rows = (long long)g.B * g.Tn;
    float* skip = (float*)(c.acts + c.AL.skip); float* a1 = (float*)(c.acts + c.AL.a1); float* z = (float*)(c.scratch + c.SL.z);
    if (mvn_wide_supported(g)) {      // the skip 1x1 convs of all layers: one GEMM over the layer-concatenated gated activations
        int rc2 = mvn_wide_skip_fwd(c.acts + c.AL.w_gated, skip, c.packed, c.P, g, c.st);
        if (rc2) return rc2;
    }
    if (mvn_wide_supported(g))
        return mvn_wide_head_fwd(c.packed, c.P, g, skip, 1, a1, out, c.scratch + c.SL.w_l0, c.scratch + c.SL.z, c.st);
    if (g.adt == MVN_DTYPE_BF16 && mvn_tc_head_supported(g.A, g.S))
        return mvn_tc_head_fwd(c.packed, c.P, g, skip, out, c.st);
    if (mvn_wide_head_supported(g))       // e.g. the reference's test architecture: A = 256, C = 64, S = 64
        return mvn_wide_head_fwd(c.packed, c.P, g, skip, 0, a1, out, c.scratch + c.SL.w_l0, c.scratch + c.SL.z, c.st);
    int rc;
    RowGemmArgs a = new_args(rows, g.Tn, g.A, EPI_STORE, c.packed + c.P.b1);
    a.nsrc = 1; a.src[0] = make_src(skip, MVN_F32, g.S, g.S, g.Tout, 0, 1, c.packed + c.P.w1p, g.A);
    set_out(a, a1, MVN_F32, g.A, g.Tn, 0);
    if ((rc = row_gemm(c, a))) return rc;
    RowGemmArgs b = new_args(rows, g.Tn, g.A, EPI_STORE, c.packed + c.P.b2);
    b.nsrc = 1; b.src[0] = make_src(a1, MVN_F32, g.A, g.A, g.Tn, 0, 1, c.packed + c.P.w2p, g.A);
    set_out(b, z, MVN_F32, g.A, g.Tn, 0);
    if ((rc = row_gemm(c, b))) return rc;
    dim3 grid(mvn_cdiv(g.Tn, 32), g.B);
    const size_t smem = (size_t)32 * (g.A + 1) * 4;
    MVN_REQUIRE(smem <= 200 * 1024, "head: input_channels too large (%d)", g.A);
    MVN_CUDA(cudaFuncSetAttribute(softmax_nct_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    softmax_nct_fwd_kernel<<<grid, 256, smem, c.st>>>(z, out, g.A, g.Tn, g.logits);
    return mvn_check_launch("softmax_nct_fwd");
}

extern "C" int mvn_input_fwd(const mvn_shape_t* s, const void* packed, const float* audio, void* acts, void* stream) {
    Ctx c; int rc = ctx_init(c, s, packed, acts, nullptr, stream, "mvn_input_fwd"); if (rc) return rc;
    return input_fwd(c, audio);
}
extern "C" int mvn_video_fwd(const mvn_shape_t* s, const void* packed, const float* video, void* acts, void* stream) {
    Ctx c; int rc = ctx_init(c, s, packed, acts, nullptr, stream, "mvn_video_fwd"); if (rc) return rc;
    MVN_REQUIRE(c.g.video && video, "mvn_video_fwd: shape has no video");
    return video_fwd(c, video);
}
extern "C" int mvn_layer_fwd(const mvn_shape_t* s, const void* packed, int layer, void* acts, void* scratch, void* stream) {
    Ctx c; int rc = ctx_init(c, s, packed, acts, scratch, stream, "mvn_layer_fwd"); if (rc) return rc;
    MVN_REQUIRE(layer >= 0 && layer < c.g.N, "mvn_layer_fwd: layer %d out of range", layer);
    if (layer == 0) MVN_CUDA(cudaMemsetAsync(c.acts + c.AL.skip, 0, (size_t)c.g.B * c.g.Tout * c.g.S * 4, c.st));
    return layer_fwd(c, layer);
}
extern "C" int mvn_head_fwd(const mvn_shape_t* s, const void* packed, void* acts, float* out, void* scratch, void* stream) {
    Ctx c; int rc = ctx_init(c, s, packed, acts, scratch, stream, "mvn_head_fwd"); if (rc) return rc;
    return head_fwd(c, out);
}

extern "C" int mvn_wavenet_forward(const mvn_shape_t* s, const void* packed, const float* audio, const float* video,
                                   void* acts, float* out, void* scratch, void* stream) {
    Ctx c; int rc = ctx_init(c, s, packed, acts, scratch, stream, "mvn_wavenet_forward"); if (rc) return rc;
    MVN_REQUIRE(out && packed && acts && scratch, "mvn_wavenet_forward: null buffer");
    MVN_REQUIRE(!c.g.video || video, "mvn_wavenet_forward: shape says video but video is null");
    if (c.g.video && (rc = video_fwd(c, video))) return rc;
    if ((rc = input_fwd(c, audio))) return rc;
    MVN_CUDA(cudaMemsetAsync(c.acts + c.AL.skip, 0, (size_t)c.g.B * c.g.Tout * c.g.S * 4, c.st));
    for (int l = 0; l < c.g.N; ++l) if ((rc = layer_fwd(c, l))) return rc;
    return head_fwd(c, out);
}

extern "C" int mvn_read_activation(const mvn_shape_t* s, const void* acts, int which, int layer, float* dst, void* stream) {
    Ctx c; int rc = ctx_init(c, s, nullptr, acts, nullptr, stream, "mvn_read_activation"); if (rc) return rc;
    const Geo& g = c.g; const void* src; int dt; long long n;
    if (which == 0) { MVN_REQUIRE(layer >= 0 && layer < g.N, "mvn_read_activation: bad layer"); src = c.x(layer); dt = g.adt; n = (long long)g.B * g.T * g.C; }
    else if (which == 1) { src = c.acts + c.AL.skip; dt = MVN_F32; n = (long long)g.B * g.Tout * g.S; }
    else if (which == 2) { MVN_REQUIRE(g.video, "mvn_read_activation: no context"); src = c.acts + c.AL.ctx; dt = g.adt; n = (long long)g.B * g.T * g.C; }
    else { mvn_set_error("mvn_read_activation: bad selector"); return -1; }
    to_f32_kernel<<<1184, 256, 0, c.st>>>(src, dt, dst, n);
    return mvn_check_launch("read_activation");
}

// ------------------------------------------------------------------------------------------------
// backward
static int head_bwd(const Ctx& c, const float* out, const float* dout, const long long* target, const float* grad_loss, float* pg) {
    const Geo& g = c.g;
    float* skip = (float*)(c.acts + c.AL.skip); float* a1 = (float*)(c.acts + c.AL.a1);
    float* dzh = (float*)(c.scratch + c.SL.z); float* da1 = (float*)(c.scratch + c.SL.da1); float* dskip = (float*)(c.scratch + c.SL.dskip);
    MVN_CUDA(cudaMemsetAsync(dskip, 0, (size_t)g.B * g.Tout * g.S * 4, c.st));
    if (g.Tn <= 0) return 0;
    if (g.adt == MVN_DTYPE_BF16 && mvn_tc_head_supported(g.A, g.S))
    {
        c.pend_head = c.defer;
        return mvn_tc_head_bwd(c.packed, c.P, g, skip, out, dout, target, grad_loss, dskip, pg, c.tc_part(0), c.st, c.defer);
    }
    if (mvn_wide_head_supported(g)) {
        const size_t half = (size_t)g.B * g.Tout * g.A * 2;
        return mvn_wide_head_bwd(c.packed, c.P, g, skip, 0, a1, out, dout, target, grad_loss, c.scratch + c.SL.z, c.scratch + c.SL.da1,
                                 c.scratch + c.SL.w_l0, c.scratch + c.SL.z + half, nullptr, dskip, (float*)(c.scratch + c.SL.w_colsum), pg, c.st);
    }
    MVN_REQUIRE(dout, "the fused loss backward needs the tensor-core head (mvn_fused_loss_supported)");
    const long long rows = (long long)g.B * g.Tn;
    int rc;
    dim3 grid(mvn_cdiv(g.Tn, 32), g.B);
    const size_t smem = (size_t)2 * 32 * (g.A + 1) * 4;
    MVN_REQUIRE(smem <= 200 * 1024, "head: input_channels too large (%d)", g.A);
    MVN_CUDA(cudaFuncSetAttribute(softmax_nct_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    softmax_nct_bwd_kernel<<<grid, 256, smem, c.st>>>(out, dout, dzh, g.A, g.Tn, g.logits);
    if ((rc = mvn_check_launch("softmax_nct_bwd"))) return rc;
    {   // conv2 weight/bias grads: dW2p[k][n] = sum lrelu(a1)[k] dz[n]
        TnGemmArgs t; memset(&t, 0, sizeof(t));
        t.rows = rows; t.Trow = g.Tn; t.N = g.A; t.nsrc = 1;
        t.src[0] = make_tn(a1, MVN_F32, g.A, g.A, g.Tn, 0, 1, pg + c.P.w2p, g.A);
        t.q = dzh; t.q_dtype = MVN_F32; t.ldq = g.A; t.q_T = g.Tn; t.q_shift = 0; t.dbias = pg + c.P.b2;
        if ((rc = tn_gemm(c, t))) return rc;
    }
    {   // da1 = (dz W2) * lrelu'(a1)
        RowGemmArgs a = new_args(rows, g.Tn, g.A, EPI_MUL_LRELU_GRAD, nullptr);
        a.nsrc = 1; a.src[0] = make_src(dzh, MVN_F32, g.A, g.A, g.Tn, 0, 0, c.packed + c.P.w2pT, g.A);
        set_out(a, da1, MVN_F32, g.A, g.Tn, 0); set_aux(a, a1, MVN_F32, g.A, g.Tn, 0);
        if ((rc = row_gemm(c, a))) return rc;
    }
    {   // conv1 grads: dW1p[s][a] = sum lrelu(skip)[s] da1[a]
        TnGemmArgs t; memset(&t, 0, sizeof(t));
        t.rows = rows; t.Trow = g.Tn; t.N = g.A; t.nsrc = 1;
        t.src[0] = make_tn(skip, MVN_F32, g.S, g.S, g.Tout, 0, 1, pg + c.P.w1p, g.A);
        t.q = da1; t.q_dtype = MVN_F32; t.ldq = g.A; t.q_T = g.Tn; t.q_shift = 0; t.dbias = pg + c.P.b1;
        if ((rc = tn_gemm(c, t))) return rc;
    }
    {   // dskip = (da1 W1) * lrelu'(skip_sum); the dropped last column keeps a zero gradient
        RowGemmArgs a = new_args(rows, g.Tn, g.S, EPI_MUL_LRELU_GRAD, nullptr);
        a.nsrc = 1; a.src[0] = make_src(da1, MVN_F32, g.A, g.A, g.Tn, 0, 0, c.packed + c.P.w1pT, g.S);
        set_out(a, dskip, MVN_F32, g.S, g.Tout, 0); set_aux(a, skip, MVN_F32, g.S, g.Tout, 0);
        if ((rc = row_gemm(c, a))) return rc;
    }
    return 0;
}

static int layer_bwd(const Ctx& c, int l, const void* dx_next, void* dx_cur, float* pg) {
    const Geo& g = c.g; const int C = g.C, S = g.S, d = g.dil[l], Kz = g.Kz;
    const long long rows = (long long)g.B * g.T;
    const float* lw = c.lw(l);
    float* lg = pg + c.P.layer0 + (size_t)l * c.P.layer_stride;
    float* dskip = (float*)(c.scratch + c.SL.dskip);
    void* dgated = c.scratch + c.SL.dgated; void* dz = c.scratch + c.SL.dz; void* gated = c.scratch + c.SL.gated;
    void* ctx = g.video ? c.acts + c.AL.ctx : nullptr;
    const int sshift = -(g.RF - 1);
    int rc;
    {   // d(gated) = Wr^T d(residual) + Ws^T d(skip)
        RowGemmArgs a = new_args(rows, g.T, C, EPI_STORE, nullptr);
        int n = 0;
        if (dx_next) a.src[n++] = make_src(dx_next, g.adt, C, C, g.T, 0, 0, lw + c.P.oWrsT, C);
        a.src[n++] = make_src(dskip, MVN_F32, S, S, g.Tout, sshift, 0, lw + c.P.oWrsT + (size_t)C * C, C);
        a.nsrc = n;
        set_out(a, dgated, g.adt, C, g.T, 0);
        if ((rc = row_gemm(c, a))) return rc;
    }
    {   // recompute the pre-activations, then d(pre) = d(gated) * d(tanh*sigmoid)
        RowGemmArgs a = new_args(rows, g.T, 2 * C, EPI_GATE_BWD, lw + c.P.obz);
        a.nsrc = 2;
        a.src[0] = make_src(c.x(l), g.adt, C, C, g.T, -d, 0, lw + c.P.oWz, 2 * C);
        a.src[1] = make_src(c.x(l), g.adt, C, C, g.T, 0, 0, lw + c.P.oWz + (size_t)C * 2 * C, 2 * C);
        if (g.video) { a.nsrc = 3; a.src[2] = make_src(ctx, g.adt, C, C, g.T, 0, 0, lw + c.P.oWz + (size_t)2 * C * 2 * C, 2 * C); }
        set_out(a, dz, g.adt, 2 * C, g.T, 0); set_out2(a, gated, g.adt, C, g.T, 0); set_aux(a, dgated, g.adt, C, g.T, 0);
        if ((rc = row_gemm(c, a))) return rc;
    }
    {   // weight grads of the dilated convs (+ context convs and their biases)
        TnGemmArgs t; memset(&t, 0, sizeof(t));
        t.rows = rows; t.Trow = g.T; t.N = 2 * C; t.nsrc = 2;
        t.src[0] = make_tn(c.x(l), g.adt, C, C, g.T, -d, 0, lg + c.P.oWz, 2 * C);
        t.src[1] = make_tn(c.x(l), g.adt, C, C, g.T, 0, 0, lg + c.P.oWz + (size_t)C * 2 * C, 2 * C);
        if (g.video) { t.nsrc = 3; t.src[2] = make_tn(ctx, g.adt, C, C, g.T, 0, 0, lg + c.P.oWz + (size_t)2 * C * 2 * C, 2 * C); }
        t.q = dz; t.q_dtype = g.adt; t.ldq = 2 * C; t.q_T = g.T; t.q_shift = 0;
        t.dbias = g.video ? lg + c.P.obz : nullptr;
        if ((rc = tn_gemm(c, t))) return rc;
    }
    if (dx_next) {   // residual 1x1 conv grads
        TnGemmArgs t; memset(&t, 0, sizeof(t));
        t.rows = rows; t.Trow = g.T; t.N = C; t.nsrc = 1;
        t.src[0] = make_tn(gated, g.adt, C, C, g.T, 0, 0, lg + c.P.oWrs, C + S);
        t.q = dx_next; t.q_dtype = g.adt; t.ldq = C; t.q_T = g.T; t.q_shift = 0; t.dbias = lg + c.P.obrs;
        if ((rc = tn_gemm(c, t))) return rc;
    }
    {   // skip 1x1 conv grads
        TnGemmArgs t; memset(&t, 0, sizeof(t));
        t.rows = rows; t.Trow = g.T; t.N = S; t.nsrc = 1;
        t.src[0] = make_tn(gated, g.adt, C, C, g.T, 0, 0, lg + c.P.oWrs + C, C + S);
        t.q = dskip; t.q_dtype = MVN_F32; t.ldq = S; t.q_T = g.Tout; t.q_shift = sshift; t.dbias = lg + c.P.obrs + C;
        if ((rc = tn_gemm(c, t))) return rc;
    }
    {   // d(x_l)[t] = d(x_{l+1})[t] + W1^T dz[t] + W0^T dz[t+d]
        RowGemmArgs a = new_args(rows, g.T, C, EPI_ADD_AUX, nullptr);
        a.nsrc = 2;
        a.src[0] = make_src(dz, g.adt, 2 * C, 2 * C, g.T, 0, 0, lw + c.P.oWzT + C, Kz);
        a.src[1] = make_src(dz, g.adt, 2 * C, 2 * C, g.T, d, 0, lw + c.P.oWzT, Kz);
        set_out(a, dx_cur, g.adt, C, g.T, 0);
        if (dx_next) set_aux(a, dx_next, g.adt, C, g.T, 0);
        if ((rc = row_gemm(c, a))) return rc;
    }
    if (g.video) {   // d(ctx) += V^T dz
        RowGemmArgs a = new_args(rows, g.T, C, EPI_ACCUM, nullptr);
        a.nsrc = 1; a.src[0] = make_src(dz, g.adt, 2 * C, 2 * C, g.T, 0, 0, lw + c.P.oWzT + 2 * C, Kz);
        set_out(a, c.scratch + c.SL.dctx, MVN_F32, C, g.T, 0);
        if ((rc = row_gemm(c, a))) return rc;
    }
    return 0;
}

static int input_bwd(const Ctx& c, const float* audio, const void* dh0, const void* dh0b, int shift_b, float* pg) {
    const Geo& g = c.g;
    const long long rows = (long long)g.B * g.T;
    const int n = 2 * g.A * g.C;
    MVN_REQUIRE(g.C <= 1024, "input conv gradient: residual_channels too large (%d)", g.C);
    int nz;
    const size_t table = (size_t)n * 4;
    if (table <= 100 * 1024 && g.C <= 1024) {
        // the table fits in shared memory (twice per SM): one block per row slice, rows applied in order
        nz = 4 * mvn_sm_count();
        while (nz > 1 && (size_t)nz * n > MVN_DET_WS_FLOATS) nz /= 2;
        if (nz > rows) nz = (int)rows;
        const int rpz = (int)((rows + nz - 1) / nz);
        static MvnSmemAttr attr;
        MVN_CUDA(mvn_ensure_smem(input_bwd_table_kernel, (int)table, attr));
        MVN_CUDA(mvn_launch_pdl(input_bwd_table_kernel, dim3(nz), dim3(((g.C + 31) / 32) * 32), table, c.st, audio,
                                (const int*)(c.acts + c.AL.codes), (const unsigned char*)(c.acts + c.AL.dense), dh0, dh0b, shift_b, g.adt,
                                c.det_ws(), g.A, g.C, g.T, rows, rpz));
    } else {
        nz = 64;                                   // row slices: 2 A x 64 blocks scan 1/64 of the rows each
        while (nz > 1 && (size_t)nz * n > MVN_DET_WS_FLOATS) nz /= 2;
        if (nz > rows) nz = (int)rows;
        const int rpz = (int)((rows + nz - 1) / nz);
        MVN_CUDA(mvn_launch_pdl(input_bwd_det_kernel, dim3(2 * g.A, nz), dim3(256), (size_t)(0), c.st, audio,
                                (const int*)(c.acts + c.AL.codes), (const unsigned char*)(c.acts + c.AL.dense), dh0, dh0b, shift_b, g.adt,
                                c.det_ws(), g.A, g.C, g.T, rows, rpz));
    }
    int rc = mvn_check_launch("input_bwd");
    if (rc) return rc;
    MVN_CUDA(mvn_launch_pdl(input_bwd_reduce_kernel, dim3(mvn_cdiv(n, 256)), dim3(256), (size_t)(0), c.st, (const float*)c.det_ws(),
                            pg + c.P.win, n, nz));
    return mvn_check_launch("input_bwd_reduce");
}

static int video_bwd(const Ctx& c, const float* video, const void* dctx, int dctx_dtype, float* pg) {
    const Geo& g = c.g; const int C = g.C;
    const float* enc = (const float*)(c.acts + c.AL.enc); const float* u1 = (const float*)(c.acts + c.AL.u1);
    const float* u2 = (const float*)(c.acts + c.AL.u2);
    float* du2 = (float*)(c.scratch + c.SL.du2);
    float* du1 = (float*)(c.scratch + c.SL.du1); float* denc = (float*)(c.scratch + c.SL.denc);
    const float* in[3] = {enc, u1, u2}; const void* dout[3] = {du1, du2, dctx}; float* din[3] = {denc, du1, du2};
    const int ddt[3] = {MVN_F32, MVN_F32, dctx_dtype};
    const int len[3] = {160, 1600, 16000};
    int rc;
    const bool tc = dctx_dtype == MVN_BF16 && g.adt == MVN_DTYPE_BF16 && mvn_tc_upsample_supported(C);
    for (int i = 2; i >= 0; --i) {
        const int rows = g.B * len[i];
        if (tc) {      // every level on tensor cores: bf16 inputs (enc16, u1, u2) and bf16 gradients between the levels
            const void* in16 = i == 0 ? (const void*)((const char*)u1 + (size_t)g.B * 1600 * C * 2) : (const void*)in[i];
            const float* img = c.packed + (i == 2 ? c.P.tc_up : c.P.tc_up01[i]);
            if ((rc = mvn_tc_upsample_bwd(img, in16, dout[i], din[i], i > 0, pg + c.P.wt[i], pg + c.P.bt[i],
                                          c.tc_part(2 + i), rows, c.st, c.defer))) return rc;
            if (c.defer) c.pend_up |= 1 << i;
            continue;
        }
        TnGemmArgs t; memset(&t, 0, sizeof(t));
        t.rows = rows; t.Trow = rows; t.N = 10 * C; t.nsrc = 1;
        t.src[0] = make_tn(in[i], MVN_F32, C, C, rows, 0, 0, pg + c.P.wt[i], 10 * C);
        t.q = dout[i]; t.q_dtype = ddt[i]; t.ldq = 10 * C; t.q_T = rows; t.q_shift = 0; t.dbias = pg + c.P.bt[i];
        if ((rc = tn_gemm(c, t))) return rc;
        RowGemmArgs a = new_args(rows, rows, C, EPI_STORE, nullptr);
        a.nsrc = 1; a.src[0] = make_src(dout[i], ddt[i], 10 * C, 10 * C, rows, 0, 0, c.packed + c.P.wtT[i], C);
        set_out(a, din[i], MVN_F32, C, rows, 0);
        a.allow_ksplit = 1;
        if ((rc = row_gemm(c, a))) return rc;
    }
    const int rows = g.B * 160, K = 4096 * g.Cin;
    if (tc && mvn_tc_video_supported(C, K) && mvn_tc_video_partial_floats(rows, K) * 4 <= MVN_TC_PARTIAL_SLOT_BYTES) {
        c.pend_video = c.defer;
        return mvn_tc_video_bwd(video, denc, c.tc_part(5), pg + c.P.wv, pg + c.P.bv, rows, K, c.st, c.defer);
    }
    TnGemmArgs t; memset(&t, 0, sizeof(t));
    t.rows = rows; t.Trow = rows; t.N = C; t.nsrc = 1;
    t.src[0] = make_tn(video, MVN_F32, K, K, rows, 0, 0, pg + c.P.wv, C);
    t.q = denc; t.q_dtype = MVN_F32; t.ldq = C; t.q_T = rows; t.q_shift = 0; t.dbias = pg + c.P.bv;
    return tn_gemm(c, t);
}

// one layer of the backward pass on whatever gradient state the scratch buffer holds (profiling / roofline timing)
extern "C" int mvn_layer_bwd(const mvn_shape_t* s, const void* packed, int layer, const void* acts, void* packed_grads,
                             void* scratch, void* stream) {
    Ctx c; int rc = ctx_init(c, s, packed, acts, scratch, stream, "mvn_layer_bwd"); if (rc) return rc;
    const Geo& g = c.g;
    MVN_REQUIRE(layer >= 0 && layer < g.N && packed_grads, "mvn_layer_bwd: bad arguments");
    float* pg = (float*)packed_grads;
    if (mvn_wide_supported(g)) {
        float* cs = (float*)(c.scratch + c.SL.w_colsum);
        return mvn_wide_layer_bwd(c.x(layer), layer + 1 < g.N ? c.scratch + c.SL.dxa : nullptr, c.scratch + c.SL.dxb, c.scratch + c.SL.w_ds16,
                                  c.acts + c.AL.w_gab, c.acts + c.AL.w_gated, c.scratch + c.SL.dz, cs + 512 * 1024, c.packed, pg, cs,
                                  (float*)(c.scratch + c.SL.w_wgpart), c.P, g, layer, c.st);
    }
    if (g.adt == MVN_DTYPE_BF16 && mvn_tc_layer_supported(g.C, g.S, g.video)) {
        const size_t nb = (size_t)g.B * g.T * g.C * g.es;
        float* lg = pg + c.P.layer0 + (size_t)layer * c.P.layer_stride;
        const bool pair_in = layer + 1 < g.N && !mvn_tc_bwd_sum_out(g, layer + 1), sum_out = mvn_tc_bwd_sum_out(g, layer);
        float* part = (float*)(c.scratch + c.SL.tc_layer_partial + (size_t)layer * mvn_tc_bwd_partial_bytes());
        if (mvn_tc_bwd_db_supported(g, layer))
            return mvn_tc_layer_bwd_db(c.x(layer), g.video ? c.acts + c.AL.ctx : nullptr, c.scratch + c.SL.dxa, pair_in ? c.scratch + c.SL.dgated : nullptr,
                                       c.scratch + c.SL.dxb, c.scratch + c.SL.dz, (const float*)(c.scratch + c.SL.dskip), c.scratch + c.SL.dctx,
                                       c.lw(layer), part, c.P, g, layer, c.st);
        return mvn_tc_layer_bwd(c.x(layer), g.video ? c.acts + c.AL.ctx : nullptr, c.scratch + c.SL.dxa, pair_in ? c.scratch + c.SL.dgated : nullptr,
                                c.scratch + c.SL.dxb, sum_out ? nullptr : c.scratch + c.SL.dz, (const float*)(c.scratch + c.SL.dskip),
                                c.scratch + c.SL.dctx, c.scratch + c.SL.dctx + nb, c.lw(layer), lg, part, c.P, g, layer,
                                c.st);
    }
    return layer_bwd(c, layer, layer + 1 < g.N ? c.scratch + c.SL.dxa : nullptr, c.scratch + c.SL.dxb, pg);
}

static int backward_impl(const mvn_shape_t* s, const void* packed, const float* audio, const float* video, const void* acts,
                         const float* out, const float* dout, const long long* target, const float* grad_loss,
                         void* packed_grads, void* scratch, void* stream, const char* who);

extern "C" int mvn_wavenet_backward(const mvn_shape_t* s, const void* packed, const float* audio, const float* video,
                                    const void* acts, const float* out, const float* dout, void* packed_grads,
                                    void* scratch, void* stream) {
    MVN_REQUIRE(dout, "mvn_wavenet_backward: null buffer");
    return backward_impl(s, packed, audio, video, acts, out, dout, nullptr, nullptr, packed_grads, scratch, stream, "mvn_wavenet_backward");
}

extern "C" int mvn_fused_loss_supported(const mvn_shape_t* s) {
    Geo g; if (!s || geo_init(g, s)) return 0;
    return !g.logits && !g.no_grad && g.adt == MVN_DTYPE_BF16 && (mvn_tc_head_supported(g.A, g.S) || mvn_wide_head_supported(g)) && g.Tn > 0;
}

extern "C" int mvn_wavenet_backward_loss(const mvn_shape_t* s, const void* packed, const float* audio, const float* video,
                                         const void* acts, const float* out, const int64_t* target, const float* grad_loss,
                                         void* packed_grads, void* scratch, void* stream) {
    MVN_REQUIRE(target && grad_loss && out, "mvn_wavenet_backward_loss: null buffer");
    MVN_REQUIRE(mvn_fused_loss_supported(s), "mvn_wavenet_backward_loss: not supported for this shape (mvn_fused_loss_supported)");
    return backward_impl(s, packed, audio, video, acts, out, nullptr, (const long long*)target, grad_loss, packed_grads, scratch,
                         stream, "mvn_wavenet_backward_loss");
}

static int backward_impl(const mvn_shape_t* s, const void* packed, const float* audio, const float* video, const void* acts,
                         const float* out, const float* dout, const long long* target, const float* grad_loss,
                         void* packed_grads, void* scratch, void* stream, const char* who) {
    Ctx c; int rc = ctx_init(c, s, packed, acts, scratch, stream, who); if (rc) return rc;
    MVN_REQUIRE(packed && acts && scratch && packed_grads, "%s: null buffer", who);
    MVN_REQUIRE(!c.g.no_grad, "%s: the forward pass ran with shape.no_grad = 1 (inference only): its activations cannot be differentiated", who);
    MVN_REQUIRE(c.g.logits || out, "%s: the probabilities returned by forward are required", who);
    const Geo& g = c.g;
    float* pg = (float*)packed_grads;
    MVN_CUDA(cudaMemsetAsync(pg, 0, c.P.total * 4, c.st));
    if (mvn_wide_supported(g)) {
        // wide-channel path (wide.cu): every GEMM on the weight-streaming tcgen05 engine, weight gradients through cuBLAS
        const size_t half = (size_t)g.B * g.Tout * g.A * 2;
        void* dzh = c.scratch + c.SL.z; void* l1 = c.scratch + c.SL.z + half;
        void* ds16 = c.scratch + c.SL.w_ds16;
        float* cs = (float*)(c.scratch + c.SL.w_colsum); float* dbs = cs + 512 * 1024;
        if ((rc = mvn_wide_head_bwd(c.packed, c.P, g, (const float*)(c.acts + c.AL.skip), 1, (const float*)(c.acts + c.AL.a1), out, dout, target,
                                    grad_loss, dzh, c.scratch + c.SL.da1, c.scratch + c.SL.w_l0, l1, ds16, nullptr, cs, pg, c.st))) return rc;
        if ((rc = mvn_wide_skip_bias_grad(ds16, g, cs, dbs, c.st))) return rc;
        void* bufs[2] = {c.scratch + c.SL.dxa, c.scratch + c.SL.dxb};
        const void* dx_next = nullptr; int cur = 0;
        for (int l = g.N - 1; l >= 0; --l) {
            if ((rc = mvn_wide_layer_bwd(c.x(l), dx_next, bufs[cur], ds16, c.acts + c.AL.w_gab, c.acts + c.AL.w_gated, c.scratch + c.SL.dz,
                                         dbs, c.packed, pg, cs, (float*)(c.scratch + c.SL.w_wgpart), c.P, g, l, c.st))) return rc;
            dx_next = bufs[cur]; cur ^= 1;
        }
        return mvn_wide_input_bwd(audio, (const int*)(c.acts + c.AL.codes), (const unsigned char*)(c.acts + c.AL.dense), dx_next,
                                  c.scratch + c.SL.w_oh16, pg, c.P, g, c.st);
    }
    const bool tc_layers = g.adt == MVN_DTYPE_BF16 && mvn_tc_layer_supported(g.C, g.S, g.video);
    c.defer = tc_layers && mvn_side_stream(0) != nullptr;
    if ((rc = head_bwd(c, out, dout, target, grad_loss, pg))) return rc;
    if (g.video && !tc_layers) MVN_CUDA(cudaMemsetAsync(c.scratch + c.SL.dctx, 0, (size_t)g.B * g.T * g.C * 4, c.st));
    const void* dctx_final = c.scratch + c.SL.dctx; int dctx_dtype = MVN_F32;
    if (tc_layers) {
        // tensor-core path: the stream gradient travels as (P, U), see layer_tc_bwd.cu
        void* Pb[2] = {c.scratch + c.SL.dxa, c.scratch + c.SL.dxb};
        void* Ub[2] = {c.scratch + c.SL.dgated, c.scratch + c.SL.dz};
        const size_t nb = (size_t)g.B * g.T * g.C * g.es;
        // the last layer's residual output is discarded: its incoming (P, U, Q) are zero and are never read (null pointers)
        // running sum of the context gradient, bf16, ping-pong inside the fp32 dctx slot
        void* Qb[2] = {c.scratch + c.SL.dctx, c.scratch + c.SL.dctx + nb};
        // a layer with dilation <= 128 writes ONE summed stream into its P slot (mvn_tc_bwd_sum_out) and no U
        // (the double-buffered kernel of layer_tc_bwd_db.cu -- dilation <= 8 -- adds to the running sum in place, the other one
        // reads one buffer and writes the other)
        int cur = 0, pair = 0, qcur = 0;
        for (int l = g.N - 1; l >= 0; --l) {
            float* lg = pg + c.P.layer0 + (size_t)l * c.P.layer_stride;
            const bool first = l == g.N - 1;
            const int sum_out = mvn_tc_bwd_sum_out(g, l);
            float* part = (float*)(c.scratch + c.SL.tc_layer_partial + (size_t)l * mvn_tc_bwd_partial_bytes());
            const void* ctx = g.video ? c.acts + c.AL.ctx : nullptr;
            if (mvn_tc_bwd_db_supported(g, l)) {
                if ((rc = mvn_tc_layer_bwd_db(c.x(l), ctx, first ? nullptr : Pb[cur], first || !pair ? nullptr : Ub[cur], Pb[cur ^ 1], Ub[cur ^ 1],
                                              (const float*)(c.scratch + c.SL.dskip), Qb[qcur], c.lw(l), part, c.P, g, l, c.st))) return rc;
            } else {
                if ((rc = mvn_tc_layer_bwd(c.x(l), ctx, first ? nullptr : Pb[cur], first || !pair ? nullptr : Ub[cur],
                                           Pb[cur ^ 1], sum_out ? nullptr : Ub[cur ^ 1],
                                           (const float*)(c.scratch + c.SL.dskip), Qb[qcur], Qb[qcur ^ 1], c.lw(l), lg, part, c.P, g, l,
                                           c.st))) return rc;
                qcur ^= 1;
            }
            cur ^= 1; pair = !sum_out;
        }
        // the layers' partials (176 MB at cfg01: a 29 us reduction) and the head's are reduced on a side stream, under the input /
        // upsampler / video kernels that follow
        cudaStream_t red0 = c.defer ? mvn_side_stream(0) : c.st;
        if ((rc = mvn_stream_after(red0, c.st))) return rc;
        if (c.pend_head) { if ((rc = mvn_tc_head_reduce(c.P, g, pg, c.tc_part(0), red0))) return rc; c.pend_head = 0; }
        if ((rc = mvn_tc_bwd_reduce_all((const float*)(c.scratch + c.SL.tc_layer_partial), pg, c.P, g, red0))) return rc;
        if (mvn_tc_input_supported(g.A, g.C)) {
            c.pend_input = c.defer;
            if ((rc = mvn_tc_input_bwd(audio, (const int*)(c.acts + c.AL.codes), (const unsigned char*)(c.acts + c.AL.dense), Pb[cur],
                                       pair ? Ub[cur] : nullptr, pg + c.P.win, c.tc_part(1), g, c.st, c.defer))) return rc;
        } else if ((rc = input_bwd(c, audio, Pb[cur], pair ? Ub[cur] : nullptr, pair ? g.dil[0] : 0, pg))) return rc;
        dctx_final = Qb[qcur]; dctx_dtype = MVN_BF16;
    } else {
        void* bufs[2] = {c.scratch + c.SL.dxa, c.scratch + c.SL.dxb};
        const void* dx_next = nullptr; int cur = 0;
        for (int l = g.N - 1; l >= 0; --l) {
            if ((rc = layer_bwd(c, l, dx_next, bufs[cur], pg))) return rc;
            dx_next = bufs[cur]; cur ^= 1;
        }
        if ((rc = input_bwd(c, audio, dx_next, nullptr, 0, pg))) return rc;
    }
    if (g.video && (rc = video_bwd(c, video, dctx_final, dctx_dtype, pg))) return rc;
    if (c.defer) {
        // every producer has run: the remaining small reductions side by side (three streams), then the caller's stream waits
        cudaStream_t red0 = mvn_side_stream(0), red1 = mvn_side_stream(1), red2 = mvn_side_stream(2);
        if (c.pend_head) { if ((rc = mvn_stream_after(red0, c.st)) || (rc = mvn_tc_head_reduce(c.P, g, pg, c.tc_part(0), red0))) return rc; }
        if (c.pend_input) { if ((rc = mvn_stream_after(red1, c.st)) || (rc = mvn_tc_input_reduce(pg + c.P.win, c.tc_part(1), g, red1))) return rc; }
        if (c.pend_up) {
            const int len[3] = {160, 1600, 16000};
            if ((rc = mvn_stream_after(red2, c.st))) return rc;
            for (int i = 2; i >= 0; --i)
                if ((c.pend_up >> i) & 1)
                    if ((rc = mvn_tc_upsample_reduce(pg + c.P.wt[i], pg + c.P.bt[i], c.tc_part(2 + i), (long long)g.B * len[i], red2))) return rc;
        }
        if (c.pend_video)
            if ((rc = mvn_tc_video_reduce((const float*)(c.scratch + c.SL.denc), c.tc_part(5), pg + c.P.wv, pg + c.P.bv, g.B * 160, 4096 * g.Cin, c.st))) return rc;
        if ((rc = mvn_stream_after(c.st, red0)) || (rc = mvn_stream_after(c.st, red1)) || (rc = mvn_stream_after(c.st, red2))) return rc;
    }
    return 0;
}
