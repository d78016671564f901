// tcgen05 / TMEM / TMA (bf16, fp32 accumulate) fused layer kernels -- see layer_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include "layout.h"

// 1 when the tensor-core layer kernel covers this (C, S, video) combination
int mvn_tc_layer_supported(int C, int S, int video);
// GatedResidualConv1d.forward (movenet/modules.py:67-93) as one fused kernel.
// x_out may be null for the last layer (its residual output is discarded).
int mvn_tc_layer_fwd(const void* x_in, const void* ctx, void* x_out, float* skip_sum, const float* layer_weights,
                     const PackedLayout& P, const Geo& g, int layer, cudaStream_t st);
// build the per-layer shared-memory weight images inside the packed buffer (called by mvn_pack_weights)
int mvn_tc_pack(const float* const* param_ptrs_dev, float* packed, const PackedLayout& P, const Geo& g, cudaStream_t st);
// Backward of one layer on tensor cores (layer_tc_bwd.cu).  The residual-stream gradient travels as the pair
// (P, U): d(x_{l+1})[t] = P[t] + U[t + dilation_{l+1}].  lg = this layer's slot of the packed gradients.
// q_in / q_out: running sum over layers of the context gradient, bf16 (B,T,C) (video only).
// u_in == null (p_in given): the incoming gradient is ONE summed stream; u_out == null: write the summed stream (dilation <= 128).
// mvn_tc_bwd_sum_out(g, l): does layer l write the summed stream?  (decided from the top layer down, layer_tc_bwd.cu)
int mvn_tc_bwd_sum_out(const Geo& g, int layer);
size_t mvn_tc_bwd_partial_bytes();          // per layer; the layers' partial slots are consecutive
// reduce every layer's per-CTA partials into the packed gradients (one launch, after the backward sweep)
int mvn_tc_bwd_reduce_all(const float* partial_all, float* pg, const PackedLayout& P, const Geo& g, cudaStream_t st);
int mvn_tc_layer_bwd(const void* x_in, const void* ctx, const void* p_in, const void* u_in, void* p_out, void* u_out,
                     const float* dskip, const void* q_in, void* q_out, const float* lw, float* lg, float* partial, const PackedLayout& P,
                     const Geo& g, int layer, cudaStream_t st);
// The same backward with a second x / ctx buffer (layer_tc_bwd_db.cu): layers with dilation <= 8, (P, U) pair output.  The
// context-gradient running sum is updated IN PLACE in q_sum (bf16 TMA add-reduction; the top layer -- p_in == null -- stores).
int mvn_tc_bwd_db_supported(const Geo& g, int layer);
int mvn_tc_layer_bwd_db(const void* x_in, const void* ctx, const void* p_in, const void* u_in, void* p_out, void* u_out,
                        const float* dskip, void* q_sum, const float* lw, float* partial, const PackedLayout& P, const Geo& g, int layer,
                        cudaStream_t st);
// DenseConv head + softmax on tensor cores (head_tc.cu), input_channels == 64 or 128
int mvn_tc_head_supported(int A, int S);
size_t mvn_tc_head_partial_bytes();
int mvn_tc_head_pack(const float* w2_ref, float* packed, const PackedLayout& P, int A, cudaStream_t st);
int mvn_tc_head_fwd(const float* packed, const PackedLayout& P, const Geo& g, const float* skip, float* out, cudaStream_t st);
int mvn_tc_head_bwd(const float* packed, const PackedLayout& P, const Geo& g, const float* skip, const float* probs,
                    const float* dout, const long long* target, const float* grad_loss, float* dskip, float* pg, float* partial,
                    cudaStream_t st, int defer_reduce);
// (defer_reduce: only the kernel that writes the per-CTA partials is launched; the caller launches mvn_tc_*_reduce later, possibly
// on a side stream, so that the small fixed-order reduction does not sit between two persistent kernels)
int mvn_tc_head_reduce(const PackedLayout& P, const Geo& g, float* pg, const float* partial, cudaStream_t st);
// causal input conv weight gradient on tensor cores (input_tc.cu), A <= 64, C == 64, gradient given as (P, U)
int mvn_tc_input_supported(int A, int C);
int mvn_tc_input_bwd(const float* audio, const int* codes, const unsigned char* dense, const void* p, const void* u,
                     float* dwin, float* partial, const Geo& g, cudaStream_t st, int defer_reduce);
int mvn_tc_input_reduce(float* dwin, const float* partial, const Geo& g, cudaStream_t st);
// last level of the video upsampler on tensor cores (upsample_tc.cu), C == 64
int mvn_tc_upsample_supported(int C);
size_t mvn_tc_upsample_img_floats();
int mvn_tc_upsample_pack(const float* wt, const float* bt, float* img, cudaStream_t st);
int mvn_tc_upsample_fwd(const float* img, const void* u2_bf16, void* ctx_bf16, long long rows, cudaStream_t st);
int mvn_tc_upsample_bwd(const float* img, const void* u_bf16, const void* dout_bf16, void* du, int du_bf16, float* dwt, float* dbt,
                        float* partial, long long rows, cudaStream_t st, int defer_reduce);
int mvn_tc_upsample_reduce(float* dwt, float* dbt, const float* partial, long long rows, cudaStream_t st);

// video encoder (Conv3d = one 4096 Cin -> C linear map per frame) and its weight gradient on tensor cores (video_tc.cu), C == 64
int mvn_tc_video_supported(int C, int K);
size_t mvn_tc_video_partial_floats(int rows, int K);
int mvn_tc_video_fwd(const float* video, const float* wv, const float* bv, float* part, float* enc, void* enc16_or_null, int rows, int K,
                     cudaStream_t st);
int mvn_tc_video_bwd(const float* video, const float* denc, float* part, float* dwv, float* dbv, int rows, int K, cudaStream_t st,
                     int defer_reduce);
int mvn_tc_video_reduce(const float* denc, const float* part, float* dwv, float* dbv, int rows, int K, cudaStream_t st);

// ---- wide-channel path (wide.cu): weight-streaming tcgen05 GEMMs with fused epilogues, residual_channels >= 128 -------------
int mvn_wide_supported(const Geo& g);
int mvn_wide_pack(const float* const* param_ptrs_dev, float* packed, const PackedLayout& P, const Geo& g, cudaStream_t st);
int mvn_wide_layer_fwd(const void* x_in, void* x_out, void* gated_all, void* gab_all, const float* packed, const PackedLayout& P,
                       const Geo& g, int l, cudaStream_t st);
int mvn_wide_skip_fwd(const void* gated_all, float* skip, const float* packed, const PackedLayout& P, const Geo& g, cudaStream_t st);
int mvn_wide_head_supported(const Geo& g);     // the wide engine's head kernels (also for C <= 64 models with A = 128 / 256, S % 64 == 0)
int mvn_wide_head_fwd(const float* packed, const PackedLayout& P, const Geo& g, const float* skip, int skip_on_T, float* a1, float* out,
                      void* l0, void* l1, cudaStream_t st);
int mvn_wide_head_bwd(const float* packed, const PackedLayout& P, const Geo& g, const float* skip, int skip_on_T, const float* a1,
                      const float* probs, const float* dout, const long long* target, const float* grad_loss, void* dzh, void* da1,
                      void* l0, void* l1, void* ds16, float* dskip32, float* colsum_ws, float* pg, cudaStream_t st);
int mvn_wide_skip_bias_grad(const void* ds16, const Geo& g, float* colsum_ws, float* out_S, cudaStream_t st);
int mvn_wide_layer_bwd(const void* x_in, const void* dx_next, void* dx_cur, const void* ds16, const void* gab_all, const void* gated_all, void* dz,
                       const float* dbs, const float* packed, float* pg, float* colsum_ws, float* wgpart, const PackedLayout& P, const Geo& g,
                       int l, cudaStream_t st);
int mvn_wide_input_bwd(const float* audio, const int* codes, const unsigned char* dense, const void* dh0, void* oh16, float* pg,
                       const PackedLayout& P, const Geo& g, cudaStream_t st);
