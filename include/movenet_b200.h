/* movenet_b200 -- C ABI of the B200-native WaveNet hot path.
 *
 * This is the drop-in boundary.  The reference (cosmicBboy/movenet) is pure
 * Python and has no FFI of its own: its "plugin interface" for this path is the
 * nn.Module surface of movenet/wavenet.py:50-239 and movenet/modules.py:15-142.
 * Each entry point below names the reference lines whose arithmetic it
 * replaces.  Plain pointers and sizes only: no torch types, no exceptions, no
 * allocation (every buffer is caller-owned device memory; sizes come from
 * mvn_*_bytes).  Every launcher takes the CUDA stream explicitly, returns 0 on
 * success and a non-zero code on failure, with a message in mvn_last_error().
 * There is no CPU fallback anywhere behind this interface.
 *
 * Layouts
 *   audio / probabilities / logits : (B, A, T) fp32, channels-first, exactly the
 *       reference's AudioTensor (movenet/types.py:4) -- API tensors.
 *   video                          : (B, 160, 64, 64, Cin) fp32 (movenet/types.py:5)
 *   internal activations           : time-major (B, T, C), fp32 or bf16 (act_dtype)
 *   parameters                     : the reference's own state_dict tensors, fp32,
 *       passed as a DEVICE array of device pointers in state_dict order
 *       (MVN_PARAM_* indices below); gradients likewise.
 */
#ifndef MOVENET_B200_H
#define MOVENET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVN_DTYPE_F32 0
#define MVN_DTYPE_BF16 1

#define MVN_MAX_AUDIO_FRAMES 160000 /* movenet/wavenet.py:27 */
#define MVN_MAX_VIDEO_FRAMES 160    /* movenet/wavenet.py:28 */
#define MVN_VIDEO_HW 64             /* movenet/wavenet.py:29 */
#define MVN_UPSAMPLE_STRIDE 10      /* movenet/wavenet.py:31 */

/* WaveNet.__init__ arguments (movenet/wavenet.py:75-83) + the batch geometry
 * and the forward() flags (movenet/wavenet.py:158-165). */
typedef struct mvn_shape {
    int layer_size, stack_size;
    int input_channels;     /* A */
    int residual_channels;  /* C */
    int skip_channels;      /* S */
    int context_in_channels;
    int batch;              /* B */
    int frames;             /* T */
    int has_video;          /* 1: video-conditioned (frames must be 160000) */
    int act_dtype;          /* MVN_DTYPE_F32 (exact mode) | MVN_DTYPE_BF16 (tensor-core mode) */
    int remove_last;        /* forward(remove_last=...) */
    int output_logits;      /* 1: raw logits (output_unnormalized=False, sic), 0: softmax probabilities */
    int no_grad;            /* 1: inference-only forward -- what only a backward pass needs is neither computed nor stored
                               (the wide path's gate-derivative factors); mvn_wavenet_backward* then refuses the activations */
} mvn_shape_t;

/* index of each reference parameter in the pointer tables (state_dict order) */
#define MVN_PARAM_VIDEO_CONV_W 0
#define MVN_PARAM_VIDEO_CONV_B 1
#define MVN_PARAM_VT_W(i) (2 + 2 * (i))
#define MVN_PARAM_VT_B(i) (3 + 2 * (i))
#define MVN_PARAM_CAUSAL_W 8
#define MVN_PARAM_LAYER(l, j) (9 + 10 * (l) + (j))
/* j: 0 conv_filter.conv.weight 1 conv_gate.conv.weight 2 context_conv_filter.weight 3 .bias
 *    4 context_conv_gate.weight 5 .bias 6 conv_residual.weight 7 .bias 8 conv_skip.weight 9 .bias */
#define MVN_PARAM_DENSE(n_layers, j) (9 + 10 * (n_layers) + (j)) /* conv1.w conv1.b conv2.w conv2.b */
#define MVN_PARAM_COUNT(n_layers) (13 + 10 * (n_layers))

const char* mvn_last_error(void);
int mvn_version(void);
/* number of kernel launches this library has issued in this process (bench bookkeeping) */
unsigned long long mvn_launch_count(void);

/* which kernel family serves a shape: 0 CUDA-core (fp32 exact mode, or bf16 shapes without a tensor-core kernel),
 * 1 fused tcgen05 layer kernels with shared-memory-resident weights (residual_channels <= 64, bf16),
 * 2 wide-channel weight-streaming tcgen05 GEMMs with fused epilogues (residual/skip channels multiples of 128, bf16) */
int mvn_kernel_path(const mvn_shape_t* s);

/* geometry helpers: WaveNet.receptive_fields (movenet/wavenet.py:125-134) and
 * compute_output_size (movenet/wavenet.py:136-147; returns <1 when the
 * reference raises ValueError). */
int mvn_receptive_fields(int layer_size, int stack_size);
int mvn_output_size(int layer_size, int stack_size, int frames);

/* buffer sizes (bytes) for one (shape) */
size_t mvn_packed_bytes(const mvn_shape_t* s);   /* packed weights; packed grads use the same size */
size_t mvn_acts_bytes(const mvn_shape_t* s);     /* activations kept from forward to backward */
size_t mvn_scratch_bytes(const mvn_shape_t* s);  /* reusable workspace for forward / backward */

/* mu-law companding: torchaudio.functional.mu_law_encoding / mu_law_decoding
 * (call sites movenet/dataset.py:284, movenet/callbacks.py:66-76).  Encoding is
 * a search over A-1 decision thresholds held in the input dtype (built on the
 * host, see movenet_b200/mulaw.py), which is what makes the integer codes
 * bit-identical to the CPU function; inputs outside [-1, 1] and NaN take the
 * direct formula.  dtype: 0 fp32, 2 fp64 input. */
int mvn_mulaw_encode(const void* x, int x_is_f64, const void* thresholds, int n_channels, int64_t* codes,
                     int64_t n, void* stream);
int mvn_mulaw_decode(const int64_t* codes, const float* lut, int n_channels, float* x, int64_t n, void* stream);
/* codes (B,T) int64 -> one-hot (B,A,T) fp32 (movenet/dataset.py:285-288) */
int mvn_one_hot(const int64_t* codes, float* audio, int B, int A, int T, void* stream);

/* re-layout the reference parameters for the kernels (once per optimizer step) */
int mvn_pack_weights(const mvn_shape_t* s, const void* const* param_ptrs_dev, void* packed, void* stream);
/* packed gradients -> reference-layout gradient tensors, all living in ONE flat fp32 buffer (so the
 * data-parallel all-reduce is a single message): offsets_dev[i] is the element offset of parameter i
 * (state_dict order) inside flat_grads, or -1 for a parameter that gets no gradient.  Every gradient is multiplied by
 * `scale` on the way (1 / world_size under data parallelism: the all-reduce that follows is then a plain sum and the average
 * needs no pass of its own, movenet/trainer.py:230-234). */
int mvn_unpack_grads(const mvn_shape_t* s, const void* packed_grads, float* flat_grads,
                     const int64_t* offsets_dev, float scale, void* stream);

/* WaveNet.forward (movenet/wavenet.py:158-191): causal conv (modules.py:15-30),
 * video encoder + upsampler (wavenet.py:149-156), the gated residual stack
 * (modules.py:67-130), the skip sum (wavenet.py:181), the dense head
 * (modules.py:133-142) and the softmax (wavenet.py:191).
 * out: (B, A, T-RF+1-remove_last) fp32. */
int mvn_wavenet_forward(const mvn_shape_t* s, const void* packed, const float* audio, const float* video,
                        void* acts, float* out, void* scratch, void* stream);

/* Integer-code input: instead of the one-hot float tensor the caller may hand over the class indices
 * themselves -- (B, T) int64, 1/(4A) of the bytes.  Call this with the buffer that will be passed as `acts`, then
 * mvn_wavenet_forward / mvn_wavenet_backward with audio == NULL. */
int mvn_codes_input(const mvn_shape_t* s, const int64_t* codes, void* acts, void* stream);

/* autograd of the above for d(loss)/d(out) = dout; fills packed_grads (same
 * layout as the packed weights), to be followed by mvn_unpack_grads. */
int mvn_wavenet_backward(const mvn_shape_t* s, const void* packed, const float* audio, const float* video,
                         const void* acts, const float* out, const float* dout, void* packed_grads,
                         void* scratch, void* stream);

/* mvn_wavenet_backward with the trainer's loss folded in: instead of d(out) it takes the class targets (B, Tn) int64 and
 * d(loss) (1 float on the device) and forms d(out) = d(loss)/(B Tn) * (softmax_c(out) - onehot(target)) -- the backward of
 * F.cross_entropy(forward(...), target) on the PROBABILITIES (movenet/pytorch_lightning_trainer.py:62-65) -- inside the
 * head kernel, so the (B, A, Tn) gradient tensor is never written or read.  Same gradients as mvn_softmax_ce_bwd followed by
 * mvn_wavenet_backward.  Only for shapes where mvn_fused_loss_supported() is non-zero (bf16 mode, probabilities). */
int mvn_fused_loss_supported(const mvn_shape_t* s);
int mvn_wavenet_backward_loss(const mvn_shape_t* s, const void* packed, const float* audio, const float* video,
                              const void* acts, const float* out, const int64_t* target, const float* grad_loss,
                              void* packed_grads, void* scratch, void* stream);

/* The trainer's loss, F.cross_entropy(forward(...), target) (movenet/pytorch_lightning_trainer.py:62-65), i.e. a
 * cross-entropy over the PROBABILITIES (a second softmax), mean over B*T columns, fused:
 *   fwd: loss (1 float) from probs (B,A,T) fp32 and int64 targets (B,T); partials: mvn_softmax_ce_partials floats
 *   bwd: dprobs = grad_loss[0]/(B*T) * (softmax_c(probs) - onehot(target)) */
size_t mvn_softmax_ce_partials(int B, int T);
int mvn_softmax_ce_fwd(const float* probs, const int64_t* target, int B, int A, int T, float* partials, float* loss,
                       void* stream);
int mvn_softmax_ce_bwd(const float* probs, const int64_t* target, const float* grad_loss, int B, int A, int T,
                       float* dprobs, void* stream);

/* AdamW for every parameter tensor in one launch: torch.optim.AdamW's update as built by
 * movenet/pytorch_lightning_trainer.py:128-202 (amsgrad = False), optionally preceded by the global gradient-norm clip of
 * :233-243 (max_grad_norm > 0; coefficient min(1, max/(norm + 1e-6)) like torch.nn.utils.clip_grad_norm_).
 *   segments_dev: array of { float* param; const float* grad; float* exp_avg; float* exp_avg_sq; long long numel }
 *                 (mvn_adamw_segment_bytes() each), chunks_dev: array of { int segment; int first_chunk_of_segment },
 *                 one entry per mvn_adamw_chunk_elems() elements of a segment; both on the device, built once by the host.
 *   bias_correction1/2 = 1 - beta^step; the scalars are doubles because torch forms 1 - beta, lr / bias_correction1, ... in
 *   double before they reach its kernels.  sq_partials: n_chunks floats of scratch (only read/written when clipping),
 *   grad_norm_out: optional 1 float that receives the pre-clip gradient norm. */
size_t mvn_adamw_segment_bytes(void);
int mvn_adamw_chunk_elems(void);
int mvn_adamw_step(const void* segments_dev, const void* chunks_dev, int n_chunks, double lr, double beta1, double beta2,
                   double eps, double weight_decay, double bias_correction1, double bias_correction2, double max_grad_norm,
                   float* sq_partials, float* grad_norm_out, void* stream);

/* stage-level entry points (the same kernels the two calls above launch) */
int mvn_onehot_to_codes(const float* audio, int B, int A, int T, int* codes, unsigned char* dense, void* stream);
int mvn_input_fwd(const mvn_shape_t* s, const void* packed, const float* audio, void* acts, void* stream);
int mvn_video_fwd(const mvn_shape_t* s, const void* packed, const float* video, void* acts, void* stream);
int mvn_layer_fwd(const mvn_shape_t* s, const void* packed, int layer, void* acts, void* scratch, void* stream);
int mvn_head_fwd(const mvn_shape_t* s, const void* packed, void* acts, float* out, void* scratch, void* stream);
/* one layer of the backward pass on the gradient state currently in `scratch` (profiling / roofline timing) */
int mvn_layer_bwd(const mvn_shape_t* s, const void* packed, int layer, const void* acts, void* packed_grads,
                  void* scratch, void* stream);
/* copy an internal activation out as fp32: which = 0 layer input x_l (B,T,C), 1 skip sum (B,Tout,S), 2 upsampled context (B,T,C)
 * -- the context read is the product's WaveNet.upsample_video (movenet/wavenet.py:134-156), the others serve tests */
int mvn_read_activation(const mvn_shape_t* s, const void* acts, int which, int layer, float* dst, void* stream);

/* ---- data-parallel gradient averaging over NVLink peer memory (one node, one process per GPU; csrc/peer.cu) -------------
 * Replaces DistributedDataParallel's NCCL bucket all-reduce (movenet/trainer.py:230-234) by ONE kernel fused in front of the
 * gradient unpack.  Every rank allocates one shareable exchange buffer of stage_bytes + recv_bytes (mvn_peer_layout,
 * mvn_peer_alloc: zero-filled), hands the 64-byte handle to its peers by any means (the Python host uses torch.distributed)
 * and maps theirs (mvn_peer_open).  mvn_peer_reduce_unpack(epoch), epoch = 1, 2, ... counted per exchange buffer and called by
 * every rank once per step, sums the packed gradients of all ranks in rank order IN PLACE (bit-identical on every rank) and
 * leaves `scale` times the sum, unpacked, in `flat_grads`.  peer_base[r] = rank r's exchange buffer as mapped in THIS process
 * (own rank: the allocation itself).  A rank that never calls makes the others trap after ~5 s. */
#define MVN_PEER_MAX 8
int mvn_peer_layout(const mvn_shape_t* s, size_t* stage_bytes, size_t* recv_bytes);
int mvn_peer_alloc(size_t bytes, void** ptr, void* handle64);
int mvn_peer_open(const void* handle64, void** ptr);
int mvn_peer_close(void* ptr);
int mvn_peer_free(void* ptr);
int mvn_peer_reduce_unpack(const mvn_shape_t* s, void* const* peer_base, int rank, int world, unsigned epoch, void* packed_grads,
                           float* flat_grads, const int64_t* offsets_dev, float scale, void* stream);

/* byte offset of an internal activation inside the `acts` buffer: which = 0 layer input x_l (B,T,C) act dtype,
 * 2 upsampled context (B,T,C) act dtype (has_video shapes only; what mvn_decode_steps takes as `ctx`) */
size_t mvn_acts_offset(const mvn_shape_t* s, int which, int layer);

/* cached autoregressive decoding: WaveNet.generate (movenet/wavenet.py:193-239)
 * with per-layer dilation queues instead of a window recompute per sample.
 * mode:
 *   MVN_DECODE_REFERENCE  the reference's function exactly.  generate() evaluates an RF-long window per sample and
 *       CausalConv1d zero-pads the window's left edge (movenet/modules.py:15-30); with stack_size == 1 that edge
 *       reaches the output, so besides the queue update one extra "edge" column per layer is evaluated
 *       per step from rings that hold each layer's inputs back to the window's left edge (decode.cu).  With
 *       stack_size >= 2 the edge never reaches the output and this mode is identical to MVN_DECODE_CAUSAL.
 *   MVN_DECODE_CAUSAL     the true causal model (d-deep rings only).
 * state: mvn_decode_state_bytes; prefill fills the rings from the layer inputs of a forward pass over the prompt
 * (`acts` of mvn_wavenet_forward, or of mvn_input_fwd + mvn_layer_fwd per layer; the prompt is the first
 * prompt_frames <= shape.frames columns, 0 = all of them: with video the pass runs at 160000 frames);
 * steps generates n_new samples starting at absolute position t_start
 * and writes int32 codes (B, n_new) and optionally the logits (B, n_new, A).
 * ctx: the upsampled context (B, frames, C) in act dtype for has_video shapes (acts + mvn_acts_offset(s, 2, 0)):
 * column t-1 conditions the prediction of sample t, as in forward() -- the reference itself cannot run generate()
 * with video, this is the window [i-RF, i) definition of oracle/wavenet_oracle.py.
 * temperature == 0: argmax, ties to the lowest index like torch.argmax;
 * temperature > 0: a draw from softmax(softmax(z)/temperature) (movenet/wavenet.py:227-231)
 * from a counter-based generator keyed by (seed, clip, position). */
#define MVN_DECODE_CAUSAL 0
#define MVN_DECODE_REFERENCE 1
size_t mvn_decode_state_bytes(const mvn_shape_t* s, int mode);
int mvn_decode_prefill(const mvn_shape_t* s, const void* acts, void* state, int mode, int prompt_frames, void* stream);
int mvn_decode_steps(const mvn_shape_t* s, const void* packed, void* state, const void* ctx, int mode, int t_start,
                     int n_new, int* out_codes, float* out_logits, float temperature, unsigned seed,
                     void* stream);

/* Throughput mode of the cached decoder on tensor cores (bf16 operands and queues, fp32 accumulation): 128 clips
 * advance in lock-step as the rows of tcgen05 MMAs, weights resident in shared memory.  Available when
 * mvn_decode_tc_supported(shape) (no video, skip_channels == 8, residual_channels 16 or 32, input_channels <= 128:
 * power-of-two dilations: the receptive-field configuration, experiments/04).  Prefill as mvn_decode_prefill, plus the packed
 * weights: the queues hold the activations minus the residual biases accumulated below each layer (the step kernel keeps a
 * bias-free residual track).  t_start must be past every dilation (a prompt of at least receptive_fields columns).
 * out_codes_t is STEP-major (n_new, B) so a step's tokens are one coalesced store; forced (B, n_new) optionally
 * overrides the chosen token (teacher forcing, used by the parity tests to compare logits step by step). */
int mvn_decode_tc_supported(const mvn_shape_t* s);
size_t mvn_decode_tc_state_bytes(const mvn_shape_t* s);
/* `packed`: the weights of mvn_pack_weights (the queues hold the activations minus the accumulated residual biases). */
int mvn_decode_tc_prefill(const mvn_shape_t* s, const void* packed, const void* acts, void* state, void* stream);
int mvn_decode_tc_steps(const mvn_shape_t* s, const void* packed, void* state, int t_start, int n_new,
                        int* out_codes_t, float* out_logits, const int* forced, float temperature, unsigned seed,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif
