"""Cached autoregressive generation (host side of mvn_decode_*).

Restates the contract of ``WaveNet.generate`` (movenet/wavenet.py:193-239):
output is a (B, A, n) one-hot tensor whose first RF columns are the prompt and
whose remaining columns are generated one sample at a time.

Decoder modes (``WaveNet.decode_mode``):

* ``"exact"``  (default) fp32 CUDA-core decoder computing the REFERENCE's function: for
  ``stack_size >= 2`` the dilation-queue recursion is identical to the reference's window
  recompute; for ``stack_size == 1`` the reference's zero-padded window edge reaches its
  output and the decoder additionally evaluates that edge column per layer
  (MVN_DECODE_REFERENCE, csrc/decode.cu).  Token-exact against ``generate(temperature=0)``.
* ``"causal"`` fp32 decoder of the true causal model (no window edge); differs from
  ``"exact"`` only when ``stack_size == 1``.
* ``"fast"``   tensor-core throughput decoder (bf16 queues / operands, true causal model).
"""
import ctypes as C

import torch

from . import _lib

MODES = ("exact", "causal", "fast")


def _stream():
    return torch.cuda.current_stream().cuda_stream


class DecodeState:
    """queues + bookkeeping of one batch of clips being generated"""

    def __init__(self, shape, bufs, state, ctx, batch, channels, fast=False, mode=_lib.DECODE_REFERENCE, keep=None):
        self.shape, self.bufs, self.state, self.ctx, self.batch, self.channels = shape, bufs, state, ctx, batch, channels
        self.fast = fast
        self.mode = mode
        self.keep = keep        # buffers the context pointer points into


def fast_mode_available(model, batch, n_prompt):
    """tensor-core throughput decoder (bf16 queues/operands): see mvn_decode_tc_supported"""
    shape = model._shape(batch, n_prompt, False, False, True, _lib.F32)
    return bool(_lib.load().mvn_decode_tc_supported(C.byref(shape)))


def _abi_mode(mode):
    if mode not in MODES:
        raise ValueError(f"decode mode must be one of {MODES}, found {mode!r}")
    return _lib.DECODE_CAUSAL if mode == "causal" else _lib.DECODE_REFERENCE


def prefill(model, prompt, video, fast=False, mode="exact"):
    """Run the prompt through the stack once and fill the per-layer rings.

    ``prompt`` is the (B, A, n_prompt) one-hot prompt; ``video`` the optional (B,160,64,64,Cin) conditioning clip.
    fast=True selects the tensor-core throughput decoder (audio-only)."""
    B, A, n_prompt = prompt.shape
    dev = prompt.device
    if video is not None:
        return _prefill_video(model, prompt, video, fast, mode)
    shape = model._shape(B, n_prompt, False, False, True, _lib.F32)
    bufs = model._buffers_for(shape, dev)
    model._pack(bufs, model._param_list())
    acts = torch.empty(bufs.acts_bytes, dtype=torch.uint8, device=dev)
    logits = torch.empty(B, A, n_prompt - model.receptive_fields + 1, dtype=torch.float32, device=dev)
    _lib.call("mvn_wavenet_forward", C.byref(shape), bufs.packed.data_ptr(), prompt.data_ptr(), 0,
              acts.data_ptr(), logits.data_ptr(), bufs.get_scratch().data_ptr(), _stream())
    if fast:
        if not _lib.load().mvn_decode_tc_supported(C.byref(shape)):
            raise _lib.MovenetB200Error("the tensor-core decoder does not support this model shape")
        state = torch.zeros(_lib.size("mvn_decode_tc_state_bytes", shape), dtype=torch.uint8, device=dev)
        _lib.call("mvn_decode_tc_prefill", C.byref(shape), bufs.packed.data_ptr(), acts.data_ptr(), state.data_ptr(), _stream())
        return DecodeState(shape, bufs, state, None, B, A, fast=True)
    m = _abi_mode(mode)
    state = torch.zeros(_lib.size("mvn_decode_state_bytes", shape, m), dtype=torch.uint8, device=dev)
    _lib.call("mvn_decode_prefill", C.byref(shape), acts.data_ptr(), state.data_ptr(), m, 0, _stream())
    return DecodeState(shape, bufs, state, None, B, A, mode=m)


def _prefill_video(model, prompt, video, fast, mode):
    """Video-conditioned prefill.  The reference cannot run generate() with video at all (its upsampled
    context is always 160000 frames long while the window is RF long); the definition implemented here is the oracle's
    (oracle/wavenet_oracle.py generate(context=...)): context column t-1 conditions the prediction of sample t, exactly
    as in forward().  The upsampler only exists at the full clip length, so the stack runs once over a 160000-frame
    pass whose first n_prompt columns are the prompt (causality: the padding after it cannot reach them)."""
    from .wavenet import MAX_AUDIO_FRAMES
    if fast:
        raise _lib.MovenetB200Error("the tensor-core decoder has no video conditioning; use decode_mode 'exact'")
    B, A, n_prompt = prompt.shape
    dev = prompt.device
    video = model._check_video(video)
    assert video.shape[0] == B, "expected video and audio tensors to have equal batch sizes"
    shape = model._shape(B, MAX_AUDIO_FRAMES, True, False, True, _lib.F32)
    bufs = model._buffers_for(shape, dev)
    model._pack(bufs, model._param_list())
    acts = torch.empty(bufs.acts_bytes, dtype=torch.uint8, device=dev)
    codes = torch.zeros(B, MAX_AUDIO_FRAMES, dtype=torch.int64, device=dev)
    codes[:, :n_prompt] = prompt.argmax(1)
    st = _stream()
    _lib.call("mvn_codes_input", C.byref(shape), codes.data_ptr(), acts.data_ptr(), st)
    _lib.call("mvn_video_fwd", C.byref(shape), bufs.packed.data_ptr(), video.data_ptr(), acts.data_ptr(), st)
    _lib.call("mvn_input_fwd", C.byref(shape), bufs.packed.data_ptr(), 0, acts.data_ptr(), st)
    scratch = bufs.get_scratch()
    for l in range(model.layer_size * model.stack_size):
        _lib.call("mvn_layer_fwd", C.byref(shape), bufs.packed.data_ptr(), l, acts.data_ptr(), scratch.data_ptr(), st)
    m = _abi_mode(mode)
    state = torch.zeros(_lib.size("mvn_decode_state_bytes", shape, m), dtype=torch.uint8, device=dev)
    _lib.call("mvn_decode_prefill", C.byref(shape), acts.data_ptr(), state.data_ptr(), m, n_prompt, st)
    ctx_ptr = acts.data_ptr() + _lib.size("mvn_acts_offset", shape, 2, 0)
    return DecodeState(shape, bufs, state, ctx_ptr, B, A, mode=m, keep=acts)


def run_steps(model, st, t_start, n_new, temperature=0.0, return_logits=False, forced=None):
    """generate n_new samples for every clip, starting at absolute position t_start; int32 codes (B, n_new)"""
    dev = st.state.device
    logits = torch.empty(st.batch, n_new, st.channels, dtype=torch.float32, device=dev) if return_logits else None
    seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if temperature > 0 else 0
    if st.fast:
        codes_t = torch.empty(n_new, st.batch, dtype=torch.int32, device=dev)
        forced = None if forced is None else forced.to(torch.int32).contiguous()
        _lib.call("mvn_decode_tc_steps", C.byref(st.shape), st.bufs.packed.data_ptr(), st.state.data_ptr(), t_start, n_new,
                  codes_t.data_ptr(), 0 if logits is None else logits.data_ptr(), 0 if forced is None else forced.data_ptr(),
                  C.c_float(float(temperature)), seed, _stream())
        codes = codes_t.t()
        return (codes, logits) if return_logits else codes
    if forced is not None:
        raise ValueError("teacher forcing is only wired into the tensor-core decoder")
    codes = torch.empty(st.batch, n_new, dtype=torch.int32, device=dev)
    _lib.call("mvn_decode_steps", C.byref(st.shape), st.bufs.packed.data_ptr(), st.state.data_ptr(),
              0 if st.ctx is None else st.ctx, st.mode, t_start, n_new, codes.data_ptr(),
              0 if logits is None else logits.data_ptr(), C.c_float(float(temperature)), seed, _stream())
    return (codes, logits) if return_logits else codes


def cached_generate(model, audio, video, n_samples, temperature, return_logits=False, fast=False, mode="exact"):
    audio = model._check_audio(audio)
    B, A, T_in = audio.shape
    RF = model.receptive_fields
    n = T_in if n_samples is None else int(n_samples)
    if T_in < RF:
        raise ValueError(f"generate() needs at least receptive_fields={RF} prompt columns, found {T_in}")
    if video is not None and n > 160000:
        raise ValueError("video-conditioned generate() cannot run past the 160000 context frames")
    out = torch.zeros(B, A, n, dtype=audio.dtype, device=audio.device)
    keep = min(RF, n)
    out[:, :, :keep] = audio[:, :, :keep]
    n_new = n - RF
    if n_new <= 0:
        return (out, None) if return_logits else out
    with torch.cuda.device(audio.device):
        st = prefill(model, audio[:, :, :RF].contiguous(), video, fast=fast, mode=mode)
        logits = steps_into(model, st, RF, n_new, out[:, :, RF:], temperature, return_logits)
    return (out, logits) if return_logits else out


def steps_into(model, st, t_start, n_new, out_columns, temperature=0.0, return_logits=False):
    """generate n_new samples per clip and write them as one-hot columns into ``out_columns`` (B, A, n_new), which must
    be zero on entry: the reference's output format (movenet/wavenet.py:211-236), one 1.0 per generated sample"""
    res = run_steps(model, st, t_start, n_new, temperature, return_logits)
    codes, logits = res if return_logits else (res, None)
    out_columns.scatter_(1, codes.long().unsqueeze(1), 1.0)
    return logits
