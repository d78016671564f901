"""Cached autoregressive generation (host side of mvn_decode_*).

Restates the contract of ``WaveNet.generate`` (movenet/wavenet.py:193-239):
output is a (B, A, n) one-hot tensor whose first RF columns are the prompt and
whose remaining columns are generated one sample at a time.
"""
import ctypes as C

import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def prefill(model, prompt, video):
    """Run the prompt through the stack once and fill the per-layer dilation queues.

    Returns (shape, buffers, state, ctx): everything mvn_decode_steps needs.
    Decoding always uses the fp32 (exact) kernels, whatever the training dtype is.
    """
    B, A, n_prompt = prompt.shape
    dev = prompt.device
    has_video = video is not None
    # with video the context covers the whole 160000-frame clip; the prompt forward still only
    # needs its own columns, so run it audio-only for the queues and add the context per step
    if has_video:
        raise NotImplementedError("video-conditioned generate() is not wired up yet")
    shape = model._shape(B, n_prompt, False, False, True, _lib.F32)
    bufs = model._buffers_for(shape, dev)
    model._pack(bufs, model._param_list())
    acts = torch.empty(bufs.acts_bytes, dtype=torch.uint8, device=dev)
    logits = torch.empty(B, A, n_prompt - model.receptive_fields + 1, dtype=torch.float32, device=dev)
    _lib.call("mvn_wavenet_forward", C.byref(shape), bufs.packed.data_ptr(), prompt.data_ptr(), 0,
              acts.data_ptr(), logits.data_ptr(), bufs.get_scratch().data_ptr(), _stream())
    state = torch.zeros(_lib.size("mvn_decode_state_bytes", shape), dtype=torch.uint8, device=dev)
    _lib.call("mvn_decode_prefill", C.byref(shape), acts.data_ptr(), state.data_ptr(), _stream())
    return shape, bufs, state, None


def cached_generate(model, audio, video, n_samples, temperature, return_logits=False):
    audio = model._check_audio(audio)
    B, A, T_in = audio.shape
    RF = model.receptive_fields
    n = T_in if n_samples is None else int(n_samples)
    if T_in < RF:
        raise ValueError(f"generate() needs at least receptive_fields={RF} prompt columns, found {T_in}")
    out = torch.zeros(B, A, n, dtype=audio.dtype, device=audio.device)
    keep = min(RF, n)
    out[:, :, :keep] = audio[:, :, :keep]
    n_new = n - RF
    if n_new <= 0:
        return (out, None) if return_logits else out
    with torch.cuda.device(audio.device):
        prompt = audio[:, :, :RF].contiguous()
        shape, bufs, state, ctx = prefill(model, prompt, video)
        codes = torch.empty(B, n_new, dtype=torch.int32, device=audio.device)
        logits = torch.empty(B, n_new, A, dtype=torch.float32, device=audio.device) if return_logits else None
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if temperature > 0 else 0
        _lib.call("mvn_decode_steps", C.byref(shape), bufs.packed.data_ptr(), state.data_ptr(),
                  0 if ctx is None else ctx.data_ptr(), RF, n_new, codes.data_ptr(),
                  0 if logits is None else logits.data_ptr(), C.c_float(float(temperature)), seed, _stream())
        out[:, :, RF:].scatter_(1, codes.long().unsqueeze(1), 1.0)
    return (out, logits) if return_logits else out
