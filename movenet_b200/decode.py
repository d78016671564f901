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


class DecodeState:
    """queues + bookkeeping of one batch of clips being generated"""

    def __init__(self, shape, bufs, state, ctx, batch, channels, fast=False):
        self.shape, self.bufs, self.state, self.ctx, self.batch, self.channels = shape, bufs, state, ctx, batch, channels
        self.fast = fast


def fast_mode_available(model, batch, n_prompt):
    """tensor-core throughput decoder (bf16 queues/operands): see mvn_decode_tc_supported"""
    shape = model._shape(batch, n_prompt, False, False, True, _lib.F32)
    return bool(_lib.load().mvn_decode_tc_supported(C.byref(shape)))


def prefill(model, prompt, video, fast=False):
    """Run the prompt through the stack once and fill the per-layer dilation queues.

    fast=False: the fp32 (exact) decoder, token-exact against the reference (the default of generate()).
    fast=True : the tensor-core throughput decoder (bf16 queues and MMA operands).
    """
    B, A, n_prompt = prompt.shape
    dev = prompt.device
    if video is not None:
        # finding F4: the reference cannot run generate() with video at all (its upsampled context is
        # always 160000 frames long while the window is RF long).  Not wired up here either yet.
        raise NotImplementedError("video-conditioned generate() is not implemented (it raises in the reference too)")
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
    state = torch.zeros(_lib.size("mvn_decode_state_bytes", shape), dtype=torch.uint8, device=dev)
    _lib.call("mvn_decode_prefill", C.byref(shape), acts.data_ptr(), state.data_ptr(), _stream())
    return DecodeState(shape, bufs, state, None, B, A)


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
              0 if st.ctx is None else st.ctx.data_ptr(), t_start, n_new, codes.data_ptr(),
              0 if logits is None else logits.data_ptr(), C.c_float(float(temperature)), seed, _stream())
    return (codes, logits) if return_logits else codes


def cached_generate(model, audio, video, n_samples, temperature, return_logits=False, fast=False):
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
        st = prefill(model, audio[:, :, :RF].contiguous(), video, fast=fast)
        logits = steps_into(model, st, RF, n_new, out[:, :, RF:], temperature, return_logits)
    return (out, logits) if return_logits else out


def steps_into(model, st, t_start, n_new, out_columns, temperature=0.0, return_logits=False):
    """generate n_new samples per clip and write them as one-hot columns into ``out_columns`` (B, A, n_new), which must
    be zero on entry: the reference's output format (movenet/wavenet.py:211-236), one 1.0 per generated sample"""
    res = run_steps(model, st, t_start, n_new, temperature, return_logits)
    codes, logits = res if return_logits else (res, None)
    out_columns.scatter_(1, codes.long().unsqueeze(1), 1.0)
    return logits
