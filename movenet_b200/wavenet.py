"""B200-native WaveNet with the reference's nn.Module surface.

Drop-in for ``movenet.wavenet.WaveNet`` (cosmicBboy/movenet, movenet/wavenet.py:50-239):
same constructor arguments, attributes, sub-module / parameter names and shapes
(``state_dict`` interchanges with reference checkpoints), same ``forward`` and
``generate`` signatures and semantics -- including the inverted
``output_unnormalized`` flag (movenet/wavenet.py:189-191: the default returns
softmax probabilities).  The arithmetic runs in the hand-written sm_100a CUDA
library behind include/movenet_b200.h; there is no PyTorch or CPU fallback.

Deviations from the reference as shipped (all documented in DESIGN.md):
* video conditioning: the reference raises at movenet/modules.py:76 (a length-T
  context is added to a length-(T-d) tensor); here the context is right-aligned
  the way the same function aligns the residual (movenet/modules.py:84).
* ``generate`` runs a cached (dilation-queue) decoder instead of recomputing a
  window per sample; see ``generate`` for what that means when stack_size == 1.
"""
import ctypes as C
import functools
import math
import os
from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .loss import ProbabilityTensor
from .modules import CausalConv1d, DenseConv, ResidualConvStack
from .types import AudioTensor, VideoTensor

# fixed clip geometry (movenet/wavenet.py:27-31): 10 s at 16 kHz, 16 video frames / s
MAX_AUDIO_FRAMES = 160000
MAX_VIDEO_FRAMES = 160
VIDEO_KERNEL_SIZE = (1, 64, 64)
UPSAMPLE_STRIDE = 10

_DTYPES = {"fp32": _lib.F32, "float32": _lib.F32, "bf16": _lib.BF16, "bfloat16": _lib.BF16}


def upsample_kernel_size_solver(in_size, out_size, stride=1, padding=0, output_padding=0, dilation=1):
    """Kernel size that makes ConvTranspose1d map in_size -> out_size (movenet/wavenet.py:34-47)."""
    k = out_size - 1 - output_padding - (in_size - 1) * stride + 2 * padding
    return (int(k / dilation + 1),)


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _Buffers:
    """Caller-owned device buffers for one (device, shape): the C ABI never allocates."""

    def __init__(self, shape: _lib.Shape, device):
        self.shape = shape
        self.device = device
        self.packed = torch.empty(_lib.size("mvn_packed_bytes", shape), dtype=torch.uint8, device=device)
        self.packed_stamp = None     # identifies the parameter values the packed image was built from (WaveNet._pack)
        self.packed_grads = None
        self.scratch = None
        self.acts_bytes = _lib.size("mvn_acts_bytes", shape)

    def get_scratch(self):
        if self.scratch is None:
            self.scratch = torch.empty(_lib.size("mvn_scratch_bytes", self.shape), dtype=torch.uint8, device=self.device)
        return self.scratch

    def get_packed_grads(self):
        if self.packed_grads is None:
            self.packed_grads = torch.empty_like(self.packed)
        return self.packed_grads


class _ForwardState:
    """what a backward pass needs from the forward pass; shared by the network's autograd node and, when the trainer's loss
    is taken straight from the returned probabilities, by the fused loss node (loss.py)"""
    # (never the output tensor itself: output -> grad_fn -> this object -> output would be a reference cycle that keeps a
    # whole step's activations alive until the garbage collector runs)
    __slots__ = ("module", "bufs", "acts", "audio", "video", "out_ptr", "has_video", "fused_loss_ok", "done")


class _WaveNetFunction(torch.autograd.Function):
    """forward()/backward() of the whole network as ONE autograd node."""

    @staticmethod
    def forward(ctx, module, audio, video, remove_last, output_logits, no_grad, *params):
        bufs = module._engine_buffers(audio, video is not None, remove_last, output_logits, no_grad)
        shape = bufs.shape
        module._pack(bufs, params)
        Tn = shape.frames - module.receptive_fields + 1 - (1 if remove_last else 0)
        out = torch.empty(shape.batch, shape.input_channels, max(Tn, 0), dtype=torch.float32, device=audio.device)
        acts = torch.empty(bufs.acts_bytes, dtype=torch.uint8, device=audio.device)
        is_codes = audio.dim() == 2
        if is_codes:
            _lib.call("mvn_codes_input", C.byref(shape), audio.data_ptr(), acts.data_ptr(), _stream())
        _lib.call("mvn_wavenet_forward", C.byref(shape), bufs.packed.data_ptr(), 0 if is_codes else audio.data_ptr(),
                  0 if video is None else video.data_ptr(), acts.data_ptr(), out.data_ptr(),
                  bufs.get_scratch().data_ptr(), _stream())
        st = _ForwardState()
        st.module, st.bufs, st.acts, st.audio, st.video, st.out_ptr = module, bufs, acts, audio, video, out.data_ptr()
        st.has_video = video is not None
        st.done = set()          # autograd nodes of this forward pass that have run their backward (each may run once)
        st.fused_loss_ok = bool(_lib.load().mvn_fused_loss_supported(C.byref(shape)))
        ctx.state = st
        ctx.save_for_backward(out)
        module._fwd_state = st
        return out

    @staticmethod
    def backward(ctx, dout):
        st = ctx.state
        module, bufs, audio, video = st.module, st.bufs, st.audio, st.video
        (out,) = ctx.saved_tensors
        # The activations stay alive as long as the graph does (like autograd's saved tensors): the same output may feed
        # several loss terms -- e.g. the fused cross-entropy node (loss.py) plus another differentiable use -- and each node
        # runs its backward once per pass; a SECOND pass through the same node raises like torch does without retain_graph.
        if "net" in st.done:
            raise RuntimeError("Trying to backward through the WaveNet graph a second time (the reference would need retain_graph=True)")
        st.done.add("net")
        dout = dout.contiguous().float()
        def run(pg_ptr):
            _lib.call("mvn_wavenet_backward", C.byref(bufs.shape), bufs.packed.data_ptr(), 0 if audio.dim() == 2 else audio.data_ptr(),
                      0 if video is None else video.data_ptr(), st.acts.data_ptr(), out.data_ptr(), dout.data_ptr(),
                      pg_ptr, bufs.get_scratch().data_ptr(), _stream())
        views = module._backward_and_average(bufs, st.has_video, audio.device, run)
        return (None, None, None, None, None, None, *views)


class WaveNet(nn.Module):
    """WaveNet with local (video) conditioning -- see the module docstring.

    Extra, defaulted, keyword-only knob (not in the reference): ``compute_dtype``
    = "fp32" (exact mode: CUDA-core fp32 arithmetic everywhere, the mode the
    token-exact decode guarantee is stated for) or "bf16" (tensor-core mode:
    bf16 activations, fp32 accumulation).  Default: $MOVENET_B200_DTYPE or fp32.
    """

    def __init__(self, layer_size: int, stack_size: int, input_channels: int, residual_channels: int = 16,
                 skip_channels: int = 16, context_in_channels: int = 1, *, compute_dtype: Optional[str] = None):
        super().__init__()
        self.layer_size = layer_size
        self.stack_size = stack_size
        self.input_channels = input_channels
        self.residual_channels = residual_channels
        self.skip_channels = skip_channels
        self.context_in_channels = context_in_channels
        compute_dtype = compute_dtype or os.environ.get("MOVENET_B200_DTYPE", "fp32")
        if compute_dtype not in _DTYPES:
            raise ValueError(f"compute_dtype must be one of {sorted(_DTYPES)}")
        self.compute_dtype = compute_dtype

        # video encoder: one 64x64 "pixel" linear map per frame (movenet/wavenet.py:94-98)
        self.video_conv = nn.Conv3d(context_in_channels, residual_channels, kernel_size=VIDEO_KERNEL_SIZE)
        # learned upsampling 160 -> 1600 -> 16000 -> 160000 frames (movenet/wavenet.py:100-118)
        sizes = np.geomspace(MAX_VIDEO_FRAMES, MAX_AUDIO_FRAMES,
                             num=math.ceil(np.log10(MAX_AUDIO_FRAMES / MAX_VIDEO_FRAMES) + 1)).astype(int)
        self.video_transpose = nn.Sequential(*[
            nn.ConvTranspose1d(residual_channels, residual_channels,
                               kernel_size=upsample_kernel_size_solver(a, b, stride=UPSAMPLE_STRIDE),
                               stride=UPSAMPLE_STRIDE)
            for a, b in zip(sizes[:-1], sizes[1:])])
        assert len(self.video_transpose) == 3 and all(m.kernel_size == (10,) for m in self.video_transpose)
        self.causal_conv = CausalConv1d(input_channels, residual_channels)
        self.residual_conv_stack = ResidualConvStack(layer_size, stack_size, residual_channels, skip_channels)
        self.dense_conv = DenseConv(skip_channels, input_channels)

        #: "exact" : fp32 CUDA-core decoder, the reference's function incl. its window edge, token-exact (default);
        #: "causal": fp32 decoder of the true causal model (differs from "exact" only for stack_size == 1);
        #: "fast"  : tensor-core decoder (bf16 queues / operands) where mvn_decode_tc_supported -- throughput mode
        self.decode_mode = os.environ.get("MOVENET_B200_DECODE", "exact")
        self._bufs = {}
        self._ptr_tables = {}
        self._dp_group = None
        self._dp_world = 1
        self._dp_peer = {}
        self._flat_in_pass = set()
        self._weights_epoch = 0      # bumped by anything that rewrites parameters behind autograd's back (see _pack)

    # ------------------------------------------------------------------ reference surface
    @property
    def receptive_fields(self) -> int:
        """sum of dilations + one per stack (movenet/wavenet.py:125-134)."""
        return sum(self.residual_conv_stack.dilations) + self.residual_conv_stack.stack_size

    def compute_output_size(self, x) -> int:
        """T - RF + 1, ValueError when the input is too short (movenet/wavenet.py:136-147)."""
        output_size = int(x.size(2)) - self.receptive_fields + 1
        if output_size < 1:
            raise ValueError(
                "input time steps must be larger than the number of receptive fields. "
                f"Number of input timesteps = {x.size(2)}, receptive fields = {self.receptive_fields}")
        return output_size

    def upsample_video(self, video: VideoTensor) -> torch.Tensor:
        """(B,160,64,64,Cin) -> (B,C,160000) fp32, channels-first (movenet/wavenet.py:149-156).

        Inspection helper: ``forward`` runs the same kernels internally and keeps the
        result time-major; this copy is detached from autograd.
        """
        video = self._check_video(video)
        B = video.shape[0]
        shape = self._shape(B, MAX_AUDIO_FRAMES, True, True, False, _lib.F32)
        bufs = self._buffers_for(shape, video.device)
        with torch.cuda.device(video.device):
            self._pack(bufs, self._param_list())
            acts = torch.empty(bufs.acts_bytes, dtype=torch.uint8, device=video.device)
            _lib.call("mvn_video_fwd", C.byref(shape), bufs.packed.data_ptr(), video.data_ptr(), acts.data_ptr(), _stream())
            ctx = torch.empty(B, MAX_AUDIO_FRAMES, self.residual_channels, dtype=torch.float32, device=video.device)
            _lib.call("mvn_read_activation", C.byref(shape), acts.data_ptr(), 2, 0, ctx.data_ptr(), _stream())
        out = ctx.permute(0, 2, 1).contiguous()
        assert out.shape[-1] == MAX_AUDIO_FRAMES
        return out

    def forward(self, audio: AudioTensor, video: Optional[VideoTensor] = None, global_features=None,
                output_unnormalized: bool = True, remove_last: bool = True):
        """movenet/wavenet.py:158-191.  NOTE the reference's flag polarity: the default
        (``output_unnormalized=True``) returns softmax PROBABILITIES over dim 1, raw logits come
        back only for ``output_unnormalized=False``.  ``global_features`` is unused there too.

        Non-breaking overload: ``audio`` may also be the (batch, frames) INTEGER tensor of mu-law
        codes themselves instead of their one-hot expansion -- 1/(4A) of the bytes to move to the device; the
        training target is then simply ``audio[:, receptive_fields:]``."""
        audio = self._check_audio(audio)
        frames = audio.shape[-1]
        if video is not None:
            video = self._check_video(video)
            assert frames == MAX_AUDIO_FRAMES and video.shape[0] == audio.shape[0], (
                "expected video and audio tensors to have equal sizes, found "
                f"{(video.shape[0], self.residual_channels, MAX_AUDIO_FRAMES)}, "
                f"{(audio.shape[0], self.residual_channels, frames)}")
        if frames - self.receptive_fields + 1 < 1:
            raise ValueError(
                "input time steps must be larger than the number of receptive fields. "
                f"Number of input timesteps = {frames}, receptive fields = {self.receptive_fields}")
        with torch.cuda.device(audio.device):
            params = self._param_list()
            # inference-only passes skip what only a backward needs (the wide path's gate-derivative factors)
            no_grad = not (torch.is_grad_enabled() and any(p.requires_grad for p in params))
            out = _WaveNetFunction.apply(self, audio, video, bool(remove_last), not output_unnormalized, no_grad, *params)
        # probabilities know the fused route for the trainer's F.cross_entropy(output, target) (loss.py)
        st = self.__dict__.pop("_fwd_state", None)
        if not output_unnormalized:
            return out
        out = out.as_subclass(ProbabilityTensor)
        out._mvn_state = st          # lets F.cross_entropy(out, target) run the loss-fused backward (loss.py)
        return out

    @torch.no_grad()
    def generate(self, audio: AudioTensor, video: Optional[VideoTensor] = None, global_features=None,
                 n_samples: Optional[int] = None, temperature: float = 1.0):
        """movenet/wavenet.py:193-239: keep the first RF columns of ``audio`` as the prompt and
        generate up to ``n_samples`` TOTAL columns; returns the (B, A, n) one-hot tensor.

        The reference recomputes an RF-long window per sample; this runs the cached decoder
        (per-layer rings, O(layers) per sample).  With stack_size >= 2 the two are the same
        function of the prompt.  With stack_size == 1 the reference's zero-padded window edge
        reaches its output; the default ``decode_mode = "exact"`` reproduces it with
        one extra edge column per layer per step (movenet_b200/decode.py), ``"causal"`` evaluates
        the true causal model instead, ``"fast"`` is the tensor-core throughput decoder.
        Only ``temperature == 0`` (argmax) is deterministic in the reference; ``temperature > 0``
        draws from softmax(probs / temperature).

        ``video``: the reference raises for it; here context column t-1 conditions
        sample t exactly as in ``forward`` (the oracle's window definition).
        """
        from .decode import cached_generate
        self.eval()
        return cached_generate(self, audio, video, n_samples, temperature, fast=(self.decode_mode == "fast"),
                               mode="exact" if self.decode_mode == "fast" else self.decode_mode)

    # ------------------------------------------------------------------ data parallel
    def enable_data_parallel(self, process_group=None):
        """Average gradients over ``process_group`` once per backward (the role DistributedDataParallel plays at
        movenet/trainer.py:230-234): one NCCL all-reduce of the flat gradient buffer, or with $MOVENET_B200_DP=peer (one
        node) the peer-memory kernel of csrc/peer.cu fused in front of the gradient unpack."""
        import torch.distributed as dist
        self._dp_group = process_group if process_group is not None else dist.group.WORLD
        self._dp_world = dist.get_world_size(self._dp_group)
        self._dp_peer = {}
        # the replicas must start identical: like DistributedDataParallel's constructor, take rank 0's parameters
        # (ranks may have been built with different RNG state, or only some may have loaded a checkpoint)
        if self._dp_world > 1:
            src = dist.get_global_rank(self._dp_group, 0)
            with torch.no_grad():
                for p in self.parameters():
                    dist.broadcast(p.data, src=src, group=self._dp_group)
            self._weights_epoch += 1
        return self

    def _reduce_grads(self, flat):
        # the gradients left mvn_unpack_grads already multiplied by 1 / world: a plain sum completes the average (ncclAvg was
        # measured slower at N = 2 -- 2.54 vs 2.35 ms per step -- and a separate scaling pass costs a launch)
        if self._dp_world > 1:
            import torch.distributed as dist
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self._dp_group)

    def _peer_gradients(self, bufs, has_video, device):
        """the NVLink peer-memory exchange of this (model, video flag), set up by a collective at its first backward;
        None when it was not asked for ($MOVENET_B200_DP=peer) or does not apply (one rank, gloo, several hosts): NCCL all-reduce then"""
        if self._dp_world <= 1:
            return None
        key = (has_video, str(device))
        if key not in self._dp_peer:
            from . import peer
            pgr = None
            if peer.available(self._dp_group, device):
                pgr = peer.PeerGradients(bufs.shape, device, self._dp_group)
                if not pgr.ok:
                    import warnings
                    warnings.warn(f"movenet_b200: peer-memory gradient exchange unavailable ({pgr.why}); using the NCCL all-reduce")
                    pgr = None
            self._dp_peer[key] = pgr
        return self._dp_peer[key]

    def _backward_and_average(self, bufs, has_video, device, run):
        """run(pg_ptr) launches the backward kernels into a packed-gradient buffer; returns the per-parameter gradient
        views (averaged over the data-parallel group) in parameter order"""
        peer = self._peer_gradients(bufs, has_video, device)
        pg = bufs.get_packed_grads()
        run(pg.data_ptr())
        flat, views = self._flat_grads(has_video, device)
        offs = self._grad_offsets(has_video, device)
        if peer is not None:
            # one kernel: push slices, rank-ordered sum, push the sums (all NVLink traffic is stores); then the unpack
            peer.reduce_unpack(bufs.shape, pg.data_ptr(), flat, offs, _stream())
        else:
            _lib.call("mvn_unpack_grads", C.byref(bufs.shape), pg.data_ptr(), flat.data_ptr(), offs.data_ptr(),
                      C.c_float(1.0 / self._dp_world), _stream())
            self._reduce_grads(flat)
        return views

    # ------------------------------------------------------------------ plumbing
    def _param_list(self):
        """parameters in state_dict order == the C ABI's MVN_PARAM_* order (cached: walking the module tree every
        step costs more host time than some of the kernels take)"""
        cache = self.__dict__.get("_param_cache")
        if cache is None:
            cache = self.__dict__["_param_cache"] = [p for _, p in self.named_parameters()]
        return cache

    def _apply(self, fn, *args, **kwargs):
        self.__dict__.pop("_param_cache", None)
        self.__dict__.pop("_grad_cache", None)
        self._ptr_tables = {}
        return super()._apply(fn, *args, **kwargs)

    def _check_audio(self, audio):
        if isinstance(audio, torch.Tensor) and audio.dim() == 2 and not audio.is_floating_point():
            if not audio.is_cuda:
                raise RuntimeError("movenet_b200.WaveNet runs on CUDA tensors only (there is no CPU path)")
            return audio.detach().to(torch.int64).contiguous()          # integer codes (batch, frames)
        if not isinstance(audio, torch.Tensor) or audio.dim() != 3:
            raise ValueError("audio must be a (batch, channels, frames) tensor")
        if not audio.is_cuda:
            raise RuntimeError("movenet_b200.WaveNet runs on CUDA tensors only (there is no CPU path)")
        if audio.shape[1] != self.input_channels:
            raise RuntimeError(f"expected {self.input_channels} audio channels, found {audio.shape[1]}")
        if audio.requires_grad:
            raise RuntimeError("gradients with respect to the audio input are not implemented")
        return audio.detach().float().contiguous()

    def _check_video(self, video):
        if not video.is_cuda:
            raise RuntimeError("movenet_b200.WaveNet runs on CUDA tensors only (there is no CPU path)")
        assert video.dim() == 5 and tuple(video.shape[1:]) == (MAX_VIDEO_FRAMES, 64, 64, self.context_in_channels), (
            f"expected video of shape (B, {MAX_VIDEO_FRAMES}, 64, 64, {self.context_in_channels}), found {tuple(video.shape)}")
        return video.detach().float().contiguous()

    def _shape(self, B, T, has_video, remove_last, output_logits, act_dtype=None, no_grad=False):
        return _lib.Shape(self.layer_size, self.stack_size, self.input_channels, self.residual_channels,
                          self.skip_channels, self.context_in_channels, B, T, int(has_video),
                          _DTYPES[self.compute_dtype] if act_dtype is None else act_dtype,
                          int(remove_last), int(output_logits), int(no_grad))

    def _buffers_for(self, shape, device):
        key = (shape.key(), str(device))
        bufs = self._bufs.get(key)
        if bufs is None:
            if len(self._bufs) >= 8:          # bound the cache: drop the oldest geometry
                self._bufs.pop(next(iter(self._bufs)))
            bufs = self._bufs[key] = _Buffers(shape, device)
        return bufs

    def _engine_buffers(self, audio, has_video, remove_last, output_logits, no_grad=False):
        shape = self._shape(audio.shape[0], audio.shape[-1], has_video, remove_last, output_logits, no_grad=no_grad)
        return self._buffers_for(shape, audio.device)

    def _pack(self, bufs, params):
        """re-layout the reference parameters for the kernels -- only when they changed since this buffer set was last
        packed: every in-place update through torch bumps a tensor's ``_version``; movenet_b200.optim.AdamW (which writes
        through raw pointers) and enable_data_parallel bump ``_lib.weights_epoch`` / ``_weights_epoch`` instead"""
        ptrs = tuple(p.data_ptr() for p in params)
        key = ("w", str(bufs.device))
        cached = self._ptr_tables.get(key)
        if cached is None or cached[0] != ptrs:
            for p in params:
                if p.dtype != torch.float32 or not p.is_contiguous() or p.device != bufs.device:
                    raise RuntimeError("movenet_b200.WaveNet parameters must be contiguous fp32 tensors on the input's device")
            table = torch.tensor(ptrs, dtype=torch.int64).to(bufs.device)
            cached = self._ptr_tables[key] = (ptrs, table)
        stamp = (ptrs, _lib.weights_epoch[0], self._weights_epoch, sum(p._version for p in params),
                 torch.cuda.current_stream().cuda_stream)
        if bufs.packed_stamp == stamp:
            return
        _lib.call("mvn_pack_weights", C.byref(bufs.shape), cached[1].data_ptr(), bufs.packed.data_ptr(), _stream())
        bufs.packed_stamp = stamp

    def _grad_layout(self, has_video):
        """element offset of every parameter's gradient in the flat buffer (-1: no gradient, as in the
        reference: video/context parameters without video, and the last layer's conv_residual whose
        output is discarded, movenet/modules.py:125-130)."""
        cache = self.__dict__.setdefault("_grad_cache", {})
        if has_video in cache:
            return cache[has_video][0], cache[has_video][1]
        last = f"residual_conv_stack.conv_layers.{self.layer_size * self.stack_size - 1}.conv_residual."
        offsets, off = [], 0
        for name, p in self.named_parameters():
            no_grad = (not p.requires_grad or name.startswith(last)
                       or (not has_video and (name.startswith("video_") or ".context_conv_" in name)))
            if no_grad:
                offsets.append(-1)
            else:
                offsets.append(off)
                off += p.numel()           # dense, no padding: the views come from ONE unflatten call (host time, see _flat_grads)
        with_grad = [p for o, p in zip(offsets, self._param_list()) if o >= 0]
        cache[has_video] = (offsets, off, with_grad)
        return offsets, off

    def _grad_offsets(self, has_video, device):
        key = ("g", has_video, str(device))
        if key not in self._ptr_tables:
            offsets, _ = self._grad_layout(has_video)
            self._ptr_tables[key] = torch.tensor(offsets, dtype=torch.int64).to(device)
        return self._ptr_tables[key]

    def _flat_grads(self, has_video, device):
        """the flat fp32 gradient buffer and, per parameter, a FRESH view of it that becomes param.grad (None: no gradient).

        Host time matters here (103 parameters; eight trainer processes share one host): the views are made by one
        ``unflatten_dense_tensors`` call, and the buffer itself is REUSED from step to step as long as every parameter's
        ``.grad`` is None when the backward runs (``zero_grad(set_to_none=True)``, torch's default) -- so the gradient pointers
        stay the same and ``movenet_b200.optim.AdamW`` keeps its device tables (like DDP's ``gradient_as_bucket_view``: a
        gradient tensor the caller kept from an earlier step is overwritten).  Otherwise (gradient accumulation over several
        backward passes) every pass gets a new buffer."""
        offsets, total = self._grad_layout(has_video)
        with_grad = self._grad_cache[has_video][2]
        key = ("flat", has_video, str(device))
        flat = self._ptr_tables.get(key)
        # reusable: nothing aliases the kept buffer any more -- no parameter holds a gradient, and it has not already been handed
        # out in THIS backward pass (several loss terms / forward passes in one backward: their gradients are still on their way
        # to .grad; the flag is cleared by an engine callback when the pass is over)
        reusable = key not in self._flat_in_pass and all(p.grad is None for p in with_grad)
        if flat is None or not reusable:
            flat = torch.empty(total, dtype=torch.float32, device=device)
            if reusable:
                self._ptr_tables[key] = flat
        if reusable:
            self._flat_in_pass.add(key)
            try:
                torch.autograd.Variable._execution_engine.queue_callback(lambda: self._flat_in_pass.discard(key))
            except RuntimeError:          # not inside a backward pass (a direct call): nothing to wait for
                self._flat_in_pass.discard(key)
        it = iter(torch._C._nn.unflatten_dense_tensors(flat, with_grad))
        views = [None if o < 0 else next(it) for o in offsets]
        return flat, views
