"""Fused version of the reference's training loss.

The trainers compute ``F.cross_entropy(model(audio, video), target)`` on the PROBABILITIES that
``forward`` returns (movenet/pytorch_lightning_trainer.py:62-65, movenet/trainer.py:127-129).
``WaveNet.forward`` hands its output back as a ``ProbabilityTensor`` -- a ``torch.Tensor`` subclass that
behaves like any tensor but recognises exactly that call (default arguments, class-index targets) and
routes it to one fused CUDA kernel pair instead of torch's log_softmax + nll_loss chain.  Any other use,
and any other argument combination, takes torch's ordinary path, so the trainer code is unchanged and
the numbers are the same function of the inputs.  ``MOVENET_B200_FUSED_CE=0`` switches the routing off.

Limits of the fused route (it is taken for the default arguments only): targets must be valid class indices in
[0, channels) -- torch's default ``ignore_index=-100`` is accepted as an argument but such targets are NOT ignored (the
trainers never produce them: ``target = audio[:, :, RF:].argmax(1)``); an out-of-range target contributes
``logsumexp`` without a picked probability instead of raising.  The output of one ``forward`` may feed several loss
terms (the fused node plus any other differentiable use): every autograd node of the pass runs once, a second pass
through the same node raises like torch without ``retain_graph``.
"""
import ctypes as C
import os

import torch
import torch.nn.functional as F

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _SoftmaxCrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs, target):
        B, A, T = probs.shape
        loss = torch.empty((), dtype=torch.float32, device=probs.device)
        partials = torch.empty(_lib.load().mvn_softmax_ce_partials(B, T), dtype=torch.float32, device=probs.device)
        with torch.cuda.device(probs.device):
            _lib.call("mvn_softmax_ce_fwd", probs.data_ptr(), target.data_ptr(), B, A, T, partials.data_ptr(),
                      loss.data_ptr(), _stream())
        ctx.save_for_backward(probs, target)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        probs, target = ctx.saved_tensors
        B, A, T = probs.shape
        dprobs = torch.empty_like(probs)
        g = grad_loss.contiguous().float()
        with torch.cuda.device(probs.device):
            _lib.call("mvn_softmax_ce_bwd", probs.data_ptr(), target.data_ptr(), g.data_ptr(), B, A, T,
                      dprobs.data_ptr(), _stream())
        return dprobs, None


class _FusedLoss(torch.autograd.Function):
    """loss = cross_entropy(probabilities, target) as ONE autograd node over the network's parameters: its backward runs
    the whole network backward with the loss gradient formed inside the head kernel (mvn_wavenet_backward_loss), so the
    (B, A, T) gradient of the probabilities is never materialised.  The probabilities enter detached, so their own autograd
    node is not part of this loss's graph (it still serves any other use of the tensor)."""

    @staticmethod
    def forward(ctx, state, probs, target, *params):
        B, A, T = probs.shape
        loss = torch.empty((), dtype=torch.float32, device=probs.device)
        partials = torch.empty(_lib.load().mvn_softmax_ce_partials(B, T), dtype=torch.float32, device=probs.device)
        with torch.cuda.device(probs.device):
            _lib.call("mvn_softmax_ce_fwd", probs.data_ptr(), target.data_ptr(), B, A, T, partials.data_ptr(),
                      loss.data_ptr(), _stream())
        ctx.state = state
        ctx.save_for_backward(target, probs)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        st = ctx.state
        target, probs = ctx.saved_tensors
        key = ("loss", id(ctx))
        if key in st.done:
            raise RuntimeError("Trying to backward through the fused loss node a second time (the reference would need retain_graph=True)")
        st.done.add(key)
        module, bufs, audio, video = st.module, st.bufs, st.audio, st.video
        g = grad_loss.contiguous().float()
        def run(pg_ptr):
            _lib.call("mvn_wavenet_backward_loss", C.byref(bufs.shape), bufs.packed.data_ptr(),
                      0 if audio.dim() == 2 else audio.data_ptr(), 0 if video is None else video.data_ptr(),
                      st.acts.data_ptr(), probs.data_ptr(), target.data_ptr(), g.data_ptr(), pg_ptr,
                      bufs.get_scratch().data_ptr(), _stream())
        with torch.cuda.device(audio.device):
            views = module._backward_and_average(bufs, st.has_video, audio.device, run)
        return (None, None, None, *views)


def softmax_cross_entropy(probs: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """mean_{b,t} [ logsumexp_c probs[b,c,t] - probs[b,target[b,t],t] ]  ==  F.cross_entropy(probs, target)"""
    return _SoftmaxCrossEntropy.apply(probs.as_subclass(torch.Tensor).contiguous(), target.contiguous())


def _fast_path_ok(args, kwargs):
    if os.environ.get("MOVENET_B200_FUSED_CE", "1") == "0" or len(args) != 2:
        return False
    defaults = {"weight": None, "size_average": None, "ignore_index": -100, "reduce": None, "reduction": "mean",
                "label_smoothing": 0.0}
    for k, v in kwargs.items():
        if k not in defaults or v != defaults[k]:
            return False
    x, t = args
    return (isinstance(x, torch.Tensor) and isinstance(t, torch.Tensor) and x.is_cuda and t.is_cuda
            and x.dim() == 3 and x.dtype == torch.float32 and t.dtype == torch.int64
            and t.shape == (x.shape[0], x.shape[2]) and x.numel() > 0)


class ProbabilityTensor(torch.Tensor):
    """what WaveNet.forward returns: an ordinary tensor that knows the fused route for the trainer's loss"""

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        if func is F.cross_entropy and _fast_path_ok(args, kwargs):
            x, t = args
            st = getattr(x, "_mvn_state", None)
            if (st is not None and st.fused_loss_ok and st.acts is not None and st.out_ptr == x.data_ptr()
                    and torch.is_grad_enabled() and x.requires_grad
                    and os.environ.get("MOVENET_B200_FUSED_LOSS_BWD", "1") != "0"):
                # the probabilities enter detached: this node differentiates the loss w.r.t. the parameters itself
                return _FusedLoss.apply(st, x.detach().as_subclass(torch.Tensor), t.contiguous(), *st.module._param_list())
            return softmax_cross_entropy(x, t)
        return super().__torch_function__(func, types, args, kwargs)
