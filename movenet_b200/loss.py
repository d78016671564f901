"""Fused version of the reference's training loss.

The trainers compute ``F.cross_entropy(model(audio, video), target)`` on the PROBABILITIES that
``forward`` returns (movenet/pytorch_lightning_trainer.py:62-65, movenet/trainer.py:127-129).
``WaveNet.forward`` hands its output back as a ``ProbabilityTensor`` -- a ``torch.Tensor`` subclass that
behaves like any tensor but recognises exactly that call (default arguments, class-index targets) and
routes it to one fused CUDA kernel pair instead of torch's log_softmax + nll_loss chain.  Any other use,
and any other argument combination, takes torch's ordinary path, so the trainer code is unchanged and
the numbers are the same function of the inputs.  ``MOVENET_B200_FUSED_CE=0`` switches the routing off.
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
            return softmax_cross_entropy(args[0], args[1])
        return super().__torch_function__(func, types, args, kwargs)
