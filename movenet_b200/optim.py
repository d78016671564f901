"""AdamW for the whole model in one kernel launch.

``movenet_b200.optim.AdamW(model.parameters(), lr=...)`` is a ``torch.optim.Optimizer`` with the update of
``torch.optim.AdamW`` (the optimizer ``movenet/pytorch_lightning_trainer.py:128-202`` builds by default) and, optionally, the
global gradient-norm clip of ``:233-243`` (``max_grad_norm``) folded into the same pass.  The model has 10 N + 13 small
tensors; torch's fused multi-tensor AdamW needs three launches for them, this one needs one (two when clipping), and nothing
is read back to the host.  Parameters without a gradient are skipped, like in torch, and -- like in torch -- every
parameter keeps its OWN step counter (``state[p]["step"]``), so a parameter that only intermittently receives a gradient
(the video / context tensors when batches mix video and no-video) gets its own bias correction: parameters are grouped by
step count, one launch per distinct count (one in the usual case).  The clip norm is global over all groups' gradients.
A ``torch.optim.AdamW`` state_dict loads (its per-parameter ``step`` is taken over).  CUDA fp32 parameters only.
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _lib


class AdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=None):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm))
        self._tables = {}          # group index -> (key, segments_dev, chunks_dev, n_chunks, partials)
        self.grad_norm = None      # device scalar holding the last pre-clip gradient norm (when clipping)

    def _table(self, key_id, active):
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in active)
        cached = self._tables.get(key_id)
        if cached is not None and cached[0] == key:
            return cached
        for p in active:          # (checked when a table is built: the same tensors at the same addresses were checked before)
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()
                    and p.grad.dtype == torch.float32 and not p.grad.is_sparse):
                raise RuntimeError("movenet_b200.optim.AdamW: contiguous CUDA fp32 parameters and gradients only")
        dev = active[0].device
        chunk = _lib.load().mvn_adamw_chunk_elems()
        assert _lib.load().mvn_adamw_segment_bytes() == 40
        seg = np.zeros((len(active), 5), dtype=np.int64)
        chunks = []
        for i, p in enumerate(active):
            st = self.state[p]
            seg[i] = (p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel())
            chunks += [(i, c) for c in range((p.numel() + chunk - 1) // chunk)]
        seg_dev = torch.from_numpy(seg).to(dev)
        chunks_dev = torch.tensor(chunks, dtype=torch.int32, device=dev)
        partials = torch.empty(len(chunks), dtype=torch.float32, device=dev)
        self._tables[key_id] = (key, seg_dev, chunks_dev, len(chunks), partials)
        return self._tables[key_id]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            active = [p for p in group["params"] if p.grad is not None]
            if not active:
                continue
            for p in active:
                st = self.state[p]
                if "exp_avg" not in st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                # per-parameter step (torch keeps a tensor; a loaded torch state_dict is taken over)
                st["step"] = int(st.get("step", 0)) + 1
            b1, b2 = group["betas"]
            clip = group.get("max_grad_norm") or 0.0
            by_step = {}
            for p in active:
                by_step.setdefault(self.state[p]["step"], []).append(p)
            if clip > 0 and len(by_step) > 1:
                raise RuntimeError("movenet_b200.optim.AdamW: max_grad_norm needs all parameters of a group to have the same "
                                   "step count (the clip norm is computed inside the one launch)")
            if clip > 0 and self.grad_norm is None:
                self.grad_norm = torch.zeros(1, dtype=torch.float32, device=active[0].device)
            for si, (t, plist) in enumerate(sorted(by_step.items())):
                _, seg_dev, chunks_dev, n_chunks, partials = self._table((gi, si), plist)
                with torch.cuda.device(plist[0].device):
                    _lib.call("mvn_adamw_step", seg_dev.data_ptr(), chunks_dev.data_ptr(), n_chunks, float(group["lr"]), float(b1),
                              float(b2), float(group["eps"]), float(group["weight_decay"]), 1.0 - math.pow(b1, t),
                              1.0 - math.pow(b2, t), float(clip), partials.data_ptr(),
                              self.grad_norm.data_ptr() if clip > 0 else 0, torch.cuda.current_stream().cuda_stream)
            _lib.weights_epoch[0] += 1       # the parameters were rewritten behind autograd's version counters
        return loss
