"""fp32 CPU restatement of the movenet WaveNet hot path (test infrastructure).

See oracle/__init__.py for the rules about who may import this and for the
parity status.  Every function cites the reference lines it restates
(paths relative to /root/reference).  Parameters travel as a plain
``dict[str, Tensor]`` whose keys and shapes are exactly the reference
``WaveNet.state_dict()`` ones, so a reference checkpoint drops in.

The arithmetic is deliberately the same ATen CPU ops the reference calls
(``conv1d`` / ``conv3d`` / ``conv_transpose1d`` / ``softmax`` /
``cross_entropy``) in the same order, which is what makes the audio-only
outputs bit-identical to the reference's on a CPU.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

MAX_AUDIO_FRAMES = 160000   # movenet/wavenet.py:27
MAX_VIDEO_FRAMES = 160      # movenet/wavenet.py:28
VIDEO_HW = 64               # movenet/wavenet.py:29  VIDEO_KERNEL_SIZE = (1, 64, 64)
UPSAMPLE_STRIDE = 10        # movenet/wavenet.py:31
LRELU_SLOPE = 0.01          # F.leaky_relu default, movenet/modules.py:140-141


@dataclass(frozen=True)
class Shape:
    """The six constructor arguments of WaveNet (movenet/wavenet.py:75-83)."""
    layer_size: int
    stack_size: int
    input_channels: int
    residual_channels: int = 16
    skip_channels: int = 16
    context_in_channels: int = 1

    @property
    def dilations(self):
        # movenet/modules.py:113-117
        return [2 ** x for _ in range(self.stack_size) for x in range(self.layer_size)]

    @property
    def receptive_fields(self) -> int:
        # movenet/wavenet.py:125-134 : sum of dilations plus one per stack
        return sum(self.dilations) + self.stack_size

    @property
    def n_layers(self) -> int:
        return self.layer_size * self.stack_size


def layer_prefix(i: int) -> str:
    return f"residual_conv_stack.conv_layers.{i}."


def init_params(shape: Shape, seed: int = 0, video: bool = True) -> Dict[str, torch.Tensor]:
    """Random parameters with the reference's key names / shapes.

    The distribution is PyTorch's default conv init (uniform(+-1/sqrt(fan_in))
    for weight and bias), but the stream is our own: parity tests always copy
    the SAME tensors into both implementations, they never rely on RNG parity.
    """
    g = torch.Generator().manual_seed(seed)
    A, C, S = shape.input_channels, shape.residual_channels, shape.skip_channels

    def u(*size, fan_in):
        bound = 1.0 / math.sqrt(fan_in)
        return (torch.rand(*size, generator=g, dtype=torch.float32) * 2 - 1) * bound

    p: Dict[str, torch.Tensor] = {}
    if video:
        fan = shape.context_in_channels * VIDEO_HW * VIDEO_HW
        p["video_conv.weight"] = u(C, shape.context_in_channels, 1, VIDEO_HW, VIDEO_HW, fan_in=fan)
        p["video_conv.bias"] = u(C, fan_in=fan)
        for k in range(3):
            # ConvTranspose1d weight is (in, out, k); torch computes fan_in from dim 1
            p[f"video_transpose.{k}.weight"] = u(C, C, UPSAMPLE_STRIDE, fan_in=C * UPSAMPLE_STRIDE)
            p[f"video_transpose.{k}.bias"] = u(C, fan_in=C * UPSAMPLE_STRIDE)
    p["causal_conv.conv.weight"] = u(C, A, 2, fan_in=2 * A)
    for i in range(shape.n_layers):
        pre = layer_prefix(i)
        p[pre + "conv_filter.conv.weight"] = u(C, C, 2, fan_in=2 * C)
        p[pre + "conv_gate.conv.weight"] = u(C, C, 2, fan_in=2 * C)
        p[pre + "context_conv_filter.weight"] = u(C, C, 1, fan_in=C)
        p[pre + "context_conv_filter.bias"] = u(C, fan_in=C)
        p[pre + "context_conv_gate.weight"] = u(C, C, 1, fan_in=C)
        p[pre + "context_conv_gate.bias"] = u(C, fan_in=C)
        p[pre + "conv_residual.weight"] = u(C, C, 1, fan_in=C)
        p[pre + "conv_residual.bias"] = u(C, fan_in=C)
        p[pre + "conv_skip.weight"] = u(S, C, 1, fan_in=C)
        p[pre + "conv_skip.bias"] = u(S, fan_in=C)
    p["dense_conv.conv1.weight"] = u(A, S, 1, fan_in=S)
    p["dense_conv.conv1.bias"] = u(A, fan_in=S)
    p["dense_conv.conv2.weight"] = u(A, A, 1, fan_in=A)
    p["dense_conv.conv2.bias"] = u(A, fan_in=A)
    return p


def upsample_video(p, video: torch.Tensor) -> torch.Tensor:
    """movenet/wavenet.py:149-156.  (B,160,64,64,Cin) -> (B,C,160000)."""
    v = video.permute(0, 4, 1, 2, 3)
    enc = F.conv3d(v, p["video_conv.weight"], p["video_conv.bias"]).squeeze(-1).squeeze(-1)
    for k in range(3):
        enc = F.conv_transpose1d(enc, p[f"video_transpose.{k}.weight"],
                                 p[f"video_transpose.{k}.bias"], stride=UPSAMPLE_STRIDE)
    assert enc.shape[-1] == MAX_AUDIO_FRAMES
    return enc


def causal_conv(p, audio: torch.Tensor) -> torch.Tensor:
    """movenet/modules.py:15-30: Conv1d(k=2, pad=1, no bias), last column dropped."""
    return F.conv1d(audio, p["causal_conv.conv.weight"], None, padding=1)[:, :, :-1]


def gated_layer(p, i: int, dilation: int, x, context, skip_size: int):
    """movenet/modules.py:67-93, with the F3 crop (context right-aligned)."""
    pre = layer_prefix(i)
    f = F.conv1d(x, p[pre + "conv_filter.conv.weight"], None, dilation=dilation)
    g = F.conv1d(x, p[pre + "conv_gate.conv.weight"], None, dilation=dilation)
    if context is not None:
        # the reference adds the un-cropped context here and raises
        # (movenet/modules.py:75-77); crop like movenet/modules.py:84 does.
        ctx = context[:, :, -f.size(2):]
        f = f + F.conv1d(ctx, p[pre + "context_conv_filter.weight"], p[pre + "context_conv_filter.bias"])
        g = g + F.conv1d(ctx, p[pre + "context_conv_gate.weight"], p[pre + "context_conv_gate.bias"])
    gated = torch.tanh(f) * torch.sigmoid(g)
    residual = F.conv1d(gated, p[pre + "conv_residual.weight"], p[pre + "conv_residual.bias"])
    residual = residual + x[:, :, -residual.size(2):]
    skip = F.conv1d(gated, p[pre + "conv_skip.weight"], p[pre + "conv_skip.bias"])
    return residual, skip[:, :, -skip_size:]


def residual_stack(p, shape: Shape, x, context, skip_size: int) -> torch.Tensor:
    """movenet/modules.py:119-130 followed by the sum of movenet/wavenet.py:181."""
    skips = []
    for i, d in enumerate(shape.dilations):
        x, s = gated_layer(p, i, d, x, context, skip_size)
        skips.append(s)
    return torch.sum(torch.stack(skips), dim=0)


def dense_head(p, x):
    """movenet/modules.py:139-142."""
    x = F.conv1d(F.leaky_relu(x), p["dense_conv.conv1.weight"], p["dense_conv.conv1.bias"])
    return F.conv1d(F.leaky_relu(x), p["dense_conv.conv2.weight"], p["dense_conv.conv2.bias"])


def stack_from_context(p, shape: Shape, audio, context, output_unnormalized=True, remove_last=True):
    """movenet/wavenet.py:166-191 with an already-upsampled context."""
    h = causal_conv(p, audio)
    if context is not None:
        assert context.size() == h.size()
    out_size = int(h.size(2)) - shape.receptive_fields + 1    # movenet/wavenet.py:136-147
    if out_size < 1:
        raise ValueError("input time steps must be larger than the number of receptive fields")
    out = dense_head(p, residual_stack(p, shape, h, context, out_size))
    if remove_last:
        out = out[:, :, :-1]
    if not output_unnormalized:      # sic: movenet/wavenet.py:189-191 (flag is inverted)
        return out
    return F.softmax(out, dim=1)


def forward(p, shape: Shape, audio, video=None, output_unnormalized=True, remove_last=True):
    """movenet/wavenet.py:158-191.  Default returns PROBABILITIES (finding F1)."""
    context = None if video is None else upsample_video(p, video)
    return stack_from_context(p, shape, audio, context, output_unnormalized, remove_last)


def training_loss(p, shape: Shape, audio, video=None):
    """movenet/pytorch_lightning_trainer.py:62-66: CE on the probabilities (F2)."""
    output = forward(p, shape, audio, video)
    target = audio[:, :, shape.receptive_fields:].argmax(1)
    loss = F.cross_entropy(output, target)
    acc = (output.argmax(1) == target).float().mean()
    return loss, output, target, acc


def loss_and_grads(p, shape: Shape, audio, video=None):
    """One fwd+bwd; returns (loss, output, {name: grad or None})."""
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    loss, output, _, _ = training_loss(leaf, shape, audio, video)
    loss.backward()
    return loss.detach(), output.detach(), {k: v.grad for k, v in leaf.items()}


@torch.no_grad()
def generate(p, shape: Shape, audio, context=None, n_samples: Optional[int] = None,
             temperature: float = 0.0, return_logits: bool = False):
    """movenet/wavenet.py:193-239, window recompute per sample.

    ``context`` is an already-upsampled (B,C,>=n) tensor; the window slice
    [i-RF, i) of it is what conditions step i (our definition, finding F4: the
    reference cannot run generate() with video at all).
    Only temperature == 0 is deterministic: argmax(softmax(softmax(z))).
    """
    RF = shape.receptive_fields
    n = audio.shape[2] if n_samples is None else n_samples
    out = torch.zeros(audio.shape[0], audio.shape[1], n, dtype=audio.dtype)
    out[:, :, :RF] = audio[:, :, :RF]
    logits = []
    for i in range(RF, n):
        ctx = None if context is None else context[:, :, i - RF:i]
        z = stack_from_context(p, shape, out[:, :, i - RF:i], ctx,
                               output_unnormalized=False, remove_last=False)
        assert z.shape[2] == 1
        probs = F.softmax(z, dim=1)
        if temperature > 0:
            choice = torch.multinomial(F.softmax(probs / temperature, dim=1).squeeze(2), 1).unsqueeze(2)
        else:
            choice = F.softmax(probs, dim=1).argmax(1, keepdim=True)
        out[:, :, [i]] = torch.zeros_like(probs).scatter_(1, choice, 1)
        logits.append(z[:, :, 0])
    if return_logits:
        return out, (torch.stack(logits, dim=2) if logits else None)
    return out


@torch.no_grad()
def causal_logits(p, shape: Shape, audio, context=None):
    """Logits of the TRUE causal model for every t (zero history before t=0).

    Column t is the prediction for sample t+1 from x[..t].  Equals the
    reference's full-sequence logits on its valid region t >= RF-1; used to
    document finding F5 (the windowed generate() differs when stack_size==1).
    """
    # zero *activations* before t=0 in every layer (what a dilation-queue decoder
    # that starts from empty queues computes)
    x = causal_conv(p, audio)
    cur_ctx = context
    skips = 0
    for i, d in enumerate(shape.dilations):
        xp = F.pad(x, (d, 0))
        pre = layer_prefix(i)
        f = F.conv1d(xp, p[pre + "conv_filter.conv.weight"], None, dilation=d)
        g = F.conv1d(xp, p[pre + "conv_gate.conv.weight"], None, dilation=d)
        if cur_ctx is not None:
            f = f + F.conv1d(cur_ctx, p[pre + "context_conv_filter.weight"], p[pre + "context_conv_filter.bias"])
            g = g + F.conv1d(cur_ctx, p[pre + "context_conv_gate.weight"], p[pre + "context_conv_gate.bias"])
        gated = torch.tanh(f) * torch.sigmoid(g)
        skips = skips + F.conv1d(gated, p[pre + "conv_skip.weight"], p[pre + "conv_skip.bias"])
        x = F.conv1d(gated, p[pre + "conv_residual.weight"], p[pre + "conv_residual.bias"]) + x
    return dense_head(p, skips)


@torch.no_grad()
def window_edge_logits(p, shape: Shape, audio, context=None):
    """What generate()'s RF-long window yields at every step i in [RF, n], computed WITHOUT the window recompute.

    Restates the algorithm of the cached decoder's reference-window mode (movenet_b200/csrc/decode.cu,
    MVN_DECODE_REFERENCE) so that the derivation is pinned on the CPU against the reference's own generate() logits
    (tests/test_oracle.py).  Only meaningful for stack_size == 1 (finding F5); for stack_size >= 2 it returns
    causal_logits.  movenet/wavenet.py:217-224 feeds x[i-RF:i]; movenet/modules.py:15-30 zero-pads its left edge, so
    the window's first h0 column is e_0 = W[:,:,1] x[i-RF]; each layer drops d columns on the left
    (movenet/modules.py:36-46), hence exactly one column per layer descends from e_0:
        e_{l+1} = Wr_l gate(Wz0_l e_l + Wz1_l x_l[p_l] (+ ctx[p_l])) + br_l + x_l[p_l],  p_l = i-1 - sum_{k>l} d_k
    and only the LAST layer's skip output at i-1 is built from it.
    Returns (B, A, n-RF+1): column j is the logits for sample RF+j.
    """
    RF, dil, N = shape.receptive_fields, shape.dilations, shape.n_layers
    n = audio.shape[2]
    if shape.stack_size != 1:
        return causal_logits(p, shape, audio, context)[:, :, RF - 1:]
    # true causal layer inputs x_l[t] and skip outputs for every t
    xs, skips = [], []
    x = causal_conv(p, audio)
    for i, d in enumerate(dil):
        xs.append(x)
        xp = F.pad(x, (d, 0))
        pre = layer_prefix(i)
        f = F.conv1d(xp, p[pre + "conv_filter.conv.weight"], None, dilation=d)
        g = F.conv1d(xp, p[pre + "conv_gate.conv.weight"], None, dilation=d)
        if context is not None:
            f = f + F.conv1d(context[:, :, :n], p[pre + "context_conv_filter.weight"], p[pre + "context_conv_filter.bias"])
            g = g + F.conv1d(context[:, :, :n], p[pre + "context_conv_gate.weight"], p[pre + "context_conv_gate.bias"])
        gated = torch.tanh(f) * torch.sigmoid(g)
        skips.append(F.conv1d(gated, p[pre + "conv_skip.weight"], p[pre + "conv_skip.bias"]))
        x = F.conv1d(gated, p[pre + "conv_residual.weight"], p[pre + "conv_residual.bias"]) + x
    age = [sum(dil[l + 1:]) for l in range(N)]
    steps = torch.arange(RF, n + 1)                       # i: the sample being predicted
    tau = steps - 1
    W1 = p["causal_conv.conv.weight"][:, :, 1]            # (C, A)
    e = torch.einsum("ca,bat->bct", W1, audio[:, :, steps - RF])
    skip_sum = sum(s[:, :, tau] for s in skips[:-1]) if N > 1 else 0
    for l, d in enumerate(dil):
        pre = layer_prefix(l)
        xl = xs[l][:, :, tau - age[l]]
        wf, wg = p[pre + "conv_filter.conv.weight"], p[pre + "conv_gate.conv.weight"]
        f = torch.einsum("oc,bct->bot", wf[:, :, 0], e) + torch.einsum("oc,bct->bot", wf[:, :, 1], xl)
        g = torch.einsum("oc,bct->bot", wg[:, :, 0], e) + torch.einsum("oc,bct->bot", wg[:, :, 1], xl)
        if context is not None:
            cx = context[:, :, tau - age[l]]
            f = f + F.conv1d(cx, p[pre + "context_conv_filter.weight"], p[pre + "context_conv_filter.bias"])
            g = g + F.conv1d(cx, p[pre + "context_conv_gate.weight"], p[pre + "context_conv_gate.bias"])
        gated = torch.tanh(f) * torch.sigmoid(g)
        if l < N - 1:
            e = F.conv1d(gated, p[pre + "conv_residual.weight"], p[pre + "conv_residual.bias"]) + xl
        else:
            skip_sum = skip_sum + F.conv1d(gated, p[pre + "conv_skip.weight"], p[pre + "conv_skip.bias"])
    return dense_head(p, skip_sum)
