"""GPU parity of the wide-channel tensor-core path (residual_channels >= 128: the widened scale-up shape, BASELINE
configs[3] / SURVEY "03w"): weight-streaming tcgen05 GEMMs with fused epilogues (csrc/wide_gemm.cuh, csrc/wide.cu), both as
CTA pairs (cta_group::2, the default) and as single-CTA tiles, against the oracle on the same seeded inputs."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

import movenet_b200
from movenet_b200 import _lib
from oracle import wavenet_oracle as orc

pytestmark = pytest.mark.gpu

LOGIT_RTOL = 2e-2       # north star: 2e-2 relative on bf16 logits
LOSS_RTOL = 1e-3
GRAD_RTOL = 0.10        # per-tensor relative L2 of the bf16-mode gradients against the fp32 oracle
GRAD_COS = 0.995


def rel_l2(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def make(kw, seed, B, extra):
    shape = orc.Shape(**kw)
    p = orc.init_params(shape, seed=seed, video=True)
    m = movenet_b200.WaveNet(**kw, compute_dtype="bf16")
    m.load_state_dict(p)
    m.cuda()
    T = shape.receptive_fields + extra
    codes = torch.randint(0, kw["input_channels"], (B, T), generator=torch.Generator().manual_seed(seed + 1))
    audio = torch.zeros(B, kw["input_channels"], T).scatter_(1, codes.unsqueeze(1), 1.0)
    return shape, p, m, codes, audio


@pytest.mark.parametrize("pair", ["2", "1"])
@pytest.mark.parametrize("kw,B,extra", [
    (dict(layer_size=3, stack_size=2, input_channels=256, residual_channels=128, skip_channels=128), 2, 777),
    (dict(layer_size=4, stack_size=2, input_channels=256, residual_channels=256, skip_channels=256), 1, 1300),
    (dict(layer_size=2, stack_size=1, input_channels=128, residual_channels=256, skip_channels=128), 3, 255),
])
def test_wide_path_forward_loss_and_grads_against_the_oracle(kw, B, extra, pair, monkeypatch):
    monkeypatch.setenv("MOVENET_B200_WIDE_PAIR", pair)
    shape, p, m, codes, audio = make(kw, 11, B, extra)
    s = m._shape(B, audio.shape[2], False, True, False)
    assert _lib.load().mvn_kernel_path(C.byref(s)) == 2, "this shape must run on the wide tensor-core path"
    RF = shape.receptive_fields
    with torch.no_grad():
        logits = m(audio.cuda(), output_unnormalized=False)
    ref_logits = orc.forward(p, shape, audio, output_unnormalized=False)
    assert logits.shape == ref_logits.shape
    assert (logits.cpu() - ref_logits).abs().max().item() <= LOGIT_RTOL * ref_logits.abs().max().item()
    out = m(audio.cuda())
    target = audio.cuda()[:, :, RF:].argmax(1)
    loss = F.cross_entropy(out, target)
    assert type(loss.grad_fn).__name__.startswith("_FusedLoss")
    loss.backward()
    o_loss, o_out, o_grads = orc.loss_and_grads(p, shape, audio)
    assert (out.detach().cpu() - o_out).abs().max().item() <= LOGIT_RTOL * o_out.abs().max().item()
    assert (out.detach().sum(1) - 1).abs().max().item() < 1e-4
    assert abs(loss.item() - o_loss.item()) <= LOSS_RTOL * abs(o_loss.item())
    got = dict(m.named_parameters())
    errs = {}
    for k, g in o_grads.items():
        if g is None or k.startswith("video_") or ".context_conv_" in k:
            assert got[k].grad is None, k
            continue
        assert got[k].grad is not None, k
        errs[k] = rel_l2(got[k].grad.cpu(), g)
    worst = max(errs, key=errs.get)
    assert errs[worst] < GRAD_RTOL, (worst, errs[worst])
    a = torch.cat([got[k].grad.cpu().flatten() for k in errs])
    b = torch.cat([o_grads[k].flatten() for k in errs])
    assert F.cosine_similarity(a, b, dim=0).item() >= GRAD_COS
    # the two-step route (d(out) materialised) must agree with the loss-fused one
    m.zero_grad(set_to_none=True)
    out2 = m(audio.cuda())
    F.cross_entropy(out2.as_subclass(torch.Tensor), target).backward()
    for k in errs:
        assert rel_l2(dict(m.named_parameters())[k].grad.cpu(), o_grads[k]) < GRAD_RTOL, k


def test_wide_path_pair_and_single_agree_bitwise_on_forward():
    """the CTA-pair (cta_group::2) and single-CTA variants run the same MMAs in the same order per output element"""
    import os
    kw = dict(layer_size=3, stack_size=2, input_channels=256, residual_channels=256, skip_channels=128)
    shape, p, m, codes, audio = make(kw, 5, 2, 1000)
    outs = []
    for pair in ("2", "1"):
        os.environ["MOVENET_B200_WIDE_PAIR"] = pair
        with torch.no_grad():
            outs.append(m(audio.cuda(), output_unnormalized=False).clone())
    os.environ.pop("MOVENET_B200_WIDE_PAIR")
    assert torch.equal(outs[0], outs[1])


def test_wide_path_full_clip_properties():
    """T = 160000 (the benchmark's clip length): probabilities sum to one, causality, batch independence"""
    torch.manual_seed(0)
    m = movenet_b200.WaveNet(3, 2, 256, 256, 256, compute_dtype="bf16").cuda()
    T = 160000
    codes = torch.randint(0, 256, (2, T), device="cuda")
    with torch.no_grad():
        p = m(codes)
        assert p.shape == (2, 256, T - m.receptive_fields)
        assert (p.sum(1) - 1).abs().max().item() < 1e-4
        codes2 = codes.clone(); codes2[:, 100000:] = (codes2[:, 100000:] + 1) % 256
        p2 = m(codes2)
        assert torch.equal(p[:, :, :100000 - m.receptive_fields], p2[:, :, :100000 - m.receptive_fields])
        assert not torch.equal(p[:, :, 100000:], p2[:, :, 100000:])
        assert torch.equal(m(codes[1:2])[0], p[1])


@pytest.mark.parametrize("kw", [
    dict(layer_size=4, stack_size=2, input_channels=256, residual_channels=64, skip_channels=64),    # the reference's test architecture (tests/test_model.py:42-48), shortened
    dict(layer_size=3, stack_size=1, input_channels=128, residual_channels=32, skip_channels=64),    # narrow stack (zero-padded to 64), A = 128
])
def test_reference_test_architecture_head_runs_on_the_wide_engine(kw):
    """A = 256 / S = 64 (the architecture of the reference's own test): the residual stack runs on the fused C <= 64 tcgen05
    kernels, the DenseConv head + softmax and their backward on the wide engine's head kernels (not on CUDA-core GEMMs)"""
    shape, p, m, codes, audio = make(kw, 21, 2, 900)
    s = m._shape(2, audio.shape[2], False, True, False)
    assert _lib.load().mvn_kernel_path(C.byref(s)) == 1
    assert _lib.load().mvn_fused_loss_supported(C.byref(s)) == 1
    RF = shape.receptive_fields
    with torch.no_grad():
        logits = m(audio.cuda(), output_unnormalized=False)
    ref_logits = orc.forward(p, shape, audio, output_unnormalized=False)
    assert (logits.cpu() - ref_logits).abs().max().item() <= LOGIT_RTOL * ref_logits.abs().max().item()
    n0 = _lib.load().mvn_launch_count()
    out = m(audio.cuda())
    target = audio.cuda()[:, :, RF:].argmax(1)
    loss = F.cross_entropy(out, target)
    assert type(loss.grad_fn).__name__.startswith("_FusedLoss")
    loss.backward()
    o_loss, o_out, o_grads = orc.loss_and_grads(p, shape, audio)
    assert (out.detach().cpu() - o_out).abs().max().item() <= LOGIT_RTOL * o_out.abs().max().item()
    assert abs(loss.item() - o_loss.item()) <= LOSS_RTOL * abs(o_loss.item())
    got = dict(m.named_parameters())
    errs = {k: rel_l2(got[k].grad.cpu(), g) for k, g in o_grads.items()
            if g is not None and not k.startswith("video_") and ".context_conv_" not in k}
    worst = max(errs, key=errs.get)
    assert errs[worst] < 0.15, (worst, errs[worst])
    a = torch.cat([got[k].grad.cpu().flatten() for k in errs])
    b = torch.cat([o_grads[k].flatten() for k in errs])
    assert F.cosine_similarity(a, b, dim=0).item() >= GRAD_COS
