"""GPU parity of the tensor-core (bf16) mode: tcgen05 layer kernels, bf16 activations, fp32 accumulate.

Tolerances are the north star's: 2e-2 relative on logits, 1e-3 relative on the loss.  Gradients are
compared per tensor (relative L2) against the golden fp32 gradients."""
import pytest
import torch
import torch.nn.functional as F

from conftest import golden_audio, golden_video, load_golden
import movenet_b200

pytestmark = pytest.mark.gpu

LOGIT_RTOL = 2e-2
LOSS_RTOL = 1e-3
# bf16 activations AND bf16 activation-gradients: every weight gradient is a sum over ~1e2..1e5 time steps of products of
# rounded terms with heavy cancellation.  The tolerance is CALIBRATED, not flat: tests/golden/make_autocast_floor.py runs the
# oracle (bit-identical to the reference) under torch.autocast(bfloat16) -- the reference's own mixed-precision recipe,
# movenet/trainer.py:124 -- on the same fixtures and records its per-tensor relative L2 error against the fp32 golden gradients
# (tests/golden/autocast_floor.json: 1-4 % on the plain fixtures, up to 21 % / 24 % for single bias gradients of the gain-2.5 and
# the 14-layer C = 16 fixtures).  The CUDA path must stay within max(10 %, 2 x that floor) per tensor and max(5 %, 2 x the
# floor's mean) on average, with a cosine similarity >= 0.99 over the whole gradient.
import json
import os

GRAD_RTOL = 0.30          # (only for shapes without a recorded floor: random models built inside a test)
GRAD_MEAN_RTOL = 0.15
GRAD_COS = 0.99
_FLOOR = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "autocast_floor.json")))


def grad_tolerances(name):
    """(per-tensor tolerance dict, mean tolerance) for a golden fixture"""
    fl = _FLOOR[name]
    return {k: max(0.10, 2.0 * v) for k, v in fl["per_tensor"].items()}, max(0.05, 2.0 * fl["mean"])


def build(fx, dtype):
    m = movenet_b200.WaveNet(**fx["shape"], compute_dtype=dtype)
    m.load_state_dict(fx["params"], strict=False)
    return m.cuda()


def rel_l2(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("summed,double_buffered", [("0", "1"), ("0", "0"), ("1", "1")])
@pytest.mark.parametrize("name", ["cfg00", "cfg00_gain", "cfg03", "cfg04_short", "cfg04_full", "testarch_small", "odd"])
def test_bf16_forward_loss_and_grads_against_golden(name, summed, double_buffered, monkeypatch):
    """The three backward layer kernels on the same fixtures.  Default: layers with dilation <= 8 on the double-buffered kernel
    (layer_tc_bwd_db.cu: two-tap box, in-place add-reduction of the context gradient), the others on layer_tc_bwd.cu;
    double_buffered = "0": every layer on layer_tc_bwd.cu; summed = "1": its one-stream variant (MOVENET_B200_BWD_SUM): few tiles per
    clip, so nearly every CTA run is a warm-up tile plus one tile."""
    monkeypatch.setenv("MOVENET_B200_BWD_SUM", summed)
    monkeypatch.setenv("MOVENET_B200_BWD_DB", double_buffered)
    fx = load_golden(name)
    m = build(fx, "bf16")
    audio = golden_audio(fx).cuda()
    with torch.no_grad():
        logits = m(audio, output_unnormalized=False)
    ref = fx["logits"]
    assert (logits.cpu() - ref).abs().max().item() <= LOGIT_RTOL * ref.abs().max().item()
    output = m(audio)
    target = audio[:, :, m.receptive_fields:].argmax(1)
    loss = F.cross_entropy(output, target)
    loss.backward()
    assert abs(loss.item() - fx["loss"].item()) <= LOSS_RTOL * abs(fx["loss"].item())
    got = dict(m.named_parameters())
    tol, mean_tol = grad_tolerances(name)
    errs = []
    for k, g in fx["grads"].items():
        assert got[k].grad is not None, k
        errs.append(rel_l2(got[k].grad.cpu(), g))
        assert errs[-1] < tol[k], (k, errs[-1], tol[k])
    assert sum(errs) / len(errs) < mean_tol, (sum(errs) / len(errs), mean_tol)
    for k in fx["none_grads"]:
        assert got[k].grad is None, k
    a = torch.cat([got[k].grad.cpu().flatten() for k in fx["grads"]])
    b = torch.cat([g.flatten() for g in fx["grads"].values()])
    assert F.cosine_similarity(a, b, dim=0).item() >= GRAD_COS


def test_bf16_benchmarked_shape_against_the_reference():
    """The shape bench.py times (BASELINE configs[1]: A64 C64 S8 3x3, video, T = 160000) in the mode it times (bf16 tensor
    cores: the nchunks == 3 path of the tcgen05 layer kernels, the tensor-core upsampler, head and input kernels), compared
    DIRECTLY with the fixture the patched reference produced at this shape -- not via the fp32 mode."""
    fx = load_golden("cfg01_true")
    m = build(fx, "bf16")
    audio = golden_audio(fx).cuda()
    video = golden_video(fx, 1).cuda()
    cols = fx["cols"].cuda()
    with torch.no_grad():
        logits = m(audio, video, output_unnormalized=False)
    ref = fx["logits_cols"]
    assert (logits[:, :, cols].cpu() - ref).abs().max().item() <= LOGIT_RTOL * ref.abs().max().item()
    output = m(audio, video)
    target = audio[:, :, m.receptive_fields:].argmax(1)
    loss = F.cross_entropy(output, target)
    assert type(loss.grad_fn).__name__.startswith("_FusedLoss")        # the route the bench takes
    loss.backward()
    assert (output.detach()[:, :, cols].cpu() - fx["probs_cols"]).abs().max().item() <= LOGIT_RTOL * fx["probs_cols"].abs().max().item()
    assert abs(loss.item() - fx["loss"].item()) <= LOSS_RTOL * abs(fx["loss"].item())
    got = dict(m.named_parameters())
    tol, mean_tol = grad_tolerances("cfg01_true")
    errs = {}
    for k, g in fx["grads"].items():
        assert got[k].grad is not None, k
        errs[k] = rel_l2(got[k].grad.cpu(), g)
        assert errs[k] < tol[k], (k, errs[k], tol[k])
    assert sum(errs.values()) / len(errs) < mean_tol, (sum(errs.values()) / len(errs), mean_tol)
    for k in fx["none_grads"]:
        assert got[k].grad is None, k
    a = torch.cat([got[k].grad.cpu().flatten() for k in fx["grads"]])
    b = torch.cat([g.flatten() for g in fx["grads"].values()])
    assert F.cosine_similarity(a, b, dim=0).item() >= GRAD_COS


def test_bf16_video_full_clip_against_fp32_mode():
    """C = 64 with video at the full clip length: the tensor-core path against the exact path
    (which tests/test_gpu_parity.py pins to the reference)."""
    torch.manual_seed(0)
    kw = dict(layer_size=3, stack_size=3, input_channels=64, residual_channels=64, skip_channels=8)
    m32 = movenet_b200.WaveNet(**kw, compute_dtype="fp32").cuda()
    m16 = movenet_b200.WaveNet(**kw, compute_dtype="bf16").cuda()
    m16.load_state_dict(m32.state_dict())
    codes = torch.randint(0, 64, (1, 160000), device="cuda")
    audio = movenet_b200.one_hot(codes, 64)
    video = torch.randint(0, 256, (1, 160, 64, 64, 1), device="cuda").float()
    target = codes[:, m32.receptive_fields:]
    outs, grads, losses = [], [], []
    for m in (m32, m16):
        out = m(audio, video, output_unnormalized=False)
        loss = F.cross_entropy(torch.softmax(out, 1), target)
        loss.backward()
        outs.append(out.detach()); losses.append(loss.item())
        grads.append({k: v.grad for k, v in m.named_parameters()})
    assert (outs[0] - outs[1]).abs().max().item() <= LOGIT_RTOL * outs[0].abs().max().item()
    assert abs(losses[0] - losses[1]) <= LOSS_RTOL * abs(losses[0])
    for k, g in grads[0].items():
        if g is None:
            assert grads[1][k] is None
        else:
            assert rel_l2(grads[1][k], g) < GRAD_RTOL, (k, rel_l2(grads[1][k], g))


def test_bf16_ragged_tile_edges():
    """T not a multiple of the 128-row tile, several clips: TMA zero-fill on the left edge of every
    clip (rows t-d < 0 must not read the previous clip) and clipping on the right edge."""
    fx = load_golden("cfg00")
    m32, m16 = build(fx, "fp32"), build(fx, "bf16")
    g = torch.Generator().manual_seed(1)
    for T in (24, 25, 127, 128, 129, 300, 1000):
        codes = torch.randint(0, 64, (3, T), generator=g).cuda()
        audio = movenet_b200.one_hot(codes, 64)
        with torch.no_grad():
            a = m32(audio, output_unnormalized=False, remove_last=False)
            b = m16(audio, output_unnormalized=False, remove_last=False)
        assert a.shape == b.shape == (3, 64, T - 23)
        assert (a - b).abs().max().item() <= LOGIT_RTOL * a.abs().max().item(), T


@pytest.mark.parametrize("name", ["cfg04_short", "cfg03", "cfg04_full"])
def test_tensor_core_decoder_teacher_forced_logits(name):
    """Throughput decoder (tcgen05, bf16 queues) on the receptive-field architecture (stack_size 1, C 16, A 128) and on
    the scale-up one (2 x 2 layers, C 32: the two-group variant of the kernel):
    teacher-forced with the reference's own generated tokens, its logits must match the oracle's true-causal logits
    within the bf16 tolerance at every step; free-running it must emit valid tokens = argmax of its own logits."""
    from movenet_b200.decode import fast_mode_available, prefill, run_steps
    fx = load_golden(name)
    m = build(fx, "fp32")
    RF = m.receptive_fields
    audio = golden_audio(fx).cuda()
    n = fx["gen_codes"].shape[1]
    B = audio.shape[0]
    assert fast_mode_available(m, B, RF)
    forced = fx["gen_codes"][:, RF:].cuda()
    st = prefill(m, audio[:, :, :RF].contiguous(), None, fast=True)
    codes, logits = run_steps(m, st, RF, n - RF, 0.0, return_logits=True, forced=forced)
    assert torch.equal(codes.cpu().long(), fx["gen_codes"][:, RF:].long())
    ref = fx["gen_causal_logits"]                                   # (B, A, n_new)
    got = logits.permute(0, 2, 1).cpu()
    assert (got - ref).abs().max().item() <= LOGIT_RTOL * ref.abs().max().item()
    # free-running through the public entry point
    m.decode_mode = "fast"
    gen = m.generate(audio[:, :, :RF], n_samples=n, temperature=0.0)
    assert gen.shape == (B, m.input_channels, n)
    assert torch.equal(gen.sum(1).cpu(), torch.ones(B, n))
    assert torch.equal(gen[:, :, :RF].cpu(), golden_audio(fx)[:, :, :RF])
    st = prefill(m, audio[:, :, :RF].contiguous(), None, fast=True)
    codes, logits = run_steps(m, st, RF, n - RF, 0.0, return_logits=True)
    assert torch.equal(codes.long(), logits.argmax(2))
    assert torch.equal(codes.long().cpu(), gen[:, :, RF:].argmax(1).cpu())
    # many clips (several CTAs, both groups, a ragged last group)
    big = audio[:1, :, :RF].repeat(300, 1, 1).contiguous()
    g2 = m.generate(big, n_samples=RF + 8, temperature=0.0)
    assert torch.equal(g2[0], g2[299]) and torch.equal(g2[0], g2[128]) and torch.equal(g2[0, :, :RF + 8].cpu(), gen[0, :, :RF + 8].cpu())


def test_bf16_non_one_hot_audio():
    """arbitrary float audio (not one-hot) through the tensor-core mode: the input conv takes its dense path in the
    forward kernel and the one-hot^T x d(h0) weight-gradient kernel puts the real values into its operand tile."""
    fx = load_golden("cfg00")
    m32, m16 = build(fx, "fp32"), build(fx, "bf16")
    g = torch.Generator().manual_seed(5)
    audio = torch.rand(2, 64, 200, generator=g).cuda()
    audio[0, :, 17] = 0
    audio[1, :, 18] = 0; audio[1, 5, 18] = 1
    target = audio[:, :, m32.receptive_fields:].argmax(1)
    grads = []
    for m in (m32, m16):
        out = m(audio)
        F.cross_entropy(out, target).backward()
        grads.append(m.causal_conv.conv.weight.grad.clone())
    assert rel_l2(grads[1], grads[0]) < GRAD_RTOL
    with torch.no_grad():
        a, b = m32(audio, output_unnormalized=False), m16(audio, output_unnormalized=False)
    assert (a - b).abs().max().item() <= LOGIT_RTOL * a.abs().max().item()


@pytest.mark.parametrize("name", ["cfg00", "cfg04_short"])
def test_loss_fused_backward_matches_the_two_step_route(name, monkeypatch):
    """F.cross_entropy(model(audio), target).backward() runs ONE backward with the loss gradient formed inside the head
    kernel (mvn_wavenet_backward_loss); with MOVENET_B200_FUSED_LOSS_BWD=0 the same call materialises d(probabilities)
    (mvn_softmax_ce_bwd) and feeds it to mvn_wavenet_backward.  Same function, same bf16 arithmetic downstream."""
    fx = load_golden(name)
    audio = golden_audio(fx).cuda()
    grads = []
    for fused in ("1", "0"):
        monkeypatch.setenv("MOVENET_B200_FUSED_LOSS_BWD", fused)
        m = build(fx, "bf16")
        out = m(audio)
        target = audio[:, :, m.receptive_fields:].argmax(1)
        loss = F.cross_entropy(out, target)
        assert type(loss.grad_fn).__name__.startswith("_FusedLoss" if fused == "1" else "_SoftmaxCrossEntropy")
        (3.0 * loss).backward()          # a non-unit d(loss) must flow through both routes
        grads.append({k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
    assert grads[0].keys() == grads[1].keys()
    for k in grads[0]:
        assert rel_l2(grads[0][k], grads[1][k]) < 2e-2, (k, rel_l2(grads[0][k], grads[1][k]))


def test_loss_fused_route_only_for_the_plain_call():
    """any other use of the returned probabilities (a slice, non-default arguments) takes torch's ordinary path and still
    differentiates through the network's own node"""
    fx = load_golden("cfg00")
    audio = golden_audio(fx).cuda()
    m = build(fx, "bf16")
    out = m(audio)
    target = audio[:, :, m.receptive_fields:].argmax(1)
    fused = F.cross_entropy(out, target)
    plain = F.cross_entropy(out.as_subclass(torch.Tensor), target)
    smoothed = F.cross_entropy(out, target, label_smoothing=0.1)
    assert type(fused.grad_fn).__name__.startswith("_FusedLoss")
    assert "Fused" not in type(plain.grad_fn).__name__ and "Fused" not in type(smoothed.grad_fn).__name__
    assert abs(fused.item() - plain.item()) <= 1e-5 * abs(plain.item())
    plain.backward()
    assert m.causal_conv.conv.weight.grad is not None


@pytest.mark.parametrize("video", [False, True])
def test_bf16_training_step_is_bit_reproducible(video):
    """every reduction has a fixed order (per-CTA partials + ordered sums, no atomics -- since round 2 also the video encoder's
    split-K GEMM and its weight gradient), and every hand-off between warps, kernels (programmatic dependent launch) and
    proxies is fenced: the same step on the same inputs must give the same BITS, run after run, at the full clip length,
    with and without video."""
    torch.manual_seed(3)
    m = movenet_b200.WaveNet(3, 3, 64, 64, 8, compute_dtype="bf16").cuda()
    B, T = 2, 160000
    codes = torch.randint(0, 64, (B, T), device="cuda")
    audio = movenet_b200.one_hot(codes, 64)
    vid = torch.randint(0, 256, (B, 160, 64, 64, 1), device="cuda").float() if video else None
    target = codes[:, m.receptive_fields:]
    runs = []
    for _ in range(3):
        m.zero_grad(set_to_none=True)
        out = m(audio, vid)
        loss = F.cross_entropy(out, target)
        loss.backward()
        runs.append((out.detach().clone(), loss.detach().clone(),
                     {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}))
    for out, loss, grads in runs[1:]:
        assert torch.equal(out, runs[0][0])
        assert torch.equal(loss, runs[0][1])
        for k, g in grads.items():
            assert torch.equal(g, runs[0][2][k]), k


def _grads_of(m, audio, video, target):
    for p in m.parameters():
        p.grad = None
    out = m(audio, video) if video is not None else m(audio)
    F.cross_entropy(out, target).backward()
    return {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None}


@pytest.mark.parametrize("video,layer_size", [(False, 8), (False, 9), (True, 3), (True, 5), (True, 9)])
def test_backward_gradient_stream_modes(video, layer_size, monkeypatch):
    """The residual-stream gradient between two layer kernels is the pair (P, U) by default; with MOVENET_B200_BWD_SUM=1 it
    is ONE summed tensor when every dilation from the top layer down to the producing one is small enough for the d-row
    carry (<= 128 audio-only, <= 32 with video).  layer_size 8 / 3 / 5: summed everywhere, carries of up to 128 / 4 / 16 rows;
    layer_size 9: a 256 dilation on top, the whole stack stays on the pair.  Ragged T, several clips per CTA run, warm-up
    tiles.  Checked against the exact fp32 mode; both variants of the same step must agree with each other."""
    torch.manual_seed(3)
    kw = dict(layer_size=layer_size, stack_size=2, input_channels=64, residual_channels=64, skip_channels=8)
    m32 = movenet_b200.WaveNet(**kw, compute_dtype="fp32").cuda()
    m16 = movenet_b200.WaveNet(**kw, compute_dtype="bf16").cuda()
    m16.load_state_dict(m32.state_dict())
    T, B = (160000, 1) if video else (40000 + 77, 3)
    codes = torch.randint(0, 64, (B, T), device="cuda")
    audio = movenet_b200.one_hot(codes, 64)
    vid = torch.randint(0, 256, (B, 160, 64, 64, 1), device="cuda").float() if video else None
    target = codes[:, m32.receptive_fields:]
    ref = _grads_of(m32, audio, vid, target)
    monkeypatch.setenv("MOVENET_B200_BWD_SUM", "1")
    got = _grads_of(m16, audio, vid, target)
    monkeypatch.setenv("MOVENET_B200_BWD_SUM", "0")
    pair = _grads_of(m16, audio, vid, target)
    # ... and the pair on the single-buffer kernel everywhere (default: dilations <= 8 on the double-buffered kernel, whose
    # context-gradient sum is accumulated by bf16 add-reductions in place -- a different rounding of the same sum)
    monkeypatch.setenv("MOVENET_B200_BWD_DB", "0")
    single = _grads_of(m16, audio, vid, target)
    monkeypatch.delenv("MOVENET_B200_BWD_DB")
    for k in ref:
        assert rel_l2(single[k], ref[k]) < GRAD_RTOL, (k, "single buffer", rel_l2(single[k], ref[k]))
        assert rel_l2(pair[k], single[k]) < 0.02, (k, "double vs single buffer", rel_l2(pair[k], single[k]))
    assert ref.keys() == got.keys() == pair.keys()
    errs = []
    for k in ref:
        errs.append(rel_l2(got[k], ref[k]))
        assert errs[-1] < GRAD_RTOL, (k, errs[-1])
        assert rel_l2(pair[k], ref[k]) < GRAD_RTOL, (k, "pair", rel_l2(pair[k], ref[k]))
    assert sum(errs) / len(errs) < GRAD_MEAN_RTOL
    a, b, c = (torch.cat([g[k].flatten() for k in ref]) for g in (got, ref, pair))
    assert F.cosine_similarity(a, b, dim=0).item() >= GRAD_COS
    assert F.cosine_similarity(a, c, dim=0).item() >= 0.999


@pytest.mark.parametrize("video,skip_channels,B,T", [(True, 16, 1, 160000), (True, 32, 2, 160000), (False, 32, 3, 20000 + 13),
                                                       (False, 16, 5, 3000)])
def test_double_buffered_backward_with_wider_skip(video, skip_channels, B, T, monkeypatch):
    """layer_tc_bwd_db.cu with skip_channels > 8 (the d(skip) tile then holds [ones | 16 or 32 skip channels], G2 takes two K
    steps from it and the skip rows of the weight-gradient block move), several clips per CTA (clip boundaries inside a CTA's
    tile sequence: the two-tap box zero-fills rows before a clip's start) and fewer tiles than CTAs: against the exact fp32 mode
    and against the single-buffer kernel on the same step."""
    torch.manual_seed(5)
    kw = dict(layer_size=4, stack_size=2, input_channels=64, residual_channels=64, skip_channels=skip_channels)
    m32 = movenet_b200.WaveNet(**kw, compute_dtype="fp32").cuda()
    m16 = movenet_b200.WaveNet(**kw, compute_dtype="bf16").cuda()
    m16.load_state_dict(m32.state_dict())
    codes = torch.randint(0, 64, (B, T), device="cuda")
    audio = movenet_b200.one_hot(codes, 64)
    vid = torch.randint(0, 256, (B, 160, 64, 64, 1), device="cuda").float() if video else None
    target = codes[:, m32.receptive_fields:]
    ref = _grads_of(m32, audio, vid, target)
    got = _grads_of(m16, audio, vid, target)
    monkeypatch.setenv("MOVENET_B200_BWD_DB", "0")
    single = _grads_of(m16, audio, vid, target)
    assert ref.keys() == got.keys() == single.keys()
    for k in ref:
        assert rel_l2(got[k], ref[k]) < GRAD_RTOL, (k, rel_l2(got[k], ref[k]))
        assert rel_l2(got[k], single[k]) < 0.02, (k, "double vs single buffer", rel_l2(got[k], single[k]))
    a, b = (torch.cat([g[k].flatten() for k in ref]) for g in (got, ref))
    assert F.cosine_similarity(a, b, dim=0).item() >= GRAD_COS


def test_tensor_core_decoder_sampling_is_seeded_and_valid():
    """temperature > 0 in the throughput decoder: tokens are valid codes, identical clips with the same seed draw
    per-clip streams (clip index enters the counter), the same torch seed reproduces them."""
    from movenet_b200.decode import prefill, run_steps
    fx = load_golden("cfg04_short")
    m = build(fx, "fp32")
    RF = m.receptive_fields
    audio = golden_audio(fx).cuda()[:1, :, :RF].repeat(256, 1, 1).contiguous()
    st = prefill(m, audio, None, fast=True)
    greedy = run_steps(m, st, RF, 12, 0.0)
    assert len({tuple(r.tolist()) for r in greedy.cpu()}) == 1   # argmax decoding: identical prompts, identical clips
    st = prefill(m, audio, None, fast=True)
    torch.manual_seed(11)
    hot = run_steps(m, st, RF, 12, 5.0)
    assert int(hot.min()) >= 0 and int(hot.max()) < m.input_channels
    assert len({tuple(r.tolist()) for r in hot.cpu()}) > 1   # identical prompts, different clips: different draws
    st = prefill(m, audio, None, fast=True)
    torch.manual_seed(11)
    again = run_steps(m, st, RF, 12, 5.0)
    assert torch.equal(hot, again)                        # same torch seed: same streams
