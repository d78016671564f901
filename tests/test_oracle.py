"""CPU tests: the oracle against the golden vectors produced by the reference itself
(tests/golden/make_golden.py).  These pin the oracle; the GPU tests then compare
the CUDA path with the oracle / the same vectors."""
import pytest
import torch

from conftest import AUDIO_CASES, full_params, golden_audio, golden_video, load_golden
from oracle import mulaw_oracle
from oracle import wavenet_oracle as orc


@pytest.mark.parametrize("name", AUDIO_CASES)
def test_forward_loss_grads_match_reference(name):
    fx = load_golden(name)
    shape, p = full_params(fx)
    audio = golden_audio(fx)
    loss, probs, grads = orc.loss_and_grads(p, shape, audio)
    assert torch.equal(probs, fx["probs"])
    assert torch.equal(loss, fx["loss"])
    logits = orc.forward(p, shape, audio, output_unnormalized=False)
    assert torch.equal(logits, fx["logits"])
    for k, g in fx["grads"].items():
        assert torch.equal(grads[k], g), k
    for k in fx["none_grads"]:
        assert grads[k] is None, k
    # finding F1: the default forward() returns probabilities
    assert torch.allclose(probs.sum(1), torch.ones_like(probs.sum(1)), atol=1e-5)


@pytest.mark.parametrize("name", ["cfg00", "cfg03", "cfg04_short", "odd"])
def test_generate_matches_reference(name):
    fx = load_golden(name)
    shape, p = full_params(fx)
    audio = golden_audio(fx)
    RF = shape.receptive_fields
    n = RF + fx["gen_codes"].shape[1] - RF
    gen, logits = orc.generate(p, shape, audio[:, :, :RF], None, fx["gen_codes"].shape[1], 0.0, return_logits=True)
    assert torch.equal(gen.argmax(1), fx["gen_codes"].long())
    assert torch.equal(logits, fx["gen_logits"])
    # F5: windowed generate == true causal model only when stack_size >= 2
    causal = orc.causal_logits(p, shape, gen)[:, :, RF - 1:gen.shape[2] - 1]
    diff = (causal - logits).abs().max().item()
    if shape.stack_size >= 2:
        assert diff < 1e-5
    else:
        assert diff > 1e-5   # the zero-padded window edge reaches the output


def test_video_case_matches_patched_reference():
    fx = load_golden("video")
    shape, p = full_params(fx)
    audio = golden_audio(fx)
    video = golden_video(fx, audio.shape[0])
    torch.set_num_threads(max(1, torch.get_num_threads()))
    loss, probs, grads = orc.loss_and_grads(p, shape, audio, video)
    cols = fx["cols"]
    assert torch.equal(probs[:, :, cols], fx["probs_cols"])
    assert torch.equal(loss, fx["loss"])
    for k, g in fx["grads"].items():
        assert torch.equal(grads[k], g), k
    assert fx["none_grads"] == [f"residual_conv_stack.conv_layers.{shape.n_layers - 1}.conv_residual.bias",
                                f"residual_conv_stack.conv_layers.{shape.n_layers - 1}.conv_residual.weight"]


def test_too_short_input_raises():
    shape = orc.Shape(3, 3, 16, 8, 8)
    p = orc.init_params(shape, 0)
    with pytest.raises(ValueError):
        orc.forward(p, shape, torch.zeros(1, 16, shape.receptive_fields - 1))


def test_mulaw_oracle_matches_torchaudio_vectors():
    import os
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "mulaw.pt"), weights_only=True)
    for A in (64, 128, 256):
        assert torch.equal(mulaw_oracle.mu_law_encode(fx[A]["x32"], A), fx[A]["codes32"])
        assert torch.equal(mulaw_oracle.mu_law_encode(fx[A]["x64"], A), fx[A]["codes64"])
        assert torch.equal(mulaw_oracle.mu_law_decode(torch.arange(A), A), fx[A]["decode_lut"])
    # the reference test's fixture (tests/test_model.py:20-27)
    assert fx[256]["codes64"][:8].tolist() == [128, 203, 218, 227, 233, 238, 242, 245]


@pytest.mark.parametrize("name", ["cfg04_short", "cfg04_full", "video_gen"])
def test_window_edge_chain_reproduces_the_reference_generate(name):
    """stack_size == 1 (finding F5): the cached decoder's reference-window algorithm, restated on the CPU
    (oracle.window_edge_logits), gives the logits of the reference's own generate() -- far below the gap between the
    window and the true causal model -- so the CUDA decoder has a pinned definition to be compared with."""
    fx = load_golden(name)
    shape, p = full_params(fx)
    codes = fx["gen_codes"].long()
    gen = torch.zeros(codes.shape[0], shape.input_channels, codes.shape[1]).scatter_(1, codes.unsqueeze(1), 1.0)
    ctx = orc.upsample_video(p, golden_video(fx, codes.shape[0])) if "video_seed" in fx else None
    z = orc.window_edge_logits(p, shape, gen, ctx)[:, :, :-1]
    err = (z - fx["gen_logits"]).abs().max().item()
    assert err < 2e-6, err
    gap = fx["meta"].get("window_vs_causal_maxabs")
    if gap is not None and gap > 1e-5:
        assert err < 1e-2 * gap


def test_benchmarked_shape_fixture_matches_patched_reference():
    """cfg01_true: BASELINE configs[1] at its real size (one 160000-sample clip with video)"""
    fx = load_golden("cfg01_true")
    shape, p = full_params(fx)
    audio = golden_audio(fx)
    video = golden_video(fx, 1)
    loss, probs, grads = orc.loss_and_grads(p, shape, audio, video)
    assert torch.equal(probs[:, :, fx["cols"]], fx["probs_cols"])
    assert torch.equal(loss, fx["loss"])
    for k, g in fx["grads"].items():
        assert torch.equal(grads[k], g), k
