"""CPU tests of the host layer: the C-ABI library loads and exports every symbol the header
declares, the geometry helpers agree with the reference, and the nn.Module surface (constructor,
parameter names/shapes, error behaviour) is the reference's.  No kernels run here."""
import ctypes
import inspect
import os
import re

import pytest
import torch

from conftest import ROOT, load_golden
import movenet_b200
from movenet_b200 import _lib
from oracle import wavenet_oracle as orc


def header_functions():
    text = open(os.path.join(ROOT, "include", "movenet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mvn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/movenet_b200.h but not exported"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.mvn_version() >= 100


@pytest.mark.parametrize("L,S", [(3, 3), (2, 2), (14, 1), (10, 3), (1, 1)])
def test_geometry_matches_reference_formula(L, S):
    lib = _lib.load()
    rf = orc.Shape(L, S, 8).receptive_fields
    assert lib.mvn_receptive_fields(L, S) == rf
    assert lib.mvn_output_size(L, S, 160000) == 160000 - rf + 1
    assert movenet_b200.WaveNet(L, S, 8, 8, 8).receptive_fields == rf


def test_buffer_sizes_are_reported_without_a_gpu():
    s = _lib.Shape(3, 3, 64, 64, 8, 1, 3, 160000, 1, _lib.F32, 1, 0)
    assert _lib.size("mvn_packed_bytes", s) > 4 * 64 * 4096
    assert _lib.size("mvn_acts_bytes", s) > 9 * 3 * 160000 * 64 * 4
    assert _lib.size("mvn_scratch_bytes", s) > 0
    assert _lib.size("mvn_decode_state_bytes", s, _lib.DECODE_CAUSAL) > 0
    # stack_size == 1: the reference-window mode keeps every layer's inputs back to the window edge
    s1 = _lib.Shape(14, 1, 128, 16, 8, 1, 2, 16384, 0, _lib.F32, 0, 1)
    causal = _lib.size("mvn_decode_state_bytes", s1, _lib.DECODE_CAUSAL)
    window = _lib.size("mvn_decode_state_bytes", s1, _lib.DECODE_REFERENCE)
    assert causal >= 2 * 16383 * 16 * 4 and window > 10 * causal
    s3 = _lib.Shape(3, 3, 64, 64, 8, 1, 2, 24, 0, _lib.F32, 0, 1)
    assert _lib.size("mvn_decode_state_bytes", s3, _lib.DECODE_CAUSAL) == _lib.size("mvn_decode_state_bytes", s3, _lib.DECODE_REFERENCE)


def test_peer_exchange_buffers_cover_the_gradient_regions():
    """csrc/peer.cu: staging + receive areas hold two 16-byte lines per float4 of the gradient regions; the flat gradient
    (reference shapes) can never be larger than the packed regions it is unpacked from"""
    import ctypes as C
    m = movenet_b200.WaveNet(3, 3, 64, 64, 8)
    for video in (0, 1):
        s = _lib.Shape(3, 3, 64, 64, 8, 1, 3, 160000, video, _lib.BF16, 1, 0)
        stage, recv = C.c_size_t(), C.c_size_t()
        _lib.call("mvn_peer_layout", C.byref(s), C.byref(stage), C.byref(recv))
        _, flat_floats = m._grad_layout(bool(video))
        assert stage.value == recv.value and stage.value % 256 == 0
        assert stage.value >= flat_floats * 4 * 2                  # (value, step) pairs: twice the payload
        assert stage.value <= _lib.size("mvn_packed_bytes", s) * 2 + 4096


def test_constructor_signature_is_the_reference_one():
    sig = inspect.signature(movenet_b200.WaveNet.__init__)
    names = [p for p in sig.parameters if p != "self"]
    assert names[:6] == ["layer_size", "stack_size", "input_channels", "residual_channels", "skip_channels",
                         "context_in_channels"]
    assert sig.parameters["residual_channels"].default == 16 and sig.parameters["skip_channels"].default == 16
    assert sig.parameters["context_in_channels"].default == 1
    # anything extra must be keyword-only and defaulted
    for extra in names[6:]:
        assert sig.parameters[extra].kind is inspect.Parameter.KEYWORD_ONLY
        assert sig.parameters[extra].default is not inspect.Parameter.empty
    fwd = list(inspect.signature(movenet_b200.WaveNet.forward).parameters)
    assert fwd == ["self", "audio", "video", "global_features", "output_unnormalized", "remove_last"]
    gen = list(inspect.signature(movenet_b200.WaveNet.generate).parameters)
    assert gen == ["self", "audio", "video", "global_features", "n_samples", "temperature"]


@pytest.mark.parametrize("name", ["cfg00", "video", "cfg04_short"])
def test_state_dict_keys_and_shapes_match_reference(name):
    fx = load_golden(name)
    model = movenet_b200.WaveNet(**fx["shape"])
    sd = model.state_dict()
    ref_like = orc.init_params(orc.Shape(**fx["shape"]), 0, video=True)   # key set asserted == reference in make_golden
    assert list(sd.keys()) == list(ref_like.keys())
    for k, v in ref_like.items():
        assert tuple(sd[k].shape) == tuple(v.shape), k
    for k, v in fx["params"].items():            # tensors saved from the reference itself
        assert tuple(sd[k].shape) == tuple(v.shape), k
    missing, unexpected = model.load_state_dict(fx["params"], strict=False)
    assert not unexpected and all(k.startswith("video_") for k in missing)
    n = model.layer_size * model.stack_size
    assert len(sd) == 13 + 10 * n
    assert model.residual_conv_stack.dilations == orc.Shape(**fx["shape"]).dilations


def test_errors_mirror_the_reference():
    m = movenet_b200.WaveNet(3, 3, 16, 8, 8)
    with pytest.raises(ValueError):                       # movenet/wavenet.py:141-146
        m.compute_output_size(torch.zeros(1, 8, m.receptive_fields - 1))
    assert m.compute_output_size(torch.zeros(1, 8, m.receptive_fields)) == 1
    with pytest.raises(RuntimeError, match="no CPU path"):  # the product path never falls back to the CPU
        m(torch.zeros(1, 16, 100))
    with pytest.raises(RuntimeError):
        m.causal_conv(torch.zeros(1, 16, 100))
    with pytest.raises(RuntimeError, match="CUDA"):
        movenet_b200.mu_law_encoding(torch.zeros(4), 256)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.MovenetB200Error, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "movenet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "/root/reference" not in text, f


def test_mulaw_threshold_tables_reproduce_torchaudio_codes():
    from movenet_b200 import mulaw
    fx = torch.load(os.path.join(ROOT, "tests", "golden", "mulaw.pt"), weights_only=True)
    for A in (64, 128, 256):
        for dt, key in ((torch.float32, "32"), (torch.float64, "64")):
            thr = mulaw._thresholds(A, dt)
            assert bool((thr[1:] > thr[:-1]).all())
            assert torch.equal(torch.searchsorted(thr, fx[A]["x" + key], right=True), fx[A]["codes" + key])
        assert torch.equal(mulaw._decode_lut(A), fx[A]["decode_lut"])


def test_grad_layout_marks_the_parameters_the_reference_leaves_without_grad():
    fx = load_golden("cfg00")
    m = movenet_b200.WaveNet(**fx["shape"])
    offs, total = m._grad_layout(has_video=False)
    names = [n for n, _ in m.named_parameters()]
    none = sorted(n for n, o in zip(names, offs) if o < 0)
    assert none == fx["none_grads"]
    offs_v, _ = m._grad_layout(has_video=True)
    assert sorted(n for n, o in zip(names, offs_v) if o < 0) == sorted(n for n in fx["none_grads"] if "conv_residual" in n)
    assert total >= sum(p.numel() for n, p in m.named_parameters() if n not in none)


def test_movenet_import_shim_resolves_to_this_package():
    """`from movenet.wavenet import WaveNet` (movenet/pytorch_lightning_trainer.py:16) must give the B200 module"""
    import subprocess, sys
    code = ("import sys; sys.path.insert(0, %r); from movenet.wavenet import WaveNet, MAX_AUDIO_FRAMES, MAX_VIDEO_FRAMES, "
            "VIDEO_KERNEL_SIZE; from movenet.modules import GatedResidualConv1d, ResidualConvStack, DenseConv, CausalConv1d, "
            "DilatedCausalConv1d; from movenet.types import AudioTensor, VideoTensor; import movenet_b200; "
            "assert WaveNet is movenet_b200.WaveNet and MAX_AUDIO_FRAMES == 160000 and MAX_VIDEO_FRAMES == 160 "
            "and VIDEO_KERNEL_SIZE == (1, 64, 64); print('ok')") % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/")
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr


def test_flat_gradient_buffer_is_reused_only_when_nothing_aliases_it():
    """WaveNet._flat_grads (host-time optimisation): the flat gradient buffer keeps its address from step to step while no
    parameter holds a gradient; a parameter that still holds one (gradient accumulation), or a second hand-out inside the SAME
    backward pass (several loss terms), gets a fresh buffer."""
    import torch
    m = movenet_b200.WaveNet(2, 1, 8, 8, 8)
    dev = torch.device("cpu")
    flat1, views1 = m._flat_grads(False, dev)
    offsets, total = m._grad_layout(False)
    assert flat1.numel() == total == sum(p.numel() for o, p in zip(offsets, m.parameters()) if o >= 0)
    for o, p, v in zip(offsets, m.parameters(), views1):
        assert (v is None) == (o < 0)
        if v is not None:
            assert v.shape == p.shape and v.data_ptr() == flat1.data_ptr() + 4 * o
    flat2, views2 = m._flat_grads(False, dev)
    assert flat2.data_ptr() == flat1.data_ptr() and all(a is not b for a, b in zip(views1, views2) if a is not None)   # same buffer, fresh views
    held = next(p for o, p in zip(offsets, m.parameters()) if o >= 0)
    held.grad = views2[[i for i, o in enumerate(offsets) if o >= 0][0]]
    flat3, _ = m._flat_grads(False, dev)
    assert flat3.data_ptr() != flat1.data_ptr()            # a parameter still holds a gradient: accumulate into a new buffer
    held.grad = None
    assert m._flat_grads(False, dev)[0].data_ptr() == flat1.data_ptr()

    seen = []

    class TwoHandOuts(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            return x * 2

        @staticmethod
        def backward(ctx, g):
            seen.append(m._flat_grads(False, dev)[0].data_ptr())
            seen.append(m._flat_grads(False, dev)[0].data_ptr())
            return g * 2

    x = torch.ones(3, requires_grad=True)
    TwoHandOuts.apply(x).sum().backward()
    assert seen[0] == flat1.data_ptr() and seen[1] != seen[0]          # the second hand-out of one pass must not alias the first
    assert m._flat_grads(False, dev)[0].data_ptr() == flat1.data_ptr()  # the pass is over: reusable again
