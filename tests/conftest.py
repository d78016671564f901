import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

# cfg04_full: the decode benchmark's architecture at full depth (14 x 1 layers, RF 16384)
AUDIO_CASES = ["cfg00", "cfg00_gain", "cfg03", "cfg04_short", "cfg04_full", "testarch_small", "odd"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, f"wavenet_{name}.pt"), weights_only=True)


def golden_audio(fx):
    """(B, A, T) fp32 one-hot rebuilt from the stored codes"""
    codes = fx["codes"].long()
    A = fx["shape"]["input_channels"]
    return torch.zeros(codes.shape[0], A, codes.shape[1]).scatter_(1, codes.unsqueeze(1), 1.0)


def golden_video(fx, B):
    g = torch.Generator().manual_seed(fx["video_seed"])
    return torch.randint(0, 256, (B, 160, 64, 64, fx["shape"].get("context_in_channels", 1)), generator=g).float()


def full_params(fx):
    """fixture parameters completed with (unused) video parameters so every key exists"""
    from oracle import wavenet_oracle as orc
    shape = orc.Shape(**fx["shape"])
    p = orc.init_params(shape, seed=99, video=True)
    p.update(fx["params"])
    return shape, p
