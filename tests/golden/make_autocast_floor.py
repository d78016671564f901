"""Reduced-precision floor of the gradients: what the REFERENCE's own mixed-precision recipe costs.

The reference trains under ``torch.cuda.amp.autocast`` (movenet/trainer.py:124).  This script runs the oracle (bit-identical
to the reference, tests/golden/make_golden.py) under ``torch.autocast(bfloat16)`` on the CPU on the same fixtures and records,
per parameter tensor, the relative L2 error of its gradients against the fp32 golden gradients.  tests/test_gpu_bf16.py asserts
that the CUDA tensor-core mode is within max(0.10, 2 x this floor) per tensor instead of a flat tolerance.

    python tests/golden/make_autocast_floor.py        # writes tests/golden/autocast_floor.json
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import full_params, golden_audio, golden_video, load_golden  # noqa: E402
from oracle import wavenet_oracle as orc  # noqa: E402

CASES = ["cfg00", "cfg00_gain", "cfg03", "cfg04_short", "cfg04_full", "testarch_small", "odd", "cfg01_true"]


def main():
    torch.set_num_threads(8)
    out = {}
    for name in CASES:
        fx = load_golden(name)
        shape, p = full_params(fx)
        audio = golden_audio(fx)
        video = golden_video(fx, audio.shape[0]) if "video_seed" in fx else None
        leaf = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
        with torch.autocast(device_type="cpu", dtype=torch.bfloat16):
            loss, _, _, _ = orc.training_loss(leaf, shape, audio, video)
        loss.float().backward()
        errs = {}
        for k, g in fx["grads"].items():
            got = leaf[k].grad.float()
            errs[k] = float((got - g).norm() / g.norm().clamp_min(1e-30))
        out[name] = {"loss_rel": abs(float(loss) - float(fx["loss"])) / abs(float(fx["loss"])),
                     "max": max(errs.values()), "mean": sum(errs.values()) / len(errs), "per_tensor": errs}
        worst = max(errs, key=errs.get)
        print(f"{name:16s} loss_rel {out[name]['loss_rel']:.2e}  grad rel-L2: mean {out[name]['mean']:.3f} max {out[name]['max']:.3f} ({worst})")
    json.dump(out, open(os.path.join(HERE, "autocast_floor.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
