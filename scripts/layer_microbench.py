"""Times the residual-layer kernels alone (forward and backward, CUDA events) on bench.py's workload for one or
more per-GPU batch sizes.  GPU box only.  Used for A/B runs of kernel changes:

    python scripts/layer_microbench.py [--batch 1,3,6] [--dtype bf16] [--audio-only]

Prints one line per batch size: tiles per launch, microseconds per launch, fraction of the measured HBM peak.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import movenet_b200


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", default="3")
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--audio-only", action="store_true")
    args = ap.parse_args()
    w = bench.WORKLOAD
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = movenet_b200.WaveNet(w["layer_size"], w["stack_size"], w["input_channels"], w["residual_channels"],
                                 w["skip_channels"], compute_dtype=args.dtype).to(dev)
    for B in [int(b) for b in args.batch.split(",")]:
        codes = torch.randint(0, w["input_channels"], (B, bench.T_CLIP), device=dev)
        audio = movenet_b200.one_hot(codes, w["input_channels"])
        video = None if args.audio_only else torch.randint(0, 256, (B, 160, 64, 64, 1), device=dev).float()
        bwd, fwd = bench.layer_roofline(model, audio, video, args.dtype)
        tiles = B * ((bench.T_CLIP + 127) // 128)
        print("B=%d tiles=%d  fwd %.1f us (%.3f of peak)  bwd %.1f us (%.3f of peak)" % (
            B, tiles, fwd["ms_per_launch"] * 1e3, fwd["frac"], bwd["ms_per_launch"] * 1e3, bwd["frac"]), flush=True)


if __name__ == "__main__":
    main()
