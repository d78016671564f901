"""Per-kernel device time of one bench.py training step (torch profiler, kernels running back to back as in the bench).
GPU box only:  python scripts/dev/step_profile.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from torch.profiler import ProfilerActivity, profile

import bench
import movenet_b200

w = bench.WORKLOAD
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = movenet_b200.WaveNet(w["layer_size"], w["stack_size"], w["input_channels"], w["residual_channels"],
                             w["skip_channels"], compute_dtype="bf16").to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=3e-4, fused=True)
B = w["batch_per_gpu"]
codes = torch.randint(0, w["input_channels"], (B, bench.T_CLIP), device=dev)
audio = movenet_b200.one_hot(codes, w["input_channels"])
video = torch.randint(0, 256, (B, 160, 64, 64, 1), device=dev).float()
for _ in range(3):
    bench.train_step(model, opt, audio, video)
torch.cuda.synchronize()
N = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        bench.train_step(model, opt, audio, video)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
print("device time per step: %.1f us in %d launches" % (sum(e.device_time_total for e in rows) / N, sum(e.count for e in rows) // N))
for e in rows[:40]:
    print("%9.1f us  %5.1f%%  n=%3d  %s" % (e.device_time_total / N, 100 * e.device_time_total / sum(x.device_time_total for x in rows),
                                           e.count // N, e.key[:90]))
