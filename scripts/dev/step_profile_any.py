"""One training step of a bench workload inside a cudaProfilerStart/Stop range (after two warm-up steps):
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv \
        python scripts/dev/step_profile_any.py 03w"""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
import bench
import movenet_b200
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "01"]
torch.manual_seed(0)
m = movenet_b200.WaveNet(w["layer_size"], w["stack_size"], w["input_channels"], w["residual_channels"], w["skip_channels"],
                         compute_dtype="bf16").cuda()
opt = movenet_b200.optim.AdamW(m.parameters(), lr=3e-4)
B = w["batch_per_gpu"]
codes = torch.randint(0, w["input_channels"], (B, 160000), device="cuda")
audio = movenet_b200.one_hot(codes, w["input_channels"])
video = torch.randint(0, 256, (B, 160, 64, 64, 1), device="cuda").float() if w["video"] else None
for i in range(3):
    if i == 2:
        torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart()
    bench.train_step(m, opt, audio, video)
torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
print("ok")
