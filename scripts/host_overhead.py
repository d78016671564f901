"""How long does the host take to ENQUEUE one training step (no device sync inside the loop)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import movenet_b200
import bench

w = bench.WORKLOAD
dev = torch.device("cuda", 0)
m = movenet_b200.WaveNet(w["layer_size"], w["stack_size"], w["input_channels"], w["residual_channels"], w["skip_channels"], compute_dtype="bf16").to(dev)
opt = torch.optim.AdamW(m.parameters(), lr=3e-4, fused=True)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 3
T = int(sys.argv[2]) if len(sys.argv) > 2 else 160000
codes = torch.randint(0, 64, (B, T), device=dev)
audio = movenet_b200.one_hot(codes, 64)
video = torch.randint(0, 256, (B, 160, 64, 64, 1), device=dev).float() if T == 160000 else None
for _ in range(3):
    bench.train_step(m, opt, audio, video)
torch.cuda.synchronize()
import cProfile, pstats
n = 20
t0 = time.perf_counter()
for _ in range(n):
    bench.train_step(m, opt, audio, video)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"B={B} T={T}: enqueue {1e3 * (t1 - t0) / n:.3f} ms/step, total {1e3 * (t2 - t0) / n:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    bench.train_step(m, opt, audio, video)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
