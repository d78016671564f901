"""Host time to ENQUEUE one training step of the cfg01 model (tiny clips: the GPU is never the limiter), the number that decides
whether 8 processes on one host can keep 8 GPUs busy:  python scripts/dev/host_step_time.py [video 0|1]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
import movenet_b200

w = bench.WORKLOADS["01"]
video_on = (sys.argv[1] if len(sys.argv) > 1 else "1") == "1"
m = movenet_b200.WaveNet(w["layer_size"], w["stack_size"], w["input_channels"], w["residual_channels"], w["skip_channels"],
                         compute_dtype="bf16").cuda()
opt = movenet_b200.optim.AdamW(m.parameters(), lr=3e-4)
B, T = 1, 160000 if video_on else 4096
codes = torch.randint(0, 64, (B, T), device="cuda")
audio = movenet_b200.one_hot(codes, 64)
video = torch.randint(0, 256, (B, 160, 64, 64, 1), device="cuda").float() if video_on else None
for _ in range(5):
    bench.train_step(m, opt, audio, video)
torch.cuda.synchronize()
res = []
for rep in range(3):
    n = 30
    t0 = time.perf_counter()
    for _ in range(n):
        bench.train_step(m, opt, audio, video)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    res.append((1e3 * (t1 - t0) / n, 1e3 * (t2 - t0) / n))
print("video=%d side_streams=%s: enqueue %.3f ms/step (device-complete %.3f ms/step)" %
      (video_on, os.environ.get("MOVENET_B200_SIDE_STREAMS", "1"), min(r[0] for r in res), min(r[1] for r in res)))
if os.environ.get("HOST_PROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(20):
        bench.train_step(m, opt, audio, video)
    pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(28)
