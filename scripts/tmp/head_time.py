import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
import bench, movenet_b200
from torch.profiler import profile, ProfilerActivity
w = bench.WORKLOAD
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = movenet_b200.WaveNet(w["layer_size"], w["stack_size"], w["input_channels"], w["residual_channels"], w["skip_channels"], compute_dtype="bf16").to(dev)
B = 3
codes = torch.randint(0, 64, (B, bench.T_CLIP), device=dev)
audio = movenet_b200.one_hot(codes, 64)
video = torch.randint(0, 256, (B, 160, 64, 64, 1), device=dev).float()
def step():
    out = model(audio, video)
    loss = F.cross_entropy(out, codes[:, model.receptive_fields:])
    loss.backward()
for i in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(5): step()
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in rows) / 5
print("kernel time per step (us): %.1f" % tot)
for e in rows[:28]:
    print("%9.1f us  n=%3d  %s" % (e.device_time_total / 5, e.count // 5, e.key[:70]))
