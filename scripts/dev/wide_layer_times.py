"""03w-shaped model with few layers: one forward + backward, for an ncu launch list of the wide kernels
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file x.csv python scripts/dev/wide_layer_times.py"""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
import movenet_b200
torch.manual_seed(0)
m = movenet_b200.WaveNet(4, 1, 256, 256, 256, compute_dtype="bf16").cuda()
codes = torch.randint(0, 256, (1, 160000), device="cuda")
for _ in range(2):
    m.zero_grad(set_to_none=True)
    out = m(codes)
    F.cross_entropy(out, codes[:, m.receptive_fields:]).backward()
torch.cuda.synchronize()
print("ok")
