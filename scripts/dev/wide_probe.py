"""tiny forward (+ backward) through the wide path, for compute-sanitizer / debugging"""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
import movenet_b200
torch.manual_seed(0)
m = movenet_b200.WaveNet(2, 1, 128, 128, 128, compute_dtype="bf16").cuda()
codes = torch.randint(0, 128, (1, 600), device="cuda")
with torch.no_grad():
    out = m(codes)
torch.cuda.synchronize()
print("fwd ok", out.sum().item())
if len(sys.argv) > 1:
    out = m(codes)
    F.cross_entropy(out, codes[:, m.receptive_fields:]).backward()
    torch.cuda.synchronize()
    print("bwd ok")
