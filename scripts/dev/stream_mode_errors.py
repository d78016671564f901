"""per-tensor gradient error of the bf16 mode against the exact fp32 mode, summed-stream vs forced (P, U) pair backward"""
import os, sys
import torch, torch.nn.functional as F
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import movenet_b200

def grads_of(m, audio, video, target):
    for p in m.parameters(): p.grad = None
    out = m(audio, video) if video is not None else m(audio)
    F.cross_entropy(out, target).backward()
    return {k: v.grad.clone() for k, v in m.named_parameters() if v.grad is not None}

rel = lambda a, b: ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
video = len(sys.argv) > 1 and sys.argv[1] == "video"
torch.manual_seed(3)
kw = dict(layer_size=int(os.environ.get("LS", 9)), stack_size=2, input_channels=64, residual_channels=64, skip_channels=8)
m32 = movenet_b200.WaveNet(**kw, compute_dtype="fp32").cuda()
m16 = movenet_b200.WaveNet(**kw, compute_dtype="bf16").cuda()
m16.load_state_dict(m32.state_dict())
T, B = (160000, 1) if video else (40000 + 77, 3)
codes = torch.randint(0, 64, (B, T), device="cuda")
audio = movenet_b200.one_hot(codes, 64)
vid = torch.randint(0, 256, (B, 160, 64, 64, 1), device="cuda").float() if video else None
target = codes[:, m32.receptive_fields:]
ref = grads_of(m32, audio, vid, target)
os.environ["MOVENET_B200_BWD_SUM"] = "1"
got = grads_of(m16, audio, vid, target)
os.environ["MOVENET_B200_BWD_SUM"] = "0"
pair = grads_of(m16, audio, vid, target)
for k in ref:
    print("%-70s sum %.4f pair %.4f sum-vs-pair %.4f |g| %.3e" % (k, rel(got[k], ref[k]), rel(pair[k], ref[k]), rel(got[k], pair[k]), ref[k].norm().item()))
