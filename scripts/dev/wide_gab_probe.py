import sys, ctypes as C
import torch
sys.path.insert(0, ".")
import movenet_b200
from movenet_b200 import _lib
torch.manual_seed(0)
kw = dict(layer_size=2, stack_size=1, input_channels=128, residual_channels=128, skip_channels=128)
m = movenet_b200.WaveNet(**kw, compute_dtype="bf16").cuda()
codes = torch.randint(0, 128, (1, 600), device="cuda")
out = m(codes)
st = out._mvn_state
shape = st.bufs.shape
print("no_grad", shape.no_grad, "acts bytes", st.acts.numel())
og = _lib.size("mvn_acts_offset", shape, 3, 0); ogt = _lib.size("mvn_acts_offset", shape, 4, 0)
N, Cc, T = 2, 128, 600
gab = st.acts[og:og + T * N * 2 * Cc * 2].view(torch.bfloat16).view(T, N * 2 * Cc).float()
gated = st.acts[ogt:ogt + T * N * Cc * 2].view(torch.bfloat16).view(T, N * Cc).float()
print("gab abs mean per layer", gab[:, :256].abs().mean().item(), gab[:, 256:].abs().mean().item(), "gated", gated.abs().mean().item())
print(gab[300, :8], gab[300, 256:264])
