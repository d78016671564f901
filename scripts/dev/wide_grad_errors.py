"""per-tensor gradient errors of the wide path against the oracle (debugging)"""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import movenet_b200
from oracle import wavenet_oracle as orc
C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
kw = dict(layer_size=2, stack_size=1, input_channels=128, residual_channels=C, skip_channels=C)
shape = orc.Shape(**kw)
p = orc.init_params(shape, seed=11, video=True)
m = movenet_b200.WaveNet(**kw, compute_dtype="bf16"); m.load_state_dict(p); m.cuda()
T = shape.receptive_fields + 500
codes = torch.randint(0, 128, (1, T), generator=torch.Generator().manual_seed(3))
audio = torch.zeros(1, 128, T).scatter_(1, codes.unsqueeze(1), 1.0)
out = m(audio.cuda())
loss = F.cross_entropy(out, audio.cuda()[:, :, shape.receptive_fields:].argmax(1))
loss.backward()
o_loss, o_out, o_grads = orc.loss_and_grads(p, shape, audio)
for k, v in m.named_parameters():
    if v.grad is None or o_grads[k] is None: continue
    e = ((v.grad.cpu() - o_grads[k]).norm() / o_grads[k].norm()).item()
    print(f"{e:8.4f} {k}")
