import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
import bench, movenet_b200
w = bench.WORKLOAD
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = movenet_b200.WaveNet(w["layer_size"], w["stack_size"], w["input_channels"], w["residual_channels"], w["skip_channels"], compute_dtype="bf16").to(dev)
B = 3
codes = torch.randint(0, 64, (B, bench.T_CLIP), device=dev)
audio = movenet_b200.one_hot(codes, 64)
for i in range(2):
    out = model(audio, None)
    loss = F.cross_entropy(out, codes[:, model.receptive_fields:])
    loss.backward()
torch.cuda.synchronize()
