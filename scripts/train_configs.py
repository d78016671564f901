"""Training-step throughput of the other BASELINE.json shapes (not the bench line): audio-only cfg00/02, cfg03, cfg04."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import movenet_b200

CONFIGS = {
    "cfg00_audio": dict(layer_size=3, stack_size=3, input_channels=64, residual_channels=64, skip_channels=8, B=3),
    "cfg03": dict(layer_size=2, stack_size=2, input_channels=128, residual_channels=32, skip_channels=8, B=3),
    "cfg04": dict(layer_size=14, stack_size=1, input_channels=128, residual_channels=16, skip_channels=8, B=2),
    "testarch": dict(layer_size=10, stack_size=3, input_channels=256, residual_channels=64, skip_channels=64, B=4),
}
dev = torch.device("cuda", 0)
for name, kw in CONFIGS.items():
    kw = dict(kw); B = kw.pop("B")
    for dtype in ("fp32", "bf16"):
        m = movenet_b200.WaveNet(**kw, compute_dtype=dtype).to(dev)
        opt = torch.optim.AdamW(m.parameters(), lr=3e-4, fused=True)
        T = 160000
        audio = movenet_b200.one_hot(torch.randint(0, kw["input_channels"], (B, T), device=dev), kw["input_channels"])

        def step():
            opt.zero_grad(set_to_none=True)
            out = m(audio)
            loss = F.cross_entropy(out, audio[:, :, m.receptive_fields:].argmax(1))
            loss.backward(); opt.step()
            return loss
        for _ in range(2): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): loss = step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f"{name:12s} {dtype}: {ms:8.3f} ms/step  {B * T / ms / 1e3:9.1f} M samples/s  loss {loss.item():.4f}", flush=True)
        del m, opt, audio
