"""bf16-mode accuracy table against the golden fp32 fixtures: logits, loss, gradient cosine and the three worst per-tensor
relative L2 gradient errors per fixture (GPU box only; MOVENET_B200_GATE_F32=1 selects the fp32 gate epilogue)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from conftest import golden_audio, load_golden
import movenet_b200
for name in ["cfg00", "cfg00_gain", "cfg03", "cfg04_short", "testarch_small", "odd"]:
    fx = load_golden(name)
    m = movenet_b200.WaveNet(**fx["shape"], compute_dtype="bf16"); m.load_state_dict(fx["params"], strict=False); m = m.cuda()
    audio = golden_audio(fx).cuda()
    with torch.no_grad():
        logits = m(audio, output_unnormalized=False)
    ref = fx["logits"]
    lerr = (logits.cpu() - ref).abs().max().item() / ref.abs().max().item()
    out = m(audio); target = audio[:, :, m.receptive_fields:].argmax(1)
    loss = F.cross_entropy(out, target); loss.backward()
    got = dict(m.named_parameters())
    errs = sorted(((((got[k].grad.cpu() - g).norm() / g.norm().clamp_min(1e-30)).item(), k) for k, g in fx["grads"].items()), reverse=True)
    a = torch.cat([got[k].grad.cpu().flatten() for k in fx["grads"]]); b = torch.cat([g.flatten() for g in fx["grads"].values()])
    print(f"{name:15s} logit_rel {lerr:.2e} loss_rel {abs(loss.item()-fx['loss'].item())/abs(fx['loss'].item()):.1e} cos {F.cosine_similarity(a,b,dim=0).item():.5f} worst3 " + " ".join(f"{e:.3f}" for e, _ in errs[:3]))
