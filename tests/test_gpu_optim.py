"""movenet_b200.optim.AdamW (one launch for the whole model) against torch.optim.AdamW / clip_grad_norm_ on the same
parameters and gradients.  Same fp32 formula, scalars formed in double like torch's; what is left is the rounding of
p (1 - lr wd) - step m / denom as one fused expression vs torch's separate ops: 1 ulp of the parameter per step (measured
3e-8 on |p| ~ 0.3 after one step, 6e-8 after three).  Tolerance: 1e-6 relative + 2e-8 absolute after four steps."""
import copy

import pytest
import torch
import torch.nn.functional as F

from conftest import golden_audio, load_golden
import movenet_b200

pytestmark = pytest.mark.gpu


def _models():
    fx = load_golden("cfg00")
    a = movenet_b200.WaveNet(**fx["shape"], compute_dtype="fp32")
    a.load_state_dict(fx["params"], strict=False)
    a = a.cuda()
    b = copy.deepcopy(a)
    return fx, a, b


@pytest.mark.parametrize("clip", [None, 1e-4])
def test_adamw_matches_torch(clip):
    fx, ma, mb = _models()
    audio = golden_audio(fx).cuda()
    target = audio[:, :, ma.receptive_fields:].argmax(1)
    start = {k: p.detach().clone() for k, p in ma.named_parameters()}
    ours = movenet_b200.optim.AdamW(ma.parameters(), lr=3e-3, weight_decay=0.05, max_grad_norm=clip)
    ref = torch.optim.AdamW(mb.parameters(), lr=3e-3, weight_decay=0.05)
    for _ in range(4):
        for m, opt in ((ma, ours), (mb, ref)):
            opt.zero_grad(set_to_none=True)
            F.cross_entropy(m(audio), target).backward()
        # same gradients into both optimizers (the models drift apart by rounding otherwise)
        for pa, pb in zip(ma.parameters(), mb.parameters()):
            assert (pa.grad is None) == (pb.grad is None)
            if pb.grad is not None:
                pb.grad.copy_(pa.grad)
        if clip:
            norm = torch.nn.utils.clip_grad_norm_(mb.parameters(), clip)
        ours.step()
        ref.step()
        if clip:
            assert abs(ours.grad_norm.item() - norm.item()) <= 1e-5 * norm.item()
        for (k, pa), pb in zip(ma.named_parameters(), mb.parameters()):
            assert torch.allclose(pa, pb, rtol=1e-6, atol=2e-8), (k, (pa - pb).abs().max().item())
    # parameters without a gradient (video branch, last residual conv) are untouched, like in torch
    assert len(fx["none_grads"]) > 0
    for k in fx["none_grads"]:
        assert torch.equal(dict(ma.named_parameters())[k], start[k]), k


def test_adamw_is_an_optimizer():
    _, ma, _ = _models()
    opt = movenet_b200.optim.AdamW(ma.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.5)
    audio = torch.zeros(1, 64, 100, device="cuda"); audio[:, 3] = 1
    F.cross_entropy(ma(audio), audio[:, :, ma.receptive_fields:].argmax(1)).backward()
    opt.step(); sched.step()
    assert abs(opt.param_groups[0]["lr"] - 5e-4) < 1e-12
    sd = opt.state_dict()
    assert len(sd["state"]) > 0 and "exp_avg" in next(iter(sd["state"].values()))
    cpu = torch.nn.Parameter(torch.zeros(3))
    cpu.grad = torch.ones(3)
    with pytest.raises(RuntimeError):
        movenet_b200.optim.AdamW([cpu], lr=1e-3).step()


def test_adamw_per_parameter_step_counts_match_torch():
    """a parameter that only intermittently receives a gradient (the video / context tensors when batches mix video and
    no-video) keeps its own step counter and bias correction, like torch.optim.AdamW; a torch state_dict loads"""
    torch.manual_seed(0)
    pa = [torch.nn.Parameter(torch.randn(257, device="cuda")), torch.nn.Parameter(torch.randn(33, 5, device="cuda"))]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    ours = movenet_b200.optim.AdamW(pa, lr=1e-2, weight_decay=0.01)
    ref = torch.optim.AdamW(pb, lr=1e-2, weight_decay=0.01)
    for step in range(5):
        for i, (a, b) in enumerate(zip(pa, pb)):
            if i == 1 and step % 2 == 1:          # the second parameter skips every other step
                a.grad = None; b.grad = None
                continue
            g = torch.randn_like(a)
            a.grad = g.clone(); b.grad = g.clone()
        ours.step(); ref.step()
        for a, b in zip(pa, pb):
            assert torch.allclose(a, b, rtol=1e-6, atol=2e-8), (step, (a - b).abs().max().item())
    assert ours.state[pa[0]]["step"] == 5 and ours.state[pa[1]]["step"] == 3
    # warm start from torch's optimizer state: moments AND step counters are taken over
    fresh = movenet_b200.optim.AdamW(pa, lr=1e-2, weight_decay=0.01)
    fresh.load_state_dict(copy.deepcopy(ref.state_dict()))     # (load_state_dict may alias the tensors it is given)
    for a, b in zip(pa, pb):
        g = torch.randn_like(a)
        a.grad = g.clone(); b.grad = g.clone()
    fresh.step(); ref.step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=1e-6, atol=2e-8)
