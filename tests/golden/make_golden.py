"""Generate the golden fixtures in this directory from the REFERENCE itself.

Run in the authoring container only (needs /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py

It imports cosmicBboy/movenet's own ``movenet.wavenet.WaveNet`` (with a stub for
the missing ``torchtyping`` package), runs it on CPU fp32 with fixed seeds and
stores inputs, parameters and outputs.  While doing so it also checks the
oracle (oracle/wavenet_oracle.py, oracle/mulaw_oracle.py) against the
reference and refuses to write fixtures if they disagree.

Video fixtures use the reference plus the one-line right-aligned context crop
(the unmodified reference raises at movenet/modules.py:76); that is recorded
in each fixture under ``meta['reference_patch']``.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("MOVENET_REF", "/root/reference")


def import_reference():
    if "torchtyping" not in sys.modules:
        stub = types.ModuleType("torchtyping")

        class TensorType:  # only used in annotations by the reference
            def __class_getitem__(cls, item):
                return cls

        stub.TensorType = TensorType
        sys.modules["torchtyping"] = stub
    sys.path.insert(0, REF)
    import movenet.modules as ref_modules
    import movenet.wavenet as ref_wavenet
    return ref_wavenet, ref_modules


def patch_video_crop(ref_modules):
    """Reference + right-aligned context crop (finding F3), out-of-place adds."""

    def forward(self, input, context, skip_size):
        f, g = self.conv_filter(input), self.conv_gate(input)
        if context is not None:
            ctx = context[:, :, -f.size(2):]
            f = f + self.context_conv_filter(ctx)
            g = g + self.context_conv_gate(ctx)
        gated = torch.tanh(f) * torch.sigmoid(g)
        residual = self.conv_residual(gated)
        residual = residual + input[:, :, -residual.size(2):]
        skip = self.conv_skip(gated)
        return residual, skip[:, :, -skip_size:]

    ref_modules.GatedResidualConv1d.forward = forward


def synth_codes(batch, n, A, seed):
    """Two tones + noise, min-max normalised, mu-law coded."""
    from torchaudio.functional import mu_law_encoding
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n, dtype=torch.float32) / 16000.0
    rows = []
    for b in range(batch):
        w = (0.6 * torch.sin(2 * np.pi * (220.0 + 30 * b) * t)
             + 0.3 * torch.sin(2 * np.pi * 3520.0 * t)
             + 0.1 * (torch.rand(n, generator=g) * 2 - 1))
        w = 2 * (w - w.min()) / (w.max() - w.min()) - 1
        rows.append(mu_law_encoding(w, A))
    return torch.stack(rows)


def one_hot(codes, A):
    return torch.zeros(codes.shape[0], A, codes.shape[1]).scatter_(1, codes.unsqueeze(1), 1.0)


def make_mulaw(out_dir):
    from torchaudio.functional import mu_law_decoding, mu_law_encoding
    from oracle import mulaw_oracle
    fx = {"meta": {"torchaudio": str(__import__("torchaudio").__version__), "torch": str(torch.__version__)}}
    sine64 = torch.from_numpy(np.sin(np.arange(0, 400, 0.1)))          # tests/test_model.py:20-27
    g = torch.Generator().manual_seed(7)
    for A in (64, 128, 256):
        grid = torch.linspace(-1, 1, 4001, dtype=torch.float32)
        rnd = torch.rand(4096, generator=g, dtype=torch.float32) * 2 - 1
        edge = torch.tensor([-1.0, 1.0, 0.0, -0.0, 1e-8, -1e-8, 1e-3, -1e-3, 0.5, -0.5,
                             0.999999, -0.999999], dtype=torch.float32)
        x32 = torch.cat([grid, rnd, edge])
        x64 = torch.cat([sine64, x32.double()])
        c32 = mu_law_encoding(x32, A)
        c64 = mu_law_encoding(x64, A)
        assert torch.equal(c32, mulaw_oracle.mu_law_encode(x32, A))
        assert torch.equal(c64, mulaw_oracle.mu_law_encode(x64, A))
        lut = mu_law_decoding(torch.arange(A), A)
        assert torch.equal(lut, mulaw_oracle.mu_law_decode(torch.arange(A), A))
        fx[A] = {"x32": x32, "codes32": c32, "x64": x64, "codes64": c64, "decode_lut": lut}
    assert fx[256]["codes64"][:8].tolist() == [128, 203, 218, 227, 233, 238, 242, 245]
    torch.save(fx, os.path.join(out_dir, "mulaw.pt"))
    print("mulaw.pt ok")


def make_wavenet_case(ref_wavenet, name, shape_kw, B, extra_T, seed, out_dir, gain=1.0,
                      gen_new=0, video=False, sample_cols=None):
    from oracle import wavenet_oracle as orc
    torch.manual_seed(seed)
    model = ref_wavenet.WaveNet(**shape_kw)
    shape = orc.Shape(**shape_kw)
    assert model.receptive_fields == shape.receptive_fields
    assert model.residual_conv_stack.dilations == shape.dilations
    if gain != 1.0:
        with torch.no_grad():
            for k, v in model.named_parameters():
                if k.endswith("weight"):
                    v.mul_(gain)
    RF = shape.receptive_fields
    A = shape.input_channels
    T = orc.MAX_AUDIO_FRAMES if video else RF + extra_T
    codes = synth_codes(B, T, A, 1234 + seed)
    audio = one_hot(codes, A)
    vid = None
    if video:
        g = torch.Generator().manual_seed(4321)
        vid = torch.randint(0, 256, (B, 160, 64, 64, shape.context_in_channels), generator=g).float()

    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    if not video:
        sd = {k: v for k, v in sd.items() if not k.startswith("video_")}
    # oracle param key set == reference key set
    assert set(orc.init_params(shape, 0, video=True).keys()) == set(model.state_dict().keys())

    model.zero_grad()
    probs = model(audio, vid)
    target = audio[:, :, RF:].argmax(1)
    loss = F.cross_entropy(probs, target)
    loss.backward()
    grads = {k: (None if v.grad is None else v.grad.detach().clone()) for k, v in model.named_parameters()}
    with torch.no_grad():
        logits = model(audio, vid, output_unnormalized=False)

    # --- oracle vs reference ---
    p = {k: v.detach().clone() for k, v in model.state_dict().items()}
    o_loss, o_probs, o_grads = orc.loss_and_grads(p, shape, audio, vid)
    with torch.no_grad():
        o_logits = orc.forward(p, shape, audio, vid, output_unnormalized=False)
    assert torch.equal(o_logits, logits), f"{name}: oracle logits differ from the reference"
    assert torch.equal(o_probs, probs.detach()), f"{name}: oracle probs differ"
    assert torch.equal(o_loss, loss.detach()), f"{name}: oracle loss differs"
    for k, gref in grads.items():
        go = o_grads[k]
        if gref is None:
            assert go is None, k
        else:
            assert torch.equal(go, gref), f"{name}: oracle grad {k} differs"

    fx = {
        "meta": {"name": name, "torch": str(torch.__version__), "seed": seed, "gain": gain,
                 "reference_patch": "context right-aligned crop (F3)" if video else "none"},
        "shape": dict(shape_kw), "codes": codes.to(torch.int16), "params": sd,
        "loss": loss.detach(), "target": target.to(torch.int16),
        "grads": {k: v for k, v in grads.items() if v is not None and (video or not k.startswith("video_"))},
        "none_grads": sorted(k for k, v in grads.items() if v is None),
    }
    if video:
        fx["video_seed"] = 4321
        cols = torch.from_numpy(np.random.RandomState(0).choice(T - RF, sample_cols, replace=False)).sort().values
        fx["cols"] = cols
        fx["logits_cols"] = logits[:, :, cols].clone()
        fx["probs_cols"] = probs.detach()[:, :, cols].clone()
        with torch.no_grad():
            ctx = model.upsample_video(vid)
            assert torch.equal(ctx, orc.upsample_video(p, vid))
        fx["ctx_cols"] = ctx[:, :, cols].clone()
    else:
        fx["logits"] = logits
        fx["probs"] = probs.detach()

    if gen_new and video:
        # finding F4: the reference's generate() cannot run with video (its upsampled context is always 160000 frames
        # long, the window RF).  OUR definition (oracle.generate): the window [i-RF, i) of the upsampled context
        # conditions step i -- the same alignment forward() uses for column i-1.  Pinned by the oracle only.
        with torch.no_grad():
            n = RF + gen_new
            o_gen, o_glog = orc.generate(p, shape, audio[:, :, :RF], ctx[:, :, :n], n, 0.0, return_logits=True)
            fx["gen_codes"] = o_gen.argmax(1).to(torch.int16)
            fx["gen_logits"] = o_glog
            fx["meta"]["generate"] = "oracle only (reference raises, F4): context window [i-RF, i)"
    elif gen_new:
        with torch.no_grad():
            n = RF + gen_new
            gen = model.generate(audio[:, :, :RF], n_samples=n, temperature=0.0)
            o_gen, o_glog = orc.generate(p, shape, audio[:, :, :RF], None, n, 0.0, return_logits=True)
            assert torch.equal(gen, o_gen), f"{name}: oracle generate differs"
            fx["gen_codes"] = gen.argmax(1).to(torch.int16)
            fx["gen_logits"] = o_glog
            # finding F5: windowed logits vs the true causal model on the same tokens
            cl = orc.causal_logits(p, shape, gen)[:, :, RF - 1:n - 1].clone()
            fx["gen_causal_logits"] = cl
            fx["meta"]["window_vs_causal_maxabs"] = float((cl - o_glog).abs().max())
    path = os.path.join(out_dir, f"wavenet_{name}.pt")
    torch.save(fx, path)
    print(f"{name}: loss={loss.item():.6f} RF={RF} T={T} "
          f"{os.path.getsize(path) / 1e6:.2f} MB", fx["meta"].get("window_vs_causal_maxabs"))


def main():
    torch.set_num_threads(8)
    only = None
    for a in sys.argv[1:]:
        if a.startswith("--only="):
            only = set(a[len("--only="):].split(","))
    ref_wavenet, ref_modules = import_reference()
    if only is not None:
        # round-2 additions, generated without touching the round-1 files
        if "cfg04_full" in only:
            # the decode benchmark's architecture at FULL depth (experiments/04_kinetics_receptive_field.mk:58-71):
            # 14 x 1 layers, RF 16384 -- ring slots up to d = 8192 and the stack_size == 1 window edge (finding F5)
            make_wavenet_case(ref_wavenet, "cfg04_full", dict(layer_size=14, stack_size=1, input_channels=128,
                              residual_channels=16, skip_channels=8), B=2, extra_T=300, seed=8, out_dir=HERE, gen_new=32,
                              gain=1.5)
        patch_video_crop(ref_modules)
        if "cfg01_true" in only:
            # the benchmarked training shape itself (experiments/01_audio_video_debug.mk:10-17), one full-length clip
            make_wavenet_case(ref_wavenet, "cfg01_true", dict(layer_size=3, stack_size=3, input_channels=64,
                              residual_channels=64, skip_channels=8), B=1, extra_T=0, seed=9, out_dir=HERE,
                              video=True, sample_cols=1024, gen_new=32)
        if "video_gen" in only:
            make_wavenet_case(ref_wavenet, "video_gen", dict(layer_size=3, stack_size=1, input_channels=32,
                              residual_channels=16, skip_channels=8), B=2, extra_T=0, seed=10, out_dir=HERE,
                              video=True, sample_cols=256, gen_new=40)
        return
    make_mulaw(HERE)
    # audio-only cases run on the UNMODIFIED reference
    make_wavenet_case(ref_wavenet, "cfg00", dict(layer_size=3, stack_size=3, input_channels=64,
                      residual_channels=64, skip_channels=8), B=2, extra_T=257, seed=0, out_dir=HERE, gen_new=48)
    make_wavenet_case(ref_wavenet, "cfg00_gain", dict(layer_size=3, stack_size=3, input_channels=64,
                      residual_channels=64, skip_channels=8), B=2, extra_T=130, seed=1, out_dir=HERE, gain=2.5,
                      gen_new=48)
    make_wavenet_case(ref_wavenet, "cfg03", dict(layer_size=2, stack_size=2, input_channels=128,
                      residual_channels=32, skip_channels=8), B=3, extra_T=200, seed=2, out_dir=HERE, gen_new=32)
    make_wavenet_case(ref_wavenet, "cfg04_short", dict(layer_size=6, stack_size=1, input_channels=128,
                      residual_channels=16, skip_channels=8), B=2, extra_T=150, seed=3, out_dir=HERE, gen_new=40,
                      gain=2.0)
    make_wavenet_case(ref_wavenet, "testarch_small", dict(layer_size=4, stack_size=3, input_channels=256,
                      residual_channels=32, skip_channels=32), B=2, extra_T=100, seed=4, out_dir=HERE, gen_new=24)
    make_wavenet_case(ref_wavenet, "odd", dict(layer_size=2, stack_size=2, input_channels=40,
                      residual_channels=24, skip_channels=12), B=1, extra_T=77, seed=5, out_dir=HERE, gen_new=16)
    # video case: reference + crop patch (audio-only results are unchanged by the patch)
    patch_video_crop(ref_modules)
    make_wavenet_case(ref_wavenet, "video", dict(layer_size=3, stack_size=2, input_channels=32,
                      residual_channels=16, skip_channels=8), B=1, extra_T=0, seed=6, out_dir=HERE,
                      video=True, sample_cols=1024)


if __name__ == "__main__":
    main()
