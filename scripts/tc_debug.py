"""Per-layer comparison of the tensor-core (bf16) path against the fp32 exact path (GPU box only)."""
import ctypes as C
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import movenet_b200
from movenet_b200 import _lib


def run(model, audio, video, dtype):
    B, A, T = audio.shape
    shape = model._shape(B, T, video is not None, True, True, _lib.F32 if dtype == "fp32" else _lib.BF16)
    bufs = model._buffers_for(shape, audio.device)
    model._pack(bufs, model._param_list())
    acts = torch.zeros(bufs.acts_bytes, dtype=torch.uint8, device=audio.device)
    out = torch.empty(B, A, T - model.receptive_fields, device=audio.device)
    st = torch.cuda.current_stream().cuda_stream
    _lib.call("mvn_wavenet_forward", C.byref(shape), bufs.packed.data_ptr(), audio.data_ptr(),
              0 if video is None else video.data_ptr(), acts.data_ptr(), out.data_ptr(), bufs.get_scratch().data_ptr(), st)
    torch.cuda.synchronize()
    xs = []
    for l in range(model.layer_size * model.stack_size):
        x = torch.empty(B, T, model.residual_channels, device=audio.device)
        _lib.call("mvn_read_activation", C.byref(shape), acts.data_ptr(), 0, l, x.data_ptr(), st)
        xs.append(x)
    skip = torch.empty(B, T - model.receptive_fields + 1, model.skip_channels, device=audio.device)
    _lib.call("mvn_read_activation", C.byref(shape), acts.data_ptr(), 1, 0, skip.data_ptr(), st)
    torch.cuda.synchronize()
    return out, xs, skip


def main():
    video_on = "--video" in sys.argv
    T = 160000 if video_on else 1000
    torch.manual_seed(0)
    m = movenet_b200.WaveNet(3, 3, 64, 64, 8).cuda()
    codes = torch.randint(0, 64, (2, T), device="cuda")
    audio = movenet_b200.one_hot(codes, 64)
    video = torch.randint(0, 256, (2, 160, 64, 64, 1), device="cuda").float() if video_on else None
    o32, x32, s32 = run(m, audio, video, "fp32")
    o16, x16, s16 = run(m, audio, video, "bf16")
    RF = m.receptive_fields
    for l, (a, b) in enumerate(zip(x32, x16)):
        # compare on the region the reference keeps (t >= sum of earlier dilations)
        err = (a - b)[:, RF:, :].abs().max().item()
        print(f"x[{l}]: max|fp32|={a.abs().max().item():.4f} max|diff|={err:.5f} nan={bool(torch.isnan(b).any())}")
    print(f"skip: max|fp32|={s32.abs().max().item():.4f} max|diff|={(s32 - s16).abs().max().item():.5f}")
    print(f"logits: max|fp32|={o32.abs().max().item():.4f} max|diff|={(o32 - o16).abs().max().item():.5f}")
    rows = (x32[1] - x16[1])[0, RF:RF + 260].abs().max(1).values
    print("row errors of x[1], first 260 rows:", [round(v, 4) for v in rows.tolist()][:40])


if __name__ == "__main__":
    main()
