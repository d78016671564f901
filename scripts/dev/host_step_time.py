"""Host time to ENQUEUE one training step of the cfg01 model (tiny clips: the GPU is never the limiter), the number that decides
whether 8 processes on one host can keep 8 GPUs busy:  python scripts/dev/host_step_time.py [video 0|1]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import bench
import movenet_b200

w = bench.WORKLOADS["01"]
video_on = (sys.argv[1] if len(sys.argv) > 1 else "1") == "1"
m = movenet_b200.WaveNet(w["layer_size"], w["stack_size"], w["input_channels"], w["residual_channels"], w["skip_channels"],
                         compute_dtype="bf16").cuda()
opt = movenet_b200.optim.AdamW(m.parameters(), lr=3e-4)
B, T = 1, 160000 if video_on else 4096
codes = torch.randint(0, 64, (B, T), device="cuda")
audio = movenet_b200.one_hot(codes, 64)
video = torch.randint(0, 256, (B, 160, 64, 64, 1), device="cuda").float() if video_on else None
for _ in range(5):
    bench.train_step(m, opt, audio, video)
torch.cuda.synchronize()
res = []
for rep in range(3):
    n = 30
    t0 = time.perf_counter()
    for _ in range(n):
        bench.train_step(m, opt, audio, video)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    res.append((1e3 * (t1 - t0) / n, 1e3 * (t2 - t0) / n))
print("video=%d side_streams=%s: enqueue %.3f ms/step (device-complete %.3f ms/step)" %
      (video_on, os.environ.get("MOVENET_B200_SIDE_STREAMS", "1"), min(r[0] for r in res), min(r[1] for r in res)))
if os.environ.get("HOST_PROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
    for _ in range(20):
        bench.train_step(m, opt, audio, video)
    pr.disable(); torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats("tottime").print_stats(28)
if os.environ.get("HOST_SECTIONS"):
    import torch.nn.functional as F
    names = ["zero_grad", "forward", "target", "loss", "backward", "optimizer"]
    acc = [0.0] * len(names)
    n = 50
    RF = m.receptive_fields
    for _ in range(n):
        t = [time.perf_counter()]
        opt.zero_grad(set_to_none=True); t.append(time.perf_counter())
        out = m(audio, video) if video is not None else m(audio); t.append(time.perf_counter())
        target = audio[:, :, RF:].argmax(1); t.append(time.perf_counter())
        loss = F.cross_entropy(out, target); t.append(time.perf_counter())
        loss.backward(); t.append(time.perf_counter())
        opt.step(); t.append(time.perf_counter())
        for i in range(len(names)):
            acc[i] += t[i + 1] - t[i]
    torch.cuda.synchronize()
    print("host ms per step:", {k: round(1e3 * v / n, 3) for k, v in zip(names, acc)}, "sum", round(1e3 * sum(acc) / n, 3))
if os.environ.get("HOST_FORWARD"):
    import ctypes as C
    from movenet_b200 import _lib
    from movenet_b200.wavenet import _stream
    n = 50
    acc = {}
    def tick(name, t0):
        t1 = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t1 - t0); return t1
    for _ in range(n):
        opt.step()                                  # (the weights change: the pack must run)
        t = time.perf_counter()
        a2 = m._check_audio(audio); v2 = m._check_video(video) if video is not None else None; t = tick("check", t)
        params = m._param_list(); ng = not (torch.is_grad_enabled() and any(p.requires_grad for p in params)); t = tick("params", t)
        bufs = m._engine_buffers(a2, v2 is not None, True, False, ng); t = tick("buffers", t)
        m._pack(bufs, params); t = tick("pack", t)
        out = torch.empty(bufs.shape.batch, bufs.shape.input_channels, bufs.shape.frames - m.receptive_fields, dtype=torch.float32, device="cuda")
        acts = torch.empty(bufs.acts_bytes, dtype=torch.uint8, device="cuda"); t = tick("alloc", t)
        _lib.call("mvn_wavenet_forward", C.byref(bufs.shape), bufs.packed.data_ptr(), a2.data_ptr(), 0 if v2 is None else v2.data_ptr(),
                  acts.data_ptr(), out.data_ptr(), bufs.get_scratch().data_ptr(), _stream()); t = tick("c_forward", t)
        o = m(audio, video) if video is not None else m(audio); t = tick("whole_forward_call", t)
    torch.cuda.synchronize()
    print("forward host us:", {k: round(1e6 * v / n, 1) for k, v in acc.items()})
