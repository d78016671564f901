#!/usr/bin/env python
"""Benchmark of the WaveNet hot path (BASELINE.json: "WaveNet train audio-samples/s/GPU;
generation real-time factor at 16 kHz").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype fp32|bf16]

One JSON line on stdout (rank 0).  A "step" is one training step over one batch of synthetic
10-second clips exactly as the reference trainer writes it (pytorch_lightning_trainer.py:52-66):
forward (probabilities) -> target = argmax of the shifted one-hot audio -> F.cross_entropy ->
backward (-> gradient all-reduce when N > 1) -> AdamW step.

* value  : audio samples / s over all N GPUs, inputs already resident in HBM (CUDA events, max over ranks)
* e2e    : the same step driven from pinned HOST buffers: H2D copy of the one-hot audio and the video
           inside the timed region, D2H read of the loss every step
* roofline: the residual-layer forward kernel(s), timed live with CUDA events on the launching stream
* cpu_baseline: the oracle (a torch-CPU restatement of the reference, bit-identical to it; the
           reference is pure Python and cannot travel to the GPU box) on the host cores
* generation: cached autoregressive decode on the receptive-field config (experiments/04)

--impl reference times that same CPU port with all host threads (there is no compiled reference).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

METRIC = "WaveNet train audio-samples/s"
UNIT = "audio-samples/s"
T_CLIP = 160000

# BASELINE.json configs[1]: experiments/01_audio_video_debug.mk:10-17 (skip_channels: parser default 8,
# batch: parser default 3 -- SURVEY section 8)
WORKLOAD = dict(name="01_audio_video_debug", layer_size=3, stack_size=3, input_channels=64,
                residual_channels=64, skip_channels=8, batch_per_gpu=3, video=True)
# BASELINE.json configs[3] "scale-up model (03) with wider residual/skip channels, bf16": the architecture of the reference's own
# test (tests/test_model.py:42-48: 10 x 3 layers, A = 256) widened to C = S = 256 -- SURVEY section 8 "03w", the one shape of
# the path that is tensor-bound.  Audio-only (as SURVEY's flop count), one clip per GPU.  `python bench.py --workload 03w`
WORKLOAD_03W = dict(name="03w_scale_up_wide", layer_size=10, stack_size=3, input_channels=256,
                    residual_channels=256, skip_channels=256, batch_per_gpu=1, video=False)
# the architecture of the reference's own test (tests/test_model.py:14-17,42-48): 10 x 3 layers, A = 256, C = S = 64, batch 4
WORKLOAD_TESTARCH = dict(name="reference_test_architecture", layer_size=10, stack_size=3, input_channels=256,
                         residual_channels=64, skip_channels=64, batch_per_gpu=4, video=False)
WORKLOADS = {"01": WORKLOAD, "03w": WORKLOAD_03W, "testarch": WORKLOAD_TESTARCH}
# BASELINE.json configs[4]: experiments/04_kinetics_receptive_field.mk:58-71
DECODE = dict(name="04_kinetics_receptive_field", layer_size=14, stack_size=1, input_channels=128,
              residual_channels=16, skip_channels=8)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}


def synth_codes(B, T, A, seed, device):
    """two tones + noise, min-max normalised, mu-law coded: SURVEY 8(d)"""
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(T, dtype=torch.float32) / 16000.0
    rows = []
    for b in range(B):
        w = 0.6 * torch.sin(2 * torch.pi * (220.0 + 7 * b) * t) + 0.3 * torch.sin(2 * torch.pi * 3520.0 * t) \
            + 0.1 * (torch.rand(T, generator=g) * 2 - 1)
        rows.append(2 * (w - w.min()) / (w.max() - w.min()) - 1)
    return torch.stack(rows)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region"""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, enabled=True):
        self.gpu = gpu_index
        self.proc = None
        self.enabled = enabled        # rank 0 only: one NVML client per GPU polling during a 50 ms timed region disturbs the launches
        self.t0 = None                # samples before mark() (NVML start-up, warm-up) are not "under load" and are dropped

    def mark(self):
        import datetime
        self.t0 = datetime.datetime.now()

    def start(self):
        if not self.enabled:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        out, _ = self.proc.communicate()
        import datetime
        sm, mx, reasons = [], None, set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            if self.t0 is not None:
                try:          # "2026/10/19 03:12:45.123" (local time, like datetime.now()); an unparsable stamp keeps the sample
                    if datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f") < self.t0:
                        continue
                except ValueError:
                    pass
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_steps(w, steps, warmup, video=True, B=1, T=None):
    """the reference's CPU implementation of the step (oracle port), all host threads"""
    T_CLIP = T or globals()["T_CLIP"]
    from oracle import wavenet_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    shape = orc.Shape(w["layer_size"], w["stack_size"], w["input_channels"], w["residual_channels"], w["skip_channels"])
    p = {k: v.requires_grad_(True) for k, v in orc.init_params(shape, 0, video=video).items()}
    from movenet_b200.mulaw import _encode_formula
    codes = _encode_formula(synth_codes(B, T_CLIP, shape.input_channels, 1234, "cpu"), shape.input_channels)
    audio = torch.zeros(B, shape.input_channels, T_CLIP).scatter_(1, codes.unsqueeze(1), 1.0)
    vid = torch.randint(0, 256, (B, 160, 64, 64, 1), generator=torch.Generator().manual_seed(4321)).float() if video else None
    opt = torch.optim.AdamW(list(p.values()), lr=3e-4)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss, _, _, _ = orc.training_loss(p, shape, audio, vid)
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return B * T_CLIP, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    n, times = cpu_reference_steps(w, max(1, args.steps), max(0, args.warmup), video=w["video"], B=1)
    ms = 1e3 * sum(times) / len(times)
    value = n / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "sample": "1 clip (160000 samples) per step of the batch-3 workload"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                             "sample": "1 clip (160000 samples) per step, fwd+CE+bwd+AdamW, torch CPU fp32"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def train_step(model, opt, audio, video):
    opt.zero_grad(set_to_none=True)
    output = model(audio, video)
    target = audio[:, :, model.receptive_fields:].argmax(1)
    loss = F.cross_entropy(output, target)
    loss.backward()
    opt.step()
    return loss


def timed(fn, steps, warmup, sync):
    for _ in range(warmup):
        fn()
    sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        fn()
    ev1.record()
    sync()
    return ev0.elapsed_time(ev1)


def _ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/), or None"""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        table = json.load(open(path))
    except (OSError, ValueError):
        return None
    if kernel in table:
        return table[kernel]
    for k, v in table.items():          # template instantiations: "name<...>"
        if k.startswith(kernel + "<"):
            return v
    return None


def layer_roofline(model, audio, video, dtype):
    """time the residual-layer kernels live (CUDA events on the launching stream): backward (the dominant kernel
    of the step) and forward"""
    import ctypes as C
    from movenet_b200 import _lib
    B = audio.shape[0]
    shape = model._shape(B, T_CLIP, video is not None, True, False)
    bufs = model._buffers_for(shape, audio.device)
    model._pack(bufs, model._param_list())
    acts = torch.empty(bufs.acts_bytes, dtype=torch.uint8, device=audio.device)
    out = torch.empty(B, model.input_channels, T_CLIP - model.receptive_fields, device=audio.device)
    st = torch.cuda.current_stream().cuda_stream
    vp = 0 if video is None else video.data_ptr()
    scratch = bufs.get_scratch()
    pg = bufs.get_packed_grads()
    _lib.call("mvn_wavenet_forward", C.byref(shape), bufs.packed.data_ptr(), audio.data_ptr(), vp, acts.data_ptr(),
              out.data_ptr(), scratch.data_ptr(), st)
    dout = torch.randn_like(out) * 1e-6
    _lib.call("mvn_wavenet_backward", C.byref(shape), bufs.packed.data_ptr(), audio.data_ptr(), vp, acts.data_ptr(),
              out.data_ptr(), dout.data_ptr(), pg.data_ptr(), scratch.data_ptr(), st)
    layers = list(range(1, model.layer_size * model.stack_size - 1))     # interior layers
    lib = _lib.load()

    def time_stage(fn):
        n0 = lib.mvn_launch_count()
        for l in layers:
            fn(l)
        per_layer = (lib.mvn_launch_count() - n0) / len(layers)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        ev0.record()
        for _ in range(reps):
            for l in layers:     # each layer streams > L2 worth of activations, so every launch starts cold
                fn(l)
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / (reps * len(layers)), per_layer

    ms_f, n_f = time_stage(lambda l: _lib.call("mvn_layer_fwd", C.byref(shape), bufs.packed.data_ptr(), l, acts.data_ptr(),
                                               scratch.data_ptr(), st))
    ms_b, n_b = time_stage(lambda l: _lib.call("mvn_layer_bwd", C.byref(shape), bufs.packed.data_ptr(), l, acts.data_ptr(),
                                               pg.data_ptr(), scratch.data_ptr(), st))
    e = 2 if dtype == "bf16" else 4
    Cc, S = model.residual_channels, model.skip_channels
    vid = video is not None
    if lib.mvn_kernel_path(C.byref(shape)) == 2:
        # wide-channel path: the layer is tensor-bound.  Algorithmic flops per audio sample and layer, 2 per MAC,
        # the dense formulation exactly as the reference computes it: forward 10 C^2 + 2 C S; backward = data + weight gradient
        # = twice that (the backward does not recompute the gate GEMM: the forward keeps the gate's derivative factors)
        pk = peaks()
        n = B * T_CLIP
        f_fwd = 10 * Cc * Cc + 2 * Cc * S

        def tobj(name, ms, flops, launches, kernels, traffic_of):
            ach = flops * n / (ms * 1e-3) / 1e12
            return {"bound": "tensor", "kernel": "%s (%s, %d launches per layer: %s)" % (name, dtype, round(launches), kernels),
                    "achieved": ach, "peak": pk["bf16_tflops"], "peak_source": pk["source"] + " (cuBLAS bf16, sustained)",
                    "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"], "traffic": _ncu_traffic(traffic_of), "traffic_kernel": traffic_of,
                    "ms_per_launch": ms, "flops_per_sample": flops, "samples_per_launch": n}
        return (tobj("residual layer backward", ms_b, 2 * f_fwd, n_b,
                     "2 wide_gemm_kernel (d(gated) x kept gate-derivative factors -> dz; d(x) + bias column sums) + wide_wgrad_kernel (K = time weight gradients) + its reduction",
                     "wide_gemm_kernel<2, 4>"),
                # (the skip 1x1 convs of all layers run as ONE GEMM after the stack: their 2 C S flops are not in this stage)
                tobj("residual layer forward", ms_f, 10 * Cc * Cc, n_f, "2 wide_gemm_kernel (gate + derivative factors kept for the backward; residual)",
                     "wide_gemm_kernel<2, 0>"))
    # algorithmic bytes per audio sample of one layer
    fwd_b = Cc * e + Cc * e + (Cc * e if vid else 0) + 8 * S                 # read x, write x', read ctx, RMW skip_sum
    # backward, algorithmic bytes of any layer-at-a-time backward: read x, the stream gradient D, d(skip) (+ ctx and the
    # ctx-gradient running sum Q); write D' (+ Q').  The default kernel moves the stream gradient as the pair (P, U) (one more
    # read and one more write of C*e bytes: `design_bytes_per_sample`); MOVENET_B200_BWD_SUM=1 selects the one-stream variant
    # (same algorithmic bytes, measured slower: DESIGN.md section 3).
    summed = bool(int(os.environ.get("MOVENET_B200_BWD_SUM", "0"))) and max(model.residual_conv_stack.dilations) <= 128
    # the double-buffered kernel (layer_tc_bwd_db.cu: every layer of this model when all dilations are <= 8) does not read Q:
    # its contribution is added in place by a TMA reduction
    dbuf = (not summed and os.environ.get("MOVENET_B200_BWD_DB", "1") != "0" and max(model.residual_conv_stack.dilations) <= 8
            and S <= 32)
    bwd_b = 3 * Cc * e + 4 * S + (3 * Cc * e if vid else 0)
    bwd_design = bwd_b + (0 if summed else 2 * Cc * e) - (Cc * e if (dbuf and vid) else 0)
    pk = peaks()
    n = B * T_CLIP

    def obj(name, ms, per_sample, launches, kernel, design=None):
        ach = per_sample * n / (ms * 1e-3) / 1e9
        extra = {} if design is None else {"design_bytes_per_sample": design,
                                           "design_gbs": design * n / (ms * 1e-3) / 1e9,
                                           "stream_gradient": "one summed stream" if summed else "pair (P, U)"}
        return {**extra, "bound": "hbm", "kernel": "%s (%s, %d launch(es) per layer)" % (name, dtype, round(launches)),
                "achieved": ach, "peak": pk["hbm_gbs"], "peak_source": pk["source"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "traffic": _ncu_traffic(kernel), "ms_per_launch": ms,
                "bytes_per_sample": per_sample, "samples_per_launch": n}

    bwd_kernel = "layer_bwd_tc_kernel<1, 0>" if summed else ("layer_bwd_db_kernel<1>" if dbuf else "layer_bwd_tc_kernel<0, 1>")
    return (obj("residual layer backward: %s" % bwd_kernel.split("<")[0], ms_b, bwd_b, n_b, bwd_kernel, bwd_design),
            obj("residual layer forward", ms_f, fwd_b, n_f, "layer_fwd_tc_kernel"))


def model_state_repeat(m, st, reps):
    """replicate a prefilled decode state `reps` times along the clip axis (queues are (layer, slot, clip, channel))"""
    from movenet_b200.decode import DecodeState
    from movenet_b200 import _lib
    import ctypes as C
    B = st.batch
    shape = m._shape(B * reps, st.shape.frames, False, False, True, _lib.F32)
    bufs = m._buffers_for(shape, st.state.device)
    m._pack(bufs, m._param_list())
    big = torch.zeros(_lib.size("mvn_decode_tc_state_bytes", shape) if st.fast else
                      _lib.size("mvn_decode_state_bytes", shape, st.mode), dtype=torch.uint8, device=st.state.device)
    es = 2 if st.fast else 4
    Cc = m.residual_channels
    dil = list(m.residual_conv_stack.dilations)
    edge = (not st.fast) and st.mode == _lib.DECODE_REFERENCE and m.stack_size == 1
    if edge:                                        # reference-window mode: rings reach back to the window edge (decode.cu)
        dil = [max(d, sum(dil[l + 1:])) for l, d in enumerate(dil)]
    src_off = dst_off = 0
    for d in dil:                                   # per layer: (slot, clip, channel)
        n_src = d * B * Cc * es
        src = st.state[src_off:src_off + n_src].view(d, B * Cc * es)
        big[dst_off:dst_off + n_src * reps].view(d, reps, B * Cc * es).copy_(src.unsqueeze(1).expand(d, reps, B * Cc * es))
        src_off += n_src; dst_off += n_src * reps
    a256 = lambda x: (x + 255) // 256 * 256
    src_l2 = a256(src_off); dst_l2 = a256(dst_off)
    l2 = st.state[src_l2:src_l2 + B * 8].view(B, 8)
    big[dst_l2:dst_l2 + B * reps * 8].view(reps, B, 8).copy_(l2.unsqueeze(0).expand(reps, B, 8))
    if edge:                                        # code ring [RF][clip]
        RF = m.receptive_fields
        src_c = src_l2 + a256(B * 8); dst_c = dst_l2 + a256(B * reps * 8)
        ring = st.state[src_c:src_c + RF * B * 4].view(RF, B * 4)
        big[dst_c:dst_c + RF * B * reps * 4].view(RF, reps, B * 4).copy_(ring.unsqueeze(1).expand(RF, reps, B * 4))
    return DecodeState(shape, bufs, big, None, B * reps, st.channels, fast=st.fast, mode=st.mode)


def decode_bench(dev, rank=0, world=1):
    """cached generation on the receptive-field config: RTF = generated seconds of 16 kHz audio per wall second.

    world > 1: independent clips are sharded over the ranks (movenet_b200.parallel.shard_range) with NO communication in
    the data path; every rank times its shard with CUDA events between barriers, the slowest rank's time
    is the job's.  Only the throughput legs run there."""
    import torch.distributed as dist
    import movenet_b200
    from movenet_b200 import _lib
    from movenet_b200.parallel import shard_range
    d = DECODE
    m = movenet_b200.WaveNet(d["layer_size"], d["stack_size"], d["input_channels"], d["residual_channels"],
                             d["skip_channels"], compute_dtype="fp32").to(dev)
    if world > 1:
        m.enable_data_parallel()            # same weights on every rank (rank 0's); no gradient traffic in decode
    RF = m.receptive_fields
    N, Cc, A_ = m.layer_size * m.stack_size, d["residual_channels"], d["input_channels"]
    res = {}
    from movenet_b200.decode import fast_mode_available, prefill, run_steps, steps_into

    def job_seconds(fn):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(); fn(); ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = t.item()
        return ms * 1e-3

    # exact decoders: one warp per clip, 16 clips per CTA -> 148 x 16 clips fill the machine once; tensor-core decoder: 512 per CTA.
    # "reference_window": the default exact mode -- stack_size == 1 here, so the reference's window edge is evaluated too
    # (about twice the arithmetic and deeper rings); "causal": the true causal model; "fast": tensor cores, bf16 queues.
    legs = [("reference_window_f32", "exact", 1, 2000), ("reference_window_f32", "exact", 148 * 16, 400),
            ("causal_f32", "causal", 148 * 16, 400), ("fast_bf16", "fast", 148 * 512, 400)]
    if world > 1:
        legs = [l for l in legs if l[2] > 1 and l[1] != "exact"]
    for label, mode, clips_per_gpu, n_new in legs:
        fast = mode == "fast"
        lo, hi = shard_range(clips_per_gpu * world, rank, world)       # this rank's clips of the global batch
        clips = hi - lo
        if fast and not fast_mode_available(m, clips, RF):
            continue
        per = min(clips, 2368)                      # build the prompt batch in slices: the one-hot prompt is large
        codes = torch.randint(0, A_, (per, RF), device=dev, generator=torch.Generator(device=dev).manual_seed(100 + lo))
        prompt = movenet_b200.one_hot(codes, A_)
        state = prefill(m, prompt, None, fast=fast, mode="exact" if fast else mode)
        slice_tokens = None
        if clips > per:                             # big batch: prefill one slice and replicate its queues
            pristine = state.state.clone()
            slice_tokens = run_steps(m, state, RF, 8).clone()
            state.state.copy_(pristine)
            del pristine
            state = model_state_repeat(m, state, clips // per)
        nclips = state.batch
        warm = run_steps(m, state, RF, 8)           # warm-up (re-running positions afterwards is harmless for timing)
        replicas_ok = None
        if slice_tokens is not None:                # every replica of the slice must emit the slice's own tokens
            replicas_ok = bool((warm.reshape(clips // per, per, 8) == slice_tokens.unsqueeze(0)).all().item())
        e = 2 if fast else 4
        edge = mode == "exact" and m.stack_size == 1
        # bytes per generated sample: ring pop + push per layer (+ the edge tap of the reference-window mode)
        queue_b = ((3 * (N - 1) + 1) if edge else 2 * N) * Cc * e
        # (a) the decode kernel alone: int32 codes out
        sec_k = job_seconds(lambda: run_steps(m, state, RF, n_new))
        # (b) the reference's output format: a zeroed (B, A, n_new) fp32 tensor with one 1.0 per generated sample
        # (movenet/wavenet.py:211-236) -- zero fill + steps + scatter, all inside the timed region.  SURVEY 8(d)'s
        # algorithmic bytes (queue rows + the 4A-byte one-hot column) are quoted on THIS region.
        out = torch.empty(nclips, A_, n_new, dtype=torch.float32, device=dev)
        out.zero_(); steps_into(m, state, RF, 8, out[:, :, :8])     # warm-up of the allocator / scatter kernels

        def onehot_region():
            out.zero_()
            steps_into(m, state, RF, n_new, out)
        sec = job_seconds(onehot_region)
        del out
        total = nclips * world                       # equal shards
        per_clip_rtf = (n_new / 16000.0) / sec
        bytes_per_sample = queue_b + 4 * A_
        gbs = nclips * n_new * bytes_per_sample / sec / 1e9           # per GPU
        gbs_k = nclips * n_new * (queue_b + 4) / sec_k / 1e9
        res[f"{label}_clips_{total}"] = {
            "decoder": {"exact": "fp32 CUDA cores, reference-window mode (token-exact vs the reference's generate)",
                        "causal": "fp32 CUDA cores, true causal model",
                        "fast": "tcgen05, bf16 queues, true causal model"}[mode],
            "clips_total": total, "clips_per_gpu": nclips, "n_gpus": world, "new_samples_per_clip": n_new,
            "rtf_per_clip": per_clip_rtf, "aggregate_samples_per_s": total * n_new / sec, "aggregate_rtf": total * per_clip_rtf,
            "output": "one-hot (B, A, n) fp32, zero fill + scatter inside the timed region",
            "bytes_per_sample": bytes_per_sample, "hbm_gbs_algorithmic_per_gpu": gbs, "hbm_frac": gbs / peaks()["hbm_gbs"],
            "replicas_emit_the_slice_tokens": replicas_ok,
            "kernel_only": {"aggregate_samples_per_s": total * n_new / sec_k, "rtf_per_clip": (n_new / 16000.0) / sec_k,
                            "output": "int32 codes", "bytes_per_sample": queue_b + 4,
                            "hbm_gbs_per_gpu": gbs_k, "hbm_frac": gbs_k / peaks()["hbm_gbs"]}}
        del state, prompt
        torch.cuda.empty_cache()
    return {"workload": d["name"], "receptive_fields": RF, "dtype": "f32", "sharding": "clips over ranks, no communication",
            **res}


def dp_selfcheck(dev, rank, world):
    """Data-parallel gradient averaging, checked where the driver records it (tests/test_gpu_dp.py needs >= 2 GPUs and is
    skipped on the 1-GPU test box): every rank builds the model from a DIFFERENT seed (enable_data_parallel must hand out
    rank 0's weights), runs its shard of one global batch, and the averaged gradients must (a) be bit-identical on all
    ranks and (b) equal the gradients a single process computes on the whole batch (equal shards, mean loss)."""
    import torch.distributed as dist
    import movenet_b200
    kw = dict(layer_size=2, stack_size=2, input_channels=32, residual_channels=16, skip_channels=8)
    torch.manual_seed(1000 + rank)
    m = movenet_b200.WaveNet(**kw, compute_dtype="fp32").to(dev).enable_data_parallel()
    w0 = torch.cat([p.detach().flatten() for p in m.parameters()])
    ref_w = w0.clone()
    dist.broadcast(ref_w, src=0)
    broadcast_ok = bool(torch.equal(w0, ref_w))
    per = 2
    single = movenet_b200.WaveNet(**kw, compute_dtype="fp32").to(dev)       # no data parallelism: the whole batch here
    single.load_state_dict(m.state_dict())
    ranks_equal, err = True, 0.0
    for step in range(3):       # three steps: the peer-memory exchange alternates its buffers by step (csrc/peer.cu)
        codes = torch.randint(0, 32, (per * world, 300), generator=torch.Generator().manual_seed(1 + step)).to(dev)
        mine = codes[per * rank:per * (rank + 1)]
        m.zero_grad(set_to_none=True); single.zero_grad(set_to_none=True)
        F.cross_entropy(m(mine), mine[:, m.receptive_fields:]).backward()
        g_dp = torch.cat([p.grad.flatten() for p in m.parameters() if p.grad is not None])
        lo, hi = g_dp.clone(), g_dp.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        ranks_equal = ranks_equal and bool(torch.equal(lo, hi))
        F.cross_entropy(single(codes), codes[:, single.receptive_fields:]).backward()
        g_one = torch.cat([p.grad.flatten() for p in single.parameters() if p.grad is not None])
        err = max(err, ((g_dp - g_one).norm() / g_one.norm().clamp_min(1e-30)).item())
    exchange = "nvlink-peer (csrc/peer.cu)" if any(v is not None for v in m._dp_peer.values()) else "nccl all-reduce"
    ok = broadcast_ok and ranks_equal and err < 1e-4
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    return {"ok": bool(flag.item() == 1.0), "weights_broadcast_from_rank0": broadcast_ok, "gradients_equal_on_all_ranks": ranks_equal,
            "rel_err_vs_single_process": err, "ranks": world, "steps": 3, "exchange": exchange}


def run_ours(args):
    import torch.distributed as dist
    import movenet_b200
    from movenet_b200 import _lib
    from movenet_b200.parallel import init_from_env
    rank, local, world = init_from_env("nccl")
    assert world == args.gpus or world == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    w = WORKLOADS[args.workload]
    B = w["batch_per_gpu"]
    torch.manual_seed(0)
    model = movenet_b200.WaveNet(w["layer_size"], w["stack_size"], w["input_channels"], w["residual_channels"],
                                 w["skip_channels"], compute_dtype=args.dtype).to(dev)
    if world > 1:
        model.enable_data_parallel()
    # the reference builds getattr(torch.optim, config.optimizer)(params, lr) (pytorch_lightning_trainer.py:128-202);
    # movenet_b200.optim.AdamW is the same update for all 100-odd tensors in one launch (parity-tested against
    # torch.optim.AdamW); MOVENET_B200_TORCH_ADAMW=1 times torch's fused AdamW instead
    if os.environ.get("MOVENET_B200_TORCH_ADAMW"):
        opt = torch.optim.AdamW(model.parameters(), lr=3e-4, fused=True)
    else:
        opt = movenet_b200.optim.AdamW(model.parameters(), lr=3e-4)

    wave = synth_codes(B, T_CLIP, w["input_channels"], 1234 + rank, dev)
    host_audio = torch.zeros(B, w["input_channels"], T_CLIP).scatter_(
        1, movenet_b200.mulaw._encode_formula(wave, w["input_channels"]).unsqueeze(1), 1.0).pin_memory()
    host_video = torch.randint(0, 256, (B, 160, 64, 64, 1), generator=torch.Generator().manual_seed(4321 + rank)).float().pin_memory() \
        if w["video"] else torch.zeros(0)
    audio = host_audio.to(dev, non_blocking=True)
    video = host_video.to(dev, non_blocking=True) if w["video"] else None

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- device-resident steps ----
    # (the sampler starts BEFORE the warm-up steps: nvidia-smi needs ~1 s to initialise NVML over all GPUs of the box, which must
    # not fall into the 50 ms timed region; it then samples every 100 ms through warm-up and timed steps, i.e. under load)
    sampler = ClockSampler(local, enabled=(rank == 0))
    sampler.start()
    if rank == 0:
        time.sleep(1.5)           # NVML initialisation is over before any step is enqueued (the other ranks wait at the barrier in sync())
    for _ in range(args.warmup):
        train_step(model, opt, audio, video)
    sync()
    n0 = _lib.load().mvn_launch_count()
    sampler.mark()            # clocks are reported from here on: the timed device-resident and end-to-end legs
    ms_total = timed(lambda: train_step(model, opt, audio, video), args.steps, 0, sync)
    launches = int(_lib.load().mvn_launch_count() - n0)
    ms_total = max_over_ranks(ms_total)
    ms_step = ms_total / args.steps
    value = world * B * T_CLIP / (ms_step * 1e-3)

    # ---- end to end from pinned host buffers ----
    # every step's inputs cross PCIe inside the timed region; the copy of step i+1 is issued on a side stream
    # while step i computes (an ordinary prefetching input pipeline).  Every step's loss is read back inside the timed
    # region too: it is copied to pinned host memory asynchronously and consumed one step later, while the next step is
    # being enqueued (the way a training loop logs its loss without stalling the device); the last one is drained
    # before the closing event.
    loss_box = [0.0]
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    pending = {"ev": None, "slot": 0, "read": 0}

    def post_loss(loss):
        # consume the previous step's loss (its copy finished long ago), then queue this step's copy
        drain_loss()
        k = pending["slot"] ^ 1
        loss_host[k:k + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        pending["ev"], pending["slot"] = ev, k

    def drain_loss():
        if pending["ev"] is not None:
            pending["ev"].synchronize()
            loss_box[0] = float(loss_host[pending["slot"]])
            pending["ev"] = None
            pending["read"] += 1

    def timed_e2e(fn, steps):
        fn()
        drain_loss()
        sync()
        pending["read"] = 0
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(steps):
            fn()
        drain_loss()
        assert pending["read"] == steps, "every step's loss must be read inside the timed region"
        ev1.record()
        sync()
        return ev0.elapsed_time(ev1)

    copy_stream = torch.cuda.Stream(device=dev)
    # two device-side input slots, allocated once: no allocator traffic inside the timed region
    slots = [{"a": torch.empty_like(audio), "v": torch.empty_like(video) if w["video"] else None,
              "ready": None, "free": None} for _ in range(2)]

    def issue_copy(slot):
        with torch.cuda.stream(copy_stream):
            if slot["free"] is not None:
                copy_stream.wait_event(slot["free"])        # the step that last used this slot has finished with it
            slot["a"].copy_(host_audio, non_blocking=True)
            if w["video"]:
                slot["v"].copy_(host_video, non_blocking=True)
            slot["ready"] = torch.cuda.Event()
            slot["ready"].record(copy_stream)

    state = {"i": 0}

    def e2e_step():
        cur = slots[state["i"] & 1]
        if cur["ready"] is None:
            issue_copy(cur)
        issue_copy(slots[(state["i"] + 1) & 1])          # next step's inputs cross PCIe while this step computes
        torch.cuda.current_stream().wait_event(cur["ready"])
        loss = train_step(model, opt, cur["a"], cur["v"])
        cur["free"] = torch.cuda.Event()
        cur["free"].record(torch.cuda.current_stream())
        cur["ready"] = None
        post_loss(loss)
        state["i"] += 1

    ms_onehot = max_over_ranks(timed_e2e(e2e_step, args.steps)) / args.steps

    # the same loop fed with the integer mu-law codes instead of their one-hot expansion (the non-breaking input
    # overload of forward()): 1/(4A) of the audio bytes cross PCIe
    host_codes = host_audio.argmax(1).pin_memory()
    cslots = [{"a": torch.empty(B, T_CLIP, dtype=torch.int64, device=dev), "v": torch.empty_like(video) if w["video"] else None,
               "ready": None, "free": None} for _ in range(2)]
    cstate = {"i": 0}

    def issue_codes(slot):
        with torch.cuda.stream(copy_stream):
            if slot["free"] is not None:
                copy_stream.wait_event(slot["free"])
            slot["a"].copy_(host_codes, non_blocking=True)
            if w["video"]:
                slot["v"].copy_(host_video, non_blocking=True)
            slot["ready"] = torch.cuda.Event()
            slot["ready"].record(copy_stream)

    def codes_step():
        cur = cslots[cstate["i"] & 1]
        if cur["ready"] is None:
            issue_codes(cur)
        issue_codes(cslots[(cstate["i"] + 1) & 1])
        torch.cuda.current_stream().wait_event(cur["ready"])
        opt.zero_grad(set_to_none=True)
        out = model(cur["a"], cur["v"])
        loss = F.cross_entropy(out, cur["a"][:, model.receptive_fields:])
        loss.backward()
        opt.step()
        cur["free"] = torch.cuda.Event()
        cur["free"].record(torch.cuda.current_stream())
        cur["ready"] = None
        post_loss(loss)
        cstate["i"] += 1

    ms_codes = max_over_ranks(timed_e2e(codes_step, args.steps)) / args.steps
    h2d = host_audio.numel() * 4 + (host_video.numel() * 4 if w["video"] else 0)
    clocks = sampler.stop()

    # ---- sustained rate: the same device-resident step for >= 1000 steps (seconds, not milliseconds, under load) ----
    sustained = None
    if args.sustained_steps > 0:
        s2 = ClockSampler(local, enabled=(rank == 0))
        s2.start()
        ms_sus = max_over_ranks(timed(lambda: train_step(model, opt, audio, video), args.sustained_steps, 0, sync))
        c2 = s2.stop()
        sustained = {"steps": args.sustained_steps, "ms_per_step": ms_sus / args.sustained_steps,
                     "value": world * B * T_CLIP / (ms_sus / args.sustained_steps * 1e-3), "unit": UNIT,
                     "sm_mhz_median": c2["sm_mhz"], "sm_max_mhz": c2["sm_max_mhz"], "reasons": c2["reasons"],
                     "clock_samples": c2.get("samples")}

    # ---- N > 1: data-parallel self-check (recorded in the JSON line) and the sharded decode leg ----
    dp_check = dp_selfcheck(dev, rank, world) if world > 1 else None
    generation = None
    if world > 1 and not args.no_decode and args.workload == "01":
        generation = decode_bench(dev, rank, world)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    roof, roof_fwd = layer_roofline(model, audio, video, args.dtype)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.dtype == "fp32" else "bf16", "data": "synthetic",
            "per_gpu": value / world, "loss": loss_box[0],
            "config": {"workload": w["name"], "layer_size": w["layer_size"], "stack_size": w["stack_size"],
                       "input_channels": w["input_channels"], "residual_channels": w["residual_channels"],
                       "skip_channels": w["skip_channels"], "video": w["video"], "clips_per_gpu": B,
                       "samples_per_clip": T_CLIP, "global_clips": B * world,
                       "parallelism": f"dp{world}" if world > 1 else "single",
                       "step": "fwd(probs)+argmax target+cross_entropy+bwd+allreduce+AdamW",
                       "optimizer": "torch.optim.AdamW(fused=True)" if os.environ.get("MOVENET_B200_TORCH_ADAMW") else "movenet_b200.optim.AdamW (same update, one launch)",
                       "l2": "inputs and activations (>1 GB per step) exceed the 126 MB L2; no explicit flush"},
            "clocks": clocks, "gpu_launches": launches // max(1, args.steps) * args.steps,
            "gpu_launches_per_step": launches / max(1, args.steps),
            # headline end-to-end number: the repo's own public API fed the way SURVEY 8(f).1 asks -- forward(codes), the
            # integer mu-law codes (what the dataset produces before its one-hot expansion, movenet/dataset.py:284) instead of
            # the 4A-times larger one-hot fp32 tensor
            "e2e": {"value": world * B * T_CLIP / (ms_codes * 1e-3), "unit": UNIT, "ms_per_step": ms_codes,
                    "h2d_bytes_per_step": B * T_CLIP * 8 + (host_video.numel() * 4 if w["video"] else 0),
                    "d2h_bytes_per_step": 4, "input": "forward(codes): (B, T) int64 mu-law codes + (B,160,64,64,1) fp32 video",
                    "note": "inputs: pinned host -> device on a copy stream, one step ahead; loss: every step's value is copied "
                            "to pinned host memory asynchronously and read on the host one step later, all inside the timed region"},
            # the reference's own input format (one-hot fp32, 4A bytes per sample): host-DRAM / PCIe bound -- 130.7 MB per step
            # and GPU, eight of them share one NUMA node's memory system at N = 8
            "e2e_onehot": {"value": world * B * T_CLIP / (ms_onehot * 1e-3), "unit": UNIT, "ms_per_step": ms_onehot,
                           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                           "input": "forward(one_hot): (B, A, T) fp32, the reference's format"},
            "sustained": sustained,
            "roofline": roof, "roofline_layer_forward": roof_fwd}
    if args.workload == "03w":
        A_, Cc, S_, N_ = w["input_channels"], w["residual_channels"], w["skip_channels"], w["layer_size"] * w["stack_size"]
        F_ = 4 * A_ * Cc + N_ * (10 * Cc * Cc + 2 * Cc * S_) + 2 * S_ * A_ + 2 * A_ * A_         # SURVEY 8(d), audio-only
        tf = 3 * F_ * value / world / 1e12
        line["step_tensor_roofline"] = {"flops_per_sample": 3 * F_, "achieved_tflops_per_gpu": tf, "peak": peaks()["bf16_tflops"],
                                        "frac": tf / peaks()["bf16_tflops"],
                                        "note": "whole training step on SURVEY 8(d)'s algorithmic 3F flops per sample (fwd + dgrad + wgrad)"}
    if dp_check is not None:
        line["dp_selfcheck"] = dp_check
    if generation is not None:
        line["generation"] = generation
    if args.workload != "01":
        args.no_decode = True
    if world == 1:
        if not args.no_cpu_baseline:
            if args.workload == "01":
                n, times = cpu_reference_steps(w, 2, 1, video=w["video"], B=1)
                sample = "1 clip (160000 samples) per step, 1 warm-up + 2 timed steps, torch CPU fp32"
            else:       # 72 MFLOP per sample: a tenth of a clip keeps the CPU leg within seconds
                n, times = cpu_reference_steps(w, 1, 1, video=w["video"], B=1, T=16000)
                sample = "16000 samples (a tenth of a clip) per step, 1 warm-up + 1 timed step, torch CPU fp32"
            cpu = n / (sum(times) / len(times))
            line["cpu_baseline"] = {"value": cpu, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample}
        if not args.no_decode:
            line["generation"] = decode_bench(dev)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dtype", default=os.environ.get("MOVENET_B200_DTYPE", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--workload", default="01", choices=sorted(WORKLOADS),
                    help="01: BASELINE configs[1] (the headline); 03w: the widened scale-up shape (tensor-bound)")
    ap.add_argument("--sustained-steps", type=int, default=1000,
                    help="extra device-resident steps timed separately for the sustained rate (0: skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
