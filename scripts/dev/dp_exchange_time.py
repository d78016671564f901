"""torchrun --nproc-per-node N scripts/dev/dp_exchange_time.py : device time of ONE gradient exchange of the cfg01 model,
peer-memory kernel (csrc/peer.cu) against unpack + NCCL all-reduce, 200 back-to-back exchanges each (ranks in lock-step)."""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import movenet_b200
from movenet_b200 import _lib, peer
from movenet_b200.parallel import init_from_env

rank, local, world = init_from_env("nccl")
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
torch.manual_seed(0)
m = movenet_b200.WaveNet(3, 3, 64, 64, 8, compute_dtype="bf16").to(dev).enable_data_parallel()
shape = m._shape(3, 160000, True, True, False)
bufs = m._buffers_for(shape, dev)
pg = bufs.get_packed_grads()
pg.view(torch.float32).normal_()
st = torch.cuda.current_stream().cuda_stream
flat, views = m._flat_grads(True, dev)
offs = m._grad_offsets(True, dev)
pgr = peer.PeerGradients(shape, dev, m._dp_group)
assert pgr.ok, pgr.why
def timed(fn, n=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def f_peer():
    pgr.reduce_unpack(shape, pg.data_ptr(), flat, offs, st)


def f_nccl():
    _lib.call("mvn_unpack_grads", C.byref(shape), pg.data_ptr(), flat.data_ptr(), offs.data_ptr(), C.c_float(1.0 / world), st)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)


def f_unpack_only():
    _lib.call("mvn_unpack_grads", C.byref(shape), pg.data_ptr(), flat.data_ptr(), offs.data_ptr(), C.c_float(1.0 / world), st)


f_peer(); a = flat.clone()
f_nccl(); b = flat.clone()
res = {"world": world, "flat_MB": flat.numel() * 4 / 1e6, "peer_vs_nccl_max_abs_diff": (a - b).abs().max().item(),
       "peer_us": timed(f_peer), "nccl_us": timed(f_nccl), "unpack_only_us": timed(f_unpack_only)}
if rank == 0:
    print(json.dumps(res))
dist.barrier()
dist.destroy_process_group()
