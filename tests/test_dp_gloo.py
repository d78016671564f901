"""world_size-2 CPU (gloo) test of the data-parallel gradient path: one all-reduce of the flat
gradient buffer, averaged over ranks (the semantics of the reference's DDP wrap,
movenet/trainer.py:230-234)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import movenet_b200
    torch.manual_seed(100 + rank)            # replicas built from different RNG states ...
    m = movenet_b200.WaveNet(2, 2, 16, 8, 8)
    before = torch.cat([p.detach().flatten() for p in m.parameters()])
    m.enable_data_parallel()                 # ... must leave enable_data_parallel with rank 0's parameters (DDP's constructor)
    assert m._dp_world == world
    mine = torch.cat([p.detach().flatten() for p in m.parameters()])
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    same = all(torch.equal(g, gathered[0]) for g in gathered)
    kept = torch.equal(mine, before)         # true on rank 0 only
    offs, total = m._grad_layout(has_video=False)
    flat = torch.full((total,), float(rank + 1)) / world     # (mvn_unpack_grads hands the gradients over scaled by 1 / world)
    m._reduce_grads(flat)
    ok = bool(torch.allclose(flat, torch.full((total,), (1 + world) / 2)))
    # sharding helper: disjoint, covering
    from movenet_b200.parallel import shard_range
    lo, hi = shard_range(10, rank, world)
    out[rank] = (ok and same and (kept == (rank == 0)), lo, hi)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_averages_over_ranks():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert all(out[r][0] for r in range(world))
    assert (out[0][1], out[0][2], out[1][1], out[1][2]) == (0, 5, 5, 10)
