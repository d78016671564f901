"""Data-parallel gradient averaging, peer-memory kernel and NCCL (needs >= 2 GPUs; skipped otherwise): two ranks on two half-batches
must end up with the gradients one process computes on the whole batch (equal shard sizes, mean loss) --
the semantics of the reference's DistributedDataParallel wrap (movenet/trainer.py:230-234)."""
import os
import socket

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir, exchange):
    import torch.distributed as dist
    import movenet_b200
    os.environ["MOVENET_B200_DP"] = exchange
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    torch.manual_seed(0 if rank == 0 else 77 + rank)      # only rank 0 holds the weights the single process will use
    m = movenet_b200.WaveNet(2, 2, 32, 16, 8, compute_dtype="fp32").cuda().enable_data_parallel()
    g = torch.Generator().manual_seed(1)
    for step in range(3):       # the third step's gradients are compared (the peer exchange alternates two buffers by step)
        codes = torch.randint(0, 32, (4, 300), generator=g)
        mine = codes[2 * rank:2 * rank + 2].cuda()
        m.zero_grad(set_to_none=True)
        out = m(mine)
        F.cross_entropy(out, mine[:, m.receptive_fields:]).backward()
    used_peer = any(v is not None for v in m._dp_peer.values())
    assert used_peer == (exchange == "peer"), "the gradient exchange that ran is not the one asked for"
    torch.save({k: v.grad.cpu() for k, v in m.named_parameters() if v.grad is not None}, os.path.join(out_dir, f"g{rank}.pt"))
    opt = movenet_b200.optim.AdamW(m.parameters(), lr=1e-3)
    opt.step()                                              # replicas must still agree after an optimizer step
    torch.save({k: v.detach().cpu() for k, v in m.named_parameters()}, os.path.join(out_dir, f"w{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("exchange", ["peer", "nccl"])        # csrc/peer.cu over NVLink peer memory (default) / NCCL all-reduce
def test_gradient_average_matches_single_process(tmp_path, exchange):
    import torch.multiprocessing as mp
    import movenet_b200
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), exchange), nprocs=2, join=True)
    g0 = torch.load(tmp_path / "g0.pt"); g1 = torch.load(tmp_path / "g1.pt")
    w0 = torch.load(tmp_path / "w0.pt"); w1 = torch.load(tmp_path / "w1.pt")
    assert all(torch.equal(w0[k], w1[k]) for k in w0)
    torch.manual_seed(0)
    m = movenet_b200.WaveNet(2, 2, 32, 16, 8, compute_dtype="fp32").cuda()
    gen = torch.Generator().manual_seed(1)
    for step in range(3):
        codes = torch.randint(0, 32, (4, 300), generator=gen).cuda()
    F.cross_entropy(m(codes), codes[:, m.receptive_fields:]).backward()
    for k, v in m.named_parameters():
        if v.grad is None:
            continue
        assert torch.equal(g0[k], g1[k]), k                       # both ranks hold the same averaged gradient
        err = ((g0[k] - v.grad.cpu()).norm() / v.grad.norm().clamp_min(1e-30)).item()
        assert err < 1e-4, (k, err)
