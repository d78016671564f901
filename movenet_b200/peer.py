"""Gradient averaging over NVLink peer memory (csrc/peer.cu): the data-parallel exchange of one node.

`PeerGradients` owns this rank's cudaIpc-shared exchange buffer (staging slots and the receive area of the sums) and the mappings
of every peer's.  Set up once per (model, video flag) by a collective
handle exchange over the process group; after that a step's exchange is one kernel launch, no NCCL call and no host
synchronisation.  Opt-in with $MOVENET_B200_DP=peer: its sums are taken in rank order (bit-identical results whatever
algorithm NCCL would pick), but NCCL's all-reduce is the faster exchange on the B200 / NVSwitch boxes measured (csrc/peer.cu)
and stays the default.  Any failure during the set-up (ranks on different hosts, peer access unavailable) is agreed on by all ranks
and leaves the NCCL all-reduce in charge (`WaveNet._reduce_grads`): both are device paths, neither is a CPU fallback.
"""
import ctypes as C
import os
import socket

import torch

from . import _lib

MAX_PEERS = 8


def available(group, device) -> bool:
    """opt-in ($MOVENET_B200_DP=peer; NCCL's all-reduce measured faster on NVSwitch B200 boxes, see csrc/peer.cu); needs a CUDA
    device, the NCCL backend (one process per GPU) and 2..8 ranks"""
    import torch.distributed as dist
    if os.environ.get("MOVENET_B200_DP", "nccl") != "peer" or device.type != "cuda":
        return False
    world = dist.get_world_size(group)
    return 2 <= world <= MAX_PEERS and dist.get_backend(group) == "nccl"


class PeerGradients:
    def __init__(self, shape, device, group):
        import torch.distributed as dist
        self.group, self.device = group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.epoch = 0
        self.base = None
        self.peers = []
        stage, recv = C.c_size_t(), C.c_size_t()
        _lib.call("mvn_peer_layout", C.byref(shape), C.byref(stage), C.byref(recv))
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        ok, why = True, ""
        with torch.cuda.device(device):
            try:
                _lib.call("mvn_peer_alloc", C.c_size_t(stage.value + recv.value), C.byref(ptr), handle)
                self.base = ptr.value
            except RuntimeError as e:
                ok, why = False, str(e)
            mine = (socket.gethostname(), os.getpid(), handle.raw if ok else None)
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            if any(h is None for _, _, h in everyone) or len({host for host, _, _ in everyone}) != 1:
                ok, why = False, why or "ranks on several hosts (or a peer could not allocate)"
            table = (C.c_void_p * self.world)()
            if ok:
                for r, (_, _, h) in enumerate(everyone):
                    if r == self.rank:
                        table[r] = self.base
                        self.peers.append(None)
                        continue
                    p = C.c_void_p()
                    try:
                        _lib.call("mvn_peer_open", C.create_string_buffer(h, 64), C.byref(p))
                    except RuntimeError as e:
                        ok, why = False, str(e)
                        break
                    table[r] = p.value
                    self.peers.append(p.value)
            agreed = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
            dist.all_reduce(agreed, op=dist.ReduceOp.MIN, group=group)       # (also: nobody uses a mapping before everybody has it)
            self.ok = bool(agreed.item())
            self.why = why
            self.table = table
            if not self.ok:
                self.close()

    def reduce_unpack(self, shape, packed_grads_ptr, flat, offs, stream):
        """flat <- (1 / world) * sum over ranks of the packed gradients (summed in place), unpacked to the parameter shapes"""
        self.epoch += 1
        _lib.call("mvn_peer_reduce_unpack", C.byref(shape), self.table, self.rank, self.world, C.c_uint(self.epoch),
                  packed_grads_ptr, flat.data_ptr(), offs.data_ptr(), C.c_float(1.0 / self.world), stream)

    def close(self):
        lib = _lib.load()
        with torch.cuda.device(self.device):
            for p in self.peers:
                if p:
                    lib.mvn_peer_close(C.c_void_p(p))
            self.peers = []
            if self.base:
                lib.mvn_peer_free(C.c_void_p(self.base))
                self.base = None

    def __del__(self):
        try:
            if torch.cuda.is_available():
                torch.cuda.synchronize(self.device)
                self.close()
        except Exception:
            pass
