"""Collective plumbing between column shards.

One process per GPU, ``torch.distributed`` for rendezvous and cold-path
collectives (NCCL on GPUs, Gloo for the CPU plumbing configuration of
BASELINE.json).  The HOT path (partial-dot all-reduce inside ``csrc/sgns_pairs.cu``)
does not go through this module on GPUs: it uses symmetric-memory peer
pointers from ``parallel.symm`` inside the kernel.

This replaces the reference's Akka/Aeron star topology: worker
and server are the same S processes; the "client sums partials" step of
``BigWord2VecMatrix.dotprod`` [G] is an all-reduce.
"""
from __future__ import annotations

import datetime
import os
from typing import List, Optional

import torch
import torch.distributed as dist


class Comm:
    """Single-shard communicator (world size 1): every collective is a no-op."""
    rank = 0
    world = 1
    group = None

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        return t

    def all_gather_cols(self, t: torch.Tensor) -> torch.Tensor:
        """[R, K] per rank -> [R, S*K] (column concatenation in rank order)."""
        return t

    def broadcast_object(self, obj, src: int = 0):
        return obj

    def barrier(self):
        pass

    def gather_objects(self, obj, dst: int = 0):
        return [obj]


class TorchDistComm(Comm):
    """Communicator over a ``torch.distributed`` process group."""

    def __init__(self, group=None, control_group=None):
        self.group = group if group is not None else dist.group.WORLD
        # object collectives always run on a CPU-capable (gloo) group when given
        self.control_group = control_group if control_group is not None else self.group
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)

    def all_reduce_sum(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather_cols(self, t):
        t = t.contiguous()
        out = torch.empty((self.world * t.shape[0], t.shape[1]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=self.group)
        # [S*R, K] -> [S, R, K] -> [R, S*K]
        return out.view(self.world, t.shape[0], t.shape[1]).permute(1, 0, 2).reshape(
            t.shape[0], self.world * t.shape[1])

    def broadcast_object(self, obj, src=0):
        lst = [obj]
        dist.broadcast_object_list(lst, src=src, group=self.control_group)
        return lst[0]

    def barrier(self):
        dist.barrier(group=self.group)

    def gather_objects(self, obj, dst=0):
        out: List = [None] * self.world if self.rank == dst else None
        dist.gather_object(obj, out, dst=dst, group=self.control_group)
        return out


def init_process_group(backend: Optional[str] = None, rank: Optional[int] = None,
                       world: Optional[int] = None, master_addr: str = "127.0.0.1",
                       master_port: Optional[int] = None, timeout_s: float = 600.0,
                       device: Optional[torch.device] = None):
    """Initialise ``torch.distributed`` from explicit arguments or the torchrun
    environment.  Always rendezvous on 127.0.0.1 unless told otherwise (the
    container hostname may not resolve)."""
    if dist.is_initialized():
        return
    rank = int(os.environ.get("RANK", 0)) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", 1)) if world is None else world
    os.environ.setdefault("MASTER_ADDR", master_addr)
    if master_port is not None:
        os.environ["MASTER_PORT"] = str(master_port)
    os.environ.setdefault("MASTER_PORT", "29512")
    if backend is None:
        backend = "cpu:gloo,cuda:nccl" if torch.cuda.is_available() else "gloo"
    kwargs = {}
    if device is not None and device.type == "cuda":
        kwargs["device_id"] = device
    dist.init_process_group(backend=backend, rank=rank, world_size=world,
                            timeout=datetime.timedelta(seconds=timeout_s), **kwargs)


def comm_from_env() -> Comm:
    """``TorchDistComm`` when running under an initialised process group with
    more than one rank, else the single-shard ``Comm``."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return TorchDistComm()
    return Comm()
