"""Symmetric (peer-mapped) device memory for in-kernel NVLink communication.

The fused SGNS kernel exchanges partial dot products by storing straight into
its peers' memory and spinning on flags in its own -- no NCCL call on the hot
path.  This module allocates the buffers with
``torch.distributed._symmetric_memory`` (CUDA VMM + fabric/fd handle exchange
under the hood) and exposes raw peer pointers (and the NVLS multicast pointer
when the driver grants one) for the kernels.

Replaces: the Akka ask-pattern fan-out/fan-in between Spark workers and Glint
servers.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import torch
import torch.distributed as dist


@dataclass
class SymmBuffer:
    local: torch.Tensor            # this rank's buffer (uint8 view of the symmetric allocation)
    ptrs: List[int]                # device pointers of every rank's buffer, mapped into THIS process
    multicast_ptr: int             # 0 when NVLS multicast is unavailable
    handle: object

    def barrier(self):
        self.handle.barrier()


def alloc_symmetric(nbytes: int, device: torch.device, group=None) -> SymmBuffer:
    """Collective: allocate ``nbytes`` of zeroed symmetric memory on every rank."""
    import torch.distributed._symmetric_memory as symm_mem
    group = group if group is not None else dist.group.WORLD
    nbytes = (nbytes + 15) // 16 * 16
    t = symm_mem.empty(nbytes, dtype=torch.uint8, device=device)
    hdl = symm_mem.rendezvous(t, group)
    t.zero_()
    torch.cuda.synchronize(device)
    hdl.barrier()
    mc = 0
    try:
        if hdl.has_multicast_support:
            mc = int(hdl.multicast_ptr)
    except Exception:
        mc = 0
    return SymmBuffer(t, [int(p) for p in hdl.buffer_ptrs], mc, hdl)
