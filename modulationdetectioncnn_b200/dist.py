"""Frame sharding across GPUs (one process per GPU) and the class-histogram reduction.

Frames (and FWHT spectra) are independent, so the stream is cut into contiguous
ranges ``[r*N/W, (r+1)*N/W)`` with no data-path collective; the only exchange is one
all-reduce(SUM) of the ``int64[C]`` class histogram, i.e. the
multi-GPU form of the confusion/accuracy bookkeeping of cnn.py:200-255.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np

__all__ = ["shard_range", "init_process_group", "allreduce_histogram", "rank_world", "bind_to_gpu_numa_node"]


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced (sizes differ by at most 1), order-preserving partition."""
    if not 0 <= rank < world:
        raise ValueError("rank outside [0, world)")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)),
            int(os.environ.get("LOCAL_RANK", 0)))


def init_process_group(backend: str = "auto"):
    import torch
    import torch.distributed as dist
    if dist.is_initialized():
        return
    if backend == "auto":
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    rank, world, local = rank_world()
    if backend == "nccl":
        torch.cuda.set_device(local)
        dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend, rank=rank, world_size=world)


def allreduce_histogram(hist):
    """Sum an int64[C] histogram over all ranks (NCCL on GPU, gloo on CPU); returns numpy."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(hist.cpu() if hasattr(hist, "cpu") else hist, dtype=np.int64)
    # always reduce a private copy: the caller's array must keep its local counts
    t = (hist.detach().clone() if hasattr(hist, "is_cuda") else torch.tensor(np.array(hist, dtype=np.int64)))
    t = t.to(torch.int64)
    if dist.get_backend() == "nccl" and not t.is_cuda:
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def _parse_cpulist(text: str):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device: int = 0):
    """Pin this process (and, by first touch, the host buffers it allocates afterwards) to the NUMA node the GPU
    hangs off.  With one rank per GPU streaming ~50 GB/s of frames each, pinned buffers that land on the other
    socket put every copy on the inter-socket link.  Returns the node, or None when the topology is not exposed
    (no sysfs entry, single node) - binding is an optimisation, never a requirement."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read())
        if node < 0:
            return None
        cpus = set(_parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None
