"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch).

* MATCH shards the pair list; there is no data-path collective — results are
  gathered on the host with `all_gather_object`.
* BA shards observations by point inside the C library; the one exchange step per LM iteration
  (sum of the partial reduced camera systems S_g, rhs_g and a few scalars) is done by the
  all-reduce callable built here, which wraps the library's raw device pointer in a tensor
  and calls `torch.distributed.all_reduce` on the library's stream.

With the `gloo` backend (CPU tests) the same code path runs on host pointers.
"""
from __future__ import annotations

import ctypes
import os
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np


def init_from_env(backend: Optional[str] = None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* (torchrun)."""
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


class _CudaPtr:
    """Zero-copy view of `count` float64 at a raw device pointer."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3,
                                         "strides": None}


def make_allreduce(group=None, device: Optional[int] = None) -> Callable[[int, int, int], None]:
    """Returns allreduce(ptr, count, stream): in-place sum over ranks of float64[count] at `ptr`."""
    import torch
    import torch.distributed as dist

    backend = dist.get_backend(group)
    views: Dict[Tuple[int, int], Any] = {}    # (ptr, count) -> tensor view (the library reuses its buffers)
    streams: Dict[int, Any] = {}

    def allreduce(ptr: int, count: int, stream: int) -> None:
        if backend == "nccl":
            dev = torch.cuda.current_device() if device is None else device
            t = views.get((ptr, count))
            if t is None:
                if len(views) > 256:
                    views.clear()
                t = views[(ptr, count)] = torch.as_tensor(_CudaPtr(ptr, count), device="cuda:%d" % dev)
            ext = streams.get(stream)
            if ext is None:
                ext = streams[stream] = (torch.cuda.ExternalStream(stream, device=dev) if stream
                                         else torch.cuda.current_stream(dev))
            with torch.cuda.stream(ext):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        else:  # gloo: host memory
            arr = np.ctypeslib.as_array((ctypes.c_double * count).from_address(ptr))
            t = torch.from_numpy(arr)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    return allreduce


def gather_pair_results(local: Dict[Tuple[Any, Any], np.ndarray], world: int, group=None) -> Dict[Tuple[Any, Any], np.ndarray]:
    """Host-side gather of per-pair match arrays from every rank (no device collective)."""
    if world == 1:
        return dict(local)
    import torch.distributed as dist

    parts: List[Optional[Dict]] = [None] * world
    dist.all_gather_object(parts, local, group=group)
    out: Dict[Tuple[Any, Any], np.ndarray] = {}
    for p in parts:
        out.update(p)
    return out
