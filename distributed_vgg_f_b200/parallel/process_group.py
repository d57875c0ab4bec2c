"""Rendezvous and rank -> device mapping.

Reference (distributedVggf.py:283-289, :296): ``init_process_group(backend="gloo",
init_method=<-iu url>, world_size, rank)`` when ``-ws > 1``, after pinning gloo to ``eth0``; no
device selection at all, so on a multi-GPU host every rank would sit on ``cuda:0``.

Here torch.distributed is kept for exactly two things: the rendezvous (TCP store behind the same
``tcp://host:port`` URL) and the exchange of peer-memory handles for the symmetric gradient arena.
Each rank binds to ``cuda:(rank % device_count)``; the backend is NCCL when every rank has a GPU
and gloo otherwise (``-nc``).  Gradient traffic never goes through this process group on the GPU
path -- it goes through ``parallel.symm`` + the fused all-reduce kernels.
"""
from __future__ import annotations

import datetime
import os
from typing import Optional

import torch
import torch.distributed as dist

from ..config import CPU_BACKEND, GPU_BACKEND


def distributed_is_initialized() -> bool:
    """Reference: distributedVggf.py:27-32."""
    return dist.is_available() and dist.is_initialized()


def world_size() -> int:
    return dist.get_world_size() if distributed_is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if distributed_is_initialized() else 0


def pick_device(rank_: int, no_cuda: bool = False) -> torch.device:
    if torch.cuda.is_available() and not no_cuda:
        local = int(os.environ.get("LOCAL_RANK", rank_ % torch.cuda.device_count()))
        torch.cuda.set_device(local)
        return torch.device("cuda", local)
    return torch.device("cpu")


def init_distributed(init_url: Optional[str], rank_: int, world_size_: int, device: torch.device,
                     backend: Optional[str] = None, timeout_s: int = 600) -> None:
    """Join the job.  ``init_url`` is the reference's ``-iu`` value (``tcp://host:port``) or
    ``env://`` when launched by torchrun."""
    if world_size_ <= 1 or distributed_is_initialized():
        return
    if backend is None:
        backend = GPU_BACKEND if device.type == "cuda" else CPU_BACKEND
    # The reference hard-codes GLOO_SOCKET_IFNAME=eth0 (distributedVggf.py:296); only honour an
    # explicit user choice, otherwise let gloo auto-detect (loopback runs break on 'eth0').
    kwargs = dict(backend=backend, init_method=init_url or "env://", world_size=world_size_,
                  rank=rank_, timeout=datetime.timedelta(seconds=timeout_s))
    if backend == "nccl":
        kwargs["device_id"] = device
    dist.init_process_group(**kwargs)


def shutdown(graceful: bool = True) -> None:
    """Leave the job.  ``graceful=False`` (this rank failed): no barrier -- the peers are not going to
    reach one -- and no blocking teardown."""
    if distributed_is_initialized():
        if not graceful:
            return
        try:
            dist.barrier()
        except Exception:
            pass
        dist.destroy_process_group()


def device_name(device: torch.device) -> str:
    """The reference prints torch.cuda.get_device_name(device) unconditionally and crashes on
    CPU-only hosts (distributedVggf.py:218); print 'cpu' there instead."""
    return torch.cuda.get_device_name(device) if device.type == "cuda" else "cpu"
