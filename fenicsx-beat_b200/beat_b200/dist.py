"""Rank plumbing for multi-GPU runs: one process per GPU, torch.distributed carries the NCCL unique id."""

from __future__ import annotations

import os

_UID_CACHE: dict = {}


def init_comm(ctx, comm) -> None:
    """Create the context's NCCL communicator (idempotent per context)."""
    if getattr(ctx, "_comm_ready", False) or comm.size == 1:
        return
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=comm.rank, world_size=comm.size)
    uid = ctx.comm_unique_id() if comm.rank == 0 else bytes(128)
    t = torch.tensor(list(uid), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0)
    ctx.comm_init(comm.size, comm.rank, bytes(t.cpu().tolist()))
    ctx._comm_ready = True
