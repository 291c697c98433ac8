"""Multi-GPU layout of the path: one process per GPU, PPO environments sharded as independent
replicas with no data-path collective; the only exchange is the gradient all-reduce (rl/ppo_trainer.py)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_replicas(total: int, world: int, rank: int):
    """(first replica, count) owned by `rank` when `total` environment replicas are split over `world` ranks: the
    first total % world ranks take one extra."""
    if total < world:
        raise ValueError("fewer replicas than ranks")
    base, extra = divmod(total, world)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def init_from_env(backend: str | None = None):
    """Joins the process group the launcher (torchrun) described through RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*.
    Returns (rank, world, device). A single process needs no group."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local) if torch.cuda.is_available() else torch.device("cpu")
    if dev.type == "cuda":
        torch.cuda.set_device(dev)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group(backend or ("nccl" if dev.type == "cuda" else "gloo"),
                                **({"device_id": dev} if dev.type == "cuda" else {}))
    return rank, world, dev
