"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink; gloo in CPU tests).

The DDPM path shards naturally: training shards the batch and needs ONE exchange per step,
a SUM all-reduce of the flat fp32 gradient arena (each rank's gradients are pre-scaled by 1/global_batch, so the
sum is the gradient of the global-mean loss); sampling shards sample_size and needs no per-step exchange."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> None:
    """Initialise the default process group when launched by torchrun (RANK / WORLD_SIZE / LOCAL_RANK set)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or dist.is_initialized():
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    else:
        dist.init_process_group(backend)


def shutdown() -> None:
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def shard_size(global_batch: int) -> int:
    w = world_size()
    if global_batch % w != 0:
        raise ValueError(f"global batch {global_batch} is not divisible by the number of ranks {w}")
    return global_batch // w


def shard_rows(batch, r: int = None, w: int = None):
    """Rows [r*B/w, (r+1)*B/w) of a global batch (numpy array or tensor)."""
    w = world_size() if w is None else w
    r = rank() if r is None else r
    if w == 1:
        return batch
    n = batch.shape[0]
    if n % w != 0:
        raise ValueError(f"batch of {n} rows is not divisible by {w} ranks")
    per = n // w
    return batch[r * per:(r + 1) * per]


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def gather_rows(t: torch.Tensor) -> torch.Tensor:
    """Concatenate equally sized per-rank row blocks on every rank (sampling output)."""
    if world_size() == 1:
        return t
    out = [torch.empty_like(t) for _ in range(world_size())]
    dist.all_gather(out, t.contiguous())
    return torch.cat(out, dim=0)
