"""Batch sharding over the GPUs of one box (one process per GPU, torch.distributed as plumbing).

The path shards trivially: image b's logits depend only on image b and the replicated (<= 13 kB) weights
(SURVEY.md 8e).  Each rank takes one contiguous slice, runs the single-GPU engine on it, and the only exchange
is the optional gather of the int32 class scores (or just the labels) at the end -- NCCL over NVLink on GPUs,
gloo in the CPU tests.  There is no collective on the data path.
"""
from __future__ import annotations

import os
from typing import Callable, Optional, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [begin, end) of rank `rank`; sizes differ by at most one; concatenation in rank order = batch."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def env_rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when launched directly."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_process_group(backend: Optional[str] = None):
    """Initialise torch.distributed from the torchrun environment (127.0.0.1 rendezvous by default)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_rank_world()
    if world == 1 or dist.is_initialized():
        return rank, local_rank, world
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kwargs = {}
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        kwargs["device_id"] = torch.device("cuda", local_rank)
    dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, local_rank, world


def gather_rows(local, total_rows: int, group=None):
    """All-gather per-rank row blocks ([n_r, ...] tensors, n_r from shard_range) into the full [total_rows, ...] tensor,
    rank order = batch order.  Ragged shards are padded to the largest shard for the collective and trimmed after."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(total_rows, r, world) for r in range(world)]
    max_rows = max(e - b for b, e in sizes)
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    if all(e - b == max_rows for b, e in sizes):
        return out
    return torch.cat([out[r * max_rows: r * max_rows + (e - b)] for r, (b, e) in enumerate(sizes)])


def sharded_infer(infer_local: Callable, images, gather: str = "logits"):
    """Run `infer_local(images_shard) -> (logits, labels)` on this rank's slice of `images` and (optionally) gather.

    gather = "logits": every rank gets (logits [n, C], labels [n]);  "labels": only labels are exchanged
    (4 B/image instead of 4*C: the NVLink-ingest bound of SURVEY.md 8e);  "none": results stay sharded.

    Device tensors stay device tensors end to end: when `infer_local` returns torch tensors (e.g. ``Engine.infer_tensor`` on
    a CUDA shard) they go into the collective as they are and the gathered results come back as tensors on the same device --
    no host bounce.  NumPy in, NumPy out (the gloo CPU tests; under NCCL the arrays are staged through the GPU once).
    For the exchange fused into the kernel's epilogue (no collective at all) see ``bitnetmcu_b200.gather``.
    """
    import torch
    import torch.distributed as dist
    rank, _, world = env_rank_world()
    if dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    n = images.shape[0]
    b, e = shard_range(n, rank, world)
    logits, labels = infer_local(images[b:e])
    if world == 1 or gather == "none":
        return logits, labels
    as_numpy = not torch.is_tensor(logits)
    if as_numpy:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t_logits = torch.from_numpy(np.ascontiguousarray(logits)).to(dev)
        t_labels = torch.from_numpy(np.ascontiguousarray(labels).astype(np.int32)).to(dev)
    else:
        t_logits, t_labels = logits, labels.view(torch.int32) if labels.dtype != torch.int32 else labels
    all_labels = gather_rows(t_labels, n)
    all_logits = gather_rows(t_logits, n) if gather == "logits" else t_logits
    if as_numpy:
        return all_logits.cpu().numpy(), all_labels.cpu().numpy().astype(np.uint32)
    return all_logits, all_labels
