"""Replica data-parallel inference across the GPUs of one box (one process per GPU, torch.distributed/NCCL).

The reference's multi-GPU mode is per-rank prompt slicing with no tensor collective at all
(eval/gen/gen_images_mp.py:188-191, 238). Samples of a packed batch never attend to each other, so the path
shards by sample: weights replicated, each rank denoises its slice, and the only exchange is ONE all-gather of
the final latents [B_local, tokens, 64] fp32 (1 MiB per 1024^2 sample) over NVLink.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced slice [lo, hi) of n_items for `rank` (first n_items % world ranks get one extra)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_list(items: Sequence, rank: int = None, world: int = None) -> List:
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    lo, hi = shard_range(len(items), rank, world)
    return list(items[lo:hi])


def gather_latents(local: torch.Tensor) -> torch.Tensor:
    """[B_local, T, C] on every rank -> [world*B_local, T, C] on every rank (rank-major), one all_gather."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


def gather_ragged_latents(local: Sequence[torch.Tensor], all_token_counts: Sequence[Sequence[int]]) -> List[torch.Tensor]:
    """Ragged variant: every rank contributes a list of [T_i, C] latents; `all_token_counts[r]` lists rank r's T_i
    (known on the host from the image sizes). Pads to the per-rank max total, one all_gather, then splits."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return list(local)
    world = dist.get_world_size()
    C = local[0].shape[1] if len(local) else 64
    totals = [sum(c) for c in all_token_counts]
    pad_to = max(totals)
    dev = local[0].device if len(local) else torch.device("cuda", torch.cuda.current_device())
    buf = torch.zeros((pad_to, C), dtype=torch.float32, device=dev)
    if len(local):
        cat = torch.cat(list(local), 0)
        buf[: cat.shape[0]] = cat
    out = torch.empty((world * pad_to, C), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(out, buf)
    res = []
    for r in range(world):
        chunk = out[r * pad_to: r * pad_to + totals[r]]
        res.extend(chunk.split(list(all_token_counts[r])))
    return res
