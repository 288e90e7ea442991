"""Batch sharding over the ranks of one box and the single result gather.

Clips are independent, weights are replicated: rank r encodes/decodes the contiguous shard ``[r*n, (r+1)*n)`` and
the only exchange is one all-gather of codes (int64) and reconstructed audio (fp32).  Works on NCCL (GPU tensors)
and gloo (CPU tensors, used by the world-size-2 CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous equal shards; ``total`` must divide evenly (288 = 8 x 36 in BASELINE config 5)."""
    if total % world:
        raise ValueError(f"batch of {total} clips does not split evenly over {world} ranks; pad the last shard")
    n = total // world
    return rank * n, (rank + 1) * n


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_results(codes: torch.Tensor, audio: torch.Tensor, out_codes: torch.Tensor = None,
                   out_audio: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather the per-rank results along the batch axis (rank order = clip order)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return codes, audio
    world = dist.get_world_size()
    if out_codes is None:
        out_codes = codes.new_empty((world * codes.shape[0],) + tuple(codes.shape[1:]))
    if out_audio is None:
        out_audio = audio.new_empty((world * audio.shape[0],) + tuple(audio.shape[1:]))
    dist.all_gather_into_tensor(out_codes, codes.contiguous())
    dist.all_gather_into_tensor(out_audio, audio.contiguous())
    return out_codes, out_audio
