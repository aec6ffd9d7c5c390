"""Data-parallel sharding of an evaluation batch across the GPUs of one box (one process per GPU).

The u-LLaVA inference path has no cross-image dependency: every rank holds a full
weight replica, takes a contiguous slice of the global batch and runs the whole forward locally.  The
only exchange is ONE all-gather of the results the eval collator needs (token ids + bit-packed masks),
a few KB per image, so NCCL's all_gather over NVLink/NVSwitch is used as is (gloo on CPU for tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(global_batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the global batch owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_results(output_ids: torch.Tensor, mask_bits: Sequence[torch.Tensor], max_masks: int,
                 words_per_mask: int) -> torch.Tensor:
    """One int32 payload row per image: [n_masks, ids(2 words each, int64 split), bits(max_masks*words)]."""
    B, T = output_ids.shape
    dev = output_ids.device
    row = 1 + 2 * T + max_masks * words_per_mask
    out = torch.zeros((B, row), dtype=torch.int32, device=dev)
    out[:, 1:1 + 2 * T] = output_ids.contiguous().view(torch.int32).view(B, 2 * T)
    for i, bits in enumerate(mask_bits):
        n = 0 if bits is None else min(int(bits.shape[0]), max_masks)
        out[i, 0] = n
        if n:
            if bits.shape[1] != words_per_mask:
                raise ValueError("all images of a gathered batch must share the output size")
            out[i, 1 + 2 * T: 1 + 2 * T + n * words_per_mask] = bits[:n].reshape(-1)
    return out


def unpack_results(payload: torch.Tensor, T: int, max_masks: int, words_per_mask: int):
    B = payload.shape[0]
    ids = payload[:, 1:1 + 2 * T].contiguous().view(torch.int64).view(B, T)
    counts = payload[:, 0].tolist()
    bits = [payload[i, 1 + 2 * T: 1 + 2 * T + counts[i] * words_per_mask].view(counts[i], words_per_mask)
            for i in range(B)]
    return ids, bits


def all_gather_results(payload: torch.Tensor, shard_sizes: Sequence[int]) -> torch.Tensor:
    """The single collective of the path: gathers every rank's payload rows, in rank order.
    Ragged shards are padded to the largest shard and trimmed after the gather."""
    rank, ws = world()
    if ws == 1:
        return payload
    mx = max(shard_sizes)
    if payload.shape[0] < mx:
        pad = torch.zeros((mx - payload.shape[0], payload.shape[1]), dtype=payload.dtype, device=payload.device)
        payload = torch.cat([payload, pad], 0)
    out = torch.empty((ws * mx, payload.shape[1]), dtype=payload.dtype, device=payload.device)
    dist.all_gather_into_tensor(out, payload.contiguous())
    if all(s == mx for s in shard_sizes):
        return out
    return torch.cat([out[r * mx: r * mx + shard_sizes[r]] for r in range(ws)], 0)


def unpack_bits(bits: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """[n, words] int32 -> bool [n, h, w] (bit i of word j = pixel 32*j + i)."""
    n = bits.shape[0]
    shifts = torch.arange(32, device=bits.device, dtype=torch.int32)
    b = ((bits.unsqueeze(-1) >> shifts) & 1).bool().reshape(n, bits.shape[1] * 32)
    return b[:, : h * w].view(n, h, w)
