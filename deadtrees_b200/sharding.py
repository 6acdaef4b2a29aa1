"""Tile-row sharding of whole-mosaic inference across the GPUs of one box.

The reference is single-GPU (``configs/trainer/default.yaml:3``); tiles are independent (each is convolved
with zero padding at its own border), so the tile grid shards by contiguous tile rows with no data-path
collective.  With overlap > 0 the blended output rows at a shard boundary need the bottom ``overlap``
rows of the previous shard's last tile row: one point-to-point send per boundary (NCCL over NVLink on
GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def split_tile_rows(gy: int, world_size: int) -> List[Tuple[int, int]]:
    """contiguous, near-equal ranges of tile rows; earlier ranks take the remainder (45 -> 6,6,6,6,6,5,5,5)."""
    base, rem = divmod(gy, world_size)
    out, r = [], 0
    for k in range(world_size):
        n = base + (1 if k < rem else 0)
        out.append((r, r + n))
        r += n
    return out


def make_halo_hook(tile: int, overlap: int, rank: int, world_size: int, group=None, has_rows=None):
    """Returns ``hook(logits, gx, halo)`` for ``MosaicInference.run``: every shard sends the bottom
    ``overlap`` rows of its LAST tile row to the next shard, which stores them in the halo tile row it
    keeps in front of its own tile rows.  ``has_rows[k]`` tells which ranks hold any tile rows."""

    def hook(logits: torch.Tensor, gx: int, halo: int) -> None:
        if world_size == 1 or overlap == 0:
            return
        active = [k for k in range(world_size) if has_rows is None or has_rows[k]]
        if rank not in active:
            return
        i = active.index(rank)
        nxt = active[i + 1] if i + 1 < len(active) else None
        prv = active[i - 1] if i > 0 else None
        ops = []
        send_buf = recv_buf = None
        if nxt is not None:
            send_buf = logits[logits.shape[0] - gx:, tile - overlap:, :, :].contiguous()
            ops.append(dist.P2POp(dist.isend, send_buf, nxt, group))
        if prv is not None and halo:
            recv_buf = torch.empty((gx, overlap, tile, logits.shape[-1]), dtype=logits.dtype, device=logits.device)
            ops.append(dist.P2POp(dist.irecv, recv_buf, prv, group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        if recv_buf is not None:
            logits[:gx, tile - overlap:, :, :].copy_(recv_buf)

    return hook
