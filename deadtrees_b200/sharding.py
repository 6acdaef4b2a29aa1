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


# ---- tile-range shards: balanced compute, row-aligned stitching (8 GPUs: 253 / 254 tiles each instead of 6 / 5 tile rows) ----

class ShardPlan:
    """Work of one rank when the ``gy * gx`` tiles of a mosaic are split into ``world`` contiguous, near-equal tile-index
    ranges.  Whole tile rows (45 rows over 8 GPUs = 6,6,6,6,6,5,5,5) bound the speed-up at 7.5x; tile ranges differ by
    at most one tile.  Stitching stays row-aligned: tile row ``r`` is stitched by the rank that computed its FIRST tile, so

    * ``compute = [t0, t1)``            tiles this rank runs through the network,
    * ``rows = [R0, R1)``               tile rows whose mask rows this rank stitches (``mask_rows`` in mosaic pixels),
    * ``buffer = [B0, B1)``             tile-row aligned range of its logits buffer, one halo tile row in front,
    * ``recv_tail`` / ``send_head``     whole logits of the tiles of row ``R1 - 1`` that the NEXT rank computed,
    * ``recv_halo`` / ``send_halo``     bottom ``overlap`` rows of the tiles of row ``R0 - 1`` that the PREVIOUS rank computed.

    Every range is a (first tile, one-past-last tile) pair of global tile indices, or ``None``."""

    def __init__(self, gy: int, gx: int, world: int, rank: int, overlap: int):
        n = gy * gx
        if world > 1 and overlap > 0 and n // world < gx + 1:
            raise ValueError(f"{n} tiles over {world} ranks: tile-range shards need more than one tile row ({gx} tiles) per rank")
        bounds = [(n * k) // world for k in range(world + 1)]
        if overlap == 0:       # no blending: a tile's pixels are final where they are computed - keep whole tile rows together
            bounds = [min(gy, (gy * k + world // 2) // world) * gx for k in range(world + 1)]
        first_row = [-(-b // gx) for b in bounds]                 # first tile row that STARTS at or after the bound
        self.gy, self.gx, self.world, self.rank, self.overlap = gy, gx, world, rank, overlap
        self.t0, self.t1 = bounds[rank], bounds[rank + 1]
        self.R0, self.R1 = first_row[rank], (gy if rank == world - 1 else first_row[rank + 1])
        halo = 1 if (self.R0 > 0 and overlap > 0) else 0
        self.ty_base = self.R0 - halo
        self.B0 = min(self.ty_base * gx, self.t0)
        self.B1 = max(self.t1, self.R1 * gx)
        self.prev = rank - 1 if rank > 0 else None
        self.next = rank + 1 if rank + 1 < world else None
        # tiles of my last stitched row that the next rank computed (it sends them whole)
        self.recv_tail = (self.t1, self.R1 * gx) if (self.next is not None and self.R1 * gx > self.t1) else None
        # tiles at the head of my range that belong to a row the previous rank stitches
        self.send_head = (self.t0, self.R0 * gx) if (self.prev is not None and self.R0 * gx > self.t0) else None
        # halo row R0 - 1: the part the previous rank computed arrives as bottom strips
        self.recv_halo = ((self.R0 - 1) * gx, self.t0) if (halo and (self.R0 - 1) * gx < self.t0) else None
        nxt_R0 = first_row[rank + 1] if self.next is not None else None
        self.send_halo = (((nxt_R0 - 1) * gx, self.t1) if (self.next is not None and overlap > 0 and nxt_R0 > 0
                                                          and (nxt_R0 - 1) * gx < self.t1) else None)
        if self.send_halo is not None and self.send_halo[0] < self.t0:
            raise ValueError("a tile row spans more than two ranks")

    def mask_rows(self, H: int, T: int) -> Tuple[int, int]:
        s = T - self.overlap
        return min(H, self.R0 * s), (H if self.R1 >= self.gy else min(H, self.R1 * s))

    def input_rows(self, H: int, T: int) -> Tuple[int, int]:
        """mosaic rows the tiles [t0, t1) read"""
        s = T - self.overlap
        return min(H, (self.t0 // self.gx) * s), min(H, ((self.t1 - 1) // self.gx) * s + T)


def exchange_logits(plan: ShardPlan, logits: torch.Tensor, tile: int, group=None) -> None:
    """one grouped point-to-point exchange (NCCL over NVLink; gloo in the CPU tests) after a rank has computed its
    tiles: whole head tiles go to the previous rank, boundary strips to the next one.  ``logits``: (B1 - B0, T, T, K)."""
    if plan.world == 1:
        return
    ov, B0 = plan.overlap, plan.B0
    ops, copies = [], []
    if plan.send_head is not None:
        a, b = plan.send_head
        ops.append(dist.P2POp(dist.isend, logits[a - B0: b - B0], plan.prev, group))           # contiguous tiles
    if plan.recv_tail is not None:
        a, b = plan.recv_tail
        ops.append(dist.P2POp(dist.irecv, logits[a - B0: b - B0], plan.next, group))
    if plan.send_halo is not None:
        a, b = plan.send_halo
        ops.append(dist.P2POp(dist.isend, logits[a - B0: b - B0, tile - ov:].contiguous(), plan.next, group))
    if plan.recv_halo is not None:
        a, b = plan.recv_halo
        buf = torch.empty((b - a, ov, tile, logits.shape[-1]), dtype=logits.dtype, device=logits.device)
        ops.append(dist.P2POp(dist.irecv, buf, plan.prev, group))
        copies.append((a, b, buf))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for a, b, buf in copies:
        logits[a - B0: b - B0, tile - ov:].copy_(buf)


def gather_mask_rows(plans, rank: int, mask: torch.Tensor, H: int, T: int, group=None) -> None:
    """every rank's stitched mask rows -> rank 0, as ONE group of point-to-point transfers (they run concurrently over
    NVSwitch instead of one after the other on the default stream)."""
    if len(plans) == 1:
        return
    ops = []
    if rank == 0:
        for k in range(1, len(plans)):
            y0, y1 = plans[k].mask_rows(H, T)
            if y1 > y0:
                ops.append(dist.P2POp(dist.irecv, mask[y0:y1], k, group))
    else:
        y0, y1 = plans[rank].mask_rows(H, T)
        if y1 > y0:
            ops.append(dist.P2POp(dist.isend, mask[y0:y1], 0, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
