"""Oracle: numpy restatement of the reference tiler arithmetic.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows
* ``deadtrees/utils/data_handling.py:9-19``  ``make_blocks_vectorized``
* ``deadtrees/utils/data_handling.py:22-34`` ``unmake_blocks_vectorized``
* ``deadtrees/deployment/tiler.py:34-56,105-132,142-170`` (``inspect_tile`` and the array
  logic of ``Tiler.load_file / get_batches / put_batches`` around an in-memory ndarray;
  GeoTIFF I/O is out of scope).

PARITY: pinned by the reference's golden vector (``tests/test_tiler.py:56-77``) and by
``tests/golden/tiler_blocks.npz`` produced by running the reference's own functions
(``oracle/make_golden.py``).  The overlap/blend functions at the bottom are a build
extension the reference does not have: **parity unpinned**; their
degenerate case (overlap 0) is checked against the pinned functions.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np


def make_blocks(x: np.ndarray, d: int) -> np.ndarray:
    """(p, m, n) -> (m/d * n/d, p, d, d); block index = row_block * (n/d) + col_block."""
    p, m, n = x.shape
    if m % d or n % d:
        raise ValueError("tile not divisible by subtile")
    out = np.empty(((m // d) * (n // d), p, d, d), dtype=x.dtype)
    for bi in range(m // d):
        for bj in range(n // d):
            out[bi * (n // d) + bj] = x[:, bi * d:(bi + 1) * d, bj * d:(bj + 1) * d]
    return out


def unmake_blocks(x: np.ndarray, d: int, m: int, n: int) -> np.ndarray:
    """(blocks, d, d) -> (m, n), inverse of ``make_blocks`` for one channel."""
    x = np.asarray(x)
    out = np.empty((m, n), dtype=x.dtype)
    nb = n // d
    for b in range((m // d) * nb):
        bi, bj = divmod(b, nb)
        out[bi * d:(bi + 1) * d, bj * d:(bj + 1) * d] = x[b]
    return out


def inspect_shape(shape: Tuple[int, int], tile_shape, subtile_shape) -> Tuple[int, int]:
    """``inspect_tile`` on a bare (H, W): number of subtiles that hold data (tiler.py:45-54)."""
    for a, b in zip(tile_shape, subtile_shape):
        if b == 0 or a % b:
            raise ValueError(f"Shapes unaligned: {a, b}")
    return (math.ceil(shape[0] / subtile_shape[0]), math.ceil(shape[1] / subtile_shape[1]))


class TilerOracle:
    """Array logic of ``deadtrees.deployment.tiler.Tiler`` (tiler.py:59-170) without GeoTIFF I/O."""

    def __init__(self, tile_shape=(2048, 2048), subtile_shape=(256, 256)):
        if subtile_shape[0] != subtile_shape[1]:
            raise ValueError("Subtile required to have matching x/y dims")
        self.tile_shape, self.subtile_shape = tuple(tile_shape), tuple(subtile_shape)

    def load_array(self, sv: np.ndarray) -> None:
        """``sv``: (bands, H, W) uint8 as ``rioxarray.open_rasterio(...).values`` returns."""
        self.size = tuple(sv.shape[1:])
        self.subtiles = inspect_shape(self.size, self.tile_shape, self.subtile_shape)
        if self.tile_shape != self.size:  # tiler.py:106-111 (4 bands hard-coded there)
            self.indata = np.zeros((sv.shape[0], *self.tile_shape), dtype=sv.dtype)
            self.indata[:, : sv.shape[1], : sv.shape[2]] = sv
        else:
            self.indata = sv
        grid = (self.tile_shape[0] // self.subtile_shape[0], self.tile_shape[1] // self.subtile_shape[1])
        use = np.zeros(grid, dtype=bool)
        use[: self.subtiles[0], : self.subtiles[1]] = True  # tiler.py:122-132
        self.subtiles_to_use = use.ravel()

    def get_batches(self) -> np.ndarray:
        return make_blocks(self.indata, self.subtile_shape[0])[self.subtiles_to_use]

    def put_batches(self, batches: np.ndarray) -> np.ndarray:
        d = self.subtile_shape[0]
        full = np.zeros((self.subtiles_to_use.size, d, d), dtype=np.float64)  # tiler.py:150-156
        full[self.subtiles_to_use] = batches
        out = unmake_blocks(full, d, *self.tile_shape).astype(np.uint8)  # uint8 on assignment :168
        self.outdata = out
        return out[: self.size[0], : self.size[1]]


# ----------------------------------------------------------------------------------------------
# Build extension: overlapping tiles + weighted blending.  Parity unpinned.
# ----------------------------------------------------------------------------------------------

def overlap_grid(H: int, W: int, T: int, overlap: int):
    """stride = T - overlap; grid = ceil((H - T) / stride) + 1; padded = (g - 1) * stride + T."""
    if not 0 <= overlap < T:
        raise ValueError("overlap must be in [0, T)")
    s = T - overlap
    gy = max(0, math.ceil((H - T) / s)) + 1
    gx = max(0, math.ceil((W - T) / s)) + 1
    return gy, gx, (gy - 1) * s + T, (gx - 1) * s + T


def extract_tiles(mosaic_hwc: np.ndarray, T: int, overlap: int) -> np.ndarray:
    """(H, W, C) uint8 -> (gy*gx, T, T, C); pixels outside the mosaic are 0 (raw uint8 zero pad)."""
    H, W, C = mosaic_hwc.shape
    gy, gx, Hp, Wp = overlap_grid(H, W, T, overlap)
    pad = np.zeros((Hp, Wp, C), dtype=mosaic_hwc.dtype)
    pad[:H, :W] = mosaic_hwc
    s = T - overlap
    out = np.empty((gy * gx, T, T, C), dtype=mosaic_hwc.dtype)
    for ty in range(gy):
        for tx in range(gx):
            out[ty * gx + tx] = pad[ty * s: ty * s + T, tx * s: tx * s + T]
    return out


def blend_window(T: int, overlap: int) -> np.ndarray:
    """1-D blending weight: linear ramp over the overlap, 1 inside; overlap 0 -> all ones."""
    i = np.arange(T, dtype=np.float32)
    w = np.minimum(np.minimum(i + 1, T - i), np.float32(overlap + 1)) / np.float32(overlap + 1)
    return w.astype(np.float32)


def stitch_blend(logits: np.ndarray, H: int, W: int, T: int, overlap: int):
    """(gy*gx, T, T, K) fp32 -> (blended (H, W, K) fp32, mask (H, W) uint8).

    blended = sum_tiles w * logit / sum_tiles w, tiles visited in (ty, tx) ascending order with
    fp32 multiply-then-add (no fused multiply-add), w = wy[y] * wx[x]; mask = first-max argmax.
    """
    gy, gx, Hp, Wp = overlap_grid(H, W, T, overlap)
    K = logits.shape[-1]
    s = T - overlap
    w1 = blend_window(T, overlap)
    w2 = (w1[:, None] * w1[None, :]).astype(np.float32)
    acc = np.zeros((Hp, Wp, K), dtype=np.float32)
    wsum = np.zeros((Hp, Wp), dtype=np.float32)
    for ty in range(gy):
        for tx in range(gx):
            sl = (slice(ty * s, ty * s + T), slice(tx * s, tx * s + T))
            prod = (logits[ty * gx + tx].astype(np.float32) * w2[..., None]).astype(np.float32)
            acc[sl] = acc[sl] + prod
            wsum[sl] = wsum[sl] + w2
    blended = (acc / wsum[..., None]).astype(np.float32)[:H, :W]
    return blended, blended.argmax(axis=-1).astype(np.uint8)
