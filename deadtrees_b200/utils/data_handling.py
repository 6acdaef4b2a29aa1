"""Block split / merge with the reference's numpy-in / numpy-out signatures, executed on the GPU.

Mirrors ``deadtrees/utils/data_handling.py:9-34`` (``make_blocks_vectorized`` /
``unmake_blocks_vectorized``): same argument meaning, same result bit for bit; the work is done by
``dt_make_blocks`` / ``dt_unmake_blocks`` (pure index permutations).  Torch CUDA tensors are accepted
too and are returned as tensors without a host round trip.
"""
from __future__ import annotations

from typing import Union

import numpy as np
import torch

from .. import ops
from .._lib import require_device

ArrayLike = Union[np.ndarray, torch.Tensor]


def _to_device(x: ArrayLike) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x if x.is_cuda else x.cuda()
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def make_blocks_vectorized(x: ArrayLike, d: int) -> ArrayLike:
    """Dissect an array (channel, tile, tile) into subtiles (blocks, channel, d, d)."""
    require_device()
    if x.ndim != 3:
        raise ValueError("expected a 3-d array (channel, height, width)")
    out = ops.make_blocks(_to_device(x), int(d))
    return out if isinstance(x, torch.Tensor) else out.cpu().numpy()


def unmake_blocks_vectorized(x: ArrayLike, d: int, m: int, n: int) -> ArrayLike:
    """Merge subtiles (blocks, d, d) back into a 2-d array (m, n)."""
    require_device()
    if not isinstance(x, torch.Tensor):
        x = np.concatenate(list(x)) if not isinstance(x, np.ndarray) else x
    out = ops.unmake_blocks(_to_device(x).reshape(-1, d, d), int(d), int(m), int(n))
    return out if isinstance(x, torch.Tensor) else out.cpu().numpy()
