"""Dataset constants and ``val_transform`` with the albumentations calling convention.

Mirrors ``deadtrees/data/deadtreedata.py:27-34`` (``DeadtreeDatasetConfig``) and ``:148-154``
(``val_transform = A.Compose([A.Normalize(mean, std), ToTensorV2()])``): call
``val_transform(image=<HWC uint8 ndarray>)["image"]`` -> CHW float32 tensor.  The arithmetic
(``(u8 - 255*mean) * (1 / (255*std))`` in fp32) runs in ``dt_tile_gather_normalize``; the result is a
CUDA tensor (the reference moves it to the GPU right after, ``scripts/inference.py:100``).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .._lib import require_device


class DeadtreeDatasetConfig:
    """Dataset configuration (stats of the reference's train shards)."""

    mean = np.array([0.3661029729, 0.3875165941, 0.3501133538, 0.5797285859])
    std = np.array([0.2388708549, 0.2103625723, 0.2050272174, 0.2025812523])
    tile_size = 256
    fractions = [0.7, 0.2, 0.1]


def normalize_constants(channels: int = 4, mean=None, std=None, max_pixel_value: float = 255.0):
    """fp32 (offset, scale) with ``normalised = (u8 - offset) * scale``."""
    mean = DeadtreeDatasetConfig.mean if mean is None else np.asarray(mean)
    std = DeadtreeDatasetConfig.std if std is None else np.asarray(std)
    offset = mean[:channels].astype(np.float32) * np.float32(max_pixel_value)
    denom = std[:channels].astype(np.float32) * np.float32(max_pixel_value)
    return offset.astype(np.float32), np.reciprocal(denom, dtype=np.float32)


class _ValTransform:
    def __call__(self, *, image, **extra):
        require_device()
        if isinstance(image, np.ndarray):
            image = torch.from_numpy(np.ascontiguousarray(image))
        if image.dtype != torch.uint8 or image.dim() != 3:
            raise TypeError("val_transform expects an (H, W, C) uint8 image")
        H, W, Cc = image.shape
        if H != W or H % 4 or Cc > 4:
            raise ValueError("val_transform handles square tiles with side % 4 == 0 and <= 4 channels")
        offset, scale = normalize_constants(Cc)
        tiles = ops.tile_gather_normalize(image.cuda(), "hwc", Cc, H, 0, (1, 1), 0, 1, offset, scale,
                                          dtype=torch.float32)
        out = dict(extra)
        out["image"] = tiles[0, :, :, :Cc].permute(2, 0, 1).contiguous()
        return out


val_transform = _ValTransform()
