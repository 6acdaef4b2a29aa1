"""Oracle: restatement of ``val_transform`` (``deadtrees/data/deadtreedata.py:148-154``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``val_transform = A.Compose([A.Normalize(mean, std), ToTensorV2()])`` with the dataset constants at
``deadtreedata.py:31-32``.  albumentations (unpinned, ``setup.py:35``; absent here) publishes
``Normalize`` as, in float32::

    mean = np.array(mean, dtype=np.float32) * max_pixel_value          # max_pixel_value = 255
    denominator = np.reciprocal(np.array(std, dtype=np.float32) * max_pixel_value)
    img = (img.astype(np.float32) - mean) * denominator

and ``ToTensorV2`` as HWC -> CHW.  PARITY: unpinned (library absent, no reference test).
"""
import numpy as np

MEAN = np.array([0.3661029729, 0.3875165941, 0.3501133538, 0.5797285859])  # deadtreedata.py:31
STD = np.array([0.2388708549, 0.2103625723, 0.2050272174, 0.2025812523])   # deadtreedata.py:32
MAX_PIXEL = 255.0


def normalize_constants(channels: int = 4):
    """fp32 (offset, scale) so that ``out = (u8 - offset) * scale``."""
    mean = MEAN[:channels].astype(np.float32) * np.float32(MAX_PIXEL)
    std = STD[:channels].astype(np.float32) * np.float32(MAX_PIXEL)
    return mean.astype(np.float32), np.reciprocal(std, dtype=np.float32)


def val_transform(image: np.ndarray) -> np.ndarray:
    """(H, W, C) uint8 -> (C, H, W) float32."""
    off, scale = normalize_constants(image.shape[-1])
    out = (image.astype(np.float32) - off) * scale
    return np.ascontiguousarray(out.astype(np.float32).transpose(2, 0, 1))
