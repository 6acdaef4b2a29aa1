"""Oracle: restatement of ``train_transform`` + ``transform()`` (``deadtrees/data/deadtreedata.py:132-146, 156-189``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``train_transform = A.Compose([A.OneOf([A.HorizontalFlip(), A.VerticalFlip()], p=0.5), A.RandomRotate90(p=0.5),
A.RandomBrightnessContrast(p=0.5, brightness_limit=0.2, contrast_limit=0.15, brightness_by_max=False),
A.Normalize(mean, std), ToTensorV2()])``.  albumentations is a dependency that is neither vendored nor pinned
(``setup.py:35``) and is absent here; its published functions, restated with numpy for GIVEN random draws:

* ``hflip``: ``img[:, ::-1]``; ``vflip``: ``img[::-1]``; ``rot90``: ``np.rot90(img, factor)`` (image, mask and ``lu`` alike)
* ``brightness_contrast_adjust`` on uint8 (image only)::

      lut = np.arange(0, 256).astype("float32")
      if alpha != 1: lut *= alpha
      if beta != 0:  lut += beta * np.mean(img)            # brightness_by_max=False
      lut = np.clip(lut, 0, 255).astype("uint8");  img = cv2.LUT(img, lut)

  with ``alpha = 1 + U(-contrast_limit, contrast_limit)``, ``beta = U(-brightness_limit, brightness_limit)``.  (From 1.3 on
  the library adds ``alpha * beta * mean`` instead; callers select that by passing ``beta * alpha`` as beta.)
* ``Normalize`` / ``ToTensorV2``: ``oracle/ref_normalize.py``.

The additive term is rounded to float32 before it is added to the float32 table (numpy < 2 scalar casting, the reference's era).
PARITY: unpinned (library absent, no reference test for the augmentation).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from . import ref_normalize


def geometric(a: np.ndarray, flip: int, rot: int) -> np.ndarray:
    if flip == 1:
        a = a[:, ::-1]
    elif flip == 2:
        a = a[::-1]
    return np.ascontiguousarray(np.rot90(a, rot))


def brightness_contrast(img: np.ndarray, alpha: float, beta: float) -> np.ndarray:
    lut = np.arange(0, 256).astype("float32")
    if alpha != 1:
        lut *= np.float32(alpha)
    if beta != 0:
        lut += np.float32(beta * np.mean(img))
    lut = np.clip(lut, 0, 255).astype("uint8")
    return lut[img]


def train_transform(image: np.ndarray, mask: Optional[np.ndarray], lu: Optional[np.ndarray], flip: int, rot: int,
                    alpha: float, beta: float, in_channels: int = 4, classes: int = 3
                    ) -> Tuple[np.ndarray, Optional[np.ndarray], Optional[np.ndarray]]:
    """(H, W, C) uint8 image (+ (H, W) mask / lu) and the random draws -> ((in_channels, H, W) float32, int64 mask, int64 lu)."""
    img = brightness_contrast(geometric(image, flip, rot), alpha, beta)
    out = ref_normalize.val_transform(img)[:in_channels]                      # deadtreedata.py:176
    m = None if mask is None else geometric(mask, flip, rot).astype(np.int64)
    if m is not None and classes == 2:
        m[m > 1] = 1                                                          # deadtreedata.py:179-180
    l = None if lu is None else geometric(lu, flip, rot).astype(np.int64)
    return np.ascontiguousarray(out), m, l
