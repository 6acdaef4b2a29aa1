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


# ---- train_transform (deadtreedata.py:132-146) and transform() (:156-189) on the device ----------------------------

BRIGHTNESS_LIMIT, CONTRAST_LIMIT = 0.2, 0.15          # A.RandomBrightnessContrast(...), deadtreedata.py:137-142


def draw_train_params(rng: np.random.Generator, n: int, beta_times_alpha: bool = False):
    """the random draws of ``train_transform`` for ``n`` samples -> (geom int32 (n, 2) {flip, rot}, bc float64 (n, 2) {alpha, beta}).

    Distributions as albumentations composes them: ``OneOf([HorizontalFlip, VerticalFlip], p=0.5)`` = no flip with 1/2,
    either flip with 1/4; ``RandomRotate90(p=0.5)`` = with 1/2 a factor drawn from {0, 1, 2, 3}; ``RandomBrightnessContrast(p=0.5)``
    = with 1/2 ``alpha = 1 + U(-0.15, 0.15)``, ``beta = U(-0.2, 0.2)``.  The stream of a numpy Generator replaces the library's
    global ``random`` state: the distributions are the reference's, the individual draws are not reproducible against it.
    ``beta_times_alpha``: albumentations >= 1.3 adds ``alpha * beta * mean`` instead of ``beta * mean``."""
    geom = np.zeros((n, 2), dtype=np.int32)
    bc = np.tile(np.array([1.0, 0.0]), (n, 1))
    for i in range(n):
        if rng.random() < 0.5:
            geom[i, 0] = 1 if rng.random() < 0.5 else 2
        if rng.random() < 0.5:
            geom[i, 1] = int(rng.integers(0, 4))
        if rng.random() < 0.5:
            alpha = 1.0 + rng.uniform(-CONTRAST_LIMIT, CONTRAST_LIMIT)
            beta = rng.uniform(-BRIGHTNESS_LIMIT, BRIGHTNESS_LIMIT)
            bc[i] = (alpha, beta * alpha if beta_times_alpha else beta)
    return geom, bc


class BatchTrainTransform:
    """``train_transform`` + ``transform()`` for a whole batch in one pass over the bytes (``dt_train_transform``), optionally
    followed by the boundary-loss distance maps (``dt_one_hot2dist``): uint8 tiles in, the tensors of
    ``SemSegment.training_step``'s batch out - ``(img fp32 (N, in_channels, H, W), mask int64, distmap fp32 | None, lu int64)``."""

    def __init__(self, in_channels: int = 4, classes: int = 3, distmap: bool = False, seed: int = 0,
                 beta_times_alpha: bool = False):
        self.in_channels, self.classes, self.distmap = in_channels, classes, distmap
        self.rng = np.random.default_rng(seed)
        self.beta_times_alpha = beta_times_alpha

    def __call__(self, images, masks=None, lus=None, params=None):
        require_device()
        as_t = lambda a: None if a is None else (torch.from_numpy(np.ascontiguousarray(a)) if isinstance(a, np.ndarray) else a)
        images, masks, lus = as_t(images), as_t(masks), as_t(lus)
        n, Cc = images.shape[0], images.shape[3]
        geom, bc = params if params is not None else draw_train_params(self.rng, n, self.beta_times_alpha)
        offset, scale = normalize_constants(Cc)
        cuda = lambda t: None if t is None else t.cuda(non_blocking=True)
        img, mask, lu = ops.train_transform(cuda(images), cuda(masks), cuda(lus), geom, bc, offset, scale,
                                            min(self.in_channels, Cc), merge_classes=self.classes == 2)
        dist = None
        if self.distmap and mask is not None:           # transform(): class2one_hot + one_hot2dist per sample (:182-185)
            dist = ops.one_hot2dist(mask, self.classes, truncate=True)
        return img, mask, dist, lu


class _TrainTransform:
    """``train_transform(image=, mask=, lu=)`` with the albumentations calling convention, one sample (HWC uint8 in, CHW out)."""

    def __init__(self, seed: int = 0):
        self.batch = BatchTrainTransform(in_channels=4, classes=3, seed=seed)

    def __call__(self, *, image, mask=None, lu=None, **extra):
        as_np = lambda a: None if a is None else (a[None] if isinstance(a, torch.Tensor) else np.asarray(a)[None])
        img, m, _, l = self.batch(as_np(image), as_np(mask), as_np(lu))
        out = dict(extra)
        out["image"] = img[0]
        if m is not None:
            out["mask"] = m[0]
        if l is not None:
            out["lu"] = l[0]
        return out


train_transform = _TrainTransform()


def transform(sample: dict, *, transform_func=None, in_channels: int = 4, classes: int = 3, distmap: bool = False):
    """``transform()`` of the dataloader (``deadtreedata.py:156-189``) on one decoded sample."""
    if transform_func:
        t = transform_func(image=sample["image"], mask=sample["mask"], lu=sample["lu"])
        sample["image"], sample["mask"], sample["lu"] = t["image"].float(), t["mask"].long(), t["lu"]
    sample["image"] = sample["image"][0:in_channels]
    sample["lu"] = torch.as_tensor(sample["lu"], dtype=torch.long)
    if classes == 2:
        sample["mask"][sample["mask"] > 1] = 1
    sample["distmap"] = ops.one_hot2dist(sample["mask"][None].cuda(), classes, truncate=True)[0] if distmap else None
    return sample
