"""Loss callables with the reference's signatures, reductions on the GPU.

Mirrors ``deadtrees/loss/losses.py``: ``class2one_hot`` (:124-141), ``DiceLoss`` (:226-247),
``SurfaceLoss`` / ``BoundaryLoss`` (:250-270), ``FocalLoss`` (:273-291), ``EPS`` (:19).  Each callable
takes ``(probs, target)`` of shape (B, K, H, W) and returns a 0-dim fp32 tensor.  The pass over the
pixels is ``dt_prob_loss_partials`` (per (b, k) sums in fp64); only the O(B*K) epilogue is tensor code.
"""
from __future__ import annotations

import logging
from typing import List

import torch
from torch import Tensor

from .. import ops
from .._lib import require_device

logger = logging.getLogger(__name__)

EPS = 1e-10


def class2one_hot(seg: Tensor, K: int) -> Tensor:
    """(B, H, W) integer labels -> (B, K, H, W) int32 one-hot; labels must lie in [0, K)."""
    require_device()
    if seg.dim() != 3:
        raise ValueError("class2one_hot expects (B, H, W)")
    res, bad = ops.class2one_hot(seg, K)
    assert int(bad.item()) == 0, (sorted(set(torch.unique(seg).tolist())), K)  # host sync, as the reference's sset()
    return res


def one_hot2dist(seg, resolution=None, dtype=None):
    """signed distance maps of the boundary loss, drop-in for ``one_hot2dist`` (``deadtrees/loss/losses.py:159-178``):
    ``seg`` one-hot ``(K, H, W)`` ndarray -> ndarray of ``seg``'s dtype (or ``dtype``).  Computed on the device
    (``dt_one_hot2dist``, exact Euclidean distance transform); as in the reference the float64 expression is truncated
    towards zero when the result dtype is an integer type (the dataloader passes the int32 one-hot,
    ``deadtrees/data/deadtreedata.py:182-185``).  Only unit pixel spacing is implemented."""
    import numpy as np
    require_device()
    seg = np.asarray(seg)
    if resolution is not None and any(float(r) != 1.0 for r in resolution):
        raise NotImplementedError("one_hot2dist: only resolution None / [1, 1] (the dataloader's) is implemented")
    if seg.ndim != 3:
        raise ValueError("one_hot2dist expects a one-hot (K, H, W) array")
    onehot = seg.astype(bool)
    assert (onehot.sum(axis=0) == 1).all() and ((seg == 0) | (seg == 1)).all(), "one_hot2dist: seg is not one-hot"
    K = seg.shape[0]
    res_dtype = np.dtype(seg.dtype if dtype is None else dtype)
    labels = torch.from_numpy(onehot.argmax(axis=0).astype(np.int64))[None].cuda()
    dist = ops.one_hot2dist(labels, K, truncate=np.issubdtype(res_dtype, np.integer))[0]
    return dist.cpu().numpy().astype(res_dtype)


class DiceLoss:
    def __init__(self, **kwargs):
        self.idc: List[int] = kwargs["idc"]
        logger.debug(f"Initialized {self.__class__.__name__} with {kwargs}")

    def __call__(self, probs: Tensor, target: Tensor) -> Tensor:
        s = ops.prob_loss_partials(probs, target)[:, self.idc]
        inter = s[..., 0].float()
        union = s[..., 1].float() + s[..., 2].float()
        return (torch.ones_like(inter) - (2 * inter + EPS) / (union + EPS)).mean()


class SurfaceLoss:
    def __init__(self, **kwargs):
        self.idc: List[int] = kwargs["idc"]
        logger.debug(f"Initialized {self.__class__.__name__} with {kwargs}")

    def __call__(self, probs: Tensor, dist_maps: Tensor) -> Tensor:
        s = ops.prob_loss_partials(probs, dist_maps.float())[:, self.idc]
        n = probs.shape[0] * len(self.idc) * probs.shape[2] * probs.shape[3]
        return (s[..., 0].sum() / n).float()


BoundaryLoss = SurfaceLoss


class FocalLoss:
    def __init__(self, **kwargs):
        self.idc: List[int] = kwargs["idc"]
        self.gamma: float = kwargs["gamma"]
        logger.debug(f"Initialized {self.__class__.__name__} with {kwargs}")

    def __call__(self, probs: Tensor, target: Tensor) -> Tensor:
        s = ops.prob_loss_partials(probs, target, gamma=self.gamma)[:, self.idc]
        loss = -s[..., 3].sum().float()
        return loss / (s[..., 2].sum().float() + EPS)
