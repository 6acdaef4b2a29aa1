"""Oracle: restatement of ``smp.utils.metrics.Fscore`` as the reference uses it.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Call sites: ``deadtrees/network/segmodel.py:145-149`` (``Fscore(ignore_channels=[0])`` and
``Fscore()``) and ``:202-208`` (applied to softmax probabilities and the int32 one-hot target).
``segmentation_models_pytorch`` (>=0.2.1, ``setup.py:47``) is absent; its published
``functional.f_score`` is: threshold the prediction at 0.5, drop ``ignore_channels``, then over the
WHOLE tensor ``tp = sum(gt * pr); fp = sum(pr) - tp; fn = sum(gt) - tp;
score = ((1 + beta^2) tp + eps) / ((1 + beta^2) tp + beta^2 fn + fp + eps)`` with beta=1, eps=1e-7.

PARITY: partially pinned by ``tests/test_dice_metric.py:16,38-52`` — the "without background"
answers (1.0, 0.6154, 0.2) equal ``Fscore(ignore_channels=[0])`` on the same inputs.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch


def fscore_counts(pr: torch.Tensor, gt: torch.Tensor, ignore_channels: Optional[Sequence[int]] = None,
                  threshold: float = 0.5) -> Tuple[int, int, int]:
    """Integer (tp, sum_pr, sum_gt) over the whole batch tensor."""
    pr = (pr > threshold)
    if ignore_channels:
        keep = [c for c in range(pr.shape[1]) if c not in ignore_channels]
        pr, gt = pr[:, keep], gt[:, keep]
    gt = gt.to(torch.int64)
    tp = int((gt * pr).sum())
    return tp, int(pr.sum()), int(gt.sum())


def fscore(pr: torch.Tensor, gt: torch.Tensor, ignore_channels: Optional[Sequence[int]] = None,
           eps: float = 1e-7, threshold: float = 0.5) -> torch.Tensor:
    tp, spr, sgt = fscore_counts(pr, gt, ignore_channels, threshold)
    tp_f = torch.tensor(float(tp), dtype=torch.float32)
    fp = torch.tensor(float(spr - tp), dtype=torch.float32)
    fn = torch.tensor(float(sgt - tp), dtype=torch.float32)
    return (2 * tp_f + eps) / (2 * tp_f + fn + fp + eps)
