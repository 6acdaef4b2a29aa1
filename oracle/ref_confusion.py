"""Oracle: confusion matrices of the validation / test epoch (``deadtrees/network/segmodel.py:291-407``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference calls ``torchmetrics.functional.confusion_matrix(prediction, target, num_classes=K[, normalize="true"])``
(``segmodel.py:302-312, 346-366``) once on all pixels and once on the forest pixels ``lu == 1`` (``:297-300, :341-344``).
``torchmetrics`` is a dependency that is neither vendored nor pinned (``setup.py``) and is absent here; its published
algorithm is ``bincount(target * K + preds, minlength=K*K).reshape(K, K)`` (rows = target) and, for ``normalize="true"``,
a division of every row by its sum with the NaN rows of absent classes set to 0.

PARITY: unpinned against torchmetrics itself; cross-checked in ``tests/test_oracle.py`` against
``sklearn.metrics.confusion_matrix`` (same definition, independent implementation).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np


def confusion_matrix(pred: np.ndarray, target: np.ndarray, K: int, normalize: Optional[str] = None) -> np.ndarray:
    pred, target = np.asarray(pred).ravel().astype(np.int64), np.asarray(target).ravel().astype(np.int64)
    cm = np.bincount(target * K + pred, minlength=K * K).reshape(K, K)
    if normalize == "true":
        with np.errstate(invalid="ignore", divide="ignore"):
            cm = cm / cm.sum(axis=1, keepdims=True)
        cm = np.nan_to_num(cm, nan=0.0)
    return cm


def epoch_matrices(pred: np.ndarray, target: np.ndarray, lu: np.ndarray, K: int) -> Dict[str, np.ndarray]:
    """the four matrices of ``test_epoch_end`` (``segmodel.py:337-377``); ``validation_epoch_end`` uses the two normalised."""
    pred, target, lu = np.asarray(pred).ravel(), np.asarray(target).ravel(), np.asarray(lu).ravel()
    m = lu == 1
    return {"cm_norm": confusion_matrix(pred, target, K, "true"), "cm_px": confusion_matrix(pred, target, K),
            "cm_norm_masked": confusion_matrix(pred[m], target[m], K, "true"),
            "cm_px_masked": confusion_matrix(pred[m], target[m], K)}
