"""TEST INFRASTRUCTURE ONLY - CPU restatement of the reference's signed distance maps for the boundary loss.

``one_hot2dist`` (``deadtrees/loss/losses.py:159-178``; called per sample by the dataloader,
``deadtrees/data/deadtreedata.py:182-185``, with ``resolution=[1, 1]``): per class k with at least one pixel,
``edt(~pos) * ~pos - (edt(pos) - 1) * pos`` with scipy's exact Euclidean distance transform; classes without a pixel stay 0.
The result array takes the dtype of ``seg`` unless ``dtype`` is given - the dataloader passes the int32 one-hot of
``class2one_hot``, so there the float64 expression is TRUNCATED towards zero on assignment.  (The reference spells the
boolean cast ``np.bool``, removed in numpy 1.24; the golden vectors in ``tests/golden/dist.npz`` were produced by the
reference function itself with that alias restored, ``oracle/make_golden.py``.)
"""
from typing import Optional, Sequence

import numpy as np
from scipy.ndimage import distance_transform_edt


def one_hot2dist(seg: np.ndarray, resolution: Optional[Sequence[float]] = None, dtype=None) -> np.ndarray:
    K = len(seg)
    res = np.zeros_like(seg, dtype=dtype)
    for k in range(K):
        posmask = seg[k].astype(bool)
        if posmask.any():
            negmask = ~posmask
            res[k] = (distance_transform_edt(negmask, sampling=resolution) * negmask
                      - (distance_transform_edt(posmask, sampling=resolution) - 1) * posmask)
    return res


def labels_to_dist(labels: np.ndarray, K: int, truncate: bool = True) -> np.ndarray:
    """(N, H, W) integer labels -> (N, K, H, W) float32 distance maps as the dataloader builds them (truncate=True:
    through the int32 one-hot, deadtreedata.py:182-185) or exact (truncate=False: ``dtype=np.float32``)."""
    out = np.zeros((labels.shape[0], K) + labels.shape[1:], dtype=np.float32)
    for n in range(labels.shape[0]):
        onehot = (labels[n][None] == np.arange(K)[:, None, None]).astype(np.int32)
        out[n] = one_hot2dist(onehot, resolution=[1, 1], dtype=None if truncate else np.float32).astype(np.float32)
    return out
