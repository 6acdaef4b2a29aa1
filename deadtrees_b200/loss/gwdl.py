"""``GeneralizedWassersteinDiceLoss`` with the reference's interface (``deadtrees/loss/gwdl.py:18-138``), computed by
``dt_gwdl_loss`` / ``dt_gwdl_loss_backward``.

``loss(input, target)``: ``input`` (N, C, H, W) score maps - the module applies the softmax itself (gwdl.py:104) -,
``target`` (N, H, W) or (N, 1, H, W) labels.  ``SemSegment.calculate_loss`` hands it the softmax PROBABILITIES
(``segmodel.py:176-178``), so on that path the probabilities are soft-maxed a second time; the fused training path
reproduces exactly that (``softmax_twice``).  Implemented: ``weighting_mode="default"``, ``reduction="mean"`` (what the
reference constructs, ``segmodel.py:118-124``).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops

SUPPORTED_WEIGHTING = ["default", "GDL"]


class _GwdlFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores: torch.Tensor, target: torch.Tensor, M, twice: bool):
        scores = scores.detach().float().contiguous()
        loss, coef = ops.gwdl_loss(scores, target, M, softmax_twice=twice)
        ctx.saved = (scores, target, M, coef, twice)
        return loss

    @staticmethod
    def backward(ctx, g):
        scores, target, M, coef, twice = ctx.saved
        grad = torch.zeros_like(scores)
        ops.gwdl_loss_backward(scores, target, M, coef, 1.0, grad, softmax_twice=twice)
        return grad.mul_(g), None, None, None


class GeneralizedWassersteinDiceLoss:
    """a plain callable like the other loss terms of this package (no parameters, no buffers)."""

    def __init__(self, dist_matrix, weighting_mode: str = "default", reduction: str = "mean"):
        assert weighting_mode in SUPPORTED_WEIGHTING, "weighting_mode must be in %s" % str(SUPPORTED_WEIGHTING)
        if weighting_mode != "default" or reduction != "mean":
            raise NotImplementedError("deadtrees_b200 implements the GWDL the reference builds: weighting 'default', reduction 'mean'")
        M = np.asarray(dist_matrix.cpu() if isinstance(dist_matrix, torch.Tensor) else dist_matrix, dtype=np.float64)
        if M.max() != 1:
            print("Normalize the maximum of the distance matrix used in the Generalized Wasserstein Dice Loss to 1.")
            M = M / M.max()
        self.M = torch.from_numpy(M)
        self.num_classes = M.shape[0]
        self.alpha_mode, self.reduction = weighting_mode, reduction

    def matrix(self):
        return self.M.tolist()

    def __call__(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        target = target.long()
        if target.dim() == input.dim():          # (N, 1, H, W)
            target = target[:, 0]
        return _GwdlFn.apply(input, target.contiguous(), self.matrix(), False)

    def on_logits(self, logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """the value ``SemSegment.calculate_loss`` gets: this module applied to ``softmax(logits)``."""
        return _GwdlFn.apply(logits, labels.long().contiguous(), self.matrix(), True)
