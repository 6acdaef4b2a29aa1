"""Fused logits-level loss/metric pass used by ``SemSegment`` (one read of the logits).

``SemSegment.training_step`` in the reference does ``class2one_hot`` + ``softmax`` + Dice/GDL + Focal +
two Fscore passes (``deadtrees/network/segmodel.py:214-225,169-208``).  ``dt_seg_loss_partials`` does all
the per-pixel work in one kernel, ``dt_seg_loss_finalize`` turns the sums into the scalars, and
``dt_seg_loss_backward`` produces d(loss)/d(logits).
"""
from __future__ import annotations

import torch

from .. import ops


def softmax_nchw(logits: torch.Tensor) -> torch.Tensor:
    return ops.softmax_nchw(logits)


class SegLossTerms:
    def __init__(self, logits: torch.Tensor, labels: torch.Tensor, dice_mode: int, use_focal: bool):
        if labels.dtype != torch.int64:
            labels = labels.long()
        self.logits, self.labels = logits.contiguous(), labels.contiguous()
        self.sums, self.counts, self.bad = ops.seg_loss_partials(self.logits, self.labels)
        self.out, self.coef, self.focal_scale = ops.seg_loss_finalize(self.sums, self.counts, dice_mode, use_focal)

    dice_loss = property(lambda self: self.out[0])
    focal_loss = property(lambda self: self.out[1])
    total_loss = property(lambda self: self.out[2])
    fscore = property(lambda self: self.out[3])
    fscore_with_bg = property(lambda self: self.out[4])

    def check_labels(self) -> None:
        assert int(self.bad.item()) == 0, "labels outside [0, K) (class2one_hot assert, losses.py:129)"

    def grad_logits(self, upstream: float = 1.0) -> torch.Tensor:
        return ops.seg_loss_backward(self.logits, self.labels, self.coef, self.focal_scale, upstream)


class _SegLossFn(torch.autograd.Function):
    """total loss of ``SemSegment.calculate_loss`` on the logits; backward = ``dt_seg_loss_backward``
    (+ ``dt_boundary_loss_backward`` when a boundary term is present)."""

    @staticmethod
    def forward(ctx, logits: torch.Tensor, terms: "SegLossTerms"):
        ctx.terms = terms
        total = terms.total_loss.clone()
        if terms.boundary is not None:
            total = total + terms.boundary_weight * terms.boundary
        if terms.gwdl is not None:
            total = total + terms.gwdl
        return total

    @staticmethod
    def backward(ctx, g):
        t = ctx.terms
        grad = t.grad_logits(1.0)
        if t.boundary is not None:
            ops.boundary_loss_backward(t.logits, t.distmap, t.boundary_idc, t.boundary_weight, grad)
        if t.gwdl is not None:
            ops.gwdl_loss_backward(t.logits, t.labels, t.gwdl_matrix, t.gwdl_coef, 1.0, grad, softmax_twice=True)
        return grad.mul_(g), None


def seg_loss(logits: torch.Tensor, labels: torch.Tensor, dice_mode: int, use_focal: bool, distmap: torch.Tensor = None,
             boundary_idc=None, boundary_weight: float = 1.0, gwdl_matrix=None):
    """-> (differentiable total loss, SegLossTerms with the individual scalars and metrics).  With ``distmap`` the
    boundary loss over the classes ``boundary_idc`` is added with ``boundary_weight`` (1, or alpha when ramped)."""
    terms = SegLossTerms(logits.detach(), labels, dice_mode, use_focal)
    terms.boundary = None
    if distmap is not None and boundary_idc:
        terms.distmap = distmap.float().contiguous()
        terms.boundary_idc, terms.boundary_weight = list(boundary_idc), float(boundary_weight)
        terms.boundary = ops.boundary_loss(terms.logits, terms.distmap, terms.boundary_idc)
    terms.gwdl = None
    if gwdl_matrix is not None:      # GWDICE replaces the Dice term (dice_mode 0): Wasserstein Dice on softmax(softmax(logits))
        terms.gwdl_matrix = gwdl_matrix
        terms.gwdl, terms.gwdl_coef = ops.gwdl_loss(terms.logits, terms.labels, gwdl_matrix, softmax_twice=True)
    return _SegLossFn.apply(logits, terms), terms
