"""Oracle: one training step of the reference on the CPU (fp32 autograd).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``SemSegment.training_step`` (``deadtrees/network/segmodel.py:210-229``): train-mode forward of the restated
``smp.Unet`` (``oracle/ref_unet.py``), ``class2one_hot`` + ``softmax`` + ``calculate_loss`` (``oracle/ref_losses.py``,
pinned to the reference's own loss code), ``loss.backward()``, then what the Lightning trainer does with
``configs/trainer/default.yaml:18`` and ``configure_optimizers`` (``segmodel.py:420-429``):
``clip_grad_norm_(0.5)`` and ``torch.optim.Adam(lr)``.

PARITY: the model arithmetic is unpinned by the reference (third-party smp, see ``ref_unet``); the loss terms are
pinned by ``tests/golden/losses.npz``; autograd, clipping and Adam are torch's own.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch

from . import ref_losses


def train_step(model: torch.nn.Module, img: torch.Tensor, mask: torch.Tensor, losses: Sequence[str] = ("DICE", "FOCAL"),
               lr: float = 3e-4, clip: float = 0.5, optimizer=None) -> Dict[str, object]:
    """runs forward + loss + backward (+ clip + Adam when `optimizer` is given or lr > 0) on `model` in place."""
    model.train()
    for p in model.parameters():
        p.grad = None
    logits = model(img)
    K = logits.shape[1]
    onehot = ref_losses.class2one_hot(mask, K)
    probs = logits.softmax(dim=1)
    terms = ref_losses.calculate_loss(probs, onehot, list(losses))
    loss = terms["total_loss"]
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    out = {"logits": logits.detach(), "loss": float(loss), "terms": {k: float(v) for k, v in terms.items()}, "grads": grads}
    if optimizer is None and lr > 0:
        optimizer = torch.optim.Adam(model.parameters(), lr=lr)
    if optimizer is not None:
        out["grad_norm"] = float(torch.nn.utils.clip_grad_norm_(model.parameters(), clip)) if clip > 0 else None
        optimizer.step()
    return out
