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
               lr: float = 3e-4, clip: float = 0.5, optimizer=None, distmap: torch.Tensor = None,
               alpha: float = 0.01) -> Dict[str, object]:
    """runs forward + loss + backward (+ clip + Adam when `optimizer` is given or lr > 0) on `model` in place."""
    model.train()
    for p in model.parameters():
        p.grad = None
    logits = model(img)
    K = logits.shape[1]
    onehot = ref_losses.class2one_hot(mask, K)
    probs = logits.softmax(dim=1)
    terms = ref_losses.calculate_loss(probs, onehot, list(losses), distmap=distmap, alpha=alpha)
    loss = terms["total_loss"]
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    out = {"logits": logits.detach(), "loss": float(loss.detach()), "terms": {k: float(v.detach()) for k, v in terms.items()},
           "grads": grads}
    if optimizer is None and lr > 0:
        optimizer = torch.optim.Adam(model.parameters(), lr=lr)
    if optimizer is not None:
        out["grad_norm"] = float(torch.nn.utils.clip_grad_norm_(model.parameters(), clip)) if clip > 0 else None
        optimizer.step()
    return out


# ----------------------------------------------------------------------------------------------------------------
# bf16-arithmetic restatement of the training step: the storage points of the CUDA engine in its "bf16" mode
# (deadtrees_b200/train_engine.py) reproduced with autograd on the CPU.  Activations are rounded to bf16 where the
# engine stores them (raw conv output, post-BN activation, network input); gradients are rounded to bf16 where the
# engine stores activation gradients (BatchNorm-backward output, data-gradient outputs, the logits gradient entering
# the head).  Weights enter the convolutions rounded to bf16 (straight-through to the fp32 master copy); the head,
# BatchNorm parameters and every parameter gradient stay fp32.  Like ``ref_unet.forward_bf16`` this is what "bf16
# operands, fp32 accumulate" computes - the CUDA path must match THIS closely; its distance to the fp32/fp64 step
# is intrinsic to the arithmetic the north star prescribes.
# ----------------------------------------------------------------------------------------------------------------
import torch.nn.functional as F  # noqa: E402


def _rb(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(t.dtype)


class _RoundBoth(torch.autograd.Function):
    """bf16 storage of an activation and of its gradient."""

    @staticmethod
    def forward(ctx, x):
        return _rb(x)

    @staticmethod
    def backward(ctx, g):
        return _rb(g)


class _RoundGrad(torch.autograd.Function):
    """identity forward; the gradient passing through is stored in bf16."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _rb(g)


def _w_bf16(w: torch.Tensor) -> torch.Tensor:
    return w + (_rb(w) - w).detach()


def _conv_bn_bf16(conv, bn, x, relu=True, residual=None):
    y = _RoundBoth.apply(F.conv2d(x, _w_bf16(conv.weight), None, conv.stride, conv.padding))
    z = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, True, bn.momentum, bn.eps)
    if residual is not None:
        z = z + _RoundGrad.apply(residual)
    return _RoundBoth.apply(F.relu(z) if relu else z)


def forward_train_bf16(model, x: torch.Tensor) -> torch.Tensor:
    """train-mode forward of ``ref_unet.Unet`` with the engine's bf16 storage points (autograd-enabled)."""
    enc = model.encoder
    x = _rb(x)
    f1 = _conv_bn_bf16(enc.conv1, enc.bn1, x)
    cur = enc.maxpool(f1)
    feats = [f1]
    for name in ("layer1", "layer2", "layer3", "layer4"):
        for b in getattr(enc, name):
            t = _conv_bn_bf16(b.conv1, b.bn1, cur)
            idt = cur if b.downsample is None else _conv_bn_bf16(b.downsample[0], b.downsample[1], cur, relu=False)
            cur = _conv_bn_bf16(b.conv2, b.bn2, t, residual=idt)
        feats.append(cur)
    skips = feats[::-1]
    y = skips[0]
    for i, blk in enumerate(model.decoder.blocks):
        y = F.interpolate(y, scale_factor=2, mode="nearest")
        if i + 1 < len(skips):
            y = torch.cat([y, skips[i + 1]], dim=1)
        y = _conv_bn_bf16(blk.conv1[0], blk.conv1[1], y)
        y = _conv_bn_bf16(blk.conv2[0], blk.conv2[1], y)
    head = model.segmentation_head[0]
    return _RoundGrad.apply(F.conv2d(y, head.weight, head.bias, 1, 1))


def train_step_bf16(model: torch.nn.Module, img: torch.Tensor, mask: torch.Tensor,
                    losses: Sequence[str] = ("DICE", "FOCAL")) -> Dict[str, object]:
    """forward + loss + backward with the bf16 storage points; gradients only (no optimizer step)."""
    model.train()
    for p in model.parameters():
        p.grad = None
    logits = forward_train_bf16(model, img)
    onehot = ref_losses.class2one_hot(mask, logits.shape[1])
    terms = ref_losses.calculate_loss(logits.softmax(dim=1), onehot, list(losses))
    terms["total_loss"].backward()
    return {"logits": logits.detach(), "loss": float(terms["total_loss"].detach()),
            "grads": {n: p.grad.detach().clone() for n, p in model.named_parameters()}}
