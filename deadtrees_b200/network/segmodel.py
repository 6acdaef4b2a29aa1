"""``SemSegment`` — drop-in for ``deadtrees.network.segmodel.SemSegment`` on the Unet/resnet34 path.

Mirrors ``deadtrees/network/segmodel.py:57-229,277-289,420-438``: same constructor arguments
(``network`` / ``training`` configs with attribute access), same loss composition, metrics, step
functions and checkpoint layout (Lightning ``.ckpt`` = ``{"state_dict": {"model.*"}, "hyper_parameters":
{"network", "training"}}``).  ``pytorch_lightning`` / ``omegaconf`` are optional: the class is a plain
``nn.Module`` that also subclasses ``LightningModule`` when that package is importable.
The forward pass, the loss/metric reductions and the optimizer run in the CUDA library.
"""
from __future__ import annotations

import logging
import math
from collections import Counter
from pathlib import Path
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
from torch import Tensor

from .. import ops
from ..loss import fused
from ..loss.gdl import GeneralizedDiceLoss
from ..loss.gwdl import GeneralizedWassersteinDiceLoss
from ..loss.losses import BoundaryLoss, DiceLoss, FocalLoss, class2one_hot
from .unet import Unet, UnetPlusPlus

log = logging.getLogger(__name__)

try:  # optional: keep Lightning compatibility when it is installed
    import pytorch_lightning as pl  # type: ignore

    _Base = pl.LightningModule
except Exception:  # pragma: no cover - absent in this image
    pl = None
    _Base = nn.Module


class Conf(dict):
    """dict with attribute access, ``copy()`` and ``del conf.key`` — the subset of DictConfig the reference uses."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]

    def copy(self):
        return Conf(self)


def to_conf(cfg) -> Conf:
    if cfg is None:
        return Conf()
    if isinstance(cfg, Conf):
        return cfg
    if isinstance(cfg, dict):
        return Conf({k: (to_conf(v) if isinstance(v, dict) else v) for k, v in cfg.items()})
    if hasattr(cfg, "items"):  # DictConfig
        return Conf({k: (list(v) if (hasattr(v, "__iter__") and not isinstance(v, (str, dict)) and not hasattr(v, "items")) else v)
                     for k, v in cfg.items()})
    return Conf(vars(cfg))


def concat_extra(img, mask, distmap, lu, stats, *, extra):
    extra_imgs, extra_masks, extra_distmaps, extra_lus, extra_stats = list(zip(*extra))
    img = torch.cat((img, *extra_imgs), dim=0)
    mask = torch.cat((mask, *extra_masks), dim=0)
    distmap = torch.cat((distmap, *extra_distmaps), dim=0) if distmap is not None else None
    lu = torch.cat((lu, *extra_lus), dim=0)
    stats.extend(sum(extra_stats, []))
    return img, mask, distmap, lu, stats


def create_combined_batch(batch: Dict[str, Any]):
    """``batch = {"main": (img, mask, distmap, lu, stats), "extra_*": ...}`` (segmodel.py:43-54)."""
    img, mask, distmap, lu, stats = batch["main"]
    extra = [v for k, v in batch.items() if k.startswith("extra")]
    if extra:
        img, mask, distmap, lu, stats = concat_extra(img, mask, distmap, lu, stats, extra=extra)
    return img, mask, distmap, lu, stats


class SemSegment(_Base):  # type: ignore[misc]
    def __init__(self, network, training=None):
        super().__init__()
        network, training = to_conf(network), to_conf(training)

        architecture = str(network.architecture).lower().strip()
        if architecture == "unet":
            Model = Unet
        elif architecture in ["unetplusplus", "unet++"]:
            Model = UnetPlusPlus       # nested decoder on the same fused kernels (inference)
        elif architecture in ["resunet", "resunetplusplus", "resunet++", "efficientunetplusplus", "efficientunet++"]:
            raise NotImplementedError(
                f"architecture <{architecture}> is outside the B200 hot path (SURVEY.md D1); only Unet is built")
        else:
            raise NotImplementedError(
                "Currently only Unet, ResUnet, Unet++, ResUnet++, and EfficientUnet++ architectures are supported")

        clean = network.copy()
        del clean.architecture
        losses = list(clean.pop("losses", []))
        classes = list(clean.pop("classes"))
        n_classes = len(classes)
        precision = clean.pop("precision", "bf16")

        self.model = Model(**clean, classes=n_classes, precision=precision)

        if clean.get("encoder_weights") is None:
            log.info("Initializing unset weights with Kaiming")
            self.model.apply(initialize_weights)
        else:
            log.warning("pretrained encoder weights (%s) cannot be downloaded here; expecting a checkpoint load",
                        clean.get("encoder_weights"))
        self.encoder_weights = clean.get("encoder_weights")

        self._hparams = Conf(network=Conf(network, losses=losses, classes=classes), training=training)

        self.classes = classes
        self.classes_int = list(range(n_classes))
        self.classes_int_wout_bg = [c for c in self.classes_int if c != 0]
        self.in_channels = network.get("in_channels", 3)

        self.dice_loss = None
        self.focal_loss = None
        self.boundary_loss = None
        self.initial_alpha = 0.01
        self.boundary_loss_ramped = False

        assert (("GDICE" in losses) and ("DICE" in losses)) is False, f"Only GDICE _OR_ DICE allowed {losses}"
        for comp in losses:
            if comp == "GDICE":
                self.dice_loss = GeneralizedDiceLoss()
            elif comp == "GWDICE":
                dist_mat = np.array([[0.0, 1.0, 1.0], [1.0, 0.0, 0.5], [1.0, 0.5, 0.0]])      # segmodel.py:119
                if n_classes == 2:      # the reference tests `self.classes_int == 2` (list vs int, never true) and then fails
                    dist_mat = dist_mat[0:2, 0:2]      # on the 3 x 3 matrix; the 2 x 2 cut it intends is applied here
                self.dice_loss = GeneralizedWassersteinDiceLoss(dist_matrix=dist_mat)
            elif comp == "DICE":
                self.dice_loss = DiceLoss(idc=self.classes_int_wout_bg)
            elif comp == "FOCAL":
                self.focal_loss = FocalLoss(idc=self.classes_int, gamma=2)
            elif comp == "BOUNDARY":
                self.boundary_loss = BoundaryLoss(idc=self.classes_int_wout_bg)
            elif comp == "BOUNDARY-RAMPED":
                self.boundary_loss = BoundaryLoss(idc=self.classes_int_wout_bg)
                self.boundary_loss_ramped = True
            else:
                raise NotImplementedError(f"The loss component <{comp}> is not recognized")
        self.loss_names = losses
        assert self.dice_loss is not None

        self.stats = {"train": Counter(), "val": Counter(), "test": Counter()}
        self.confusion: Dict[str, Any] = {}
        self.logged: Dict[str, Any] = {}
        self._epoch = 0

    # -- Lightning-compatible odds and ends ----------------------------------------------------
    @property
    def hparams(self):  # type: ignore[override]
        return self._hparams

    @property
    def current_epoch(self) -> int:  # type: ignore[override]
        return self._epoch

    @property
    def alpha(self):
        return min((self.current_epoch + 1) * self.initial_alpha, 0.99)

    def log(self, name, value, **kw):  # type: ignore[override]
        self.logged[name] = value

    def forward(self, x: Tensor) -> Tensor:
        return self.model(x)

    # -- fused loss + metric ---------------------------------------------------------------------
    def _dice_mode(self) -> int:
        """dice term of the fused loss kernels: 1 DiceLoss, 2 GeneralizedDiceLoss, 0 none (GWDICE has its own kernels)"""
        if isinstance(self.dice_loss, GeneralizedWassersteinDiceLoss):
            return 0
        return 2 if isinstance(self.dice_loss, GeneralizedDiceLoss) else 1

    def _fused_terms(self, logits: Tensor, mask: Tensor):
        return fused.SegLossTerms(logits, mask, dice_mode=self._dice_mode(), use_focal=self.focal_loss is not None)

    def calculate_loss(self, y_hat: Tensor, y: Tensor, stage: str, distmap: Optional[Tensor] = None) -> Tensor:
        """compound loss on probabilities + one-hot, as ``segmodel.py:169-200`` (API-compatible path)."""
        loss = 0
        if self.dice_loss:
            if isinstance(self.dice_loss, GeneralizedWassersteinDiceLoss):
                loss_gd = self.dice_loss(y_hat, torch.argmax(y, dim=1))      # "hack to make gwdice work" (segmodel.py:176-178)
            else:
                loss_gd = self.dice_loss(y_hat, y)
            if torch.isnan(loss_gd) or torch.isinf(loss_gd):
                log.warning("Train dice loss is NaN! What is going on?")
            self.log(f"{stage}/dice_loss", loss_gd, on_step=False, on_epoch=True)
            loss = loss + loss_gd
        if self.boundary_loss and distmap is not None:
            loss_bd = self.boundary_loss(y_hat, distmap)
            self.log(f"{stage}/boundary_loss", loss_bd, on_step=False, on_epoch=True)
            loss = loss + (self.alpha * loss_bd if self.boundary_loss_ramped else loss_bd)
        if self.focal_loss:
            loss_fo = self.focal_loss(y_hat, y)
            self.log(f"{stage}/focal_loss", loss_fo, on_step=False, on_epoch=True)
            loss = loss + loss_fo
        self.log(f"{stage}/total_loss", loss, on_step=False, on_epoch=True)
        return loss

    def _eval_step(self, img: Tensor, mask: Tensor, stage: str, distmap: Optional[Tensor] = None):
        """forward + fused softmax/loss/metric pass (one read of the logits)."""
        logits = self.model(img)
        terms = self._fused_terms(logits, mask)
        terms.check_labels()  # class2one_hot's assert (losses.py:129)
        loss = terms.dice_loss
        if isinstance(self.dice_loss, GeneralizedWassersteinDiceLoss):
            loss = self.dice_loss.on_logits(logits, mask)
        self.log(f"{stage}/dice_loss", loss)
        if self.boundary_loss and distmap is not None:
            y_hat = fused.softmax_nchw(logits)
            loss_bd = self.boundary_loss(y_hat, distmap)
            self.log(f"{stage}/boundary_loss", loss_bd)
            loss = loss + (self.alpha * loss_bd if self.boundary_loss_ramped else loss_bd)
        if self.focal_loss:
            self.log(f"{stage}/focal_loss", terms.focal_loss)
            loss = loss + terms.focal_loss
        self.log(f"{stage}/total_loss", loss)
        self.log(f"{stage}/dice", terms.fscore)
        self.log(f"{stage}/dice_with_bg", terms.fscore_with_bg)
        return logits, loss, terms

    def training_step(self, batch, batch_idx):
        """forward (train-mode BatchNorm) + compound loss + metrics, as ``segmodel.py:210-229``; the returned loss is
        differentiable: ``loss.backward()`` runs the CUDA backward pass and fills ``.grad`` of every parameter."""
        img, mask, distmap, _, stats = create_combined_batch(batch)
        logits = self.model(img)
        use_bd = bool(self.boundary_loss) and distmap is not None
        gw = isinstance(self.dice_loss, GeneralizedWassersteinDiceLoss)
        loss, terms = fused.seg_loss(logits, mask, self._dice_mode(), self.focal_loss is not None,
                                     distmap=distmap if use_bd else None,
                                     boundary_idc=self.boundary_loss.idc if use_bd else None,
                                     boundary_weight=self.alpha if self.boundary_loss_ramped else 1.0,
                                     gwdl_matrix=self.dice_loss.matrix() if gw else None)
        terms.check_labels()  # class2one_hot's assert (losses.py:129); also the host sync the reference has there
        self.log("train/dice_loss", terms.gwdl if gw else terms.dice_loss, on_step=False, on_epoch=True)
        if use_bd:
            self.log("train/boundary_loss", terms.boundary, on_step=False, on_epoch=True)
        if self.focal_loss:
            self.log("train/focal_loss", terms.focal_loss, on_step=False, on_epoch=True)
        self.log("train/total_loss", loss.detach(), on_step=False, on_epoch=True)
        if torch.isnan(loss) or torch.isinf(loss):
            log.warning("Train loss is NaN! What is going on?")
            return None
        self.log("train/dice", terms.fscore, on_step=False, on_epoch=True)
        self.log("train/dice_with_bg", terms.fscore_with_bg, on_step=False, on_epoch=True)
        self.stats["train"].update([x["file"] for x in stats])
        return loss

    def validation_step(self, batch, batch_idx):
        img, mask, distmap, lu, stats = create_combined_batch(batch)
        logits, loss, _ = self._eval_step(img, mask, "val", distmap)
        self.stats["val"].update([x["file"] for x in stats])
        return {"val_loss": loss, "target": mask, "prediction": ops.argmax_nchw(logits).long(), "lu": lu}

    def test_step(self, batch: Tuple[Tensor], batch_idx) -> Dict[str, Any]:
        img, mask, _, lu, stats = batch
        logits, _, _ = self._eval_step(img, mask, "test")
        self.stats["test"].update([x["file"] for x in stats])
        return {"target": mask, "prediction": ops.argmax_nchw(logits).long(), "lu": lu}

    # -- confusion matrices of an epoch (segmodel.py:291-407) --------------------------------------
    def _epoch_confusion(self, outputs, stage: str, with_px: bool):
        """one histogram pass per step output instead of ``torch.cat`` + four ``confusion_matrix`` calls: all pixels and the
        forest pixels (``lu == 1``) are counted together; the normalised matrices divide each target row by its sum (rows of
        absent classes are 0, as torchmetrics)."""
        import pandas as pd
        K = len(self.classes)
        counts = bad = None
        for out in outputs:
            pred, lu = out["prediction"], out.get("lu")
            if lu is not None and lu.device != pred.device:      # a land-use mask the trainer left on the host
                lu = lu.to(pred.device)
            counts, bad = ops.confusion_matrix(pred, out["target"].to(pred.device), K, lu=lu, counts=counts, bad=bad)
        assert int(bad.item()) == 0, "prediction / target outside [0, K)"      # one synchronisation per epoch
        cm = counts.cpu().numpy()
        norm = cm / np.maximum(cm.sum(axis=2, keepdims=True), 1)
        mats = {"cm_norm": norm[0], "cm_px": cm[0], "cm_norm_masked": norm[1], "cm_px_masked": cm[1]}
        keys = ["cm_norm", "cm_px", "cm_norm_masked", "cm_px_masked"] if with_px else ["cm_norm", "cm_norm_masked"]
        dfs = {k: pd.DataFrame(mats[k], index=self.classes, columns=self.classes) for k in keys}
        self.confusion[stage] = dfs
        return dfs

    def validation_epoch_end(self, outputs):
        """normalised confusion matrices, all pixels and forest only (``segmodel.py:291-332``); the chart / wandb upload of
        the reference is UI and not part of this package - the tables are kept in ``self.confusion["val"]``."""
        return self._epoch_confusion(outputs, "val", with_px=False)

    def test_epoch_end(self, outputs):
        """normalised and pixel-count confusion matrices, all pixels and forest only (``segmodel.py:334-407``)."""
        dfs = self._epoch_confusion(outputs, "test", with_px=True)
        log.info(f"CM - DEFAULT - NORMALIZED: {dfs['cm_norm'].to_string()}")
        log.info(f"CM - FORESTONLY - NORMALIZED: {dfs['cm_norm_masked'].to_string()}")
        log.info(f"CM - DEFAULT - PIXEL: {dfs['cm_px'].to_string()}")
        log.info(f"CM - FORESTONLY - PIXEL: {dfs['cm_px_masked'].to_string()}")
        return dfs

    def configure_optimizers(self):
        from ..optim import FusedAdam

        # the reference leaves clipping to the Lightning trainer (configs/trainer/default.yaml:18, gradient_clip_val 0.5);
        # without a trainer, `training.gradient_clip_val` folds the same global-norm clip into the fused Adam step
        opt = FusedAdam(self.parameters(), lr=self.hparams.training.learning_rate,
                        max_grad_norm=float(self.hparams.training.get("gradient_clip_val", 0.0) or 0.0))
        if next(self.parameters()).is_cuda:
            opt.attach_engine(self.model.train_engine())     # one flat clip + Adam launch per step
        sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=self.hparams.training.cosineannealing_tmax)
        return [opt], [sch]

    # -- checkpoints -----------------------------------------------------------------------------
    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, **kwargs):
        """Loads a Lightning ``.ckpt`` written by the reference (or by :meth:`save_checkpoint`)."""
        ckpt = torch.load(str(checkpoint_path), map_location=map_location or "cpu", weights_only=False)
        hp = ckpt.get("hyper_parameters", {})
        network = kwargs.get("network", hp.get("network"))
        training = kwargs.get("training", hp.get("training"))
        if network is None:
            raise KeyError("checkpoint has no hyper_parameters.network")
        model = cls(network, training)
        model.load_state_dict(ckpt["state_dict"], strict=True)
        return model

    def save_checkpoint(self, path) -> None:
        def plain(c):
            return {k: (plain(v) if isinstance(v, dict) else v) for k, v in c.items()}
        torch.save({"state_dict": self.state_dict(), "hyper_parameters": plain(self._hparams),
                    "epoch": self._epoch, "pytorch-lightning_version": "b200-drop-in"}, str(Path(path)))


def initialize_weights(m):
    """Kaiming-normal conv/linear weights, zero biases (segmodel.py:432-438)."""
    if getattr(m, "bias", None) is not None:
        torch.nn.init.constant_(m.bias, 0)
    if isinstance(m, (torch.nn.Conv2d, torch.nn.Linear)):
        torch.nn.init.kaiming_normal_(m.weight)
    for c in m.children():
        initialize_weights(c)
