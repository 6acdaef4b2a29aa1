"""``Unet`` — drop-in for ``smp.Unet(encoder_name="resnet34")`` as the reference instantiates it.

The reference builds ``self.model = smp.Unet(**network_conf, classes=n)``
(``deadtrees/network/segmodel.py:62-63,79-85``).  This module keeps smp's attribute names
(``encoder`` / ``decoder`` / ``segmentation_head``) and state-dict keys so
checkpoints load unchanged and ``list(model.parameters())[0].shape[1]`` is the channel count
(``deployment/inference.py:42``).  The ``nn`` sub-modules below only OWN the parameters; the forward
pass runs through :class:`deadtrees_b200.engine.UnetEngine` (hand-written CUDA kernels).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn

from ..engine import RESNET34_LAYERS, RESNET34_PLANES, UnetEngine, UnetPlusPlusEngine
from .. import ops


class _Block(nn.Module):
    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))


class _Encoder(nn.Module):
    def __init__(self, in_channels: int):
        super().__init__()
        self.conv1 = nn.Conv2d(in_channels, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        for li, (planes, nblk) in enumerate(zip(RESNET34_PLANES, RESNET34_LAYERS), start=1):
            blocks = []
            for b in range(nblk):
                blocks.append(_Block(cin, planes, 2 if (b == 0 and li > 1) else 1))
                cin = planes
            setattr(self, f"layer{li}", nn.Sequential(*blocks))


class _ConvBnRelu(nn.Sequential):
    def __init__(self, cin: int, cout: int):
        super().__init__(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))


class _DecoderBlock(nn.Module):
    def __init__(self, cin: int, cskip: int, cout: int):
        super().__init__()
        self.conv1 = _ConvBnRelu(cin + cskip, cout)
        self.conv2 = _ConvBnRelu(cout, cout)


class _Decoder(nn.Module):
    def __init__(self, decoder_channels: Sequence[int]):
        super().__init__()
        enc = [512, 256, 128, 64, 64]
        cins = [enc[0]] + list(decoder_channels[:-1])
        skips = enc[1:] + [0]
        self.blocks = nn.ModuleList(_DecoderBlock(i, s, o) for i, s, o in zip(cins, skips, decoder_channels))


class _DecoderPlusPlus(nn.Module):
    """parameters of smp's ``UnetPlusPlusDecoder`` in its registration order (topology:
    ``deadtrees/network/extra/efficientunetplusplus/decoder.py:116-153``)"""

    def __init__(self, decoder_channels: Sequence[int]):
        super().__init__()
        enc = [512, 256, 128, 64, 64]
        in_ch = [enc[0]] + list(decoder_channels[:-1])
        skip_ch = enc[1:] + [0]
        out_ch = list(decoder_channels)
        blocks = {}
        for layer_idx in range(len(in_ch) - 1):
            for depth_idx in range(layer_idx + 1):
                if depth_idx == 0:
                    i, s, o = in_ch[layer_idx], skip_ch[layer_idx] * (layer_idx + 1), out_ch[layer_idx]
                else:
                    o, s, i = skip_ch[layer_idx], skip_ch[layer_idx] * (layer_idx + 1 - depth_idx), skip_ch[layer_idx - 1]
                blocks[f"x_{depth_idx}_{layer_idx}"] = _DecoderBlock(i, s, o)
        blocks[f"x_0_{len(in_ch) - 1}"] = _DecoderBlock(in_ch[-1], 0, out_ch[-1])
        self.blocks = nn.ModuleDict(blocks)


class Unet(nn.Module):
    ENGINE = UnetEngine

    def __init__(self, encoder_name: str = "resnet34", encoder_depth: int = 5, encoder_weights: Optional[str] = None,
                 decoder_channels: Sequence[int] = (256, 128, 64, 32, 16), in_channels: int = 3, classes: int = 1,
                 precision: str = "bf16", **unused):
        super().__init__()
        if str(encoder_name).lower() != "resnet34":
            raise NotImplementedError("the B200 build implements encoder_name='resnet34' only")
        if int(encoder_depth) != 5 or list(decoder_channels) != [256, 128, 64, 32, 16]:
            raise NotImplementedError("the B200 build implements encoder_depth=5, decoder_channels=[256,128,64,32,16]")
        self.in_channels, self.classes, self.precision = int(in_channels), int(classes), precision
        self.encoder = _Encoder(self.in_channels)
        self.decoder = self._make_decoder(list(decoder_channels))
        self.segmentation_head = nn.Sequential(nn.Conv2d(decoder_channels[-1], self.classes, 3, padding=1),
                                               nn.Identity(), nn.Identity())
        self._engine: Optional[UnetEngine] = None
        self._engine_key = None
        self._train_engine = None

    def _make_decoder(self, decoder_channels):
        return _Decoder(decoder_channels)

    # -- engine management -------------------------------------------------------------------
    def _param_version(self):
        return tuple(p._version for p in self.parameters()) + tuple(b._version for b in self.buffers())

    def engine(self) -> UnetEngine:
        # ops.STATE_GENERATION: the optimizer / train-mode BatchNorm kernels and CUDA-graph replays write parameters and
        # running statistics through raw pointers, which never bump `_version`
        key = (self._param_version(), ops.STATE_GENERATION, self.precision, next(self.parameters()).device)
        if self._engine is None or key != self._engine_key:
            if self.training:
                raise NotImplementedError(
                    "train-mode (batch-statistics BatchNorm) forward is not implemented in this round; call .eval()")
            self._engine = self.ENGINE(self.state_dict(), self.in_channels, self.classes, precision=self.precision)
            self._engine_key = key
        return self._engine

    def set_precision(self, precision: str) -> "Unet":
        if precision != self.precision:
            self.precision = precision
            self._train_engine = None     # an optimizer attached to the old engine refuses to step (FusedAdam.step)
        return self

    def train_engine(self):
        """engine of the train-mode forward / backward (batch-statistics BatchNorm, autograd through the CUDA library)."""
        from ..train_engine import UnetTrainEngine
        dev = next(self.parameters()).device
        if self._train_engine is None or self._train_engine.device != dev or self._train_engine.precision != self.precision:
            self._train_engine = UnetTrainEngine(self, precision=self.precision)
        return self._train_engine

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """(N, C, H, W) float -> (N, classes, H, W) fp32 logits, as ``smp.Unet.forward``."""
        if not x.is_cuda:
            from .._lib import DeadtreesB200Error
            raise DeadtreesB200Error("deadtrees_b200.Unet runs on a CUDA (B200) device only; move the input with .cuda()")
        if x.dim() != 4 or x.shape[1] < self.in_channels:
            raise ValueError(f"expected (N, >={self.in_channels}, H, W), got {tuple(x.shape)}")
        if self.training:
            from ..train_engine import unet_train_forward
            return unet_train_forward(self.train_engine(), x)
        eng = self.engine()
        xin = ops.pack_input_nchw(x, self.in_channels, eng.act_dtype)
        return eng.forward(xin, want_logits_nchw=True)["logits_nchw"]


class UnetPlusPlus(Unet):
    """drop-in for ``smp.UnetPlusPlus(encoder_name="resnet34")`` (``segmodel.py:63-64``): inference on the B200 kernels
    (:class:`deadtrees_b200.engine.UnetPlusPlusEngine`); the training step is built for the Unet path only."""
    ENGINE = UnetPlusPlusEngine

    def _make_decoder(self, decoder_channels):
        return _DecoderPlusPlus(decoder_channels)

    def train_engine(self):
        raise NotImplementedError("the B200 training step implements architecture 'unet'; Unet++ runs inference (eval) only")

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            raise NotImplementedError("the B200 training step implements architecture 'unet'; call .eval() for Unet++")
        return super().forward(x)
