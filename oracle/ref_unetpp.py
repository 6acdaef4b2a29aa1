"""Oracle: CPU fp32 restatement of ``smp.UnetPlusPlus(encoder_name="resnet34")``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

The reference selects it with ``architecture in ["unetplusplus", "unet++"]`` (``deadtrees/network/segmodel.py:63-64``).  The
class lives in the absent third-party ``segmentation_models_pytorch>=0.2.1`` (``setup.py:47``); its decoder topology is
evidenced in-tree by the vendored fork ``deadtrees/network/extra/efficientunetplusplus/decoder.py``, which keeps smp's
``UnetPlusPlusDecoder`` bookkeeping verbatim and only swaps the block type:

* channel bookkeeping and block names ``x_{depth}_{layer}``: ``decoder.py:116-153``
* dense forward pass (which maps are concatenated, in which order): ``decoder.py:156-185``
* a block = nearest x2 -> ``cat([x, skip])`` -> two Conv3x3 + BN + ReLU (smp's own ``DecoderBlock``; the fork replaces the
  two convs by inverted residuals, ``decoder.py:61-100``; ``Conv2dReLU`` = ``extra/modules.py:53-92``)
* encoder / head as in ``ref_unet`` (torchvision resnet34 trunk, ``Conv2d(16 -> K, 3, padding=1)``).

PARITY: unpinned by the reference (no test touches this architecture); pinned against the CUDA engine on shared
state-dicts, and its encoder against torchvision through ``ref_unet``.
"""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

from .ref_unet import DecoderBlock, ResNet34Encoder, initialize_weights


class UnetPlusPlusDecoder(nn.Module):
    def __init__(self, encoder_channels: Sequence[int], decoder_channels: Sequence[int]):
        super().__init__()
        enc = list(encoder_channels[1:])[::-1]
        self.in_channels = [enc[0]] + list(decoder_channels[:-1])
        self.skip_channels = list(enc[1:]) + [0]
        self.out_channels = list(decoder_channels)
        blocks = {}
        for layer_idx in range(len(self.in_channels) - 1):
            for depth_idx in range(layer_idx + 1):
                if depth_idx == 0:
                    in_ch = self.in_channels[layer_idx]
                    skip_ch = self.skip_channels[layer_idx] * (layer_idx + 1)
                    out_ch = self.out_channels[layer_idx]
                else:
                    out_ch = self.skip_channels[layer_idx]
                    skip_ch = self.skip_channels[layer_idx] * (layer_idx + 1 - depth_idx)
                    in_ch = self.skip_channels[layer_idx - 1]
                blocks[f"x_{depth_idx}_{layer_idx}"] = DecoderBlock(in_ch, skip_ch, out_ch)
        blocks[f"x_0_{len(self.in_channels) - 1}"] = DecoderBlock(self.in_channels[-1], 0, self.out_channels[-1])
        self.blocks = nn.ModuleDict(blocks)
        self.depth = len(self.in_channels) - 1

    def forward(self, *features):
        features = features[1:][::-1]
        dense = {}
        for layer_idx in range(len(self.in_channels) - 1):
            for depth_idx in range(self.depth - layer_idx):
                if layer_idx == 0:
                    dense[f"x_{depth_idx}_{depth_idx}"] = self.blocks[f"x_{depth_idx}_{depth_idx}"](
                        features[depth_idx], features[depth_idx + 1])
                else:
                    li = depth_idx + layer_idx
                    cat = [dense[f"x_{idx}_{li}"] for idx in range(depth_idx + 1, li + 1)]
                    cat = torch.cat(cat + [features[li + 1]], dim=1)
                    dense[f"x_{depth_idx}_{li}"] = self.blocks[f"x_{depth_idx}_{li}"](dense[f"x_{depth_idx}_{li - 1}"], cat)
        dense[f"x_0_{self.depth}"] = self.blocks[f"x_0_{self.depth}"](dense[f"x_0_{self.depth - 1}"])
        return dense[f"x_0_{self.depth}"]


class UnetPlusPlus(nn.Module):
    def __init__(self, in_channels: int = 3, classes: int = 3, decoder_channels: Sequence[int] = (256, 128, 64, 32, 16)):
        super().__init__()
        self.encoder = ResNet34Encoder(in_channels)
        self.decoder = UnetPlusPlusDecoder((in_channels, 64, 64, 128, 256, 512), decoder_channels)
        self.segmentation_head = nn.Sequential(nn.Conv2d(decoder_channels[-1], classes, 3, padding=1), nn.Identity(), nn.Identity())

    def forward(self, x):
        return self.segmentation_head(self.decoder(*self.encoder(x)))


def build_reference_unetpp(in_channels: int = 3, classes: int = 3, seed: int = 0) -> UnetPlusPlus:
    """random-init oracle model (reference init ``segmodel.py:432-438``) with non-trivial BatchNorm statistics"""
    g = torch.Generator().manual_seed(seed)
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        model = UnetPlusPlus(in_channels, classes)
        model.apply(initialize_weights)
    finally:
        torch.random.set_rng_state(state)
    for mod in model.modules():
        if isinstance(mod, nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
            mod.weight.data.copy_(1.0 + 0.1 * torch.randn(mod.num_features, generator=g))
            mod.bias.data.copy_(0.1 * torch.randn(mod.num_features, generator=g))
    return model.eval()
