"""Inference engine for Unet(resnet34): folds BN, repacks weights for the CUDA kernels and sequences
the fused convolutions of one forward pass.

Follows what ``smp.Unet.forward`` computes in the reference (call sites
``deadtrees/network/segmodel.py:214`` and ``deadtrees/deployment/inference.py:60``; layer list in
SURVEY.md Appendix A); state-dict key layout = smp's.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import ops
from ._lib import CONV_PAIR, CONV_UPS_FOLDED, CONV_X_PAD3, require_device

BN_EPS = 1e-5
RESNET34_LAYERS = (3, 4, 6, 3)
RESNET34_PLANES = (64, 128, 256, 512)
BK = 64


@dataclass
class FusedConv:
    """conv (+ folded eval BatchNorm) (+ residual) (+ ReLU) in the kernels' weight layout."""
    name: str
    C_in: int
    C_x: int
    C_out: int
    R: int
    S: int
    stride: int
    pad: int
    relu: bool
    upsample: bool
    w: torch.Tensor
    scale: torch.Tensor
    shift: torch.Tensor
    flops_per_px_out: int = field(default=0)
    flags: int = field(default=0)          # extra DT_CONV_* flags of this layer (e.g. DT_CONV_UPS_FOLDED)


def fold_bn(sd: Dict[str, torch.Tensor], prefix: str, C_out: int, device):
    """eval BatchNorm -> per-channel (scale, shift): y = x*g/sqrt(v+eps) + (b - m*g/sqrt(v+eps))."""
    if prefix + ".weight" not in sd:
        return torch.ones(C_out, device=device), torch.zeros(C_out, device=device)
    g, b = sd[prefix + ".weight"].float(), sd[prefix + ".bias"].float()
    m, v = sd[prefix + ".running_mean"].float(), sd[prefix + ".running_var"].float()
    scale = g / torch.sqrt(v + BN_EPS)
    return scale.to(device).contiguous(), (b - m * scale).to(device).contiguous()


def pack_weight(w: torch.Tensor, precision: str, stem: bool, device) -> torch.Tensor:
    """OIHW fp32 -> kernel layout (see include/deadtrees_b200.h, dt_conv2d_fwd)."""
    C_out, C_in, R, S = w.shape
    w = w.float()
    if stem and C_in < 4:
        w = torch.cat([w, w.new_zeros(C_out, 4 - C_in, R, S)], dim=1)
        C_in = 4
    if precision == "fp32":
        return w.permute(2, 3, 1, 0).reshape(R * S, C_in, C_out).contiguous().to(device)
    if stem:
        wp = w.new_zeros(C_out, R, 8, 4)
        wp[:, :, :S, :] = w.permute(0, 2, 3, 1)
        flat = wp.reshape(C_out, R * 32)
        kpad = 256
    else:
        flat = w.permute(0, 2, 3, 1).reshape(C_out, R * S * C_in)
        kpad = (flat.shape[1] + BK - 1) // BK * BK
    out = w.new_zeros(C_out, kpad)
    out[:, : flat.shape[1]] = flat
    return out.to(torch.bfloat16).contiguous().to(device)


def fold_upsample_weights(w: torch.Tensor) -> torch.Tensor:
    """(C_out, C_in, 3, 3) fp32 -> (C_out, 4 classes, 4 effective taps, C_in) fp32: nearest-x2 up-sampling folded into the
    weights.  Output parity class (a, b) reads the 2 x 2 low-res pixels (a-1+ey, b-1+ex); the taps (fr, fs) of the 3 x 3
    filter that land on the same low-res pixel, floor((a+fr-1)/2) = a-1+ey, are summed (in fp32, from the fp32 weights)."""
    C_out, C_in = w.shape[:2]
    out = w.new_zeros(C_out, 4, 4, C_in)
    for a in range(2):
        for b in range(2):
            for fr in range(3):
                ey = (a + fr - 1) // 2 - (a - 1)
                for fs in range(3):
                    ex = (b + fs - 1) // 2 - (b - 1)
                    out[:, a * 2 + b, ey * 2 + ex] += w[:, :, fr, fs]
    return out


def pack_weight_folded(w: torch.Tensor, device, C_x: Optional[int] = None) -> torch.Tensor:
    """DT_CONV_UPS_FOLDED packing: bf16 [C_out][16 * C_x + 9 * C_s]; the first C_x input channels are the up-sampled
    operand, k = ((class * 4 + e) * C_x + ci), the remaining C_s the skip operand, k = 16 * C_x + tap * C_s + cs."""
    C_out, C_in = w.shape[:2]
    C_x = C_in if C_x is None else C_x
    w = w.float()
    parts = [fold_upsample_weights(w[:, :C_x]).reshape(C_out, 16 * C_x)]
    if C_x < C_in:
        parts.append(w[:, C_x:].permute(0, 2, 3, 1).reshape(C_out, 9 * (C_in - C_x)))
    return torch.cat(parts, dim=1).to(torch.bfloat16).contiguous().to(device)


def folds_upsample(precision: str, C_in: int, C_x: int, C_out: int, conv_flags: int = 0) -> bool:
    """up-sample layers whose up-sampling is folded into the weights of the x operand (bf16 path): the shapes the
    resident-weight parity kernel (no skip) or the class-fused kernel (64-channel slabs, C_out 32 / 64) have a folded
    form for - in Unet-resnet34 decoder.blocks.2 / 3 / 4 .conv1."""
    if precision != "bf16" or conv_flags:
        return False
    if C_x == C_in:
        return (C_in, C_out) in ((32, 16), (32, 32), (16, 16), (64, 32))
    return C_x % 64 == 0 and (C_in - C_x) % 64 == 0 and C_out in (32, 64)


class UnetEngine:
    """B200 forward pass of the reference's ``smp.Unet(resnet34, depth 5, decoder (256,128,64,32,16))``."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], in_channels: int, classes: int,
                 precision: str = "bf16", device: Optional[torch.device] = None, conv_flags: int = 0):
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        require_device()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.precision = precision
        self.act_dtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self.in_channels, self.classes = in_channels, classes
        self.conv_flags = conv_flags
        if not 1 <= in_channels <= 4:
            raise ValueError("in_channels must be 1..4")
        if not 1 <= classes <= 4:
            raise ValueError("classes must be 1..4")
        sd = {k: v.detach() for k, v in state_dict.items()}
        # wide 3x3/s1 layers on CTA pairs (tcgen05.mma.cta_group::2) where the shape allows; the library falls back itself
        import os
        self.pair_flag = CONV_PAIR if (precision == "bf16" and not conv_flags and
                                       os.environ.get("DT_CONV_PAIR", "1") != "0") else 0
        self.layers: Dict[str, FusedConv] = {}
        self.folded: Dict[str, FusedConv] = {}      # up-sample layers in the DT_CONV_UPS_FOLDED packing
        self._build(sd)
        self._ws: Dict[tuple, Dict[str, torch.Tensor]] = {}

    # ------------------------------------------------------------------------------------------
    def _conv(self, sd, name, conv_key, bn_key, stride, pad, relu, C_x=None, upsample=False, stem=False):
        w = sd[conv_key + ".weight"]
        C_out, C_in, R, S = w.shape
        scale, shift = fold_bn(sd, bn_key, C_out, self.device)
        if stem:
            C_in = 4
        cx = C_in if C_x is None else C_x
        if upsample and folds_upsample(self.precision, C_in, cx, C_out, self.conv_flags):
            # used when the low-res grid tiles into 16 x 8 regions (folded_ok); the nine-tap packing below otherwise
            self.folded[name] = FusedConv(name, C_in, cx, C_out, R, S, stride, pad, relu, upsample,
                                          pack_weight_folded(w, self.device, cx), scale, shift, 2 * C_in * R * S * C_out,
                                          flags=CONV_UPS_FOLDED)
        self.layers[name] = FusedConv(name, C_in, cx, C_out, R, S, stride, pad, relu,
                                      upsample, pack_weight(w, self.precision, stem, self.device), scale, shift,
                                      2 * C_in * R * S * C_out)

    def _build(self, sd):
        self._build_encoder(sd)
        self._build_decoder(sd)
        self._build_head(sd)

    def _build_encoder(self, sd):
        self._conv(sd, "stem", "encoder.conv1", "encoder.bn1", 2, 3, True, stem=True)
        for li, nblk in enumerate(RESNET34_LAYERS, start=1):
            for b in range(nblk):
                p = f"encoder.layer{li}.{b}"
                stride = 2 if (b == 0 and li > 1) else 1
                self._conv(sd, p + ".conv1", p + ".conv1", p + ".bn1", stride, 1, True)
                self._conv(sd, p + ".conv2", p + ".conv2", p + ".bn2", 1, 1, True)  # ReLU after the residual add
                if p + ".downsample.0.weight" in sd:
                    self._conv(sd, p + ".downsample", p + ".downsample.0", p + ".downsample.1", stride, 0, False)

    def _build_decoder(self, sd):
        self.decoder_blocks = 0
        while f"decoder.blocks.{self.decoder_blocks}.conv1.0.weight" in sd:
            self.decoder_blocks += 1
        if self.decoder_blocks != 5:
            raise NotImplementedError("only encoder_depth=5 / five decoder blocks are supported")
        x_ch = 512
        for i in range(5):
            p = f"decoder.blocks.{i}"
            self._conv(sd, p + ".conv1", p + ".conv1.0", p + ".conv1.1", 1, 1, True, C_x=x_ch, upsample=True)
            self._conv(sd, p + ".conv2", p + ".conv2.0", p + ".conv2.1", 1, 1, True)
            x_ch = self.layers[p + ".conv1"].C_out

    def _build_head(self, sd):
        hw = sd["segmentation_head.0.weight"].float()
        if hw.shape[0] != self.classes:
            raise ValueError("segmentation head does not match `classes`")
        self.head_w = hw.permute(2, 3, 1, 0).reshape(9, hw.shape[1], hw.shape[0]).contiguous().to(self.device)
        self.head_b = sd["segmentation_head.0.bias"].float().contiguous().to(self.device)
        # tensor-core head (bf16 path): weights padded to 16 output channels in the conv packing, bias to 16 floats
        self.head_w_tc = self.head_b16 = None
        if self.precision == "bf16" and hw.shape[1] == 16 and not self.conv_flags:
            hw16 = torch.zeros(16, 16, 3, 3)
            hw16[: hw.shape[0]] = hw.cpu()
            self.head_w_tc = pack_weight(hw16, "bf16", False, self.device)
            b16 = torch.zeros(16)
            b16[: hw.shape[0]] = sd["segmentation_head.0.bias"].float().cpu()
            self.head_b16 = b16.to(self.device)

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def folded_ok(H: int, W: int) -> bool:
        """the folded kernels walk 16 x 8 regions of the LOW-RES grid (H, W: the layer's output size)"""
        return (H // 2) % 16 == 0 and (W // 2) % 8 == 0

    def _run(self, name, x, N, H, W, skip=None, residual=None, out=None, flags=0):
        L = self.layers[name]
        if name in self.folded and self.folded_ok(H, W):
            L = self.folded[name]
        return ops.conv2d(x, L.w, L.scale, L.shift, N=N, H=H, W=W, C_in=L.C_in, C_x=L.C_x, C_out=L.C_out, R=L.R,
                          S=L.S, stride=L.stride, pad=L.pad, relu=L.relu, skip=skip, upsample=L.upsample,
                          residual=residual, out=out, flags=self.conv_flags | flags | L.flags | self.pair_flag,
                          algo_cin=self.in_channels if name == "stem" else (
                              (L.C_x * 4.0 / 9.0 + (L.C_in - L.C_x)) if L.flags & CONV_UPS_FOLDED else None),
                          tag=name)   # folded layers: the FLOPs the tensor pipe executes

    def _buf(self, ws, key, shape):
        t = ws.get(key)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=self.act_dtype, device=self.device)
            ws[key] = t
        return t

    def stem_padded(self, T: int) -> bool:
        """True when the stem can take its input as a zero-bordered (N, T+6, T+8, 4) frame and build the im2col
        operand with TMA (conv_stem.cu): bf16 path and an output grid that tiles into 128-pixel boxes."""
        wo = T // 2
        return self.precision == "bf16" and not self.conv_flags and (wo % 128 == 0 or wo in (16, 32, 64))

    def alloc_input(self, N: int, T: int) -> torch.Tensor:
        """input buffer for `forward(...)`: padded frame (borders zeroed once) or dense (N, T, T, 4)."""
        if self.stem_padded(T):
            return torch.zeros((N, T + 6, T + 8, 4), dtype=self.act_dtype, device=self.device)
        return torch.empty((N, T, T, 4), dtype=self.act_dtype, device=self.device)

    def forward_features(self, x: torch.Tensor, keep: Optional[Dict[str, torch.Tensor]] = None,
                         stop_before_tail: bool = False) -> torch.Tensor:
        """x: (N, T, T, 4) NHWC in the engine's activation dtype, or the padded frame (N, T+6, T+8, 4) from
        `alloc_input` -> decoder output (N, T, T, 16); with `stop_before_tail` the output of the last block's conv1
        (the input of :func:`ops.tail_fused`)."""
        N, H_in, W_in, C4 = x.shape
        padded = W_in == H_in + 2
        T = H_in - 6 if padded else H_in
        if C4 != 4 or (not padded and H_in != W_in) or T % 32 or (padded and not self.stem_padded(T)):
            raise ValueError(f"input must be (N, T, T, 4) or (N, T+6, T+8, 4) with T % 32 == 0, got {tuple(x.shape)}")
        ws = self._ws.setdefault((N, T), {})
        buf = lambda key, h, c: self._buf(ws, key, (N, h, h, c))
        f = {}
        if padded and T == 256 and self.layers["stem"].relu and os.environ.get("DT_STEM_POOL_FUSED", "1") != "0":
            L = self.layers["stem"]                 # stem + maxpool in one launch: the pooling reads the rows from smem
            f[1], cur = buf("f1", T // 2, 64), buf("pool", T // 4, 64)
            ops.stem_pool(x, L.w, L.scale, L.shift, N=N, H=T, W=T, out=f[1], pooled=cur, algo_cin=self.in_channels)
        else:
            f[1] = self._run("stem", x, N, T, T, out=buf("f1", T // 2, 64), flags=CONV_X_PAD3 if padded else 0)
            cur = ops.maxpool3x3s2(f[1], out=buf("pool", T // 4, 64))
        H = T // 4
        for li, (planes, nblk) in enumerate(zip(RESNET34_PLANES, RESNET34_LAYERS), start=1):
            for b in range(nblk):
                p = f"encoder.layer{li}.{b}"
                strided = b == 0 and li > 1
                Ho = H // 2 if strided else H
                identity = cur
                if strided:
                    identity = self._run(p + ".downsample", cur, N, H, H, out=buf(f"ds{li}", Ho, planes))
                t = self._run(p + ".conv1", cur, N, H, H, out=buf(f"t{li}", Ho, planes))
                last = b == nblk - 1
                key = f"f{li + 1}" if last else f"a{li}_{b % 2}"
                cur = self._run(p + ".conv2", t, N, Ho, Ho, residual=identity, out=buf(key, Ho, planes))
                H = Ho
            f[li + 1] = cur
        if keep is not None:
            keep.update({f"f{i}": f[i] for i in f})
        return self._decode(f, N, T, buf, keep, stop_before_tail)

    def materialize_concat(self, name: str, H: int) -> bool:
        """True when an up-sample + concat layer runs faster on a materialised input: bf16 tensor-core path, 16 x 16 output
        (the halo kernels need whole 8 x 16 tiles of the LOW-RES grid), channel counts the CTA-pair kernel takes.
        ``DT_MATERIALIZE_CONCAT=0`` keeps the virtual form."""
        L = self.layers[name]
        return (self.precision == "bf16" and not self.conv_flags and bool(self.pair_flag) and L.upsample and H == 16
                and name not in self.folded and L.C_in % 64 == 0 and L.C_out % 128 == 0 and L.C_in > L.C_x
                and os.environ.get("DT_MATERIALIZE_CONCAT", "1") != "0")

    TAIL_LAYER = "decoder.blocks.4.conv2"

    def tail_fusable(self, T: int) -> bool:
        """True when decoder.blocks.4.conv2 and the head run as one launch (conv_tail.cu): bf16 path, 16 channels into
        the head, full rows of 128 or 256 pixels.  ``DT_TAIL_FUSED=0`` keeps the two launches."""
        L = self.layers.get(self.TAIL_LAYER)
        return (self.head_w_tc is not None and L is not None and L.C_in == 16 and L.C_out == 16 and not self.conv_flags
                and T in (128, 256) and os.environ.get("DT_TAIL_FUSED", "1") != "0")

    def _decode(self, f, N, T, buf, keep, stop_before_tail=False):
        """decoder of ``smp.Unet``: five blocks, each nearest x2 + concat(skip) -> conv -> conv (never materialised)"""
        xcur, H = f[5], T // 32
        skips = [f[4], f[3], f[2], f[1], None]
        for i in range(5):
            p = f"decoder.blocks.{i}"
            H *= 2
            c_out = self.layers[p + ".conv1"].C_out
            if self.materialize_concat(p + ".conv1", H):
                # 8 x 8 low-res images (256^2 tiles, first decoder block): no 8 x 16 halo tile fits, the virtual up-sample +
                # concat would run on the per-tap gather kernel (1.0 PFLOP/s).  The concatenated tensor is small here
                # (0.4 MB per tile): write it once and run the CTA-pair kernel on it as a plain 3 x 3 conv (1.4 PFLOP/s)
                L = self.layers[p + ".conv1"]
                with ops._Timed("conv", 0.0, p + ".conv1"):      # its time belongs to the layer (no FLOPs of its own)
                    cat = ops.upsample_concat(xcur, skips[i], out=buf(f"cat{i}", H, L.C_in))
                t = ops.conv2d(cat, L.w, L.scale, L.shift, N=N, H=H, W=H, C_in=L.C_in, C_x=L.C_in, C_out=L.C_out, R=L.R, S=L.S,
                               stride=1, pad=L.pad, relu=L.relu, out=buf(f"dt{i}", H, c_out), flags=self.pair_flag,
                               tag=p + ".conv1")
                if i == 4 and stop_before_tail:
                    return t
                xcur = self._run(p + ".conv2", t, N, H, H, out=buf(f"d{i}", H, c_out))
                if keep is not None:
                    keep[f"d{i}"] = xcur
                continue
            t = self._run(p + ".conv1", xcur, N, H, H, skip=skips[i], out=buf(f"dt{i}", H, c_out))
            if i == 4 and stop_before_tail:
                return t
            xcur = self._run(p + ".conv2", t, N, H, H, out=buf(f"d{i}", H, c_out))
            if keep is not None:
                keep[f"d{i}"] = xcur
        return xcur

    def forward(self, x: torch.Tensor, *, want_logits_nchw: bool = False, want_logits_nhwc: bool = False,
                want_mask: bool = False, mask_out: Optional[torch.Tensor] = None,
                logits_nhwc_out: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        fused = type(self)._decode is UnetEngine._decode and self.tail_fusable(x.shape[1] - 6 if x.shape[2] == x.shape[1] + 2
                                                                                   else x.shape[1])
        d = self.forward_features(x, stop_before_tail=fused)
        N, T = d.shape[0], d.shape[1]
        out: Dict[str, torch.Tensor] = {}
        if want_logits_nchw:
            out["logits_nchw"] = torch.empty((N, self.classes, T, T), dtype=torch.float32, device=self.device)
        if want_logits_nhwc or logits_nhwc_out is not None:
            out["logits_nhwc"] = logits_nhwc_out if logits_nhwc_out is not None else torch.empty(
                (N, T, T, self.classes), dtype=self.act_dtype, device=self.device)
        if want_mask or mask_out is not None:
            out["mask"] = mask_out if mask_out is not None else torch.empty((N, T, T), dtype=torch.uint8,
                                                                            device=self.device)
        if fused:
            L = self.layers[self.TAIL_LAYER]
            ops.tail_fused(d, L.w, L.scale, L.shift, self.head_w_tc, self.head_b16, self.classes,
                           logits_nchw=out.get("logits_nchw"), logits_nhwc=out.get("logits_nhwc"), mask=out.get("mask"),
                           tag="tail(dec4.conv2+head)")
        elif self.head_w_tc is not None:
            ops.head_tc(d, self.head_w_tc, self.head_b16, self.classes, logits_nchw=out.get("logits_nchw"),
                        logits_nhwc=out.get("logits_nhwc"), mask=out.get("mask"))
        else:
            ops.head(d, self.head_w, self.head_b, logits_nchw=out.get("logits_nchw"),
                     logits_nhwc=out.get("logits_nhwc"), mask=out.get("mask"))
        return out



class UnetPlusPlusEngine(UnetEngine):
    """B200 forward pass of ``smp.UnetPlusPlus(resnet34)`` (``deadtrees/network/segmodel.py:63-64``): the
    resnet34 encoder and the head of :class:`UnetEngine`, and the nested decoder whose topology the reference vendors in
    ``deadtrees/network/extra/efficientunetplusplus/decoder.py:116-185`` - blocks ``x_{depth}_{layer}``, each
    nearest x2(x) ++ [denser maps ..., encoder feature] -> conv3x3+BN+ReLU -> conv3x3+BN+ReLU.

    Every block runs on the same fused kernels as the Unet decoder: the up-sampled operand is read at low resolution
    (never materialised), the concatenated skip maps are one NHWC tensor (``torch.cat`` of the block's dense inputs -
    a copy of plumbing size; a multi-operand loader is not built)."""

    def _build_decoder(self, sd):
        self.decoder_blocks = 0
        self.pp_blocks = []
        enc = [512, 256, 128, 64, 64]
        dec = []
        li = 0
        while f"decoder.blocks.x_0_{li}.conv1.0.weight" in sd:
            dec.append(sd[f"decoder.blocks.x_0_{li}.conv1.0.weight"].shape[0])
            li += 1
        if len(dec) != 5:
            raise NotImplementedError("only encoder_depth=5 / five decoder levels are supported")
        in_ch = [enc[0]] + dec[:-1]
        skip_ch = enc[1:] + [0]
        for layer_idx in range(4):
            for depth_idx in range(layer_idx + 1):
                cx = in_ch[layer_idx] if depth_idx == 0 else skip_ch[layer_idx - 1]
                self._pp_block(sd, f"x_{depth_idx}_{layer_idx}", cx)
        self._pp_block(sd, "x_0_4", in_ch[-1])

    def _pp_block(self, sd, name, cx):
        p = f"decoder.blocks.{name}"
        self._conv(sd, p + ".conv1", p + ".conv1.0", p + ".conv1.1", 1, 1, True, C_x=cx, upsample=True)
        self._conv(sd, p + ".conv2", p + ".conv2.0", p + ".conv2.1", 1, 1, True)
        self.pp_blocks.append(name)

    def _decode(self, f, N, T, buf, keep, stop_before_tail=False):
        feats = [f[5], f[4], f[3], f[2], f[1]]             # head of the encoder first (decoder.py:158-159)
        dense = {}

        def block(name, x_low, skip):
            p = f"decoder.blocks.{name}"
            H = x_low.shape[1] * 2
            c_out = self.layers[p + ".conv1"].C_out
            t = self._run(p + ".conv1", x_low, N, H, H, skip=skip, out=buf("t_" + name, H, c_out))
            return self._run(p + ".conv2", t, N, H, H, out=buf("o_" + name, H, c_out))

        depth = 4
        for layer_idx in range(4):
            for depth_idx in range(depth - layer_idx):
                if layer_idx == 0:
                    dense[f"x_{depth_idx}_{depth_idx}"] = block(f"x_{depth_idx}_{depth_idx}", feats[depth_idx], feats[depth_idx + 1])
                else:
                    li = depth_idx + layer_idx
                    cat = [dense[f"x_{idx}_{li}"] for idx in range(depth_idx + 1, li + 1)] + [feats[li + 1]]
                    skip = torch.cat(cat, dim=-1)          # NHWC: channel concat in decoder.py:171-176's order
                    dense[f"x_{depth_idx}_{li}"] = block(f"x_{depth_idx}_{li}", dense[f"x_{depth_idx}_{li - 1}"], skip)
        out = block("x_0_4", dense["x_0_3"], None)
        if keep is not None:
            keep.update(dense)
        return out


def conv_flops_per_tile(T: int, in_channels: int = 3, classes: int = 3) -> float:
    """Algorithmic conv FLOPs (2 * MACs) of one T x T tile through Unet-resnet34 incl. the head
    (15.6657 GFLOP at T=256, RGB, 3 classes)."""
    total = 2.0 * (T // 2) ** 2 * 64 * in_channels * 49
    H, cin = T // 4, 64
    for li, (planes, nblk) in enumerate(zip(RESNET34_PLANES, RESNET34_LAYERS), start=1):
        for b in range(nblk):
            if b == 0 and li > 1:
                H //= 2
                total += 2.0 * H * H * planes * cin            # 1x1 downsample
            total += 2.0 * H * H * planes * cin * 9            # conv1
            total += 2.0 * H * H * planes * planes * 9         # conv2
            cin = planes
    H, x_ch = T // 32, 512
    for skip, out in zip((256, 128, 64, 64, 0), (256, 128, 64, 32, 16)):
        H *= 2
        total += 2.0 * H * H * out * (x_ch + skip) * 9
        total += 2.0 * H * H * out * out * 9
        x_ch = out
    return total + 2.0 * T * T * classes * 16 * 9
