"""Tensor-level wrappers over the C-ABI (one Python function per ``dt_*`` entry point).

All tensors are CUDA tensors; every call runs on ``torch.cuda.current_stream()``.  Nothing here
computes on the CPU — a missing library or a non-B200 device raises ``DeadtreesB200Error``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import DT_BF16, DT_F32, ConvDesc, load, ptr, stream_ptr

ACT_DTYPES = {DT_BF16: torch.bfloat16, DT_F32: torch.float32}

# number of kernel launches issued through this module (bench.py reports it as `gpu_launches`);
LAUNCHES = 0
# bench instrumentation: when PROFILE is a dict, conv / gather / stitch launches are bracketed by CUDA
# events on the launching stream and appended as (kind, start, end, algorithmic_work) tuples
PROFILE = None
# bumped whenever a raw-pointer kernel (or a CUDA-graph replay of one) changes parameters or BatchNorm running statistics:
# those writes do not touch torch's per-tensor version counters, so caches of folded / repacked weights
# (``Unet.engine()``) carry this counter in their key
STATE_GENERATION = 0


def bump_state_generation() -> None:
    global STATE_GENERATION
    STATE_GENERATION += 1


class _Timed:
    def __init__(self, kind: str, work: float, tag: str = ""):
        self.kind, self.work, self.tag = kind, work, tag

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None and exc[0] is None:
            self.e1.record()
            PROFILE.setdefault(self.kind, []).append((self.e0, self.e1, self.work, self.tag))
        return False


def check(rc: int) -> None:
    global LAUNCHES
    _lib.check(rc)
    LAUNCHES += 1


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return DT_BF16
    if t.dtype == torch.float32:
        return DT_F32
    raise TypeError(f"activation dtype must be bfloat16 or float32, got {t.dtype}")


def _cuda(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.DeadtreesB200Error(f"{name} must be a CUDA tensor (deadtrees_b200 has no CPU fallback)")
    return t.contiguous()


# ---- tiler -----------------------------------------------------------------------------------

def make_blocks(x: torch.Tensor, d: int) -> torch.Tensor:
    """(p, m, n) -> (m/d * n/d, p, d, d) on the device; any 1/2/4/8-byte dtype."""
    x = _cuda(x, "x")
    p, m, n = x.shape
    if d <= 0 or m % d or n % d:
        raise ValueError(f"tile {(m, n)} not divisible by subtile {d}")
    out = torch.empty(((m // d) * (n // d), p, d, d), dtype=x.dtype, device=x.device)
    check(load().dt_make_blocks(x.data_ptr(), p, m, n, d, x.element_size(), out.data_ptr(), stream_ptr()))
    return out


def unmake_blocks(x: torch.Tensor, d: int, m: int, n: int) -> torch.Tensor:
    """(m/d * n/d, d, d) -> (m, n) on the device."""
    x = _cuda(x, "x")
    if d <= 0 or m % d or n % d or x.numel() != m * n:
        raise ValueError(f"cannot merge {tuple(x.shape)} into {(m, n)} with d={d}")
    out = torch.empty((m, n), dtype=x.dtype, device=x.device)
    check(load().dt_unmake_blocks(x.data_ptr(), d, m, n, x.element_size(), out.data_ptr(), stream_ptr()))
    return out


def tile_gather_normalize(mosaic: torch.Tensor, layout: str, channels: int, tile: int, overlap: int,
                          grid: Tuple[int, int], tile0: int, ntiles: int, offset: Sequence[float],
                          scale: Sequence[float], dtype: torch.dtype = torch.bfloat16,
                          out: Optional[torch.Tensor] = None, pad: int = 0) -> torch.Tensor:
    """uint8 mosaic ("hwc": (H, W, C) or "chw": (C, H, W)) -> normalised NHWC tiles (ntiles, T, T, 4);
    with pad=3 into a caller-zeroed (ntiles, T+6, T+8, 4) frame at offset (3, 3)."""
    if mosaic.dtype != torch.uint8 or not mosaic.is_cuda:
        raise TypeError("mosaic must be a CUDA uint8 tensor")
    if layout == "hwc":
        H, W, Csrc = mosaic.shape
        rs, ps, cs = mosaic.stride(0), mosaic.stride(1), mosaic.stride(2)
    elif layout == "chw":
        Csrc, H, W = mosaic.shape
        cs, rs, ps = mosaic.stride(0), mosaic.stride(1), mosaic.stride(2)
    else:
        raise ValueError("layout must be 'hwc' or 'chw'")
    if channels > Csrc:
        raise ValueError(f"model wants {channels} channels, mosaic has {Csrc}")
    if out is None:
        out = (torch.zeros((ntiles, tile + 6, tile + 8, 4), dtype=dtype, device=mosaic.device) if pad else
               torch.empty((ntiles, tile, tile, 4), dtype=dtype, device=mosaic.device))
    off = (C.c_float * 4)(*[float(v) for v in list(offset)[:channels]] + [0.0] * (4 - channels))
    sc = (C.c_float * 4)(*[float(v) for v in list(scale)[:channels]] + [0.0] * (4 - channels))
    # algorithmic bytes: N*T^2*C*(1 B in + elem out)
    with _Timed("gather", float(ntiles) * tile * tile * channels * (1 + out.element_size())):
        check(load().dt_tile_gather_normalize(mosaic.data_ptr(), H, W, channels, rs, ps, cs, tile, tile - overlap,
                                              grid[1], tile0, ntiles, off, sc, 4, _dt(out), pad,
                                              out.data_ptr(), stream_ptr()))
    return out


def pack_input_nchw(x: torch.Tensor, channels: int, dtype: torch.dtype) -> torch.Tensor:
    """(N, C_src, H, W) fp32 -> (N, H, W, 4) ``dtype`` keeping the first ``channels`` channels."""
    x = _cuda(x, "x")
    if x.dtype != torch.float32:
        x = x.float()
    N, Csrc, H, W = x.shape
    out = torch.empty((N, H, W, 4), dtype=dtype, device=x.device)
    check(load().dt_pack_input_nchw(x.data_ptr(), N, Csrc, channels, H, W, _dt(out), out.data_ptr(), stream_ptr()))
    return out


def pack_input_nchw_frame(x: torch.Tensor, channels: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(N, C_src, H, W) fp32 -> the stem's zero-bordered bf16 frame (N, H+6, W+8, 4), interior at (3, 3)."""
    x = _cuda(x, "x")
    if x.dtype != torch.float32:
        x = x.float()
    N, Csrc, H, W = x.shape
    if out is None:
        out = torch.empty((N, H + 6, W + 8, 4), dtype=torch.bfloat16, device=x.device)
    check(load().dt_pack_input_nchw_frame(x.data_ptr(), N, Csrc, channels, H, W, out.data_ptr(), stream_ptr()))
    return out


def stitch_mask(tile_masks: torch.Tensor, grid_x: int, tile0: int, mosaic_mask: torch.Tensor) -> None:
    """overlap-0 stitch of (ntiles, T, T) uint8 class ids into the (H, W) uint8 mosaic mask."""
    tile_masks = _cuda(tile_masks, "tile_masks")
    ntiles, T, _ = tile_masks.shape
    H, W = mosaic_mask.shape
    with _Timed("stitch", 2.0 * ntiles * T * T):  # 2 B / pixel
        check(load().dt_stitch_mask_u8(tile_masks.data_ptr(), T, grid_x, tile0, ntiles, mosaic_mask.data_ptr(), H,
                                       W, mosaic_mask.stride(0), stream_ptr()))


def stitch_blend_argmax(logits: torch.Tensor, overlap: int, grid: Tuple[int, int], win: torch.Tensor,
                        mosaic_mask: torch.Tensor, blended: Optional[torch.Tensor] = None, row0: int = 0,
                        nrows: Optional[int] = None, ty_base: int = 0) -> None:
    """(gy*gx, T, T, K) logits -> blended argmax over mosaic rows [row0, row0+nrows)."""
    logits = _cuda(logits, "logits")
    _, T, _, K = logits.shape
    H, W = mosaic_mask.shape
    nrows = H - row0 if nrows is None else nrows
    # algorithmic bytes of THIS call (the mask may be stitched band by band), SURVEY.md §8d: every logit that covers a
    # written mosaic row once (gather form: a row under a horizontal overlap reads two tile rows, each gx * T pixels wide)
    # + 1 B of mask per pixel; summed over the bands of a mosaic this is N_tiles * T^2 * K * elem + H * W
    step = T - overlap
    covered = 0
    if PROFILE is not None:
        for ty in range(grid[0]):
            covered += max(0, min(row0 + nrows, ty * step + T) - max(row0, ty * step))
    with _Timed("stitch", float(covered) * grid[1] * T * K * logits.element_size() + float(nrows) * W):
        check(load().dt_stitch_blend_argmax(logits.data_ptr(), _dt(logits), K, T, overlap, grid[0], grid[1],
                                            ty_base, win.data_ptr(), mosaic_mask.data_ptr(), ptr(blended), H, W,
                                            row0, nrows, stream_ptr()))


# ---- network ---------------------------------------------------------------------------------

def conv2d(x: torch.Tensor, w: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, *, N: int, H: int, W: int,
           C_in: int, C_x: int, C_out: int, R: int, S: int, stride: int, pad: int, relu: bool,
           skip: Optional[torch.Tensor] = None, upsample: bool = False, residual: Optional[torch.Tensor] = None,
           out: Optional[torch.Tensor] = None, flags: int = 0, algo_cin: Optional[int] = None, tag: str = "") -> torch.Tensor:
    Ho = (H + 2 * pad - R) // stride + 1
    Wo = (W + 2 * pad - S) // stride + 1
    if flags & _lib.CONV_TRANSPOSED:      # data gradient of a stride-2 conv: H, W are the output size, x is gy (N, Ho, Wo, C_in)
        Hs, Ws, Ho, Wo = Ho, Wo, H, W
    else:
        Hs, Ws = Ho, Wo
    if out is None:
        out = torch.empty((N, Ho, Wo, C_out), dtype=x.dtype, device=x.device)
    d = ConvDesc(N, H, W, C_in, C_x, int(upsample), C_out, R, S, stride, pad, int(relu), int(residual is not None),
                 _dt(x), flags)
    with _Timed("conv", 2.0 * N * Hs * Ws * C_out * (algo_cin or C_in) * R * S, tag):
        check(load().dt_conv2d_fwd(C.byref(d), x.data_ptr(), ptr(skip), w.data_ptr(), scale.data_ptr(),
                                   shift.data_ptr(), ptr(residual), out.data_ptr(), stream_ptr()))
    return out


def stem_pool(x: torch.Tensor, w: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, *, N: int, H: int, W: int,
              out: torch.Tensor, pooled: torch.Tensor, algo_cin: int = 4, tag: str = "stem+pool") -> None:
    """7x7/s2 stem (BN + ReLU) and the 3x3/s2 max pooling behind it in one launch (conv_stem.cu, POOL variant); x is the
    zero-bordered bf16 frame (N, H+6, W+8, 4), W / 2 must be 128.  Bit-identical to ``conv2d`` + :func:`maxpool3x3s2`."""
    d = ConvDesc(N, H, W, 4, 4, 0, 64, 7, 7, 2, 3, 1, 0, _dt(x), _lib.CONV_X_PAD3)
    with _Timed("conv", 2.0 * N * (H // 2) * (W // 2) * 64 * algo_cin * 49, tag):
        check(load().dt_stem_pool_fwd(C.byref(d), x.data_ptr(), w.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                      out.data_ptr(), pooled.data_ptr(), stream_ptr()))


def maxpool3x3s2(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    N, H, W, Cc = x.shape
    if out is None:
        out = torch.empty((N, (H - 1) // 2 + 1, (W - 1) // 2 + 1, Cc), dtype=x.dtype, device=x.device)
    check(load().dt_maxpool3x3s2(x.data_ptr(), N, H, W, Cc, _dt(x), out.data_ptr(), stream_ptr()))
    return out


def head(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, *, logits_nchw: Optional[torch.Tensor] = None,
         logits_nhwc: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None) -> None:
    N, H, W, Cc = x.shape
    K = bias.numel()
    check(load().dt_head_fwd(x.data_ptr(), _dt(x), N, H, W, Cc, K, w.data_ptr(), bias.data_ptr(), ptr(logits_nchw),
                             ptr(logits_nhwc), ptr(mask), stream_ptr()))


def head_tc(x: torch.Tensor, w_packed: torch.Tensor, bias16: torch.Tensor, K: int, *,
            logits_nchw: Optional[torch.Tensor] = None, logits_nhwc: Optional[torch.Tensor] = None,
            mask: Optional[torch.Tensor] = None) -> None:
    """tensor-core head (bf16 activations, 16 input channels)."""
    N, H, W, Cc = x.shape
    check(load().dt_head_fwd_tc(x.data_ptr(), N, H, W, K, w_packed.data_ptr(), bias16.data_ptr(), ptr(logits_nchw),
                                ptr(logits_nhwc), ptr(mask), stream_ptr()))


def tail_fused(x: torch.Tensor, w2_packed: torch.Tensor, scale2: torch.Tensor, shift2: torch.Tensor, wh_packed: torch.Tensor,
               bias16: torch.Tensor, K: int, *, logits_nchw: Optional[torch.Tensor] = None,
               logits_nhwc: Optional[torch.Tensor] = None, mask: Optional[torch.Tensor] = None,
               tag: str = "tail") -> None:
    """decoder.blocks.4.conv2 (16 -> 16, BN + ReLU) + segmentation head in one launch (conv_tail.cu); x = the output of
    decoder.blocks.4.conv1, bf16 NHWC, width 128 or 256.  Bit-identical to ``conv2d`` followed by :func:`head_tc`."""
    N, H, W, Cc = x.shape
    if Cc != 16 or x.dtype != torch.bfloat16:
        raise ValueError("tail_fused: x must be bf16 NHWC with 16 channels")
    with _Timed("conv", 2.0 * N * H * W * 16 * 9 * (16 + K), tag):
        check(load().dt_tail_fused(x.data_ptr(), N, H, W, K, w2_packed.data_ptr(), scale2.data_ptr(), shift2.data_ptr(),
                                   wh_packed.data_ptr(), bias16.data_ptr(), ptr(logits_nchw), ptr(logits_nhwc), ptr(mask),
                                   stream_ptr()))


def argmax_nchw(logits: torch.Tensor) -> torch.Tensor:
    logits = _cuda(logits, "logits")
    N, K, H, W = logits.shape
    out = torch.empty((N, H, W), dtype=torch.uint8, device=logits.device)
    check(load().dt_argmax_nchw(logits.data_ptr(), N, K, H, W, out.data_ptr(), stream_ptr()))
    return out


def mode_vote(masks: torch.Tensor, out_int64: bool = True) -> torch.Tensor:
    """(M, ...) uint8 class-id masks of M models -> per-pixel majority class (smallest on ties, as torch.mode)."""
    masks = _cuda(masks, "masks")
    if masks.dtype != torch.uint8:
        raise TypeError("masks must be uint8")
    M = masks.shape[0]
    n = masks[0].numel()
    out = torch.empty(masks.shape[1:], dtype=torch.int64 if out_int64 else torch.uint8, device=masks.device)
    check(load().dt_mode_vote(masks.data_ptr(), M, n, int(out_int64), out.data_ptr(), stream_ptr()))
    return out


# ---- loss / optimizer ------------------------------------------------------------------------

def seg_loss_partials(logits: torch.Tensor, labels: torch.Tensor):
    """-> (sums double (N, K, 4), counts int64 (K, 3), bad_label int32 (1,))."""
    logits, labels = _cuda(logits, "logits"), _cuda(labels, "labels")
    N, K, H, W = logits.shape
    sums = torch.zeros((N, K, 4), dtype=torch.float64, device=logits.device)
    counts = torch.zeros((K, 3), dtype=torch.int64, device=logits.device)
    bad = torch.zeros((1,), dtype=torch.int32, device=logits.device)
    check(load().dt_seg_loss_partials(logits.data_ptr(), labels.data_ptr(), N, K, H, W, sums.data_ptr(),
                                      counts.data_ptr(), bad.data_ptr(), stream_ptr()))
    return sums, counts, bad


def seg_loss_finalize(sums: torch.Tensor, counts: torch.Tensor, dice_mode: int, use_focal: bool):
    """-> (out float (8,), coef float (N, K, 2), focal_scale float (1,))."""
    N, K, _ = sums.shape
    out = torch.empty((8,), dtype=torch.float32, device=sums.device)
    coef = torch.empty((N, K, 2), dtype=torch.float32, device=sums.device)
    fs = torch.empty((1,), dtype=torch.float32, device=sums.device)
    check(load().dt_seg_loss_finalize(sums.data_ptr(), counts.data_ptr(), N, K, dice_mode, int(use_focal),
                                      out.data_ptr(), coef.data_ptr(), fs.data_ptr(), stream_ptr()))
    return out, coef, fs


def seg_loss_backward(logits: torch.Tensor, labels: torch.Tensor, coef: torch.Tensor, focal_scale: torch.Tensor,
                      upstream: float = 1.0) -> torch.Tensor:
    N, K, H, W = logits.shape
    grad = torch.empty_like(logits)
    check(load().dt_seg_loss_backward(logits.data_ptr(), labels.data_ptr(), N, K, H, W, coef.data_ptr(),
                                      focal_scale.data_ptr(), upstream, grad.data_ptr(), stream_ptr()))
    return grad


def _idc_mask(idc, K: int) -> int:
    m = 0
    for k in idc:
        if not 0 <= int(k) < K:
            raise ValueError(f"class {k} outside [0, {K})")
        m |= 1 << int(k)
    return m


def gwdl_loss(logits: torch.Tensor, labels: torch.Tensor, dist_matrix, softmax_twice: bool = True):
    """Generalized Wasserstein Dice loss on the logits -> (loss 0-dim fp32, coef (2N + HW,) for gwdl_loss_backward)."""
    logits, labels = _cuda(logits, "logits").contiguous(), _cuda(labels, "labels").contiguous()
    if labels.dtype != torch.int64:
        labels = labels.long()
    N, K, H, W = logits.shape
    flat = [float(v) for row in dist_matrix for v in row]
    if len(flat) != K * K:
        raise ValueError(f"class distance matrix must be {K} x {K}")
    M = (C.c_float * (K * K))(*flat)
    lib = load()
    need = lib.dt_gwdl_workspace(N, H, W)
    ws = torch.empty((need + 7) // 8, dtype=torch.float64, device=logits.device)
    loss = torch.empty((), dtype=torch.float32, device=logits.device)
    coef = torch.empty(2 * N + H * W, dtype=torch.float32, device=logits.device)
    check(lib.dt_gwdl_loss(logits.data_ptr(), labels.data_ptr(), N, K, H, W, M, int(softmax_twice), ws.data_ptr(), need,
                           loss.data_ptr(), coef.data_ptr(), stream_ptr()))
    return loss, coef


def gwdl_loss_backward(logits: torch.Tensor, labels: torch.Tensor, dist_matrix, coef: torch.Tensor, weight: float,
                       grad_logits: torch.Tensor, softmax_twice: bool = True) -> None:
    """grad_logits += weight * d(gwdl_loss)/d(logits)"""
    N, K, H, W = logits.shape
    M = (C.c_float * (K * K))(*[float(v) for row in dist_matrix for v in row])
    if labels.dtype != torch.int64:
        labels = labels.long()
    check(load().dt_gwdl_loss_backward(logits.data_ptr(), labels.contiguous().data_ptr(), N, K, H, W, M, int(softmax_twice),
                                       coef.data_ptr(), float(weight), grad_logits.data_ptr(), stream_ptr()))


def train_transform(images: torch.Tensor, masks, lus, geom, bc, offset, scale, out_channels: int, merge_classes: bool = False):
    """``dt_train_transform``: (N, H, W, C) uint8 images (+ (N, H, W) uint8 masks / lus) and per-sample draws ``geom`` (N, 2)
    {flip, rot}, ``bc`` (N, 2) {alpha, beta} -> (fp32 (N, out_channels, H, W), int64 masks, int64 lus)."""
    images = _cuda(images, "images").contiguous()
    if images.dtype != torch.uint8 or images.dim() != 4:
        raise TypeError("train_transform expects (N, H, W, C) uint8 images")
    N, H, W, Cc = images.shape
    dev = images.device

    def u8(t, name):
        if t is None:
            return None
        t = _cuda(t, name)
        if t.dtype != torch.uint8:
            t = t.to(torch.uint8)
        if tuple(t.shape) != (N, H, W):
            raise ValueError(f"{name} must be (N, H, W) = {(N, H, W)}, got {tuple(t.shape)}")
        return t.contiguous()

    masks, lus = u8(masks, "masks"), u8(lus, "lus")
    geom_t = torch.as_tensor(np.ascontiguousarray(geom, dtype=np.int32).reshape(N, 2)).to(dev, non_blocking=True)
    bc_t = torch.as_tensor(np.ascontiguousarray(bc, dtype=np.float64).reshape(N, 2)).to(dev, non_blocking=True)
    off = (C.c_float * 4)(*([float(v) for v in list(offset)[:out_channels]] + [0.0] * (4 - out_channels)))
    sc = (C.c_float * 4)(*([float(v) for v in list(scale)[:out_channels]] + [0.0] * (4 - out_channels)))
    sums = torch.empty((N,), dtype=torch.int64, device=dev)
    out = torch.empty((N, out_channels, H, W), dtype=torch.float32, device=dev)
    om = torch.empty((N, H, W), dtype=torch.int64, device=dev) if masks is not None else None
    ol = torch.empty((N, H, W), dtype=torch.int64, device=dev) if lus is not None else None
    check(load().dt_train_transform(images.data_ptr(), masks.data_ptr() if masks is not None else None,
                                    lus.data_ptr() if lus is not None else None, N, H, W, Cc, out_channels,
                                    geom_t.data_ptr(), bc_t.data_ptr(), off, sc, int(merge_classes), sums.data_ptr(),
                                    out.data_ptr(), om.data_ptr() if om is not None else None,
                                    ol.data_ptr() if ol is not None else None, stream_ptr()))
    return out, om, ol


def confusion_matrix(pred: torch.Tensor, target: torch.Tensor, K: int, lu: torch.Tensor = None,
                     counts: torch.Tensor = None, bad: torch.Tensor = None):
    """adds the (target, prediction) pairs to ``counts`` int64 (2, K, K): [0] all pixels, [1] the pixels with ``lu == 1``
    (rows = target, as torchmetrics' ``confusion_matrix``) -> ``(counts, bad)``; ``bad`` int32 (1,) becomes 1 when a value lies
    outside [0, K).  ``counts=None`` / ``bad=None`` start new accumulators; pass the returned ones back to go on."""
    pred, target = _cuda(pred, "pred").contiguous(), _cuda(target, "target").contiguous()
    if pred.dtype not in (torch.uint8, torch.int64):
        pred = pred.long()
    if target.dtype != torch.int64:
        target = target.long()
    if pred.numel() != target.numel():
        raise ValueError(f"prediction {tuple(pred.shape)} and target {tuple(target.shape)} differ in size")
    lu_ptr, lu_elem = None, 0
    if lu is not None:
        lu = _cuda(lu, "lu").contiguous()
        if lu.dtype == torch.bool:
            lu = lu.view(torch.uint8)
        elif lu.dtype not in (torch.uint8, torch.int32, torch.int64):
            lu = (lu == 1).view(torch.uint8)
        if lu.numel() != target.numel():
            raise ValueError(f"lu {tuple(lu.shape)} and target {tuple(target.shape)} differ in size")
        lu_ptr, lu_elem = lu.data_ptr(), lu.element_size()
    if counts is None:
        counts = torch.zeros((2, K, K), dtype=torch.int64, device=pred.device)
    if bad is None:
        bad = torch.zeros((1,), dtype=torch.int32, device=pred.device)
    check(load().dt_confusion_matrix(pred.data_ptr(), pred.element_size(), target.data_ptr(), lu_ptr, lu_elem, pred.numel(), K,
                                     counts.data_ptr(), bad.data_ptr(), stream_ptr()))
    return counts, bad


def one_hot2dist(labels: torch.Tensor, K: int, truncate: bool = True) -> torch.Tensor:
    """(N, H, W) int64 labels -> (N, K, H, W) fp32 signed distance maps of the boundary loss (the dataloader's
    one_hot2dist per sample); ``truncate``: integer truncation of the reference's int32 path."""
    labels = _cuda(labels, "labels")
    if labels.dtype != torch.int64:
        labels = labels.long()
    labels = labels.contiguous()
    N, H, W = labels.shape
    lib = load()
    need = lib.dt_one_hot2dist_workspace(N, K, H, W)
    ws = torch.empty((need + 3) // 4, dtype=torch.int32, device=labels.device)
    out = torch.empty((N, K, H, W), dtype=torch.float32, device=labels.device)
    check(lib.dt_one_hot2dist(labels.data_ptr(), N, K, H, W, int(truncate), out.data_ptr(), ws.data_ptr(), need, stream_ptr()))
    return out


def boundary_loss(logits: torch.Tensor, dist: torch.Tensor, idc) -> torch.Tensor:
    """mean over (b, k in idc, h, w) of softmax(logits)_k * dist_k -> 0-dim fp32 tensor."""
    logits, dist = _cuda(logits, "logits"), _cuda(dist, "dist")
    if dist.dtype != torch.float32:
        dist = dist.float()
    N, K, H, W = logits.shape
    if tuple(dist.shape) != (N, K, H, W):
        raise ValueError(f"distance maps {tuple(dist.shape)} do not match the logits {tuple(logits.shape)}")
    ws = torch.empty(64 * N, dtype=torch.float64, device=logits.device)
    out = torch.empty((), dtype=torch.float32, device=logits.device)
    check(load().dt_boundary_loss(logits.data_ptr(), dist.data_ptr(), N, K, H, W, _idc_mask(idc, K), ws.data_ptr(),
                                  out.data_ptr(), stream_ptr()))
    return out


def boundary_loss_backward(logits: torch.Tensor, dist: torch.Tensor, idc, weight: float, grad_logits: torch.Tensor) -> None:
    """grad_logits += weight * d(boundary_loss)/d(logits)"""
    if dist.dtype != torch.float32:
        dist = dist.float()
    N, K, H, W = logits.shape
    check(load().dt_boundary_loss_backward(logits.data_ptr(), dist.contiguous().data_ptr(), N, K, H, W, _idc_mask(idc, K),
                                           float(weight), grad_logits.data_ptr(), stream_ptr()))


def class2one_hot(labels: torch.Tensor, K: int):
    """-> (int32 one-hot (N, K, H, W), bad_label int32 (1,))."""
    labels = _cuda(labels, "labels")
    if labels.dtype != torch.int64:
        labels = labels.long()
    N, H, W = labels.shape
    out = torch.empty((N, K, H, W), dtype=torch.int32, device=labels.device)
    bad = torch.zeros((1,), dtype=torch.int32, device=labels.device)
    check(load().dt_class2one_hot(labels.data_ptr(), N, K, H, W, out.data_ptr(), bad.data_ptr(), stream_ptr()))
    return out, bad


def softmax_nchw(logits: torch.Tensor) -> torch.Tensor:
    logits = _cuda(logits, "logits")
    N, K, H, W = logits.shape
    out = torch.empty_like(logits)
    check(load().dt_softmax_nchw(logits.data_ptr(), N, K, H, W, out.data_ptr(), stream_ptr()))
    return out


def prob_loss_partials(probs: torch.Tensor, target: torch.Tensor, gamma: float = 2.0) -> torch.Tensor:
    """probs fp32 (N, K, H, W), target int32 one-hot or float map -> double sums (N, K, 4)."""
    probs, target = _cuda(probs, "probs"), _cuda(target, "target")
    if probs.dtype != torch.float32:
        probs = probs.float()
    if target.dtype == torch.float32:
        is_float = 1
    else:
        is_float = 0
        if target.dtype != torch.int32:
            target = target.to(torch.int32)
    N, K, H, W = probs.shape
    sums = torch.zeros((N, K, 4), dtype=torch.float64, device=probs.device)
    check(load().dt_prob_loss_partials(probs.data_ptr(), target.data_ptr(), is_float, N, K, H, W, float(gamma),
                                       sums.data_ptr(), stream_ptr()))
    return sums


def adam_step_dev(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, state: torch.Tensor, *, beta1: float,
                  beta2: float, eps: float, sumsq_acc: Optional[torch.Tensor], max_norm: float,
                  loss: Optional[torch.Tensor], scratch: torch.Tensor) -> None:
    """graph-capturable Adam: state = float (2,) {step, lr} on the device; skipped when `loss` is not finite."""
    bump_state_generation()
    check(load().dt_adam_step_dev(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), state.data_ptr(),
                                  beta1, beta2, eps, ptr(sumsq_acc), max_norm, ptr(loss), scratch.data_ptr(), stream_ptr()))


def sumsq(g: torch.Tensor, acc: torch.Tensor) -> None:
    check(load().dt_sumsq(g.data_ptr(), g.numel(), acc.data_ptr(), stream_ptr()))


def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, *, lr: float, beta1: float,
              beta2: float, eps: float, step: int, sumsq_acc: Optional[torch.Tensor], max_norm: float) -> None:
    bump_state_generation()
    check(load().dt_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1, beta2,
                              eps, step, ptr(sumsq_acc), max_norm, stream_ptr()))


# ---- training step (train-mode BatchNorm, backward) -----------------------------------------------

_WS = {}


def _reduce_ws(device) -> torch.Tensor:
    """workspace for the two-stage per-channel reductions (dt_bn_train_stats / dt_bn_train_bwd)."""
    ws = _WS.get(device)
    if ws is None:
        ws = _WS[device] = torch.empty(2 * 148 * 8 * 1024 + 2048, dtype=torch.float32, device=device)
    return ws


def bn_train_stats(y: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, running_mean: Optional[torch.Tensor],
                   running_var: Optional[torch.Tensor], eps: float = 1e-5, momentum: float = 0.1):
    """y (N, H, W, C) raw conv output -> (scale, shift, mean, invstd) float (C,); updates the running stats."""
    Cc = y.shape[-1]
    M = y.numel() // Cc
    scale, shift, mean, invstd = (torch.empty(Cc, dtype=torch.float32, device=y.device) for _ in range(4))
    if running_mean is not None or running_var is not None:
        bump_state_generation()
    check(load().dt_bn_train_stats(y.data_ptr(), M, Cc, _dt(y), gamma.data_ptr(), beta.data_ptr(), eps, momentum,
                                   ptr(running_mean), ptr(running_var), scale.data_ptr(), shift.data_ptr(),
                                   mean.data_ptr(), invstd.data_ptr(), _reduce_ws(y.device).data_ptr(), stream_ptr()))
    return scale, shift, mean, invstd


def bn_apply(y: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor, residual: Optional[torch.Tensor] = None,
             relu: bool = True, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    Cc = y.shape[-1]
    if out is None:
        out = torch.empty_like(y)
    check(load().dt_bn_apply(y.data_ptr(), y.numel() // Cc, Cc, _dt(y), scale.data_ptr(), shift.data_ptr(),
                             ptr(residual), int(relu), out.data_ptr(), stream_ptr()))
    return out


def bn_train_bwd(g: torch.Tensor, a: Optional[torch.Tensor], y: torch.Tensor, mean: torch.Tensor, invstd: torch.Tensor,
                 scale: torch.Tensor, want_gz: bool = False, dgamma: Optional[torch.Tensor] = None,
                 dbeta: Optional[torch.Tensor] = None, relu_shift: Optional[torch.Tensor] = None):
    """-> (gy, gz or None, dgamma, dbeta).  ``relu_shift`` (with ``a=None``): the layer was relu(y*scale + shift) without a
    residual - the ReLU mask is recomputed from y instead of reading the activation."""
    Cc = y.shape[-1]
    gy = torch.empty_like(y)
    gz = torch.empty_like(y) if want_gz else None
    if dgamma is None:
        dgamma = torch.empty(Cc, dtype=torch.float32, device=y.device)
    if dbeta is None:
        dbeta = torch.empty(Cc, dtype=torch.float32, device=y.device)
    if relu_shift is not None:
        if a is not None:
            raise ValueError("relu_shift replaces the activation tensor: pass a=None")
        check(load().dt_bn_train_bwd_relu(g.data_ptr(), y.data_ptr(), y.numel() // Cc, Cc, _dt(y), mean.data_ptr(),
                                          invstd.data_ptr(), scale.data_ptr(), relu_shift.data_ptr(), dgamma.data_ptr(),
                                          dbeta.data_ptr(), gy.data_ptr(), ptr(gz), _reduce_ws(y.device).data_ptr(),
                                          stream_ptr()))
        return gy, gz, dgamma, dbeta
    check(load().dt_bn_train_bwd(g.data_ptr(), ptr(a), y.data_ptr(), y.numel() // Cc, Cc, _dt(y), mean.data_ptr(),
                                 invstd.data_ptr(), scale.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(),
                                 gy.data_ptr(), ptr(gz), _reduce_ws(y.device).data_ptr(), stream_ptr()))
    return gy, gz, dgamma, dbeta


def add(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    if out is None:
        out = torch.empty_like(a)
    check(load().dt_add(a.data_ptr(), b.data_ptr(), a.numel(), _dt(a), out.data_ptr(), stream_ptr()))
    return out


def maxpool3x3s2_bwd(x: torch.Tensor, gout: torch.Tensor, addend: Optional[torch.Tensor] = None) -> torch.Tensor:
    N, H, W, Cc = x.shape
    gx = torch.empty_like(x)
    check(load().dt_maxpool3x3s2_bwd(x.data_ptr(), gout.data_ptr(), ptr(addend), N, H, W, Cc, _dt(x), gx.data_ptr(),
                                     stream_ptr()))
    return gx


def maxpool3x3s2_idx(x: torch.Tensor):
    """training-step maxpool: (y, idx) with idx (N, Ho, Wo, C) uint8 = window position of the first maximum."""
    N, H, W, Cc = x.shape
    y = torch.empty((N, (H - 1) // 2 + 1, (W - 1) // 2 + 1, Cc), dtype=x.dtype, device=x.device)
    idx = torch.empty(y.shape, dtype=torch.uint8, device=x.device)
    check(load().dt_maxpool3x3s2_idx(x.data_ptr(), N, H, W, Cc, _dt(x), y.data_ptr(), idx.data_ptr(), stream_ptr()))
    return y, idx


def maxpool3x3s2_bwd_idx(idx: torch.Tensor, gout: torch.Tensor, x_shape, addend: Optional[torch.Tensor] = None) -> torch.Tensor:
    N, H, W, Cc = x_shape
    gx = torch.empty(tuple(x_shape), dtype=gout.dtype, device=gout.device)
    check(load().dt_maxpool3x3s2_bwd_idx(idx.data_ptr(), gout.data_ptr(), ptr(addend), N, H, W, Cc, _dt(gout), gx.data_ptr(),
                                         stream_ptr()))
    return gx


def zero_insert2x(gy: torch.Tensor) -> torch.Tensor:
    """(N, Ho, Wo, C) -> (N, 2Ho, 2Wo, C) with gy at the even positions and zeros elsewhere."""
    N, Ho, Wo, Cc = gy.shape
    out = torch.empty((N, 2 * Ho, 2 * Wo, Cc), dtype=gy.dtype, device=gy.device)
    check(load().dt_zero_insert2x(gy.data_ptr(), N, Ho, Wo, Cc, _dt(gy), out.data_ptr(), stream_ptr()))
    return out


def upsample_concat(x_low: torch.Tensor, skip: Optional[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    N, Hl, Wl, Cx = x_low.shape
    Cs = 0 if skip is None else skip.shape[-1]
    if out is None:
        out = torch.empty((N, 2 * Hl, 2 * Wl, Cx + Cs), dtype=x_low.dtype, device=x_low.device)
    check(load().dt_upsample_concat(x_low.data_ptr(), ptr(skip), N, 2 * Hl, 2 * Wl, Cx, Cs, _dt(x_low), out.data_ptr(),
                                    stream_ptr()))
    return out


def upsample_concat_bwd(g_cat: torch.Tensor, Cx: int):
    N, H, W, Cc = g_cat.shape
    Cs = Cc - Cx
    g_low = torch.empty((N, H // 2, W // 2, Cx), dtype=g_cat.dtype, device=g_cat.device)
    g_skip = torch.empty((N, H, W, Cs), dtype=g_cat.dtype, device=g_cat.device) if Cs else None
    check(load().dt_upsample_concat_bwd(g_cat.data_ptr(), N, H, W, Cx, Cs, _dt(g_cat), g_low.data_ptr(), ptr(g_skip),
                                        stream_ptr()))
    return g_low, g_skip


def nchw_to_nhwc(x: torch.Tensor, Kp: int, dtype: torch.dtype) -> torch.Tensor:
    N, K, H, W = x.shape
    out = torch.empty((N, H, W, Kp), dtype=dtype, device=x.device)
    check(load().dt_nchw_to_nhwc(x.data_ptr(), N, K, H, W, Kp, _dt(out), out.data_ptr(), stream_ptr()))
    return out


def channel_sum(g: torch.Tensor, K: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sum over all pixels of the first K channels of an NHWC tensor -> float (K,); two fixed-order stages, so the result
    is the same bit pattern on every run."""
    Cc = g.shape[-1]
    if out is None:
        out = torch.empty(K, dtype=torch.float32, device=g.device)
    check(load().dt_channel_sum(g.data_ptr(), g.numel() // Cc, Cc, K, _dt(g), _reduce_ws(g.device).data_ptr(), out.data_ptr(),
                                stream_ptr()))
    return out


def pack_conv_weight(w: torch.Tensor, mode: int, out: Optional[torch.Tensor] = None,
                     cout_pad: Optional[int] = None) -> torch.Tensor:
    """fp32 OIHW master weights -> kernel layout (see dt_pack_conv_weight): 0 direct fp32, 1 tcgen05 forward,
    2 tcgen05 stem, 3 dgrad-as-forward (stride 1), 4 dgrad of a stride-2 conv (DT_CONV_TRANSPOSED).
    cout_pad: channel stride of the gradient tensor for modes 3 / 4 (>= C_out)."""
    C_out, C_in, R, S = w.shape
    cinp, kpad = C_in, 0
    if mode == 0:
        cinp = 4 if (R == 7 and C_in < 4) else C_in
        shape, dtype = (R * S, cinp, C_out), torch.float32
    elif mode == 1:
        kpad = (R * S * C_in + 63) // 64 * 64
        shape, dtype = (C_out, kpad), torch.bfloat16
    elif mode == 2:
        kpad = 256
        shape, dtype = (C_out, kpad), torch.bfloat16
    else:
        cinp = C_out if cout_pad is None else cout_pad
        kpad = (R * S * cinp + 63) // 64 * 64
        shape, dtype = (C_in, kpad), torch.bfloat16
    if out is None:
        out = torch.empty(shape, dtype=dtype, device=w.device)
    check(load().dt_pack_conv_weight(w.data_ptr(), C_out, C_in, R, S, mode, cinp, kpad, out.data_ptr(), stream_ptr()))
    return out


def pack_conv_weight_geometry(w_shape, mode: int, cout_pad: Optional[int] = None):
    """(C_in_p, Kpad, output shape, dtype) of dt_pack_conv_weight for OIHW weights of shape w_shape."""
    C_out, C_in, R, S = w_shape
    cinp, kpad = C_in, 0
    if mode == 0:
        cinp = 4 if (R == 7 and C_in < 4) else C_in
        return cinp, kpad, (R * S, cinp, C_out), torch.float32
    if mode == 1:
        kpad = (R * S * C_in + 63) // 64 * 64
        return cinp, kpad, (C_out, kpad), torch.bfloat16
    if mode == 2:
        return cinp, 256, (C_out, 256), torch.bfloat16
    cinp = C_out if cout_pad is None else cout_pad
    kpad = (R * S * cinp + 63) // 64 * 64
    return cinp, kpad, (C_in, kpad), torch.bfloat16


class WeightPacker:
    """kernel-layout copies of the fp32 master weights of many layers, refreshed by ONE launch
    (dt_pack_conv_weights_batched).  ``get`` registers a layout on first use (packing it at once);
    ``refresh`` repacks every registered layout from the current master weights."""

    def __init__(self, device):
        self.device = device
        self.entries = {}           # key -> (w, out, mode, cinp, kpad)
        self._table = None

    def get(self, key, w: torch.Tensor, mode: int, cout_pad: Optional[int] = None, rows_pad: Optional[int] = None) -> torch.Tensor:
        """``rows_pad`` (mode 1): the packed matrix gets this many rows, the rows past C_out stay zero (the tensor-core
        head takes 16 rows for its K <= 4 classes)."""
        ent = self.entries.get(key)
        if ent is not None and ent[6] == w.data_ptr():          # the address registered in the job table, not `ent[0]` (== w)
            return ent[5]
        cinp, kpad, shape, dtype = pack_conv_weight_geometry(w.shape, mode, cout_pad)
        if rows_pad is not None and mode == 1 and rows_pad > shape[0]:
            full = torch.zeros((rows_pad, shape[1]), dtype=dtype, device=w.device)
            out = pack_conv_weight(w, mode, out=full[: shape[0]], cout_pad=cout_pad)
        else:
            full = out = pack_conv_weight(w, mode, cout_pad=cout_pad)
        self.entries[key] = (w, out, mode, cinp, kpad, full, w.data_ptr())
        self._table = None
        return full

    def clear(self) -> None:
        """forget every registered layout (the master weights moved, e.g. ``UnetTrainEngine.flatten_parameters``)."""
        self.entries.clear()
        self._table = None

    def refresh(self) -> None:
        if not self.entries:
            return
        if self._table is None:
            jobs = (_lib.PackJob * len(self.entries))()
            start = 0
            for j, (w, out, mode, cinp, kpad, _, wptr) in enumerate(self.entries.values()):
                if w.data_ptr() != wptr:
                    raise _lib.DeadtreesB200Error("a master weight moved after its kernel layout was registered; call "
                                                  "WeightPacker.clear() (flatten_parameters does) before the next forward")
                C_out, C_in, R, S = w.shape
                jobs[j] = _lib.PackJob(w.data_ptr(), out.data_ptr(), C_out, C_in, R, S, mode, cinp, kpad, 0, start)
                start += out.numel()
            raw = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8)
            self._table = (raw.to(self.device), len(self.entries), start)
        table, n, total = self._table
        check(load().dt_pack_conv_weights_batched(table.data_ptr(), n, total, stream_ptr()))


def conv2d_dgrad_direct(gy: torch.Tensor, w: torch.Tensor, x_shape, stride: int, pad: int,
                        addend: Optional[torch.Tensor] = None, round_weights: bool = False) -> torch.Tensor:
    """gy (N, Ho, Wo, C_out), w fp32 OIHW -> gx of shape x_shape = (N, H, W, C_x)."""
    N, H, W, Cx = x_shape
    C_out, C_in, R, S = w.shape
    gx = torch.empty(tuple(x_shape), dtype=gy.dtype, device=gy.device)
    check(load().dt_conv2d_dgrad_direct(gy.data_ptr(), w.data_ptr(), ptr(addend), N, H, W, C_in, Cx, C_out, R, S, stride,
                                        pad, _dt(gy), int(round_weights), gx.data_ptr(), stream_ptr()))
    return gx


def conv2d_wgrad_direct(x: torch.Tensor, gy: torch.Tensor, w_shape, stride: int, pad: int, want_bias: bool = False,
                        out: Optional[torch.Tensor] = None, bias_out: Optional[torch.Tensor] = None):
    """x (N, H, W, C_x), gy (N, Ho, Wo, C_out) -> dw fp32 OIHW (and dbias)."""
    N, H, W, Cx = x.shape
    C_out, C_in, R, S = w_shape
    dw = out if out is not None else torch.empty(tuple(w_shape), dtype=torch.float32, device=x.device)
    db = bias_out if bias_out is not None else (torch.empty(C_out, dtype=torch.float32, device=x.device) if want_bias else None)
    check(load().dt_conv2d_wgrad_direct(x.data_ptr(), gy.data_ptr(), N, H, W, C_in, Cx, C_out, R, S, stride, pad, _dt(x),
                                        dw.data_ptr(), ptr(db), stream_ptr()))
    return dw, db


def wgrad_tc_supported(x: torch.Tensor, gy: torch.Tensor, w_shape, stride: int, pad: int) -> bool:
    """True when the tcgen05 weight-gradient kernel covers this layer shape (dt_conv2d_wgrad_tc)."""
    C_out, C_in, R, S = w_shape
    N, Ho, Wo, Cg = gy.shape
    kind = (R == 3 and S == 3 and pad == 1 and stride in (1, 2)) or (R == 1 and S == 1 and pad == 0 and stride == 2)
    return bool(x.dtype == torch.bfloat16 and kind and x.shape[1] == Ho * stride and x.shape[2] == Wo * stride
                and Wo % 8 == 0 and (Ho % 16 == 0 or (Ho == 8 and N % 2 == 0)) and C_in % 4 == 0
                and x.shape[-1] % 8 == 0 and Cg % 8 == 0 and x.shape[-1] >= C_in and Cg >= C_out)


def conv2d_wgrad_tc(x: torch.Tensor, gy: torch.Tensor, w_shape, stride: int = 1,
                    out: Optional[torch.Tensor] = None, tag: str = "wgrad") -> torch.Tensor:
    """tensor-core weight gradient: x (N, s*Ho, s*Wo, >=C_in), gy (N, Ho, Wo, >=C_out) bf16 -> dw fp32 OIHW."""
    C_out, C_in, R, S = w_shape
    N, Ho, Wo, Cg = gy.shape
    lib = load()
    need = lib.dt_conv2d_wgrad_tc_workspace(N, Ho, Wo, C_in, C_out, R, stride)
    if need < 0:
        raise _lib.DeadtreesB200Error(f"dt_conv2d_wgrad_tc does not support x {tuple(x.shape)} gy {tuple(gy.shape)} w {tuple(w_shape)}")
    ws = torch.empty((need + 3) // 4, dtype=torch.float32, device=x.device)
    dw = out if out is not None else torch.empty(tuple(w_shape), dtype=torch.float32, device=x.device)
    with _Timed("wgrad", 2.0 * N * Ho * Wo * C_out * C_in * R * S, tag):
        check(lib.dt_conv2d_wgrad_tc(x.data_ptr(), gy.data_ptr(), N, Ho, Wo, C_in, x.shape[-1], C_out, Cg, R, stride,
                                     dw.data_ptr(), ws.data_ptr(), need, stream_ptr()))
    return dw


def stem_wgrad_tc_supported(N: int, H: int, W: int) -> bool:
    return load().dt_stem_wgrad_tc_workspace(N, H, W) > 0


def stem_wgrad_tc(x_frame: torch.Tensor, gy: torch.Tensor, w_shape, out: Optional[torch.Tensor] = None,
                  tag: str = "wgrad.stem") -> torch.Tensor:
    """tensor-core weight gradient of the 7x7/s2 stem: x_frame (N, H+6, W+8, 4) bf16 zero-bordered, gy (N, H/2, W/2, 64)
    bf16 -> dw fp32 (64, C_in, 7, 7)."""
    C_out, C_in, R, S = w_shape
    N, Hp, Wp, _ = x_frame.shape
    H, W = Hp - 6, Wp - 8
    lib = load()
    need = lib.dt_stem_wgrad_tc_workspace(N, H, W)
    if need <= 0 or C_out != 64 or R != 7 or S != 7 or tuple(gy.shape) != (N, H // 2, W // 2, 64):
        raise _lib.DeadtreesB200Error(f"dt_stem_wgrad_tc does not support frame {tuple(x_frame.shape)} gy {tuple(gy.shape)}")
    ws = torch.empty((need + 3) // 4, dtype=torch.float32, device=gy.device)
    dw = out if out is not None else torch.empty(tuple(w_shape), dtype=torch.float32, device=gy.device)
    with _Timed("wgrad", 2.0 * N * (H // 2) * (W // 2) * C_out * C_in * R * S, tag):
        check(lib.dt_stem_wgrad_tc(x_frame.data_ptr(), gy.data_ptr(), N, H, W, C_in, dw.data_ptr(), ws.data_ptr(), need,
                                   stream_ptr()))
    return dw
