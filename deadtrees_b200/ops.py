"""Tensor-level wrappers over the C-ABI (one Python function per ``dt_*`` entry point).

All tensors are CUDA tensors; every call runs on ``torch.cuda.current_stream()``.  Nothing here
computes on the CPU — a missing library or a non-B200 device raises ``DeadtreesB200Error``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import DT_BF16, DT_F32, ConvDesc, load, ptr, stream_ptr

ACT_DTYPES = {DT_BF16: torch.bfloat16, DT_F32: torch.float32}

# number of kernel launches issued through this module (bench.py reports it as `gpu_launches`);
LAUNCHES = 0
# bench instrumentation: when PROFILE is a dict, conv / gather / stitch launches are bracketed by CUDA
# events on the launching stream and appended as (kind, start, end, algorithmic_work) tuples
PROFILE = None


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
    # algorithmic bytes: every logit of the shard once + 1 B per output pixel
    with _Timed("stitch", float(logits.numel()) * logits.element_size() + float(nrows) * W):
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
    if out is None:
        out = torch.empty((N, Ho, Wo, C_out), dtype=x.dtype, device=x.device)
    d = ConvDesc(N, H, W, C_in, C_x, int(upsample), C_out, R, S, stride, pad, int(relu), int(residual is not None),
                 _dt(x), flags)
    with _Timed("conv", 2.0 * N * Ho * Wo * C_out * (algo_cin or C_in) * R * S, tag):
        check(load().dt_conv2d_fwd(C.byref(d), x.data_ptr(), ptr(skip), w.data_ptr(), scale.data_ptr(),
                                   shift.data_ptr(), ptr(residual), out.data_ptr(), stream_ptr()))
    return out


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


def argmax_nchw(logits: torch.Tensor) -> torch.Tensor:
    logits = _cuda(logits, "logits")
    N, K, H, W = logits.shape
    out = torch.empty((N, H, W), dtype=torch.uint8, device=logits.device)
    check(load().dt_argmax_nchw(logits.data_ptr(), N, K, H, W, out.data_ptr(), stream_ptr()))
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


def sumsq(g: torch.Tensor, acc: torch.Tensor) -> None:
    check(load().dt_sumsq(g.data_ptr(), g.numel(), acc.data_ptr(), stream_ptr()))


def adam_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, *, lr: float, beta1: float,
              beta2: float, eps: float, step: int, sumsq_acc: Optional[torch.Tensor], max_norm: float) -> None:
    check(load().dt_adam_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1, beta2,
                              eps, step, ptr(sumsq_acc), max_norm, stream_ptr()))
