"""ctypes binding of ``libdeadtrees_b200.so`` (C-ABI in ``include/deadtrees_b200.h``).

There is no CPU fallback: if the shared library is missing, or the device is not sm_100-class,
every compute call raises.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Optional

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libdeadtrees_b200.so"

DT_BF16, DT_F32 = 0, 1
CONV_FORCE_GATHER, CONV_FORCE_DIRECT, CONV_NO_HALO, CONV_X_PAD3, CONV_TRANSPOSED, CONV_NO_QUAD, CONV_UPS_FOLDED = 1, 2, 4, 8, 16, 32, 64
CONV_PAIR = 128
CONV_NO_ROW = 256


class DeadtreesB200Error(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "N", "H", "W", "C_in", "C_x", "upsample", "C_out", "R", "S", "stride", "pad", "relu", "has_residual",
        "dtype", "flags")]


_i, _i64, _f, _p = C.c_int, C.c_int64, C.c_float, C.c_void_p


class PackJob(C.Structure):
    """dt_pack_job (include/deadtrees_b200.h)"""
    _fields_ = [("w", C.c_void_p), ("out", C.c_void_p)] + [(n, C.c_int32) for n in (
        "C_out", "C_in", "R", "S", "mode", "C_in_p", "Kpad", "reserved")] + [("start", C.c_int64)]

_SIGNATURES = {
    "dt_version": ([], C.c_int),
    "dt_last_error": ([C.c_char_p, C.c_size_t], C.c_int),
    "dt_device_check": ([], C.c_int),
    "dt_make_blocks": ([_p, _i, _i, _i, _i, _i, _p, _p], C.c_int),
    "dt_unmake_blocks": ([_p, _i, _i, _i, _i, _p, _p], C.c_int),
    "dt_tile_gather_normalize": ([_p, _i, _i, _i, _i64, _i64, _i64, _i, _i, _i, _i, _i,
                                  C.POINTER(_f), C.POINTER(_f), _i, _i, _i, _p, _p], C.c_int),
    "dt_pack_input_nchw": ([_p, _i, _i, _i, _i, _i, _i, _p, _p], C.c_int),
    "dt_pack_input_nchw_frame": ([_p, _i, _i, _i, _i, _i, _p, _p], C.c_int),
    "dt_stitch_mask_u8": ([_p, _i, _i, _i, _i, _p, _i, _i, _i64, _p], C.c_int),
    "dt_stitch_blend_argmax": ([_p, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p, _i, _i, _i, _i, _p], C.c_int),
    "dt_conv2d_fwd": ([C.POINTER(ConvDesc), _p, _p, _p, _p, _p, _p, _p, _p], C.c_int),
    "dt_maxpool3x3s2": ([_p, _i, _i, _i, _i, _i, _p, _p], C.c_int),
    "dt_stem_pool_fwd": ([_p, _p, _p, _p, _p, _p, _p, _p], C.c_int),
    "dt_head_fwd": ([_p, _i, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p], C.c_int),
    "dt_head_fwd_tc": ([_p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p], C.c_int),
    "dt_tail_fused": ([_p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p], C.c_int),
    "dt_argmax_nchw": ([_p, _i, _i, _i, _i, _p, _p], C.c_int),
    "dt_mode_vote": ([_p, _i, _i64, _i, _p, _p], C.c_int),
    "dt_seg_loss_partials": ([_p, _p, _i, _i, _i, _i, _p, _p, _p, _p], C.c_int),
    "dt_seg_loss_finalize": ([_p, _p, _i, _i, _i, _i, _p, _p, _p, _p], C.c_int),
    "dt_seg_loss_backward": ([_p, _p, _i, _i, _i, _i, _p, _p, _f, _p, _p], C.c_int),
    "dt_gwdl_workspace": ([_i, _i, _i], C.c_int64),
    "dt_gwdl_loss": ([_p, _p, _i, _i, _i, _i, C.POINTER(_f), _i, _p, C.c_int64, _p, _p, _p], C.c_int),
    "dt_gwdl_loss_backward": ([_p, _p, _i, _i, _i, _i, C.POINTER(_f), _i, _p, _f, _p, _p], C.c_int),
    "dt_train_transform": ([_p, _p, _p, _i, _i, _i, _i, _i, _p, _p, C.POINTER(_f), C.POINTER(_f), _i, _p, _p, _p, _p, _p], C.c_int),
    "dt_confusion_matrix": ([_p, _i, _p, _p, _i, C.c_int64, _i, _p, _p, _p], C.c_int),
    "dt_one_hot2dist_workspace": ([_i, _i, _i, _i], C.c_int64),
    "dt_one_hot2dist": ([_p, _i, _i, _i, _i, _i, _p, _p, _i64, _p], C.c_int),
    "dt_boundary_loss": ([_p, _p, _i, _i, _i, _i, C.c_uint, _p, _p, _p], C.c_int),
    "dt_boundary_loss_backward": ([_p, _p, _i, _i, _i, _i, C.c_uint, _f, _p, _p], C.c_int),
    "dt_class2one_hot": ([_p, _i, _i, _i, _i, _p, _p, _p], C.c_int),
    "dt_softmax_nchw": ([_p, _i, _i, _i, _i, _p, _p], C.c_int),
    "dt_prob_loss_partials": ([_p, _p, _i, _i, _i, _i, _i, _f, _p, _p], C.c_int),
    "dt_reduce_blocks": ([_i64, _i, _i], C.c_int),
    "dt_bn_train_stats": ([_p, _i64, _i, _i, _p, _p, _f, _f, _p, _p, _p, _p, _p, _p, _p, _p], C.c_int),
    "dt_bn_apply": ([_p, _i64, _i, _i, _p, _p, _p, _i, _p, _p], C.c_int),
    "dt_bn_train_bwd": ([_p, _p, _p, _i64, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p], C.c_int),
    "dt_bn_train_bwd_relu": ([_p, _p, _i64, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p], C.c_int),
    "dt_add": ([_p, _p, _i64, _i, _p, _p], C.c_int),
    "dt_maxpool3x3s2_bwd": ([_p, _p, _p, _i, _i, _i, _i, _i, _p, _p], C.c_int),
    "dt_maxpool3x3s2_idx": ([_p, _i, _i, _i, _i, _i, _p, _p, _p], C.c_int),
    "dt_maxpool3x3s2_bwd_idx": ([_p, _p, _p, _i, _i, _i, _i, _i, _p, _p], C.c_int),
    "dt_zero_insert2x": ([_p, _i, _i, _i, _i, _i, _p, _p], C.c_int),
    "dt_upsample_concat": ([_p, _p, _i, _i, _i, _i, _i, _i, _p, _p], C.c_int),
    "dt_upsample_concat_bwd": ([_p, _i, _i, _i, _i, _i, _i, _p, _p, _p], C.c_int),
    "dt_nchw_to_nhwc": ([_p, _i, _i, _i, _i, _i, _i, _p, _p], C.c_int),
    "dt_channel_sum": ([_p, _i64, _i, _i, _i, _p, _p, _p], C.c_int),
    "dt_pack_conv_weight": ([_p, _i, _i, _i, _i, _i, _i, _i, _p, _p], C.c_int),
    "dt_pack_conv_weights_batched": ([_p, _i, _i64, _p], C.c_int),
    "dt_conv2d_dgrad_direct": ([_p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p], C.c_int),
    "dt_conv2d_wgrad_direct": ([_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _p], C.c_int),
    "dt_conv2d_wgrad_tc_workspace": ([_i, _i, _i, _i, _i, _i, _i], C.c_int64),
    "dt_conv2d_wgrad_tc": ([_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p, _i64, _p], C.c_int),
    "dt_stem_wgrad_tc_workspace": ([_i, _i, _i], C.c_int64),
    "dt_stem_wgrad_tc": ([_p, _p, _i, _i, _i, _i, _p, _p, _i64, _p], C.c_int),
    "dt_sumsq": ([_p, _i64, _p, _p], C.c_int),
    "dt_adam_step_dev": ([_p, _p, _p, _p, _i64, _p, _f, _f, _f, _p, _f, _p, _p, _p], C.c_int),
    "dt_adam_step": ([_p, _p, _p, _p, _i64, _f, _f, _f, _f, _i, _p, _f, _p], C.c_int),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (built in-tree by ``deadtrees_b200._build`` / ``__graft_entry__.build``)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise DeadtreesB200Error(
                f"{LIB_PATH} is missing: build it with `python -m deadtrees_b200._build` "
                "(deadtrees_b200 has no CPU fallback)")
        lib = C.CDLL(str(LIB_PATH))
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes, fn.restype = argtypes, restype
        _lib = lib
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    load().dt_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc: int) -> None:
    if rc != 0:
        raise DeadtreesB200Error(f"deadtrees_b200 error {rc}: {last_error()}")


def require_device() -> None:
    """Fail loudly when there is no usable B200-class device (no CPU fallback)."""
    if not torch.cuda.is_available():
        raise DeadtreesB200Error("CUDA device required: deadtrees_b200 has no CPU fallback")
    check(load().dt_device_check())


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()
