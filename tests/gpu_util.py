"""helpers shared by the -m gpu parity tests (oracle on CPU vs CUDA path through the C-ABI)."""
import numpy as np
import torch

from oracle import ref_unet


def dev():
    return torch.device("cuda", 0)


def oracle_model(cin=3, k=3, seed=0):
    return ref_unet.build_reference_unet(cin, k, seed=seed)


def nhwc4(x_nchw: torch.Tensor, dtype) -> torch.Tensor:
    """CPU reference packing: (N, C, H, W) -> (N, H, W, 4) zero-padded channels."""
    N, C, H, W = x_nchw.shape
    out = torch.zeros(N, H, W, 4, dtype=torch.float32)
    out[..., :C] = x_nchw.permute(0, 2, 3, 1)
    return out.to(dtype)


def to_nchw(y_nhwc: torch.Tensor) -> torch.Tensor:
    return y_nhwc.float().permute(0, 3, 1, 2).contiguous().cpu()


def report(name, got: torch.Tensor, ref: torch.Tensor):
    d = (got.double() - ref.double()).abs()
    scale = ref.double().abs().max().item() + 1e-30
    msg = f"[{name}] max_abs={d.max().item():.3e} mean_abs={d.mean().item():.3e} ref_max={scale:.3e} rel={d.max().item() / scale:.3e}"
    print(msg)
    return d.max().item(), d.max().item() / scale


_TRAINED = {}


def trained_model(cin=3, k=3, seed=0):
    """the oracle network after 100 steps of the reference's training recipe on the CPU (a well-conditioned function,
    like a checkpoint of the reference; ``ref_unet.build_trained_unet``) - one per session, cached on disk."""
    key = (cin, k, seed)
    if key not in _TRAINED:
        _TRAINED[key] = ref_unet.build_trained_unet(cin, k, seed=seed)
    return _TRAINED[key]


def pattern_tiles(n, T, cin, seed=1234):
    """n synthetic-orthophoto tiles (uint8 (n, T, T, cin)) and their val_transform'ed fp32 NCHW form."""
    rng = np.random.default_rng(seed)
    u8 = np.stack([ref_unet.synthetic_pattern(int(rng.integers(0, 9000)), int(rng.integers(0, 9000)), T, T, cin, rng)[0]
                   for _ in range(n)])
    return u8, ref_unet.normalize_u8(u8)


def pattern_mosaic(H, W, cin=3, seed=21, oy=1234, ox=4321):
    return ref_unet.synthetic_pattern(oy, ox, H, W, cin, np.random.default_rng(seed))[0]
