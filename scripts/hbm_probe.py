"""Micro-benchmark of the tiler's HBM kernels on cfg2 shapes (one band of 9 tile rows, and the whole mosaic):
gather + normalise into the stem frame, blended mask-only stitch.  Prints GB/s on SURVEY.md 8d's algorithmic bytes and on
the bytes really moved.  Usage: python scripts/hbm_probe.py [label]   (kernel variants are selected by DT_* env vars)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from deadtrees_b200 import ops
from deadtrees_b200.data.deadtreedata import normalize_constants

label = sys.argv[1] if len(sys.argv) > 1 else ""
dev = torch.device("cuda:0")
T, ov, K = 256, 32, 3
step = T - ov
H = W = 10000
gy = gx = (H - T + step - 1) // step + 1
g = torch.Generator(device=dev).manual_seed(1)
mosaic = torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, device=dev, generator=g)
off, sc = normalize_constants(3, None, None)
win = torch.ones(T, dtype=torch.float32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=10):
    best = []
    for i in range(n + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 2:
            best.append(e0.elapsed_time(e1))
    return float(np.median(best)), float(min(best))


for rows in (9, gy):
    nt = rows * gx
    frame = torch.zeros((nt, T + 6, T + 8, 4), dtype=torch.bfloat16, device=dev)
    med, mn = timeit(lambda: ops.tile_gather_normalize(mosaic, "hwc", 3, T, ov, (gy, gx), 0, nt, off, sc, out=frame, pad=3))
    alg = nt * T * T * 9
    moved = nt * T * T * (3 + 8)
    print(f"[{label}] gather {nt:5d} tiles: median {1e3 * med:7.1f} us  {alg / med / 1e6:6.0f} GB/s algorithmic, {moved / med / 1e6:6.0f} GB/s moved (min {1e3 * mn:.1f} us)")
    del frame
    logits = torch.randn((nt, T, T, K), device=dev, generator=g).to(torch.bfloat16)
    nrows = min(H, (rows - 1) * step + T) if rows == gy else rows * step
    mask = torch.empty((H, W), dtype=torch.uint8, device=dev)
    med, mn = timeit(lambda: ops.stitch_blend_argmax(logits, ov, (gy, gx), win, mask, row0=0, nrows=nrows))
    cov = sum(max(0, min(nrows, ty * step + T) - ty * step) for ty in range(rows))
    alg = cov * gx * T * K * 2 + nrows * W
    print(f"[{label}] stitch {nrows:5d} rows : median {1e3 * med:7.1f} us  {alg / med / 1e6:6.0f} GB/s algorithmic (every covering logit + mask), {nrows * W * 7 / med / 1e6:6.0f} GB/s on 7 B / pixel (min {1e3 * mn:.1f} us)")
    del logits, mask
