"""GPU experiment: which UMMA descriptor addressing reads a shifted window of a swizzled halo patch correctly."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from deadtrees_b200._lib import CONV_HALO_BASEOFF, CONV_HALO_P16, CONV_NO_HALO  # noqa: E402
CONV_HALO = 0  # the halo path is the default
import test_gpu_conv as T  # noqa: E402

CASES = [
    ("64->64 @64", 2, 64, 64, 0, 64, 3, 1, 1, False, True, True),
    ("128->128 @32", 4, 32, 128, 0, 128, 3, 1, 1, False, False, True),
    ("256->256 @16", 8, 16, 256, 0, 256, 3, 1, 1, False, True, True),
    ("64->64 @128", 1, 128, 64, 0, 64, 3, 1, 1, False, False, True),
    ("32->32 @128", 1, 128, 32, 0, 32, 3, 1, 1, False, False, True),
    ("16->16 @64", 2, 64, 16, 0, 16, 3, 1, 1, False, False, True),
]
VARIANTS = [("p10 bo0", CONV_HALO), ("p10 bo1", CONV_HALO | CONV_HALO_BASEOFF),
            ("p16 bo0", CONV_HALO | CONV_HALO_P16), ("p16 bo1", CONV_HALO | CONV_HALO_P16 | CONV_HALO_BASEOFF)]

for case in CASES:
    args = T.make_case(case)
    xb, sb, wb, rb = T.bf16_round(args[0], args[1], args[2], args[5])
    ref = T.reference(case, xb, sb, wb, args[3], args[4], rb)
    base = T.run_cuda(case, *args, dtype=torch.bfloat16, flags=CONV_NO_HALO)
    for name, flags in VARIANTS:
        try:
            got = T.run_cuda(case, *args, dtype=torch.bfloat16, flags=flags)
        except Exception as e:  # noqa: BLE001
            print(f"{case[0]:14s} {name}: raised {e!r}", flush=True)
            sys.exit(0)
        d = (got - ref).abs()
        bad = d > 0.02 * ref.abs().max()
        same = (got == base).float().mean().item()
        print(f"{case[0]:14s} {name}: max_err={d.max():.4f} bad={int(bad.sum())}/{bad.numel()} identical_to_per_tap={100 * same:.2f}%"
              + (f" bad rows h: {bad.sum(dim=(0, 1, 3))[:18].tolist()} cols w: {bad.sum(dim=(0, 1, 2))[:10].tolist()}" if bad.any() else ""),
              flush=True)
