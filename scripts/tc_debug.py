"""GPU diagnostic for the tcgen05 conv kernel: every test layer through the TMA producer and through the
gather producer, against the CPU reference, with a failure-pattern summary.  Never asserts."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from deadtrees_b200._lib import CONV_FORCE_GATHER  # noqa: E402
import test_gpu_conv as T  # noqa: E402


def summarize(name, got, ref, tol=0.02):
    d = (got - ref).abs()
    bad = d > tol * (ref.abs().max() + 1e-6)
    flips = (got != ref.to(torch.bfloat16).float()).float().mean().item()  # outputs that round differently from fp32-accurate
    print(f"{name}: max_err={d.max():.4f} ref_max={ref.abs().max():.3f} bad={int(bad.sum())}/{bad.numel()} "
          f"bf16-rounding flips={100 * flips:.3f}%", flush=True)
    if bad.any():
        idx = bad.nonzero()[:5].tolist()
        print("   first bad (n,c,h,w):", idx)
        print("   got:", [round(float(got[tuple(i)]), 3) for i in idx], "ref:", [round(float(ref[tuple(i)]), 3) for i in idx])
        print("   bad per out-channel (first 16):", bad.sum(dim=(0, 2, 3))[:16].tolist())
        print("   bad per row h (first 16):", bad.sum(dim=(0, 1, 3))[:16].tolist())
        print("   bad per col w (first 16):", bad.sum(dim=(0, 1, 2))[:16].tolist())
        print("   bad per image:", bad.sum(dim=(1, 2, 3)).tolist())


def main():
    only = sys.argv[1] if len(sys.argv) > 1 else None
    for case in T.LAYERS:
        if only and only not in case[0]:
            continue
        x, skip, w, scale, shift, residual = T.make_case(case)
        xb, sb, wb, rb = T.bf16_round(x, skip, w, residual)
        ref = T.reference(case, xb, sb, wb, scale, shift, rb)
        for flags, tag in ((CONV_FORCE_GATHER, "gather"), (0, "auto  ")):
            try:
                got = T.run_cuda(case, x, skip, w, scale, shift, residual, dtype=torch.bfloat16, flags=flags)
                summarize(f"[{tag}] {case[0]}", got, ref)
            except Exception as e:  # noqa: BLE001
                print(f"[{tag}] {case[0]} raised {e!r}", flush=True)
                return


if __name__ == "__main__":
    main()
