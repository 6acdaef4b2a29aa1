"""GPU diagnostic for the tcgen05 conv kernel: small structured cases whose failure pattern localises
descriptor / swizzle / mapping bugs.  Prints a compact report; never asserts."""
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from deadtrees_b200 import ops  # noqa: E402
from deadtrees_b200._lib import CONV_FORCE_GATHER  # noqa: E402
from deadtrees_b200.engine import pack_weight  # noqa: E402


def run(x, w, stride=1, pad=None, flags=0):
    Co, Ci, R, _ = w.shape
    pad = R // 2 if pad is None else pad
    N, _, H, _ = x.shape
    xb = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    wp = pack_weight(w, "bf16", False, "cuda")
    y = ops.conv2d(xb, wp, torch.ones(Co, device="cuda"), torch.zeros(Co, device="cuda"), N=N, H=H, W=H, C_in=Ci,
                   C_x=Ci, C_out=Co, R=R, S=R, stride=stride, pad=pad, relu=False, flags=flags)
    torch.cuda.synchronize()
    return y.float().permute(0, 3, 1, 2).cpu()


def summarize(name, got, ref):
    d = (got - ref).abs()
    bad = d > 0.02 * (ref.abs().max() + 1e-6)
    print(f"{name}: max_err={d.max():.4f} ref_max={ref.abs().max():.3f} bad={int(bad.sum())}/{bad.numel()}", flush=True)
    if bad.any():
        idx = bad.nonzero()[:6].tolist()
        print("   first bad (n,c,h,w):", idx)
        print("   got:", [round(float(got[tuple(i)]), 3) for i in idx], "ref:", [round(float(ref[tuple(i)]), 3) for i in idx])
        # per-channel / per-row structure of the failure
        print("   bad per out-channel (first 16):", bad.sum(dim=(0, 2, 3))[:16].tolist())
        print("   bad per row h (first 16):", bad.sum(dim=(0, 1, 3))[:16].tolist())


def main():
    torch.manual_seed(0)
    for flags, tag in ((CONV_FORCE_GATHER, "gather"), (0, "tma")):
        # 1x1 identity: y[m, co] == x[m, co]; exercises A layout, B layout, TMEM->row mapping
        x = torch.randn(1, 64, 16, 8).to(torch.bfloat16).float()   # 128 pixels -> one M tile (non-square on purpose)
        w = torch.eye(64).reshape(64, 64, 1, 1)
        try:
            xb = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
            wp = pack_weight(w, "bf16", False, "cuda")
            y = ops.conv2d(xb, wp, torch.ones(64, device="cuda"), torch.zeros(64, device="cuda"), N=1, H=16, W=8,
                           C_in=64, C_x=64, C_out=64, R=1, S=1, stride=1, pad=0, relu=False, flags=flags)
            torch.cuda.synchronize()
            summarize(f"[{tag}] 1x1 identity 64ch", y.float().permute(0, 3, 1, 2).cpu(), x)
        except Exception as e:  # noqa: BLE001
            print(f"[{tag}] 1x1 identity raised {e!r}")
            return
        for (N, H, Ci, Co, R) in [(2, 16, 64, 64, 1), (2, 16, 64, 64, 3), (2, 16, 128, 128, 3), (4, 8, 64, 256, 3),
                                  (1, 32, 64, 32, 3), (1, 32, 64, 16, 3), (1, 128, 64, 64, 3)]:
            x = torch.randn(N, Ci, H, H).to(torch.bfloat16).float()
            w = (torch.randn(Co, Ci, R, R) * (2.0 / (Ci * R * R)) ** 0.5).to(torch.bfloat16).float()
            try:
                got = run(x, w, flags=flags)
                summarize(f"[{tag}] conv N{N} H{H} {Ci}->{Co} k{R}", got, F.conv2d(x, w, None, 1, R // 2))
            except Exception as e:  # noqa: BLE001
                print(f"[{tag}] conv N{N} H{H} {Ci}->{Co} k{R} raised {e!r}")
                return


if __name__ == "__main__":
    main()
