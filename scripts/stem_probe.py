"""stand-alone launches of the stem (+ max pooling) on the cfg2 batch shape (405 tiles of 256 x 256): fused stem + pool
(conv_stem.cu POOL variant, one or two CTAs per SM) against the stem and dt_maxpool3x3s2 launched separately"""
import os
import sys
import torch
sys.path.insert(0, '.')
from deadtrees_b200 import ops
from deadtrees_b200._lib import CONV_X_PAD3
from deadtrees_b200.engine import pack_weight
N, T = 405, 256
x = torch.zeros(N, T + 6, T + 8, 4, dtype=torch.bfloat16, device='cuda')
x[:, 3:3 + T, 3:3 + T, :3] = torch.randn(N, T, T, 3, device='cuda').to(torch.bfloat16)
wp = pack_weight(torch.randn(64, 3, 7, 7) * 0.1, 'bf16', True, 'cuda')
sc, sh = torch.ones(64, device='cuda'), torch.zeros(64, device='cuda')
y = torch.empty(N, T // 2, T // 2, 64, dtype=torch.bfloat16, device='cuda')
pooled = torch.empty(N, T // 4, T // 4, 64, dtype=torch.bfloat16, device='cuda')


def timed(fn, name):
    best = 1e9
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if it:
            best = min(best, e0.elapsed_time(e1))
    print(f"{name}: {1e3 * best:.1f} us", flush=True)


kw = dict(N=N, H=T, W=T, C_in=4, C_x=4, C_out=64, R=7, S=7, stride=2, pad=3, relu=True, flags=CONV_X_PAD3)
timed(lambda: ops.stem_pool(x, wp, sc, sh, N=N, H=T, W=T, out=y, pooled=pooled), "stem + pool fused, 2 CTAs/SM")
os.environ["DT_STEM_ONE_CTA"] = "1"
timed(lambda: ops.stem_pool(x, wp, sc, sh, N=N, H=T, W=T, out=y, pooled=pooled), "stem + pool fused, 1 CTA/SM")
timed(lambda: ops.conv2d(x, wp, sc, sh, out=y, **kw), "stem alone, 1 CTA/SM")
del os.environ["DT_STEM_ONE_CTA"]
timed(lambda: ops.conv2d(x, wp, sc, sh, out=y, **kw), "stem alone, 2 CTAs/SM")
timed(lambda: ops.maxpool3x3s2(y, out=pooled), "maxpool alone")
