"""stand-alone launches of the fused decoder tail (conv_tail.cu) and of the two launches it replaces, cfg2 shape
(405 tiles of 256 x 256); with DT_ROW_DEBUG=1 the kernels print their per-role wait clocks"""
import sys
import torch
sys.path.insert(0, '.')
from deadtrees_b200 import ops
from deadtrees_b200.engine import pack_weight
N, H, K = 405, 256, 3
x = torch.randn(N, H, H, 16, device='cuda').to(torch.bfloat16)
w2 = pack_weight(torch.randn(16, 16, 3, 3) * 0.1, 'bf16', False, 'cuda')
hw = torch.zeros(16, 16, 3, 3)
hw[:K] = torch.randn(K, 16, 3, 3) * 0.1
wh = pack_weight(hw, 'bf16', False, 'cuda')
sc, sh, b16 = torch.ones(16, device='cuda'), torch.zeros(16, device='cuda'), torch.zeros(16, device='cuda')
nhwc = torch.empty(N, H, H, K, dtype=torch.bfloat16, device='cuda')
mask = torch.empty(N, H, H, dtype=torch.uint8, device='cuda')


def timed(fn, name):
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
    print(f"{name}: {1e3 * e0.elapsed_time(e1):.1f} us", flush=True)


timed(lambda: ops.tail_fused(x, w2, sc, sh, wh, b16, K, logits_nhwc=nhwc), "fused tail (logits_nhwc)")
timed(lambda: ops.tail_fused(x, w2, sc, sh, wh, b16, K, mask=mask), "fused tail (mask only)")
mid = torch.empty_like(x)
timed(lambda: ops.conv2d(x, w2, sc, sh, N=N, H=H, W=H, C_in=16, C_x=16, C_out=16, R=3, S=3, stride=1, pad=1, relu=True, out=mid),
      "conv2 alone")
timed(lambda: ops.head_tc(mid, wh, b16, K, logits_nhwc=nhwc), "head alone (logits_nhwc)")
