"""stand-alone launches of the row-streaming conv kernel on the cfg2 layer shapes (ncu / debug-counter probe)"""
import sys
import torch
sys.path.insert(0, '.')
from deadtrees_b200 import ops
from deadtrees_b200.engine import pack_weight
shapes = [(64, 64, 405, 64), (16, 16, 405, 256), (32, 32, 405, 128)]
if len(sys.argv) > 1:
    shapes = [shapes[int(a)] for a in sys.argv[1:]]
for cin, cout, N, H in shapes:
    x = torch.randn(N, H, H, cin, device='cuda').to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3) * 0.05
    wp = pack_weight(w, 'bf16', False, 'cuda')
    sc, sh = torch.ones(cout, device='cuda'), torch.zeros(cout, device='cuda')
    for it in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = ops.conv2d(x, wp, sc, sh, N=N, H=H, W=H, C_in=cin, C_x=cin, C_out=cout, R=3, S=3, stride=1, pad=1, relu=True)
        e1.record()
        torch.cuda.synchronize()
    print(f"{cin}->{cout} N={N} {H}x{H}: {1e3 * e0.elapsed_time(e1):.1f} us")
