#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
DT_ROW_DEBUG=1 timeout 300 python - > gpurun_out/r2_rowdbg.log 2>&1 <<'P'
import torch, sys
sys.path.insert(0, '.')
from deadtrees_b200 import ops
from deadtrees_b200.engine import pack_weight
for cin, cout, N, H in ((64, 64, 405, 64), (16, 16, 405, 256), (32, 32, 405, 128)):
    x = torch.randn(N, H, H, cin, device='cuda').to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3) * 0.05
    wp = pack_weight(w, 'bf16', False, 'cuda')
    sc, sh = torch.ones(cout, device='cuda'), torch.zeros(cout, device='cuda')
    for it in range(3):
        y = ops.conv2d(x, wp, sc, sh, N=N, H=H, W=H, C_in=cin, C_x=cin, C_out=cout, R=3, S=3, stride=1, pad=1, relu=True)
    torch.cuda.synchronize()
P
echo "exit=$?"; grep "^\[row" gpurun_out/r2_rowdbg.log | cut -c1-400
