"""GPU parity of the row-streaming convolution kernel (csrc/conv_row.cu: the three vertical taps ride in the N dimension
of one tcgen05.mma, accumulators of consecutive output rows in a ring of TMEM column slots).

Per output accumulator the additions happen in the same (r, s, ci) order as in the 8 x 16 tile kernels, so the results must
be BIT-IDENTICAL to those (flag DT_CONV_NO_ROW selects them); both are also checked against torch fp32 on bf16-rounded
operands."""
import os

import pytest
import torch
import torch.nn.functional as F

from deadtrees_b200 import ops
from deadtrees_b200._lib import CONV_NO_ROW
from deadtrees_b200.engine import pack_weight
from gpu_util import report

pytestmark = pytest.mark.gpu

# cin, cout, N, H, W, residual, relu
CASES = [
    (64, 64, 3, 64, 64, True, True),       # resnet layer1: two images per M tile, odd batch (last tile half empty)
    (64, 64, 2, 64, 64, False, True),
    (64, 64, 1, 128, 128, False, True),    # decoder.blocks.2.conv2 at T = 512
    (32, 32, 2, 128, 128, False, True),    # decoder.blocks.3.conv2
    (16, 16, 2, 256, 256, False, True),    # decoder.blocks.4.conv2: two column blocks per row
    (16, 16, 1, 40, 128, True, False),     # height that is not a multiple of the row chunk, no ReLU, residual
    (64, 64, 4, 24, 64, True, True),
    (64, 32, 1, 128, 128, False, True),
    (32, 16, 1, 256, 256, False, True),
    (32, 64, 2, 64, 64, False, False),
    (16, 32, 1, 128, 128, False, True),
    (64, 16, 2, 64, 64, False, True),
    (64, 64, 37, 64, 64, True, True),      # many work items per CTA: ring wrap-around, slot reuse across items
    (16, 16, 9, 256, 256, False, True),
]


def run(x, w, scale, shift, res, relu, flags):
    N, H, W, cin = x.shape
    cout = w.shape[0]
    wp = pack_weight(w, "bf16", False, "cuda")
    y = ops.conv2d(x, wp, scale, shift, N=N, H=H, W=W, C_in=cin, C_x=cin, C_out=cout, R=3, S=3, stride=1, pad=1, relu=relu,
                   residual=res, flags=flags)
    torch.cuda.synchronize()
    return y


@pytest.mark.parametrize("cin,cout,N,H,W,res,relu", CASES)
def test_conv_row_equals_tile_kernels_and_torch(cin, cout, N, H, W, res, relu):
    g = torch.Generator().manual_seed(cin * 1000 + cout + N + H)
    x = torch.randn(N, H, W, cin, generator=g).to(torch.bfloat16)
    w = torch.randn(cout, cin, 3, 3, generator=g) * (2.0 / (cin * 9)) ** 0.5
    scale = 1.0 + 0.1 * torch.randn(cout, generator=g)
    shift = 0.1 * torch.randn(cout, generator=g)
    r = torch.randn(N, H, W, cout, generator=g).to(torch.bfloat16) if res else None
    xd, rd = x.cuda(), (r.cuda() if res else None)
    got = run(xd, w, scale.cuda(), shift.cuda(), rd, relu, 0)
    tile = run(xd, w, scale.cuda(), shift.cuda(), rd, relu, CONV_NO_ROW)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), None, 1, 1)
    ref = ref * scale[None, :, None, None] + shift[None, :, None, None]
    if res:
        ref = ref + r.float().permute(0, 3, 1, 2)
    if relu:
        ref = F.relu(ref)
    err, rel = report(f"conv_row {cin}->{cout} N={N} {H}x{W}", got.float().permute(0, 3, 1, 2).cpu(), ref)
    assert rel < 1e-2
    same = torch.equal(got.view(torch.int16), tile.view(torch.int16))
    if not same:
        d = (got.float() - tile.float()).abs()
        print(f"differs from the tile kernel: max {d.max().item():.4e}, {int((d > 0).sum())} of {d.numel()} elements")
    assert same


@pytest.mark.parametrize("N,H,W,K", [(2, 256, 256, 3), (3, 64, 64, 2), (1, 128, 128, 4), (1, 48, 128, 1)])
def test_head_row_equals_head_tile_kernel(N, H, W, K):
    g = torch.Generator().manual_seed(N + H + K)
    x = torch.randn(N, H, W, 16, generator=g).to(torch.bfloat16).cuda()
    hw = torch.zeros(16, 16, 3, 3)
    hw[:K] = torch.randn(K, 16, 3, 3, generator=g) * 0.1
    wp = pack_weight(hw, "bf16", False, "cuda")
    b16 = torch.zeros(16)
    b16[:K] = torch.randn(K, generator=g)
    b16 = b16.cuda()
    outs = []
    for row in ("1", "0"):
        os.environ["DT_CONV_ROW"] = row
        try:
            nchw = torch.empty(N, K, H, W, dtype=torch.float32, device="cuda")
            nhwc = torch.empty(N, H, W, K, dtype=torch.bfloat16, device="cuda")
            mask = torch.empty(N, H, W, dtype=torch.uint8, device="cuda")
            ops.head_tc(x, wp, b16, K, logits_nchw=nchw, logits_nhwc=nhwc, mask=mask)
            torch.cuda.synchronize()
            outs.append((nchw, nhwc, mask))
        finally:
            os.environ.pop("DT_CONV_ROW", None)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2).cpu(), hw[:K].to(torch.bfloat16).float(), b16[:K].cpu(), 1, 1)
    err, rel = report(f"head_row N={N} {H}x{W} K={K}", outs[0][0].cpu(), ref)
    assert rel < 1e-2
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][2], outs[1][2])
    assert torch.equal(outs[0][1].view(torch.int16), outs[1][1].view(torch.int16))
    assert torch.equal(outs[0][2].cpu().long(), outs[0][0].argmax(1).cpu())


@pytest.mark.parametrize("N,H,W,K", [(2, 256, 256, 3), (1, 256, 256, 2), (3, 128, 128, 4), (1, 40, 128, 1), (11, 256, 256, 3),
                                     (1, 8, 256, 3), (2, 33, 128, 3)])
def test_fused_tail_equals_conv2_then_head(N, H, W, K):
    """conv_tail.cu (decoder.blocks.4.conv2 + head in one launch, intermediate rows in shared memory) must be bit-identical
    to the two launches it replaces: same bf16 rounding of the intermediate, same accumulation order."""
    g = torch.Generator().manual_seed(7 * N + H + K)
    x = torch.randn(N, H, W, 16, generator=g).to(torch.bfloat16).cuda()
    w2 = torch.randn(16, 16, 3, 3, generator=g) * (2.0 / 144) ** 0.5
    scale = (1.0 + 0.1 * torch.randn(16, generator=g)).cuda()
    shift = (0.1 * torch.randn(16, generator=g)).cuda()
    hw = torch.zeros(16, 16, 3, 3)
    hw[:K] = torch.randn(K, 16, 3, 3, generator=g) * 0.1
    w2p, whp = pack_weight(w2, "bf16", False, "cuda"), pack_weight(hw, "bf16", False, "cuda")
    b16 = torch.zeros(16)
    b16[:K] = torch.randn(K, generator=g)
    b16 = b16.cuda()

    def outputs():
        return (torch.full((N, K, H, W), float("nan"), dtype=torch.float32, device="cuda"),
                torch.zeros(N, H, W, K, dtype=torch.bfloat16, device="cuda"),
                torch.full((N, H, W), 255, dtype=torch.uint8, device="cuda"))

    mid = run(x, w2, scale, shift, None, True, 0)
    a = outputs()
    ops.head_tc(mid, whp, b16, K, logits_nchw=a[0], logits_nhwc=a[1], mask=a[2])
    b = outputs()
    ops.tail_fused(x, w2p, scale, shift, whp, b16, K, logits_nchw=b[0], logits_nhwc=b[1], mask=b[2])
    torch.cuda.synchronize()
    d = (a[0] - b[0]).abs()
    if not torch.equal(a[0], b[0]):
        print(f"fused tail differs: max {d.max().item():.4e}, {int((d > 0).sum() + torch.isnan(d).sum())} of {d.numel()}")
    assert torch.equal(a[0], b[0])
    assert torch.equal(a[1].view(torch.int16), b[1].view(torch.int16))
    assert torch.equal(a[2], b[2])
    # only some outputs requested
    c = outputs()
    ops.tail_fused(x, w2p, scale, shift, whp, b16, K, mask=c[2])
    torch.cuda.synchronize()
    assert torch.equal(a[2], c[2])
    ref_mid = torch.relu(F.conv2d(x.float().permute(0, 3, 1, 2).cpu(), w2.to(torch.bfloat16).float(), None, 1, 1)
                         * scale.cpu().view(1, -1, 1, 1) + shift.cpu().view(1, -1, 1, 1)).to(torch.bfloat16).float()
    ref = F.conv2d(ref_mid, hw[:K].to(torch.bfloat16).float(), b16[:K].cpu(), 1, 1)
    err, rel = report(f"fused tail N={N} {H}x{W} K={K}", b[0].cpu(), ref)
    assert rel < 2e-2


def test_fused_tail_rejects_other_widths():
    x = torch.zeros(1, 64, 64, 16, dtype=torch.bfloat16, device="cuda")
    w = pack_weight(torch.zeros(16, 16, 3, 3), "bf16", False, "cuda")
    z = torch.zeros(16, device="cuda")
    with pytest.raises(RuntimeError):
        ops.tail_fused(x, w, z, z, w, z, 3, mask=torch.zeros(1, 64, 64, dtype=torch.uint8, device="cuda"))
