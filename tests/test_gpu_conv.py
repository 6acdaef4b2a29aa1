"""GPU parity: fused convolutions (fp32 check mode and the tcgen05 bf16 path) vs torch CPU fp32."""
import pytest
import torch
import torch.nn.functional as F

from deadtrees_b200 import ops
from deadtrees_b200._lib import (CONV_FORCE_DIRECT, CONV_FORCE_GATHER, CONV_NO_HALO, CONV_NO_QUAD, CONV_PAIR,
                                 CONV_UPS_FOLDED, CONV_X_PAD3)
from deadtrees_b200.engine import fold_upsample_weights, pack_weight, pack_weight_folded
from gpu_util import report, to_nchw

pytestmark = pytest.mark.gpu

# name, N, H(in, virtual), C_x, C_skip, C_out, R, stride, pad, upsample, residual, relu
LAYERS = [
    ("l1.conv 64->64 @64 (TMA)", 2, 64, 64, 0, 64, 3, 1, 1, False, True, True),
    ("l2.conv 128->128 @32 (TMA)", 4, 32, 128, 0, 128, 3, 1, 1, False, False, True),
    ("l3.conv 256->256 @16 (TMA)", 8, 16, 256, 0, 256, 3, 1, 1, False, True, True),
    ("l4.conv 512->512 @8 (TMA multi-image box)", 6, 8, 512, 0, 512, 3, 1, 1, False, True, True),
    ("d2.conv2 64->64 @128 wide rows", 1, 128, 64, 0, 64, 3, 1, 1, False, False, True),
    ("l2.0.conv1 64->128 s2", 2, 64, 64, 0, 128, 3, 2, 1, False, False, True),
    ("l2.0.downsample 1x1 s2", 2, 64, 64, 0, 128, 1, 2, 0, False, False, False),
    ("l4.0.conv1 256->512 s2 @16", 4, 16, 256, 0, 512, 3, 2, 1, False, False, True),
    ("d0.conv1 up(512)+256 -> 256 @16", 2, 16, 512, 256, 256, 3, 1, 1, True, False, True),
    ("d1.conv1 up(256)+128 -> 128 @32", 2, 32, 256, 128, 128, 3, 1, 1, True, False, True),
    ("d2.conv1 up(128)+64 -> 64 @64", 3, 64, 128, 64, 64, 3, 1, 1, True, False, True),
    ("d3.conv1 up(64)+64 -> 32 @128", 1, 128, 64, 64, 32, 3, 1, 1, True, False, True),
    ("d3.conv1 up(64)+64 -> 32 @64 x5 images", 5, 64, 64, 64, 32, 3, 1, 1, True, False, False),
    ("d3.conv2 32->32 @128", 1, 128, 32, 0, 32, 3, 1, 1, False, False, True),
    ("d4.conv1 up(32) -> 16 @64", 2, 64, 32, 0, 16, 3, 1, 1, True, False, True),
    ("d4.conv2 16->16 @64", 2, 64, 16, 0, 16, 3, 1, 1, False, False, True),
    ("ragged M tail 64->64 @24", 3, 24, 64, 0, 64, 3, 1, 1, False, True, True),
]


def make_case(case, seed=0):
    name, N, H, Cx, Cs, Co, R, stride, pad, ups, res, relu = case
    g = torch.Generator().manual_seed(seed)
    Hx = H // 2 if ups else H
    x = torch.randn(N, Cx, Hx, Hx, generator=g)
    skip = torch.randn(N, Cs, H, H, generator=g) if Cs else None
    w = torch.randn(Co, Cx + Cs, R, R, generator=g) * (2.0 / ((Cx + Cs) * R * R)) ** 0.5
    scale = 1.0 + 0.1 * torch.randn(Co, generator=g)
    shift = 0.1 * torch.randn(Co, generator=g)
    Ho = (H + 2 * pad - R) // stride + 1
    residual = torch.randn(N, Co, Ho, Ho, generator=g) if res else None
    return x, skip, w, scale, shift, residual


def reference(case, x, skip, w, scale, shift, residual):
    name, N, H, Cx, Cs, Co, R, stride, pad, ups, res, relu = case
    xin = F.interpolate(x, scale_factor=2, mode="nearest") if ups else x
    if skip is not None:
        xin = torch.cat([xin, skip], dim=1)
    y = F.conv2d(xin, w, None, stride, pad) * scale[None, :, None, None] + shift[None, :, None, None]
    if residual is not None:
        y = y + residual
    return F.relu(y) if relu else y


def run_cuda(case, x, skip, w, scale, shift, residual, dtype, flags=0):
    name, N, H, Cx, Cs, Co, R, stride, pad, ups, res, relu = case
    nhwc = lambda t: None if t is None else t.permute(0, 2, 3, 1).contiguous().to(dtype).cuda()
    wp = pack_weight(w, "fp32" if dtype == torch.float32 else "bf16", False, "cuda")
    y = ops.conv2d(nhwc(x), wp, scale.cuda(), shift.cuda(), N=N, H=H, W=H, C_in=Cx + Cs, C_x=Cx, C_out=Co, R=R, S=R,
                   stride=stride, pad=pad, relu=relu, skip=nhwc(skip), upsample=ups, residual=nhwc(residual),
                   flags=flags)
    torch.cuda.synchronize()
    return to_nchw(y)


def bf16_round(*ts):
    return [None if t is None else t.to(torch.bfloat16).float() for t in ts]


@pytest.mark.parametrize("case", LAYERS, ids=[c[0] for c in LAYERS])
def test_conv_fp32_check_mode(case):
    args = make_case(case)
    ref = reference(case, *args)
    got = run_cuda(case, *args, dtype=torch.float32)
    err, rel = report(case[0] + " fp32", got, ref)
    assert err < 1e-4


@pytest.mark.parametrize("case", LAYERS, ids=[c[0] for c in LAYERS])
def test_conv_bf16_direct_validation_path(case):
    """CUDA-core kernel on bf16 tensors with the tensor-core weight packing (validates the packing)."""
    x, skip, w, scale, shift, residual = make_case(case)
    xb, sb, wb, rb = bf16_round(x, skip, w, residual)
    ref = reference(case, xb, sb, wb, scale, shift, rb)
    got = run_cuda(case, x, skip, w, scale, shift, residual, dtype=torch.bfloat16, flags=CONV_FORCE_DIRECT)
    err, rel = report(case[0] + " bf16-direct", got, ref)
    assert rel < 1e-2  # output rounding to bf16 only


@pytest.mark.parametrize("case", LAYERS, ids=[c[0] for c in LAYERS])
def test_conv_tcgen05(case):
    """tensor-core path: bf16 operands, fp32 accumulate; reference uses the same bf16-rounded operands."""
    x, skip, w, scale, shift, residual = make_case(case)
    xb, sb, wb, rb = bf16_round(x, skip, w, residual)
    ref = reference(case, xb, sb, wb, scale, shift, rb)
    got = run_cuda(case, x, skip, w, scale, shift, residual, dtype=torch.bfloat16)
    err, rel = report(case[0] + " tcgen05", got, ref)
    assert rel < 1e-2
    # the generic gather producer and the per-tap TMA producer feed the same MMA sequence: identical bits
    got_g = run_cuda(case, x, skip, w, scale, shift, residual, dtype=torch.bfloat16, flags=CONV_FORCE_GATHER)
    got_t = run_cuda(case, x, skip, w, scale, shift, residual, dtype=torch.bfloat16, flags=CONV_NO_HALO)
    assert torch.equal(got_t, got_g)
    # the halo kernel (default for 3x3/s1 layers) sums K slab-major: equal up to fp32 summation order
    same = (got == got_g).float().mean().item()
    assert same > 0.995 and (got - got_g).abs().max() <= 2.0 ** -7 * ref.abs().max()
    if case[9]:
        # up-sample + concat: the class-fused tiles (four parity classes share the region's patches) keep every
        # class's K order - identical bits to the class-per-tile kernel
        got_nq = run_cuda(case, x, skip, w, scale, shift, residual, dtype=torch.bfloat16, flags=CONV_NO_QUAD)
        assert torch.equal(got, got_nq)


@pytest.mark.parametrize("N,H,Cx,Co", [(2, 64, 32, 16), (1, 128, 32, 16), (3, 32, 32, 32), (2, 32, 16, 16), (1, 64, 64, 32)])
def test_conv_upsample_folded(N, H, Cx, Co):
    """DT_CONV_UPS_FOLDED: nearest-x2 up-sampling folded into per-class 2 x 2 weights (resident-weight parity kernel).
    Against torch with the SAME folded bf16 weights (per-class convolutions of the low-res tensor), and against the
    nine-tap kernel - the two differ only by the rounding of the summed weights."""
    g = torch.Generator().manual_seed(N + H + Cx)
    x = torch.randn(N, Cx, H // 2, H // 2, generator=g)
    w = torch.randn(Co, Cx, 3, 3, generator=g) * (2.0 / (Cx * 9)) ** 0.5
    scale, shift = 1.0 + 0.1 * torch.randn(Co, generator=g), 0.1 * torch.randn(Co, generator=g)
    xb = x.to(torch.bfloat16).float()
    wf = fold_upsample_weights(w).to(torch.bfloat16).float()             # (Co, class, e, Cx)
    Hl = H // 2
    xp = F.pad(xb, (1, 1, 1, 1))
    ref = torch.zeros(N, Co, H, H)
    for a in range(2):
        for b in range(2):
            k = wf[:, a * 2 + b].reshape(Co, 2, 2, Cx).permute(0, 3, 1, 2)
            ref[:, :, a::2, b::2] = F.conv2d(xp[:, :, a: a + Hl + 1, b: b + Hl + 1], k)
    ref = F.relu(ref * scale[None, :, None, None] + shift[None, :, None, None])
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    kw = dict(N=N, H=H, W=H, C_in=Cx, C_x=Cx, C_out=Co, R=3, S=3, stride=1, pad=1, relu=True, upsample=True)
    y_f = ops.conv2d(nhwc(x), pack_weight_folded(w, "cuda"), scale.cuda(), shift.cuda(), flags=CONV_UPS_FOLDED, **kw)
    y_9 = ops.conv2d(nhwc(x), pack_weight(w, "bf16", False, "cuda"), scale.cuda(), shift.cuda(), **kw)
    torch.cuda.synchronize()
    err, rel = report(f"folded up-sample conv {Cx}->{Co} @{H}", to_nchw(y_f), ref)
    assert rel < 1e-2
    d = (to_nchw(y_f) - to_nchw(y_9)).abs().max().item()
    print(f"folded vs nine-tap: max diff {d:.3e} (max |y| {ref.abs().max().item():.3e})")
    assert d <= 2.0 ** -5 * ref.abs().max().item()


@pytest.mark.parametrize("N,H,Cx,Cs,Co", [(2, 64, 128, 64, 64), (1, 128, 64, 64, 32), (3, 32, 64, 64, 32), (1, 64, 64, 128, 64)])
def test_conv_upsample_folded_with_skip(N, H, Cx, Cs, Co):
    """DT_CONV_UPS_FOLDED on the class-fused kernel: folded x operand + nine-tap skip operand, against torch with the same
    bf16 weights and against the unfolded class-fused kernel."""
    g = torch.Generator().manual_seed(N + H + Cx + Cs)
    x = torch.randn(N, Cx, H // 2, H // 2, generator=g)
    skip = torch.randn(N, Cs, H, H, generator=g)
    w = torch.randn(Co, Cx + Cs, 3, 3, generator=g) * (2.0 / ((Cx + Cs) * 9)) ** 0.5
    scale, shift = 1.0 + 0.1 * torch.randn(Co, generator=g), 0.1 * torch.randn(Co, generator=g)
    rb = lambda t: t.to(torch.bfloat16).float()
    wf = rb(fold_upsample_weights(w[:, :Cx]))
    Hl = H // 2
    xp = F.pad(rb(x), (1, 1, 1, 1))
    ref = torch.zeros(N, Co, H, H)
    for a in range(2):
        for b in range(2):
            k = wf[:, a * 2 + b].reshape(Co, 2, 2, Cx).permute(0, 3, 1, 2)
            ref[:, :, a::2, b::2] = F.conv2d(xp[:, :, a: a + Hl + 1, b: b + Hl + 1], k)
    ref = ref + F.conv2d(rb(skip), rb(w[:, Cx:]), None, 1, 1)
    ref = F.relu(ref * scale[None, :, None, None] + shift[None, :, None, None])
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    kw = dict(N=N, H=H, W=H, C_in=Cx + Cs, C_x=Cx, C_out=Co, R=3, S=3, stride=1, pad=1, relu=True, upsample=True,
              skip=nhwc(skip))
    y_f = ops.conv2d(nhwc(x), pack_weight_folded(w, "cuda", Cx), scale.cuda(), shift.cuda(), flags=CONV_UPS_FOLDED, **kw)
    y_9 = ops.conv2d(nhwc(x), pack_weight(w, "bf16", False, "cuda"), scale.cuda(), shift.cuda(), **kw)
    torch.cuda.synchronize()
    err, rel = report(f"folded up-sample+skip conv {Cx}+{Cs}->{Co} @{H}", to_nchw(y_f), ref)
    assert rel < 1e-2
    assert (to_nchw(y_f) - to_nchw(y_9)).abs().max().item() <= 2.0 ** -5 * ref.abs().max().item()


PAIR_CASES = [
    ("pair 128->128 @32", 4, 32, 128, 0, 128, 3, 1, 1, False, False, True),
    ("pair 256->256 @16 +res", 8, 16, 256, 0, 256, 3, 1, 1, False, True, True),
    ("pair 128->256 @32", 3, 32, 128, 0, 256, 3, 1, 1, False, False, False),
    ("pair 256->512 @16 (two channel tiles)", 5, 16, 256, 0, 512, 3, 1, 1, False, True, True),
    ("pair 64->128 @48x48", 2, 48, 64, 0, 128, 3, 1, 1, False, False, True),
    # 8 x 8 images (resnet layer4 of 256 x 256 tiles): two whole images per CTA tile, row-interleaved patch
    ("pair 512->512 @8 x8 images +res", 8, 8, 512, 0, 512, 3, 1, 1, False, True, True),
    ("pair 512->512 @8 x6 images (half-empty last pair)", 6, 8, 512, 0, 512, 3, 1, 1, False, True, True),
    ("pair 256->128 @8 x5 images", 5, 8, 256, 0, 128, 3, 1, 1, False, False, False),
    ("pair 64->256 @8 x1 image", 1, 8, 64, 0, 256, 3, 1, 1, False, False, True),
]


@pytest.mark.parametrize("case", PAIR_CASES, ids=[c[0] for c in PAIR_CASES])
def test_conv_cta_pair(case):
    """tcgen05.mma.cta_group::2 kernel (CTA pairs, M = 256): same K order as the single-CTA halo kernel - identical bits;
    and against torch on the bf16-rounded operands."""
    x, skip, w, scale, shift, residual = make_case(case, seed=3)
    xb, sb, wb, rb = bf16_round(x, skip, w, residual)
    ref = reference(case, xb, sb, wb, scale, shift, rb)
    got_p = run_cuda(case, x, skip, w, scale, shift, residual, dtype=torch.bfloat16, flags=CONV_PAIR)
    got_h = run_cuda(case, x, skip, w, scale, shift, residual, dtype=torch.bfloat16)
    err, rel = report(case[0], got_p, ref)
    assert rel < 1e-2
    if case[2] == 8:
        # 8 x 8 images: without the flag the layer runs on the per-tap kernel (conv_tc.cu), which sums K tap-major instead
        # of slab-major - equal up to the fp32 summation order
        assert (got_p == got_h).float().mean().item() > 0.995 and (got_p - got_h).abs().max() <= 2.0 ** -7 * ref.abs().max()
    else:
        assert torch.equal(got_p, got_h)


def test_stem_tcgen05_and_fp32():
    g = torch.Generator().manual_seed(3)
    N, T, C = 2, 64, 3
    x = torch.randn(N, C, T, T, generator=g)
    w = torch.randn(64, C, 7, 7, generator=g) * (2.0 / (C * 49)) ** 0.5
    scale, shift = 1.0 + 0.1 * torch.randn(64, generator=g), 0.1 * torch.randn(64, generator=g)
    for dtype, tol_rel in ((torch.float32, 1e-5), (torch.bfloat16, 1e-2)):
        rnd = (lambda t: t) if dtype == torch.float32 else (lambda t: t.to(torch.bfloat16).float())
        ref = F.relu(F.conv2d(rnd(x), rnd(w), None, 2, 3) * scale[None, :, None, None] + shift[None, :, None, None])
        x4 = torch.zeros(N, T, T, 4); x4[..., :C] = x.permute(0, 2, 3, 1)
        wp = pack_weight(w, "fp32" if dtype == torch.float32 else "bf16", True, "cuda")
        y = ops.conv2d(x4.to(dtype).cuda(), wp, scale.cuda(), shift.cuda(), N=N, H=T, W=T, C_in=4, C_x=4, C_out=64,
                       R=7, S=7, stride=2, pad=3, relu=True)
        torch.cuda.synchronize()
        err, rel = report(f"stem {dtype}", to_nchw(y), ref)
        assert rel < tol_rel


@pytest.mark.parametrize("N,T", [(3, 64), (2, 256), (5, 32), (3, 512), (1, 1024)])
def test_stem_tma_im2col_padded_input(N, T):
    """7x7/s2 stem with the im2col operand built by TMA from a zero-bordered frame == the gather path, bit for bit;
    for output widths that are multiples of 128 (T >= 256) the default is the row form (operand read in place from the
    raw input rows through no-swizzle descriptors), which must equal the im2col-map kernel (CONV_NO_HALO) bit for bit."""
    g = torch.Generator().manual_seed(6)
    x = torch.randn(N, 3, T, T, generator=g)
    w = torch.randn(64, 3, 7, 7, generator=g) * (2.0 / 147) ** 0.5
    scale, shift = 1.0 + 0.1 * torch.randn(64, generator=g), 0.1 * torch.randn(64, generator=g)
    x4 = torch.zeros(N, T, T, 4); x4[..., :3] = x.permute(0, 2, 3, 1)
    xpad = torch.zeros(N, T + 6, T + 8, 4); xpad[:, 3:3 + T, 3:3 + T] = x4
    wp = pack_weight(w, "bf16", True, "cuda")
    kw = dict(N=N, H=T, W=T, C_in=4, C_x=4, C_out=64, R=7, S=7, stride=2, pad=3, relu=True)
    y_g = ops.conv2d(x4.to(torch.bfloat16).cuda(), wp, scale.cuda(), shift.cuda(), **kw)
    y_t = ops.conv2d(xpad.to(torch.bfloat16).cuda(), wp, scale.cuda(), shift.cuda(), flags=CONV_X_PAD3, **kw)
    torch.cuda.synchronize()
    rnd = lambda t: t.to(torch.bfloat16).float()
    ref = F.relu(F.conv2d(rnd(x), rnd(w), None, 2, 3) * scale[None, :, None, None] + shift[None, :, None, None])
    err, rel = report(f"stem TMA-im2col N={N} T={T}", to_nchw(y_t), ref)
    assert rel < 1e-2
    assert torch.equal(y_t, y_g)
    y_m = ops.conv2d(xpad.to(torch.bfloat16).cuda(), wp, scale.cuda(), shift.cuda(), flags=CONV_X_PAD3 | CONV_NO_HALO, **kw)
    torch.cuda.synchronize()
    assert torch.equal(y_t, y_m)


@pytest.mark.parametrize("N", [1, 3, 40])
def test_stem_pool_fused_equals_stem_then_maxpool(N):
    """conv_stem.cu POOL variant: the max pooling is taken from the staged stem rows in shared memory; both outputs must be
    bit-identical to dt_conv2d_fwd + dt_maxpool3x3s2 (N = 40: several items per CTA, ring reuse across items)."""
    T = 256
    g = torch.Generator().manual_seed(60 + N)
    xpad = torch.zeros(N, T + 6, T + 8, 4)
    xpad[:, 3:3 + T, 3:3 + T, :3] = torch.randn(N, T, T, 3, generator=g)
    w = torch.randn(64, 3, 7, 7, generator=g) * (2.0 / 147) ** 0.5
    scale, shift = (1.0 + 0.1 * torch.randn(64, generator=g)).cuda(), (0.1 * torch.randn(64, generator=g) - 0.05).cuda()
    wp = pack_weight(w, "bf16", True, "cuda")
    xd = xpad.to(torch.bfloat16).cuda()
    y_ref = ops.conv2d(xd, wp, scale, shift, N=N, H=T, W=T, C_in=4, C_x=4, C_out=64, R=7, S=7, stride=2, pad=3, relu=True,
                       flags=CONV_X_PAD3)
    p_ref = ops.maxpool3x3s2(y_ref)
    y = torch.full_like(y_ref, float("nan"))
    pooled = torch.full_like(p_ref, float("nan"))
    ops.stem_pool(xd, wp, scale, shift, N=N, H=T, W=T, out=y, pooled=pooled)
    torch.cuda.synchronize()
    assert torch.equal(y.view(torch.int16), y_ref.view(torch.int16))
    bad = (pooled.view(torch.int16) != p_ref.view(torch.int16))
    if bad.any():
        idx = bad.nonzero()[0].tolist()
        print(f"pooled differs at {int(bad.sum())} of {bad.numel()} elements, first {idx}")
    assert not bad.any()
    ref = F.max_pool2d(y_ref.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    assert torch.equal(pooled.float(), ref)


def test_stem_pool_fused_rejects_other_widths():
    xd = torch.zeros(1, 128 + 6, 128 + 8, 4, dtype=torch.bfloat16, device="cuda")
    wp = pack_weight(torch.zeros(64, 3, 7, 7), "bf16", True, "cuda")
    z = torch.zeros(64, device="cuda")
    with pytest.raises(RuntimeError):
        ops.stem_pool(xd, wp, z, z, N=1, H=128, W=128, out=torch.empty(1, 64, 64, 64, dtype=torch.bfloat16, device="cuda"),
                      pooled=torch.empty(1, 32, 32, 64, dtype=torch.bfloat16, device="cuda"))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_maxpool(dtype):
    g = torch.Generator().manual_seed(4)
    x = torch.randn(3, 64, 34, 34, generator=g).to(dtype).float()
    ref = F.max_pool2d(x, 3, 2, 1)
    y = ops.maxpool3x3s2(x.permute(0, 2, 3, 1).contiguous().to(dtype).cuda())
    assert torch.equal(to_nchw(y), ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("K", [2, 3])
def test_head_outputs(dtype, K):
    g = torch.Generator().manual_seed(5)
    N, H, C = 2, 32, 16
    x = torch.randn(N, C, H, H, generator=g).to(dtype).float()
    w = torch.randn(K, C, 3, 3, generator=g) * 0.1
    b = torch.randn(K, generator=g)
    ref = F.conv2d(x, w, b, 1, 1)
    wp = w.permute(2, 3, 1, 0).reshape(9, C, K).contiguous().cuda()
    nchw = torch.empty(N, K, H, H, device="cuda")
    nhwc = torch.empty(N, H, H, K, dtype=dtype, device="cuda")
    mask = torch.empty(N, H, H, dtype=torch.uint8, device="cuda")
    ops.head(x.permute(0, 2, 3, 1).contiguous().to(dtype).cuda(), wp, b.cuda(), logits_nchw=nchw, logits_nhwc=nhwc, mask=mask)
    err, _ = report(f"head {dtype}", nchw.cpu(), ref)
    assert err < 1e-4
    assert torch.equal(mask.cpu().long(), nchw.cpu().argmax(1))                 # first-max argmax of its own logits
    assert torch.equal(ops.argmax_nchw(nchw).cpu(), mask.cpu())
    assert torch.equal(nhwc.cpu(), nchw.cpu().permute(0, 2, 3, 1).to(dtype))


@pytest.mark.parametrize("K,N,H", [(3, 2, 64), (2, 1, 32), (3, 1, 16)])
def test_head_tensor_core(K, N, H):
    """tcgen05 head (bf16 weights, fp32 accumulate) with the fused argmax / layout epilogue."""
    g = torch.Generator().manual_seed(8)
    C = 16
    x = torch.randn(N, C, H, H, generator=g).to(torch.bfloat16).float()
    w = torch.randn(K, C, 3, 3, generator=g) * 0.1
    b = torch.randn(K, generator=g)
    ref = F.conv2d(x, w.to(torch.bfloat16).float(), b, 1, 1)
    w16 = torch.zeros(16, C, 3, 3); w16[:K] = w
    b16 = torch.zeros(16); b16[:K] = b
    nchw = torch.empty(N, K, H, H, device="cuda")
    nhwc = torch.empty(N, H, H, K, dtype=torch.bfloat16, device="cuda")
    mask = torch.empty(N, H, H, dtype=torch.uint8, device="cuda")
    ops.head_tc(x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda(), pack_weight(w16, "bf16", False, "cuda"),
                b16.cuda(), K, logits_nchw=nchw, logits_nhwc=nhwc, mask=mask)
    torch.cuda.synchronize()
    err, _ = report(f"head tcgen05 K={K}", nchw.cpu(), ref)
    assert err < 1e-4
    assert torch.equal(mask.cpu().long(), nchw.cpu().argmax(1))
    assert torch.equal(nhwc.cpu(), nchw.cpu().permute(0, 2, 3, 1).to(torch.bfloat16))


def test_argmax_ties_first_index():
    z = torch.zeros(1, 3, 4, 4, device="cuda")
    z[0, 1, 0, 0] = 1.0; z[0, 2, 0, 0] = 1.0
    m = ops.argmax_nchw(z).cpu()
    assert m[0, 0, 0] == 1 and m[0, 1, 1] == 0
