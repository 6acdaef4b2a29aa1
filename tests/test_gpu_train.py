"""GPU parity of the training step (S1 / K13 / O1): train-mode BatchNorm, backward kernels, clip + Adam against
torch-CPU autograd of the oracle model and the reference loss terms.

fp32 check mode: gradients within 2e-3 of the largest gradient entry of each tensor (fp32 summation order differs);
bf16 mode: loss within 2e-2 and per-tensor gradient cosine similarity >= 0.98 for the large tensors.
"""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from deadtrees_b200 import ops
from deadtrees_b200.engine import pack_weight
from deadtrees_b200.network.segmodel import SemSegment
from deadtrees_b200.optim import FusedAdam
from gpu_util import oracle_model, report
from oracle import ref_train

pytestmark = pytest.mark.gpu

NETWORK = dict(architecture="unet", encoder_name="resnet34", encoder_depth=5, encoder_weights=None,
               decoder_channels=[256, 128, 64, 32, 16], losses=["DICE", "FOCAL"],
               classes=["bg", "conifer", "broadleaf"], in_channels=4)
TRAINING = dict(learning_rate=3e-4, cosineannealing_tmax=10)


def nhwc(t, dtype=torch.float32):
    return t.permute(0, 2, 3, 1).contiguous().to(dtype).cuda()


def nchw(t):
    return t.float().permute(0, 3, 1, 2).contiguous().cpu()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C,relu,res", [(64, True, False), (16, True, True), (512, False, False)])
def test_bn_train_forward_backward(dtype, C, relu, res):
    g = torch.Generator().manual_seed(C)
    N, H, W = 3, 12, 10
    y = (torch.randn(N, C, H, W, generator=g) * 1.7 + 0.3).to(dtype).float().requires_grad_(True)
    r = torch.randn(N, C, H, W, generator=g).to(dtype).float().requires_grad_(res)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(C, generator=g)).requires_grad_(True)
    rm, rv = torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    rm_ref, rv_ref = rm.clone(), rv.clone()
    z = F.batch_norm(y, rm_ref, rv_ref, gamma, beta, True, 0.1, 1e-5)
    if res:
        z = z + r
    a = F.relu(z) if relu else z
    gout = torch.randn(N, C, H, W, generator=g).to(dtype).float()
    a.backward(gout)

    yd, rm_d, rv_d = nhwc(y.detach(), dtype), rm.cuda(), rv.cuda()
    scale, shift, mean, invstd = ops.bn_train_stats(yd, gamma.detach().cuda(), beta.detach().cuda(), rm_d, rv_d)
    ad = ops.bn_apply(yd, scale, shift, residual=nhwc(r.detach(), dtype) if res else None, relu=relu)
    gy, gz, dgamma, dbeta = ops.bn_train_bwd(nhwc(gout, dtype), ad if relu else None, yd, mean, invstd, scale, want_gz=True)
    torch.cuda.synchronize()
    tol = 1e-5 if dtype == torch.float32 else 2e-2
    assert report("bn a", nchw(ad), a.detach())[1] < tol
    assert report("bn running_mean", rm_d.cpu(), rm_ref)[0] < 1e-5
    assert report("bn running_var", rv_d.cpu(), rv_ref)[1] < 1e-5
    assert report("bn gy", nchw(gy), y.grad)[1] < (1e-4 if dtype == torch.float32 else 3e-2)
    assert report("bn dgamma", dgamma.cpu(), gamma.grad)[1] < (1e-4 if dtype == torch.float32 else 2e-2)
    assert report("bn dbeta", dbeta.cpu(), beta.grad)[1] < (1e-4 if dtype == torch.float32 else 2e-2)
    if res:
        assert report("bn gz", nchw(gz), r.grad)[1] < tol
    if relu and not res:
        # relu(bn(y)) without a residual: the mask recomputed from y (no read of the activation) gives the same bits
        gy2, gz2, dgamma2, dbeta2 = ops.bn_train_bwd(nhwc(gout, dtype), None, yd, mean, invstd, scale, want_gz=True,
                                                     relu_shift=shift)
        torch.cuda.synchronize()
        assert torch.equal(gy2, gy) and torch.equal(gz2, gz) and torch.equal(dgamma2, dgamma) and torch.equal(dbeta2, dbeta)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_maxpool_backward_first_max(dtype):
    g = torch.Generator().manual_seed(3)
    x = F.relu(torch.randn(2, 16, 14, 18, generator=g)).to(dtype).float().requires_grad_(True)  # ReLU zeros: many ties
    out = F.max_pool2d(x, 3, 2, 1)
    gout = torch.randn(out.shape, generator=g).to(dtype).float()
    out.backward(gout)
    add = torch.randn(x.shape, generator=g).to(dtype).float()
    gx = ops.maxpool3x3s2_bwd(nhwc(x.detach(), dtype), nhwc(gout, dtype), addend=nhwc(add, dtype))
    torch.cuda.synchronize()
    ref = (x.grad + add).to(dtype).float() if dtype == torch.bfloat16 else x.grad + add
    assert report("maxpool bwd", nchw(gx), ref)[1] < (1e-6 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("Cx,Cs", [(32, 0), (64, 64), (16, 8)])
def test_upsample_concat_and_backward(dtype, Cx, Cs):
    g = torch.Generator().manual_seed(5)
    xl = torch.randn(2, Cx, 6, 5, generator=g).to(dtype).float().requires_grad_(True)
    sk = torch.randn(2, Cs, 12, 10, generator=g).to(dtype).float().requires_grad_(True) if Cs else None
    up = F.interpolate(xl, scale_factor=2, mode="nearest")
    cat = torch.cat([up, sk], 1) if Cs else up
    gcat = torch.randn(cat.shape, generator=g).to(dtype).float()
    cat.backward(gcat)
    out = ops.upsample_concat(nhwc(xl.detach(), dtype), nhwc(sk.detach(), dtype) if Cs else None)
    glow, gskip = ops.upsample_concat_bwd(nhwc(gcat, dtype), Cx)
    torch.cuda.synchronize()
    assert torch.equal(nchw(out), cat.detach())
    assert report("unconcat low", nchw(glow), xl.grad)[1] < (1e-6 if dtype == torch.float32 else 1e-2)
    if Cs:
        assert torch.equal(nchw(gskip), sk.grad)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cin,cx,cout,k,stride,pad,H", [(16, 16, 32, 3, 1, 1, 12), (64, 64, 128, 3, 2, 1, 16),
                                                        (64, 64, 128, 1, 2, 0, 16), (3, 4, 64, 7, 2, 3, 32),
                                                        (16, 16, 3, 3, 1, 1, 16)])
def test_direct_dgrad_wgrad(dtype, cin, cx, cout, k, stride, pad, H):
    g = torch.Generator().manual_seed(cin * 7 + cout)
    N, W = 2, H + 4
    x = torch.randn(N, cin, H, W, generator=g).to(dtype).float().requires_grad_(True)
    w = (torch.randn(cout, cin, k, k, generator=g) * 0.1).requires_grad_(True)
    y = F.conv2d(x, w, None, stride, pad)
    gy = torch.randn(y.shape, generator=g).to(dtype).float()
    y.backward(gy)
    xs = torch.zeros(N, cx, H, W)
    xs[:, :cin] = x.detach()
    add = torch.randn(N, cx, H, W, generator=g).to(dtype).float()
    gx = ops.conv2d_dgrad_direct(nhwc(gy, dtype), w.detach().cuda(), (N, H, W, cx), stride, pad, addend=nhwc(add, dtype))
    dw, db = ops.conv2d_wgrad_direct(nhwc(xs, dtype), nhwc(gy, dtype), w.shape, stride, pad, want_bias=cout <= 4)
    torch.cuda.synchronize()
    ref_gx = add.clone()
    ref_gx[:, :cin] += x.grad
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert report("dgrad direct", nchw(gx), ref_gx)[1] < tol
    assert report("wgrad direct", dw.cpu(), w.grad)[1] < 1e-4      # fp32 accumulation of the same operands
    if db is not None:
        assert report("bias grad", db.cpu(), gy.sum((0, 2, 3)))[1] < 1e-4


def test_pack_conv_weight_modes():
    g = torch.Generator().manual_seed(9)
    w = torch.randn(32, 16, 3, 3, generator=g)
    ws = torch.randn(64, 3, 7, 7, generator=g)
    assert torch.equal(ops.pack_conv_weight(w.cuda(), 0).cpu(), pack_weight(w, "fp32", False, "cpu"))
    assert torch.equal(ops.pack_conv_weight(w.cuda(), 1).cpu(), pack_weight(w, "bf16", False, "cpu"))
    assert torch.equal(ops.pack_conv_weight(ws.cuda(), 0).cpu(), pack_weight(ws, "fp32", True, "cpu"))
    assert torch.equal(ops.pack_conv_weight(ws.cuda(), 2).cpu(), pack_weight(ws, "bf16", True, "cpu"))
    wt = w.flip(2, 3).permute(1, 0, 2, 3).contiguous()          # dgrad = conv with flipped, transposed weights
    assert torch.equal(ops.pack_conv_weight(w.cuda(), 3).cpu(), pack_weight(wt, "bf16", False, "cpu"))


def _semsegment(oracle, cin, precision):
    net = dict(NETWORK, in_channels=cin, precision=precision)
    seg = SemSegment(net, TRAINING)
    seg.model.load_state_dict(oracle.state_dict())
    return seg.cuda().train()


def _batch(n, cin, T, K, seed=11):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(n, cin, T, T, generator=g)
    yy, xx = torch.meshgrid(torch.arange(T), torch.arange(T), indexing="ij")
    mask = (((yy // 9) + (xx // 7)) % K).long()[None].repeat(n, 1, 1)
    mask[0, : T // 3] = 0
    return img, mask


def _grad_errors(params, ref_grads):
    """per-parameter (relative L2 error, cosine similarity) against the oracle gradients."""
    out = {}
    for name, p in params:
        r = ref_grads[name].double().flatten()
        g = p.grad.detach().cpu().double().flatten()
        out[name] = (float((g - r).norm() / (r.norm() + 1e-30)), float((g @ r) / (g.norm() * r.norm() + 1e-30)))
    return out


@pytest.mark.parametrize("cin,n,T,losses", [(4, 2, 64, ["DICE", "FOCAL"]), (3, 2, 64, ["GDICE", "FOCAL"]),
                                            (4, 2, 64, ["GWDICE", "FOCAL"])])
def test_training_step_fp32_matches_autograd(cin, n, T, losses):
    """fp32 check mode against float64 autograd of the oracle.  Tolerance: torch's own fp32 CPU autograd differs from
    float64 by up to 1.6e-2 (max-abs / max) on these gradients - ReLU masks flip on pre-activations within rounding
    of zero and the small per-channel batches of the deep layers amplify it - so the bar is per-tensor relative
    L2 error < 1e-2 and cosine > 0.9999, with the loss itself within 1e-5."""
    oracle = oracle_model(cin, 3)
    ref_model = copy.deepcopy(oracle).double()
    img, mask = _batch(n, cin, T, 3)
    ref = ref_train.train_step(ref_model, img.double(), mask, losses=losses, lr=0, clip=0)

    seg = _semsegment(oracle, cin, "fp32")
    if "GDICE" in losses:
        from deadtrees_b200.loss.gdl import GeneralizedDiceLoss
        seg.dice_loss = GeneralizedDiceLoss()
    if "GWDICE" in losses:      # segmodel.py:118-124
        from deadtrees_b200.loss.gwdl import GeneralizedWassersteinDiceLoss
        seg.dice_loss = GeneralizedWassersteinDiceLoss(dist_matrix=np.array([[0.0, 1.0, 1.0], [1.0, 0.0, 0.5], [1.0, 0.5, 0.0]]))
    batch = {"main": (img.cuda(), mask.cuda(), None, torch.zeros(n), [{"file": f"t{i}"} for i in range(n)])}
    loss = seg.training_step(batch, 0)
    assert abs(float(loss.detach()) - ref["loss"]) < 1e-5 * max(1.0, abs(ref["loss"]))
    loss.backward()
    torch.cuda.synchronize()
    errs = _grad_errors(seg.model.named_parameters(), ref["grads"])
    worst = max(errs.items(), key=lambda kv: kv[1][0])
    print(f"[train fp32] loss {float(loss.detach()):.7f} vs {ref['loss']:.7f}; worst rel-L2 {worst[1][0]:.3e} cos {worst[1][1]:.6f} ({worst[0]})")
    for name, (rel, cos) in errs.items():
        assert rel < 1e-2 and cos > 0.9999, (name, rel, cos)
    # running statistics follow nn.BatchNorm2d
    sd_ref = ref_model.state_dict()
    for name, b in seg.model.named_buffers():
        if name.endswith("running_var") or name.endswith("running_mean"):
            assert (b.cpu().double() - sd_ref[name]).abs().max().item() < 1e-4 * (sd_ref[name].abs().max().item() + 1)
        if name.endswith("num_batches_tracked"):
            assert int(b) == int(sd_ref[name])


def test_clip_and_adam_step_matches_torch():
    """global-norm clip 0.5 + Adam on the gradients of one training step: same update as clip_grad_norm_ + torch.optim.Adam
    applied to the SAME gradients."""
    cin, n, T = 4, 2, 64
    oracle = oracle_model(cin, 3)
    img, mask = _batch(n, cin, T, 3)
    seg = _semsegment(oracle, cin, "fp32")
    batch = {"main": (img.cuda(), mask.cuda(), None, torch.zeros(n), [{"file": f"t{i}"} for i in range(n)])}
    seg.training_step(batch, 0).backward()
    ref_model = copy.deepcopy(oracle)
    for (name, p), (_, q) in zip(ref_model.named_parameters(), seg.model.named_parameters()):
        p.grad = q.grad.detach().cpu().clone()
    torch.nn.utils.clip_grad_norm_(ref_model.parameters(), 0.5)
    torch.optim.Adam(ref_model.parameters(), lr=3e-4).step()
    FusedAdam(seg.model.parameters(), lr=3e-4, max_grad_norm=0.5).step()
    torch.cuda.synchronize()
    for (name, p), (_, q) in zip(ref_model.named_parameters(), seg.model.named_parameters()):
        assert (q.detach().cpu() - p.detach()).abs().max().item() < 1e-6, name


def test_training_step_bf16_matches_bf16_arithmetic():
    """bf16 mode (tcgen05 forward / dgrad / wgrad) against the CPU restatement with the same bf16 storage points
    (oracle/ref_train.py::train_step_bf16).  Remaining differences are fp32 summation order and the bf16 roundings it
    flips; the distance of BOTH to the float64 step is much larger (cosine ~0.6 at the stem on this random-init net)."""
    cin, n, T = 4, 2, 128
    oracle = oracle_model(cin, 3)
    img, mask = _batch(n, cin, T, 3)
    ref = ref_train.train_step_bf16(copy.deepcopy(oracle), img, mask)
    seg = _semsegment(oracle, cin, "bf16")
    batch = {"main": (img.cuda(), mask.cuda(), None, torch.zeros(n), [{"file": f"t{i}"} for i in range(n)])}
    loss = seg.training_step(batch, 0)
    loss.backward()
    torch.cuda.synchronize()
    errs = _grad_errors(seg.model.named_parameters(), ref["grads"])
    big = {k: v for k, v in errs.items() if ref["grads"][k].numel() >= 1024}
    worst = min(big.items(), key=lambda kv: kv[1][1])
    print(f"[train bf16] loss {float(loss.detach()):.5f} vs bf16 oracle {ref['loss']:.5f}; lowest cosine {worst[1][1]:.5f} ({worst[0]})")
    # whole-step agreement is bounded by the chaos of this random-init net under bf16 storage: two CPU emulations that
    # differ only in accumulation precision (fp32 vs fp64) already diverge to cosine ~0.73 at the stem and 0.45 on the
    # logits (measured); the tight statement is test_training_step_layerwise_teacher_forced
    assert abs(float(loss.detach()) - ref["loss"]) < 5e-3
    assert errs["segmentation_head.0.weight"][1] > 0.999
    assert errs["decoder.blocks.4.conv1.0.weight"][1] > 0.97
    assert worst[1][1] > 0.5, worst


@pytest.mark.parametrize("cin,cout,N,H,W,k,stride", [
    (64, 64, 2, 16, 8, 3, 1), (64, 128, 2, 32, 16, 3, 1), (128, 64, 1, 16, 16, 3, 1), (256, 256, 4, 8, 8, 3, 1),
    (192, 64, 1, 16, 24, 3, 1), (16, 16, 1, 32, 32, 3, 1), (32, 16, 2, 16, 16, 3, 1), (128, 32, 1, 16, 8, 3, 1),
    (512, 512, 2, 8, 8, 3, 1), (64, 128, 2, 16, 16, 3, 2), (256, 512, 2, 8, 8, 3, 2), (128, 256, 1, 16, 8, 3, 2),
    (64, 128, 2, 16, 16, 1, 2), (256, 512, 4, 8, 8, 1, 2),
    # narrow layers (C_out <= 32): the nine-taps-per-MMA kernel with aliased MN blocks (SWIZZLE_32B / 64B / 128B operands)
    (96, 32, 2, 32, 32, 3, 1), (32, 32, 1, 32, 16, 3, 1), (64, 16, 1, 16, 16, 3, 1), (16, 8, 3, 16, 24, 3, 1),
    (192, 32, 1, 16, 16, 3, 1), (16, 32, 2, 48, 8, 3, 1), (24, 24, 1, 16, 8, 3, 1)])
def test_wgrad_tensor_core(cin, cout, N, H, W, k, stride):
    """tcgen05 weight gradient (MN-major operands straight from NHWC; H, W = output size) against torch autograd on the
    bf16-rounded operands: 3x3 stride 1 / 2 and the 1x1 stride-2 downsample."""
    g = torch.Generator().manual_seed(cin + cout + H + k + stride)
    pad = 1 if k == 3 else 0
    x = torch.randn(N, cin, H * stride, W * stride, generator=g).to(torch.bfloat16).float()
    gy = torch.randn(N, cout, H, W, generator=g).to(torch.bfloat16).float()
    w = torch.zeros(cout, cin, k, k, requires_grad=True)
    F.conv2d(x, w, None, stride, pad).backward(gy)
    assert ops.wgrad_tc_supported(nhwc(x, torch.bfloat16), nhwc(gy, torch.bfloat16), w.shape, stride, pad)
    dw = ops.conv2d_wgrad_tc(nhwc(x, torch.bfloat16), nhwc(gy, torch.bfloat16), w.shape, stride)
    torch.cuda.synchronize()
    err, rel = report(f"wgrad tcgen05 {cin}->{cout} k{k}s{stride} N={N} {H}x{W}", dw.cpu(), w.grad)
    assert rel < 1e-4


def test_wgrad_tensor_core_padded_channels():
    """the head's shape: 3 real gradient channels stored in a 16-channel NHWC tensor."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 16, 32, 32, generator=g).to(torch.bfloat16).float()
    gy = torch.randn(2, 3, 32, 32, generator=g).to(torch.bfloat16).float()
    w = torch.zeros(3, 16, 3, 3, requires_grad=True)
    F.conv2d(x, w, None, 1, 1).backward(gy)
    gp = torch.zeros(2, 16, 32, 32)
    gp[:, :3] = gy
    dw = ops.conv2d_wgrad_tc(nhwc(x, torch.bfloat16), nhwc(gp, torch.bfloat16), w.shape, 1)
    torch.cuda.synchronize()
    assert report("wgrad tcgen05 head 16->3", dw.cpu(), w.grad)[1] < 1e-4


@pytest.mark.parametrize("cin,N,T", [(4, 2, 256), (3, 1, 256), (4, 4, 64), (4, 8, 32), (3, 2, 128)])
def test_stem_wgrad_tensor_core(cin, N, T):
    """tcgen05 weight gradient of the 7x7/s2 stem (TMA im2col map as the MN-major B operand) against torch autograd on
    the bf16-rounded operands."""
    g = torch.Generator().manual_seed(cin * 100 + N + T)
    x = torch.randn(N, cin, T, T, generator=g).to(torch.bfloat16).float()
    gy = torch.randn(N, 64, T // 2, T // 2, generator=g).to(torch.bfloat16).float()
    w = torch.zeros(64, cin, 7, 7, requires_grad=True)
    F.conv2d(x, w, None, 2, 3).backward(gy)
    assert ops.stem_wgrad_tc_supported(N, T, T)
    frame = ops.pack_input_nchw_frame(x.cuda(), cin)
    assert frame.shape == (N, T + 6, T + 8, 4)
    assert torch.equal(frame[:, 3:3 + T, 3:3 + T, :cin].float().cpu(), x.permute(0, 2, 3, 1))
    border = frame.clone()
    border[:, 3:3 + T, 3:3 + T, :cin] = 0
    assert not bool(border.any())        # zero border and zero padding channels, written by the kernel
    dw = ops.stem_wgrad_tc(frame, nhwc(gy, torch.bfloat16), w.shape)
    torch.cuda.synchronize()
    err, rel = report(f"stem wgrad tcgen05 cin={cin} N={N} T={T}", dw.cpu(), w.grad)
    assert rel < 1e-4
    dw2 = ops.stem_wgrad_tc(frame, nhwc(gy, torch.bfloat16), w.shape)
    assert torch.equal(dw, dw2)          # fixed-order split reduction: deterministic


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("N,H,W,C,add", [(2, 16, 12, 64, True), (1, 7, 9, 8, False), (3, 32, 32, 16, True)])
def test_maxpool_index_form(dtype, N, H, W, C, add):
    """maxpool 3x3/s2/p1 forward with first-maximum positions + gather backward through them, against torch
    (values in a small set so that ties are frequent: the gradient must go to the FIRST maximum)."""
    g = torch.Generator().manual_seed(N * H + C)
    x = torch.randint(-3, 4, (N, C, H, W), generator=g).float().requires_grad_(True)
    y_ref = F.max_pool2d(x, 3, 2, 1)
    gout = torch.randn(y_ref.shape, generator=g).to(dtype).float()
    y_ref.backward(gout)
    addend = torch.randn(N, C, H, W, generator=g).to(dtype).float() if add else None
    y, idx = ops.maxpool3x3s2_idx(nhwc(x.detach(), dtype))
    gx = ops.maxpool3x3s2_bwd_idx(idx, nhwc(gout, dtype), (N, H, W, C), addend=nhwc(addend, dtype) if add else None)
    old = ops.maxpool3x3s2_bwd(nhwc(x.detach(), dtype), nhwc(gout, dtype), addend=nhwc(addend, dtype) if add else None)
    torch.cuda.synchronize()
    assert torch.equal(nchw(y), y_ref.detach())
    assert int(idx.max()) <= 8
    ref = x.grad + (addend if add else 0)
    tol = 1e-6 if dtype == torch.float32 else 1e-2
    assert report("maxpool idx bwd", nchw(gx), ref)[1] < tol
    assert report("maxpool idx vs recompute kernel", nchw(gx), nchw(old))[1] < tol


@pytest.mark.parametrize("cin,cout,N,H", [(64, 64, 2, 32), (192, 64, 1, 16), (32, 16, 1, 32), (768, 256, 2, 16)])
def test_dgrad_through_forward_kernel(cin, cout, N, H):
    """data gradient of a 3x3/s1 conv = forward tcgen05 conv of gy with the flipped, transposed weights (+ addend)."""
    g = torch.Generator().manual_seed(cin + cout)
    x = torch.zeros(N, cin, H, H, requires_grad=True)
    w = (torch.randn(cout, cin, 3, 3, generator=g) * 0.05)
    gy = torch.randn(N, cout, H, H, generator=g).to(torch.bfloat16).float()
    F.conv2d(x, w.to(torch.bfloat16).float(), None, 1, 1).backward(gy)
    add = torch.randn(N, cin, H, H, generator=g).to(torch.bfloat16).float()
    wp = ops.pack_conv_weight(w.cuda(), 3)
    ones, zeros = torch.ones(cin, device="cuda"), torch.zeros(cin, device="cuda")
    gx = ops.conv2d(nhwc(gy, torch.bfloat16), wp, ones, zeros, N=N, H=H, W=H, C_in=cout, C_x=cout, C_out=cin, R=3, S=3,
                    stride=1, pad=1, relu=False, residual=nhwc(add, torch.bfloat16))
    torch.cuda.synchronize()
    ref = x.grad + add
    err, rel = report(f"dgrad via fwd kernel {cin}<-{cout}", nchw(gx), ref)
    assert rel < 1e-2


def _rel(got, ref):
    ref = ref.double()
    return float((got.double() - ref).abs().max() / (ref.abs().max() + 1e-30))


@pytest.mark.parametrize("precision,n,T", [("bf16", 2, 256), ("fp32", 2, 64)])
def test_training_step_layerwise_teacher_forced(precision, n, T):
    """Every kernel-level operation of ONE real training step (forward and backward, all 47 layers) re-computed by
    torch on the CPU from the operation's own device inputs.  This is the parity statement for the bf16 path: errors
    cannot compound across layers, so each op must agree to fp32-accumulation accuracy (statistics, weight gradients)
    or to one bf16 rounding of the stored result (activations, activation gradients)."""
    cin = 4
    oracle = oracle_model(cin, 3)
    img, mask = _batch(n, cin, T, 3)
    seg = _semsegment(oracle, cin, precision)
    eng = seg.model.train_engine()
    eng.trace = []
    batch = {"main": (img.cuda(), mask.cuda(), None, torch.zeros(n), [{"file": f"t{i}"} for i in range(n)])}
    seg.training_step(batch, 0).backward()
    torch.cuda.synchronize()
    trace, eng.trace = eng.trace, None
    params = {k: v.detach().cpu() for k, v in seg.model.named_parameters()}
    bf = precision == "bf16"
    act_tol = 1e-2 if bf else 1e-5          # one bf16 rounding of the stored tensor
    acc_tol = 2e-4                          # fp32 accumulation order
    rw = (lambda w: w.to(torch.bfloat16).float()) if bf else (lambda w: w)
    cpu = lambda t: None if t is None else nchw(t)
    worst = {}

    def note(kind, v):
        worst[kind] = max(worst.get(kind, 0.0), v)

    kinds = set()
    for kind, name, t in trace:
        kinds.add(kind)
        if kind == "conv_bn_fwd":
            w = params[name + ".weight"]
            x = cpu(t["x"])[:, : w.shape[1]]
            y_ref = F.conv2d(x, rw(w), None, t["stride"], t["pad"])
            note("conv y", _rel(cpu(t["y"]), y_ref)); assert worst["conv y"] < act_tol, name
            y = cpu(t["y"])
            mean, var = y.mean((0, 2, 3)), y.var((0, 2, 3), unbiased=False)
            note("bn mean", float((t["mean"].cpu() - mean).abs().max() / (mean.abs().max() + 1e-6)))
            note("bn invstd", _rel(t["invstd"].cpu(), 1.0 / torch.sqrt(var + 1e-5)))
            assert worst["bn mean"] < 1e-3 and worst["bn invstd"] < 1e-3, name
            z = y * t["scale"].cpu()[None, :, None, None] + t["shift"].cpu()[None, :, None, None]
            if t["residual"] is not None:
                z = z + cpu(t["residual"])
            note("bn apply", _rel(cpu(t["a"]), F.relu(z) if t["relu"] else z)); assert worst["bn apply"] < act_tol, name
        elif kind == "bn_bwd":
            g, y = cpu(t["g"]).double(), cpu(t["y"]).double()
            gz = g * (cpu(t["a"]) > 0) if t["a"] is not None else g
            mean, invstd, scale = (t[k].cpu().double()[None, :, None, None] for k in ("mean", "invstd", "scale"))
            yhat = (y - mean) * invstd
            M = y.numel() / y.shape[1]
            dbeta, dgamma = gz.sum((0, 2, 3)), (gz * yhat).sum((0, 2, 3))
            gy = scale * (gz - dbeta[None, :, None, None] / M - yhat * dgamma[None, :, None, None] / M)
            note("bn dgamma", _rel(t["dgamma"].cpu(), dgamma)); note("bn dbeta", _rel(t["dbeta"].cpu(), dbeta))
            note("bn gy", _rel(cpu(t["gy"]), gy))
            assert worst["bn dgamma"] < acc_tol and worst["bn dbeta"] < acc_tol and worst["bn gy"] < act_tol, name
            if t["gz"] is not None:
                note("bn gz", _rel(cpu(t["gz"]), gz)); assert worst["bn gz"] < act_tol, name
        elif kind == "wgrad":
            w = params[name]
            x = cpu(t["x"])[:, : w.shape[1]]
            gyc = cpu(t["gy"])[:, : w.shape[0]]          # the head's gradient is stored with padded channels
            dw = torch.nn.grad.conv2d_weight(x, w.shape, gyc, t["stride"], t["pad"])
            note("wgrad", _rel(t["dw"].cpu(), dw)); assert worst["wgrad"] < acc_tol, name
            if t["db"] is not None:
                note("bias grad", _rel(t["db"].cpu(), gyc.sum((0, 2, 3)))); assert worst["bias grad"] < acc_tol
        elif kind == "dgrad":
            w = params[name]
            gx_shape = (t["gx"].shape[0], w.shape[1], t["gx"].shape[1], t["gx"].shape[2])
            wr = w if t.get("fp32_weights") else rw(w)
            gx = torch.nn.grad.conv2d_input(gx_shape, wr, cpu(t["gy"])[:, : w.shape[0]], t["stride"], t["pad"])
            if t["addend"] is not None:
                gx = gx + cpu(t["addend"])[:, : w.shape[1]]
            note("dgrad", _rel(cpu(t["gx"])[:, : w.shape[1]], gx)); assert worst["dgrad"] < act_tol, name
        elif kind == "maxpool_fwd":
            assert torch.equal(cpu(t["y"]), F.max_pool2d(cpu(t["x"]), 3, 2, 1))
        elif kind == "maxpool_bwd":
            x = cpu(t["x"]).requires_grad_(True)
            F.max_pool2d(x, 3, 2, 1).backward(cpu(t["gout"]))
            ref = x.grad + (cpu(t["addend"]) if t["addend"] is not None else 0)
            note("maxpool bwd", _rel(cpu(t["gx"]), ref)); assert worst["maxpool bwd"] < act_tol
        elif kind == "upcat_fwd":
            up = F.interpolate(cpu(t["x"]), scale_factor=2, mode="nearest")
            assert torch.equal(cpu(t["y"]), torch.cat([up, cpu(t["skip"])], 1) if t["skip"] is not None else up)
        elif kind == "upcat_bwd":
            gc, cx = cpu(t["g_cat"]), t["cx"]
            note("unconcat", _rel(cpu(t["g_low"]), F.avg_pool2d(gc[:, :cx], 2) * 4)); assert worst["unconcat"] < act_tol
            if t["g_skip"] is not None:
                assert torch.equal(cpu(t["g_skip"]), gc[:, cx:])
        elif kind == "head_fwd":
            # bf16 path: the head runs on the tensor cores like every other layer - bf16 weight operands, fp32 accumulation
            hw = params["segmentation_head.0.weight"]
            if precision == "bf16":
                hw = hw.to(torch.bfloat16).float()
            ref = F.conv2d(cpu(t["x"]), hw, params["segmentation_head.0.bias"], 1, 1)
            note("head", _rel(t["y"].cpu(), ref)); assert worst["head"] < acc_tol
    assert kinds == {"conv_bn_fwd", "bn_bwd", "wgrad", "dgrad", "maxpool_fwd", "maxpool_bwd", "upcat_fwd", "upcat_bwd", "head_fwd"}
    assert sum(1 for k, _, _ in trace if k == "wgrad") == 47 and sum(1 for k, _, _ in trace if k == "dgrad") == 46
    print(f"[train {precision} layerwise T={T}] worst relative errors per op kind: " +
          ", ".join(f"{k}={v:.2e}" for k, v in sorted(worst.items())))


def test_flat_adam_path_equals_per_parameter_path():
    """`configure_optimizers` (flat clip + Adam over the engine's buffers) updates exactly like the per-parameter launches."""
    cin, n, T = 4, 2, 64
    oracle = oracle_model(cin, 3)
    img, mask = _batch(n, cin, T, 3)
    batch = {"main": (img.cuda(), mask.cuda(), None, torch.zeros(n), [{"file": f"t{i}"} for i in range(n)])}
    segs = []
    for flat in (False, True):
        net = dict(NETWORK, in_channels=cin, precision="fp32")
        seg = SemSegment(net, dict(TRAINING, gradient_clip_val=0.5))
        seg.model.load_state_dict(oracle.state_dict())
        seg.cuda().train()
        if flat:
            (opt,), _ = seg.configure_optimizers()
        else:
            opt = FusedAdam(seg.model.parameters(), lr=3e-4, max_grad_norm=0.5)
        for _ in range(2):
            seg.training_step(batch, 0).backward()
            opt.step()
        segs.append(seg)
    torch.cuda.synchronize()
    for (name, p), (_, q) in zip(segs[0].model.named_parameters(), segs[1].model.named_parameters()):
        # two Adam steps of 3e-4; fp32 atomics in the generic wgrad make the gradients differ in the last bits, and Adam's
        # m / sqrt(v) turns a last-bit difference of a near-zero gradient into a visible one: single entries may move by a
        # fraction of the step (bounded by 2 * lr), the tensors as a whole must coincide
        d = (p.detach() - q.detach()).abs()
        assert d.max().item() < 6.5e-4 and d.mean().item() < 2e-6, name
    # the flattened parameters still serve the inference engine
    segs[1].eval()
    with torch.no_grad():
        out = segs[1].model(img.cuda())
    assert out.shape == (n, 3, T, T) and torch.isfinite(out).all()


@pytest.mark.parametrize("losses,ramped", [(["DICE", "BOUNDARY", "FOCAL"], False), (["GDICE", "BOUNDARY-RAMPED", "FOCAL"], True)])
def test_training_step_with_boundary_loss(losses, ramped):
    """the reference's default loss family adds the boundary (surface) loss on the dataloader's distance maps
    (losses.py:250-270, segmodel.py:188-191): fp32 check mode, loss terms and gradients against oracle autograd."""
    cin, n, T = 4, 2, 64
    oracle = oracle_model(cin, 3)
    img, mask = _batch(n, cin, T, 3)
    g = torch.Generator().manual_seed(5)
    distmap = torch.randn(n, 3, T, T, generator=g) * 4.0          # signed distances (one_hot2dist), any real values
    ref_model = copy.deepcopy(oracle).double()
    ref = ref_train.train_step(ref_model, img.double(), mask, losses=losses, lr=0.0, clip=0.0, distmap=distmap.double(),
                               alpha=0.01)
    net = dict(NETWORK, in_channels=cin, precision="fp32", losses=losses)
    seg = SemSegment(net, TRAINING)
    seg.model.load_state_dict(oracle.state_dict())
    seg.cuda().train()
    batch = {"main": (img.cuda(), mask.cuda(), distmap.cuda(), torch.zeros(n), [{"file": f"t{i}"} for i in range(n)])}
    loss = seg.training_step(batch, 0)
    loss.backward()
    torch.cuda.synchronize()
    bd = float(seg.logged["train/boundary_loss"])
    print(f"[boundary] loss {float(loss.detach()):.6f} vs {ref['loss']:.6f}; boundary term {bd:.6f} vs {ref['terms']['boundary_loss']:.6f}")
    assert abs(bd - ref["terms"]["boundary_loss"]) < 1e-5 * max(1.0, abs(ref["terms"]["boundary_loss"]))
    assert abs(float(loss.detach()) - ref["loss"]) < 1e-4
    errs = _grad_errors(seg.model.named_parameters(), ref["grads"])
    worst = max(errs.items(), key=lambda kv: kv[1][0])
    print(f"[boundary] worst relative L2 gradient error {worst[1][0]:.3e} ({worst[0]})")
    assert worst[1][0] < 2e-2 and min(v[1] for v in errs.values()) > 0.999


@pytest.mark.parametrize("loss_names", [["DICE", "FOCAL"], ["GWDICE", "FOCAL"]])
def test_graphed_training_step_equals_eager(loss_names):
    """the whole step replayed from one CUDA graph (deadtrees_b200/train_graph.py) against the same step launched kernel
    by kernel: three steps, the graph fed once through __call__ and twice through the prefetch pipeline.  The losses
    must agree to the last bit; the parameters to 1e-6 (the gradient norm of the clip and the loss partials are double
    atomic sums of fixed per-block partials: the last bit of a double may differ between two runs)."""
    from deadtrees_b200.train_graph import GraphedTrainStep
    cin, n, T = 4, 2, 128
    oracle = oracle_model(cin, 3)
    img, mask = _batch(n, cin, T, 3)
    img2, mask2 = _batch(n, cin, T, 3, seed=12)
    batches = [(img, mask), (img2, mask2), (img, mask)]
    stats = [{"file": f"t{i}"} for i in range(n)]
    segs, losses = [], []
    for graphed in (False, True):
        seg = SemSegment(dict(NETWORK, in_channels=cin, precision="bf16", losses=loss_names), dict(TRAINING, gradient_clip_val=0.5))
        seg.model.load_state_dict(oracle.state_dict())
        seg.cuda().train()
        (opt,), _ = seg.configure_optimizers()
        ls = []
        if graphed:
            gs = GraphedTrainStep(seg, opt, n, T)
            pinned = [(a.pin_memory(), b.pin_memory()) for a, b in batches]
            ls.append(float(gs(*pinned[0])))
            gs.prefetch(*pinned[1])
            ls.append(float(gs.step_prefetched()))
            gs.prefetch(*pinned[2])
            ls.append(float(gs.step_prefetched()))
            assert gs.check() == ls[-1]
        else:
            for a, b in batches:
                loss = seg.training_step({"main": (a.cuda(), b.cuda(), None, torch.zeros(n), stats)}, 0)
                loss.backward()
                opt.step()
                ls.append(float(loss.detach()))
        torch.cuda.synchronize()
        segs.append(seg)
        losses.append(ls)
    print("losses eager", losses[0], "graph", losses[1])
    assert losses[0] == losses[1]
    for (name, p), (_, q) in zip(segs[0].model.named_parameters(), segs[1].model.named_parameters()):
        assert (p.detach() - q.detach()).abs().max().item() < 1e-6, name
    for (name, p), (_, q) in zip(segs[0].model.named_buffers(), segs[1].model.named_buffers()):
        assert (p.double() - q.double()).abs().max().item() < 1e-6, name


@pytest.mark.parametrize("cin,cout,k,N,H", [(64, 128, 3, 2, 32), (256, 512, 3, 2, 16), (128, 256, 1, 1, 32), (64, 128, 1, 2, 16)])
def test_dgrad_stride2_gather_kernel(cin, cout, k, N, H):
    """data gradient of the stride-2 convs (3x3 pad 1, 1x1 pad 0) through the tcgen05 gather producer (DT_CONV_TRANSPOSED)."""
    from deadtrees_b200._lib import CONV_TRANSPOSED
    g = torch.Generator().manual_seed(cin + cout + k)
    pad = 1 if k == 3 else 0
    x = torch.zeros(N, cin, H, H, requires_grad=True)
    w = torch.randn(cout, cin, k, k, generator=g) * 0.05
    gy = torch.randn(N, cout, H // 2, H // 2, generator=g).to(torch.bfloat16).float()
    F.conv2d(x, w.to(torch.bfloat16).float(), None, 2, pad).backward(gy)
    add = torch.randn(N, cin, H, H, generator=g).to(torch.bfloat16).float()
    wp = ops.pack_conv_weight(w.cuda(), 4)
    ones, zeros = torch.ones(cin, device="cuda"), torch.zeros(cin, device="cuda")
    gx = ops.conv2d(nhwc(gy, torch.bfloat16), wp, ones, zeros, N=N, H=H, W=H, C_in=cout, C_x=cout, C_out=cin, R=k, S=k,
                    stride=2, pad=pad, relu=False, residual=nhwc(add, torch.bfloat16), flags=CONV_TRANSPOSED)
    torch.cuda.synchronize()
    assert gx.shape == (N, H, H, cin)
    err, rel = report(f"dgrad s2 k{k} {cin}<-{cout}", nchw(gx), x.grad + add)
    assert rel < 1e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_channel_sum_is_reproducible(dtype):
    """dt_channel_sum (bias gradient of the head): two fixed-order stages - the same bits on every call (its float atomics
    were the one run-to-run difference of a training step and, through the clip's gradient norm, could tip a whole
    trajectory) - and the float64 sum within fp32 rounding."""
    g = torch.Generator().manual_seed(8)
    x = (torch.randn(3 * 128 * 128, 8, generator=g) * 2.0).to(dtype).cuda()
    ref = x.double().sum(dim=0)[:3].cpu()
    outs = [ops.channel_sum(x, 3).cpu() for _ in range(5)]
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    assert (outs[0].double() - ref).abs().max().item() < 1e-3 * max(1.0, ref.abs().max().item()) * 1e-2


def test_batched_weight_packing_equals_single_layer_packing():
    """dt_pack_conv_weights_batched (row-staged, every layer of the step in one launch) against dt_pack_conv_weight layer by
    layer: every layout mode, 1x1 and 3x3 and 7x7, padded gradient strides, a row too long for the staging buffer, and a
    refresh after the master weights changed.  Pure data movement + one bf16 rounding: bit-exact."""
    g = torch.Generator().manual_seed(17)
    packer = ops.WeightPacker(torch.device("cuda"))
    specs = [((64, 4, 7, 7), 2, None), ((64, 4, 7, 7), 0, None), ((64, 64, 3, 3), 1, None), ((64, 64, 3, 3), 3, None),
             ((128, 64, 3, 3), 4, None), ((128, 64, 1, 1), 1, None), ((128, 64, 1, 1), 4, None), ((256, 768, 3, 3), 1, None),
             ((256, 768, 3, 3), 3, None), ((16, 32, 3, 3), 3, 16), ((3, 16, 3, 3), 3, 16), ((3, 16, 3, 3), 1, None),
             ((8, 1024, 3, 3), 1, None), ((1024, 8, 3, 3), 3, None), ((32, 96, 3, 3), 0, None)]
    ws = [torch.randn(shape, generator=g).cuda() for shape, _, _ in specs]
    outs = [packer.get((i, mode, pad), w, mode, cout_pad=pad) for i, (w, (_, mode, pad)) in enumerate(zip(ws, specs))]
    for rnd in range(2):
        with torch.no_grad():
            for w in ws:
                w.mul_(1.5).add_(0.01 * (rnd + 1))
        packer.refresh()
        torch.cuda.synchronize()
        for w, out, (shape, mode, pad) in zip(ws, outs, specs):
            ref = ops.pack_conv_weight(w, mode, cout_pad=pad)
            assert out.shape == ref.shape and out.dtype == ref.dtype
            assert torch.equal(out.view(torch.int16) if out.dtype == torch.bfloat16 else out,
                               ref.view(torch.int16) if ref.dtype == torch.bfloat16 else ref), (shape, mode, pad)


# ---- state that raw-pointer kernels / graph replays change behind torch's back ---------------------------
def _small_seg(cin=4, precision="bf16", clip=0.5):
    oracle = oracle_model(cin, 3)
    seg = SemSegment(dict(NETWORK, in_channels=cin, precision=precision), dict(TRAINING, gradient_clip_val=clip))
    seg.model.load_state_dict(oracle.state_dict())
    return seg.cuda()


def test_eval_engine_follows_graph_replays():
    """eval -> graph replays -> eval: the replays move weights and running statistics through raw pointers (no tensor
    version changes), the cached inference engine must still be rebuilt from the NEW state."""
    from deadtrees_b200.engine import UnetEngine
    from deadtrees_b200.train_graph import GraphedTrainStep
    cin, n, T = 4, 2, 64
    seg = _small_seg(cin)
    img, mask = _batch(n, cin, T, 3)
    seg.eval()
    with torch.no_grad():
        before = seg.model(img.cuda()).clone()
    seg.train()
    (opt,), _ = seg.configure_optimizers()
    gs = GraphedTrainStep(seg, opt, n, T)
    seg.eval()
    with torch.no_grad():
        restored = seg.model(img.cuda()).clone()
    assert torch.equal(restored, before)          # capture + warm-up restore every piece of state
    seg.train()
    for _ in range(3):
        gs(img.pin_memory(), mask.pin_memory())
    seg.eval()
    with torch.no_grad():
        after = seg.model(img.cuda()).clone()
    fresh = UnetEngine(seg.model.state_dict(), cin, 3, precision="bf16")
    want = fresh.forward(ops.pack_input_nchw(img.cuda(), cin, torch.bfloat16), want_logits_nchw=True)["logits_nchw"]
    torch.cuda.synchronize()
    assert not torch.equal(after, before), "eval after training still used the engine folded from the old weights"
    assert torch.equal(after, want)


def test_graphed_step_follows_lr_changes():
    """a scheduler changes param_groups[0]['lr'] between steps (CosineAnnealingLR, segmodel.py:420-429): the replayed
    graph must use the new rate exactly like the eager step."""
    from deadtrees_b200.train_graph import GraphedTrainStep
    cin, n, T = 4, 2, 64
    img, mask = _batch(n, cin, T, 3)
    stats = [{"file": f"t{i}"} for i in range(n)]
    lrs = [3e-4, 1e-3, 5e-5]
    segs = []
    for graphed in (False, True):
        seg = _small_seg(cin).train()
        (opt,), (sch,) = seg.configure_optimizers()
        gs = GraphedTrainStep(seg, opt, n, T) if graphed else None
        for lr in lrs:
            opt.param_groups[0]["lr"] = lr
            if graphed:
                gs(img.pin_memory(), mask.pin_memory())
            else:
                seg.training_step({"main": (img.cuda(), mask.cuda(), None, torch.zeros(n), stats)}, 0).backward()
                opt.step()
        torch.cuda.synchronize()
        segs.append(seg)
    moved = 0.0
    ref = dict(oracle_model(cin, 3).named_parameters())
    for (name, p), (_, q) in zip(segs[0].model.named_parameters(), segs[1].model.named_parameters()):
        assert (p.detach() - q.detach()).abs().max().item() < 1e-6, name
        moved = max(moved, (p.detach().cpu() - ref[name].detach()).abs().max().item())
    assert moved > 1e-3          # the 1e-3 step is visible: a rate frozen at 3e-4 would move every weight less than 1e-3


def test_weight_packer_follows_flattened_parameters():
    """a train-mode forward BEFORE the optimizer is attached registers kernel layouts at the parameters' old addresses;
    attach_engine() moves the parameters into the flat buffer - the repack must read the new storage."""
    cin, n, T = 4, 2, 64
    img, mask = _batch(n, cin, T, 3)
    stats = [{"file": f"t{i}"} for i in range(n)]
    batch = {"main": (img.cuda(), mask.cuda(), None, torch.zeros(n), stats)}
    segs = []
    for early_forward in (False, True):
        seg = _small_seg(cin).train()
        if early_forward:
            snap = {k: v.clone() for k, v in seg.model.state_dict().items()}
            seg.training_step(batch, 0)                   # registers the layouts; restore the BN statistics it moved
            seg.model.load_state_dict(snap)
        (opt,), _ = seg.configure_optimizers()
        for _ in range(3):
            seg.training_step(batch, 0).backward()
            opt.step()
        torch.cuda.synchronize()
        segs.append(seg)
    for (name, p), (_, q) in zip(segs[0].model.named_parameters(), segs[1].model.named_parameters()):
        assert (p.detach() - q.detach()).abs().max().item() < 1e-6, name


def test_optimizer_survives_rebucketing_and_refuses_a_rebuilt_engine():
    cin, n, T = 4, 2, 64
    img, mask = _batch(n, cin, T, 3)
    batch = {"main": (img.cuda(), mask.cuda(), None, torch.zeros(n), [{"file": "a"}, {"file": "b"}])}
    seg = _small_seg(cin).train()
    (opt,), _ = seg.configure_optimizers()
    eng = seg.model.train_engine()
    flat = eng.reducer.flat
    eng.set_process_group(None, world_size=1, bucket_bytes=1 << 20)      # after configure_optimizers
    assert eng.reducer.flat.data_ptr() == flat.data_ptr() and len(eng.reducer.buckets) > 4
    before = seg.model.encoder.conv1.weight.detach().clone()
    seg.training_step(batch, 0).backward()
    opt.step()
    torch.cuda.synchronize()
    assert (seg.model.encoder.conv1.weight.detach() - before).abs().max().item() > 0      # stepped on live gradients
    seg.model.set_precision("fp32")                                       # rebuilds the train engine
    seg.training_step(batch, 0).backward()
    with pytest.raises(RuntimeError, match="rebuilt"):
        opt.step()
