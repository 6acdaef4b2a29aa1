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


@pytest.mark.parametrize("cin,n,T,losses", [(4, 2, 64, ["DICE", "FOCAL"]), (3, 3, 32, ["GDICE", "FOCAL"])])
def test_training_step_fp32_matches_autograd(cin, n, T, losses):
    oracle = oracle_model(cin, 3)
    ref_model = copy.deepcopy(oracle)
    img, mask = _batch(n, cin, T, 3)
    ref = ref_train.train_step(ref_model, img, mask, losses=losses, lr=3e-4, clip=0.5)

    seg = _semsegment(oracle, cin, "fp32")
    seg.loss_names = losses
    if "GDICE" in losses:
        from deadtrees_b200.loss.gdl import GeneralizedDiceLoss
        seg.dice_loss = GeneralizedDiceLoss()
    batch = {"main": (img.cuda(), mask.cuda(), None, torch.zeros(n), [{"file": f"t{i}"} for i in range(n)])}
    loss = seg.training_step(batch, 0)
    assert abs(float(loss) - ref["loss"]) < 1e-4 * max(1.0, abs(ref["loss"]))
    loss.backward()
    torch.cuda.synchronize()
    worst = 0.0
    for name, p in seg.model.named_parameters():
        r = ref["grads"][name]
        d = (p.grad.cpu() - r).abs().max().item()
        scale = r.abs().max().item() + 1e-12
        worst = max(worst, d / scale)
        assert d / scale < 2e-3, (name, d, scale)
    print(f"[train fp32] worst relative gradient error {worst:.3e}; loss {float(loss):.6f} vs {ref['loss']:.6f}")
    # running statistics follow nn.BatchNorm2d
    sd_ref = ref_model.state_dict()
    for name, b in seg.model.named_buffers():
        if name.endswith("running_var") or name.endswith("running_mean"):
            assert (b.cpu() - sd_ref[name]).abs().max().item() < 1e-4 * (sd_ref[name].abs().max().item() + 1)
        if name.endswith("num_batches_tracked"):
            assert int(b) == int(sd_ref[name])
    # clip 0.5 + Adam
    opt = FusedAdam(seg.model.parameters(), lr=3e-4, max_grad_norm=0.5)
    opt.step()
    torch.cuda.synchronize()
    for name, p in seg.model.named_parameters():
        r = dict(ref_model.named_parameters())[name].detach()
        assert (p.detach().cpu() - r).abs().max().item() < 2e-5, name


def test_training_step_bf16_tracks_fp32():
    cin, n, T = 4, 2, 64
    oracle = oracle_model(cin, 3)
    ref_model = copy.deepcopy(oracle)
    img, mask = _batch(n, cin, T, 3)
    ref = ref_train.train_step(ref_model, img, mask, lr=0, clip=0)
    seg = _semsegment(oracle, cin, "bf16")
    batch = {"main": (img.cuda(), mask.cuda(), None, torch.zeros(n), [{"file": f"t{i}"} for i in range(n)])}
    loss = seg.training_step(batch, 0)
    loss.backward()
    torch.cuda.synchronize()
    print(f"[train bf16] loss {float(loss):.5f} vs fp32 oracle {ref['loss']:.5f}")
    assert abs(float(loss) - ref["loss"]) < 2e-2
    low = []
    for name, p in seg.model.named_parameters():
        r = ref["grads"][name].flatten().double()
        gq = p.grad.cpu().flatten().double()
        cos = float((gq @ r) / (gq.norm() * r.norm() + 1e-30))
        if r.numel() >= 1024:
            low.append((cos, name))
    low.sort()
    print("[train bf16] lowest gradient cosine similarities:", low[:5])
    assert low[0][0] > 0.98, low[:5]
