"""CPU tests of the host-side mirror of the reference interface (no compute calls)."""
import math
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deadtrees_b200.deployment.inference import MosaicInference, overlap_grid, blend_window
from deadtrees_b200.deployment.tiler import Tiler, TileInfo, divisible_without_remainder, inspect_array, inspect_tile
from deadtrees_b200.engine import conv_flops_per_tile, fold_bn, pack_weight
from deadtrees_b200.network.segmodel import Conf, SemSegment, create_combined_batch, to_conf
from deadtrees_b200.sharding import make_halo_hook, split_tile_rows
from oracle import ref_tiler, ref_unet

NETWORK = dict(architecture="unet", encoder_name="resnet34", encoder_depth=5, encoder_weights=None,
               decoder_channels=[256, 128, 64, 32, 16], losses=["DICE", "FOCAL"],
               classes=["bg", "conifer", "broadleaf"], in_channels=3)
TRAINING = dict(learning_rate=3e-4, cosineannealing_tmax=10)


@pytest.mark.parametrize("a,b,result", [(10, 2, True), (5, 4, False), (2, 0, False)])
def test_divisible_without_remainder(a, b, result):  # tests/test_tiler.py:51-53
    assert divisible_without_remainder(a, b) == result


@pytest.mark.parametrize("size,expect", [((8192, 8192), (16, 16)), ((8192, 7433), (16, 15)), ((2649, 8192), (6, 16))])
def test_inspect(size, expect):
    info = inspect_array(size)
    assert info == TileInfo(size=size, subtiles=expect)
    assert inspect_tile(np.zeros((1,) + size, dtype=np.uint8)[:, :1, :1].repeat(1, 0)).size == (1, 1)


def test_tiler_errors():
    with pytest.raises(ValueError):
        Tiler(tile_shape=(8192, 8192), subtile_shape=(256, 250))  # tests/test_tiler.py:113-115
    with pytest.raises(ValueError):
        inspect_array((100, 100), subtile_shape=(512, 211))
    t = Tiler(tile_shape=(64, 64), subtile_shape=(16, 16))
    t.load_array(np.zeros((4, 40, 33), dtype=np.uint8))
    assert t._subtiles_to_use.sum() == 3 * 3 and t._indata.shape == (4, 64, 64)


def test_semsegment_constructor_contract():
    m = SemSegment(NETWORK, TRAINING)
    sd = m.state_dict()
    assert len(sd) == 278 and next(iter(sd)) == "model.encoder.conv1.weight"
    assert list(m.parameters())[0].shape[1] == 3
    assert set(sd) == {"model." + k for k in ref_unet.Unet(in_channels=3, classes=3).state_dict()}
    assert m.classes_int_wout_bg == [1, 2] and m.hparams.training.learning_rate == 3e-4
    with pytest.raises(NotImplementedError):
        SemSegment(dict(NETWORK, architecture="fancynet"), TRAINING)
    with pytest.raises(NotImplementedError):
        SemSegment(dict(NETWORK, architecture="resunet++"), TRAINING)     # depthwise / scSE forks: outside the hot path
    # Unet++ (segmodel.py:63-64): smp.UnetPlusPlus's parameter set, key for key
    from oracle import ref_unetpp
    pp = SemSegment(dict(NETWORK, architecture="UnetPlusPlus"), TRAINING)
    assert list(pp.state_dict()) == ["model." + k for k in ref_unetpp.UnetPlusPlus(3, 3).state_dict()]
    assert sum(p.numel() for p in pp.parameters()) == 26078899
    with pytest.raises(AssertionError):
        SemSegment(dict(NETWORK, losses=["GDICE", "DICE"]), TRAINING)
    with pytest.raises(NotImplementedError):
        SemSegment(dict(NETWORK, losses=["DICE", "BANANA"]), TRAINING)
    with pytest.raises(AssertionError):
        SemSegment(dict(NETWORK, losses=["FOCAL"]), TRAINING)  # a dice-type loss is required
    # GWDICE: the reference's class distances (segmodel.py:118-124), cut to 2 x 2 for two classes (what its never-true
    # `self.classes_int == 2` test intends); as in the reference the last dice-type entry of the list wins
    gw = SemSegment(dict(NETWORK, losses=["GWDICE", "FOCAL"]), TRAINING)
    assert gw.dice_loss.matrix() == [[0.0, 1.0, 1.0], [1.0, 0.0, 0.5], [1.0, 0.5, 0.0]] and gw._dice_mode() == 0
    assert SemSegment(dict(NETWORK, classes=["a", "b"], losses=["GWDICE"]), TRAINING).dice_loss.matrix() == [[0.0, 1.0], [1.0, 0.0]]
    assert SemSegment(dict(NETWORK, losses=["GWDICE", "DICE"]), TRAINING)._dice_mode() == 1


def test_checkpoint_roundtrip(tmp_path):
    m = SemSegment(dict(NETWORK, in_channels=4), TRAINING)
    p = tmp_path / "m.ckpt"
    m.save_checkpoint(p)
    ck = torch.load(p, weights_only=False)
    assert set(ck) >= {"state_dict", "hyper_parameters"} and ck["hyper_parameters"]["network"]["in_channels"] == 4
    m2 = SemSegment.load_from_checkpoint(p)
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)
    from deadtrees_b200.deployment.inference import PyTorchInference
    with pytest.raises(ValueError):
        PyTorchInference(tmp_path / "m.onnx")


def test_create_combined_batch():
    mk = lambda n: (torch.zeros(n, 3, 4, 4), torch.zeros(n, 4, 4), torch.zeros(n, 3, 4, 4), torch.zeros(n, 4, 4),
                    [{"file": f"f{i}"} for i in range(n)])
    img, mask, dist_, lu, stats = create_combined_batch({"main": mk(2), "extra_a": mk(3)})
    assert img.shape[0] == 5 and len(stats) == 5


def test_conf_access():
    c = to_conf({"a": 1, "b": {"c": 2}})
    assert c.a == 1 and c.b.c == 2
    d = c.copy(); del d.a
    assert "a" in c and "a" not in d


def test_pack_weight_layouts():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(32, 16, 3, 3, generator=g)
    p32 = pack_weight(w, "fp32", False, "cpu")
    assert p32.shape == (9, 16, 32) and torch.equal(p32[4, 5, 7], w[7, 5, 1, 1])
    pb = pack_weight(w, "bf16", False, "cpu")
    assert pb.shape == (32, 192) and pb.dtype == torch.bfloat16          # K = 144 -> padded to 192
    assert torch.equal(pb[7, (1 * 3 + 2) * 16 + 5], w[7, 5, 1, 2].to(torch.bfloat16)) and float(pb[:, 144:].abs().max()) == 0
    ws = torch.randn(64, 3, 7, 7, generator=g)
    ps = pack_weight(ws, "bf16", True, "cpu")
    assert ps.shape == (64, 256)
    assert torch.equal(ps[9, 2 * 32 + 5 * 4 + 1], ws[9, 1, 2, 5].to(torch.bfloat16))
    assert float(ps[:, 224:].abs().max()) == 0 and float(ps.view(64, 8, 8, 4)[:, :7, 7].abs().max()) == 0  # s = 7 column
    assert float(ps.view(64, 8, 8, 4)[:, :7, :7, 3].abs().max()) == 0                                        # padded channel


def test_fold_bn_matches_batchnorm_eval():
    bn = torch.nn.BatchNorm2d(8).eval()
    g = torch.Generator().manual_seed(0)
    bn.weight.data = torch.randn(8, generator=g); bn.bias.data = torch.randn(8, generator=g)
    bn.running_mean = torch.randn(8, generator=g); bn.running_var = torch.rand(8, generator=g) + 0.5
    sd = {"b." + k: v for k, v in bn.state_dict().items()}
    scale, shift = fold_bn(sd, "b", 8, "cpu")
    x = torch.randn(2, 8, 5, 5, generator=g)
    torch.testing.assert_close(x * scale[None, :, None, None] + shift[None, :, None, None], bn(x), rtol=1e-5, atol=1e-5)


def test_flops_match_survey():
    assert abs(conv_flops_per_tile(256, 3, 3) / 1e9 - 15.6657) < 1e-3
    assert abs(conv_flops_per_tile(256, 4, 3) / 1e9 - 15.7685) < 1e-3
    assert abs(conv_flops_per_tile(1024, 3, 3) / 1e9 - 250.65) < 1e-2


def test_grid_window_match_oracle():
    for H, W, T, ov in [(10000, 10000, 256, 32), (10000, 10000, 256, 0), (300, 200, 64, 16), (64, 64, 64, 8)]:
        assert overlap_grid(H, W, T, ov) == ref_tiler.overlap_grid(H, W, T, ov)[:2]
    np.testing.assert_array_equal(blend_window(256, 32, "cpu").numpy(), ref_tiler.blend_window(256, 32))


def test_split_tile_rows_partition():
    assert [b - a for a, b in split_tile_rows(45, 8)] == [6, 6, 6, 6, 6, 5, 5, 5]
    for gy in (1, 3, 40, 45):
        for ws in (1, 2, 4, 8):
            parts = split_tile_rows(gy, ws)
            assert parts[0][0] == 0 and parts[-1][1] == gy and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            rows = [MosaicInference.owned_rows(10000, 256, 32, 45, a, b) for a, b in split_tile_rows(45, ws)]
            assert rows[0][0] == 0 and rows[-1][1] == 10000 and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))


def _halo_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T, ov, gx, K, gy = 16, 4, 3, 2, 5
    parts = split_tile_rows(gy, world)
    full = torch.arange(gy * gx * T * T * K, dtype=torch.float32).reshape(gy * gx, T, T, K)
    r0, r1 = parts[rank]
    halo = 1 if r0 > 0 else 0
    local = torch.full(((r1 - r0 + halo) * gx, T, T, K), -1.0)
    local[halo * gx:] = full[r0 * gx: r1 * gx]
    make_halo_hook(T, ov, rank, world, has_rows=[b > a for a, b in parts])(local, gx, halo)
    if halo:
        expect = full[(r0 - 1) * gx: r0 * gx, T - ov:]
        assert torch.equal(local[:gx, T - ov:], expect), rank
        assert float(local[:gx, : T - ov].max()) == -1.0
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world, tmp_path):
    """the N>1 path on CPU: every shard receives the previous shard's boundary logits rows."""
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_halo_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)


# ---- data-parallel training: bucketed gradient all-reduce (deadtrees_b200/parallel.py) ----------------------------

def _param_shapes():
    """parameter names / shapes of the Unet without building tensors on a GPU."""
    from deadtrees_b200.network.unet import Unet
    m = Unet(in_channels=4, classes=3)
    return [(n, p.shape) for n, p in m.named_parameters()]


def test_backward_param_order_and_buckets():
    from deadtrees_b200.parallel import GradBucketReducer, backward_param_order
    shapes = dict(_param_shapes())
    order = backward_param_order(list(shapes))
    assert sorted(order) == sorted(shapes) and len(order) == len(shapes)
    assert order[0].startswith("segmentation_head") and order[-1] in ("encoder.conv1.weight", "encoder.bn1.weight", "encoder.bn1.bias")
    first = {k: i for i, k in enumerate(order)}
    assert first["decoder.blocks.4.conv2.0.weight"] < first["decoder.blocks.4.conv1.0.weight"] < first["decoder.blocks.0.conv1.0.weight"]
    assert first["decoder.blocks.0.conv1.0.weight"] < first["encoder.layer4.2.conv2.weight"] < first["encoder.layer4.0.downsample.0.weight"]
    assert first["encoder.layer4.0.downsample.0.weight"] < first["encoder.layer3.5.conv2.weight"] < first["encoder.layer1.0.conv1.weight"]
    red = GradBucketReducer([(n, shapes[n]) for n in order], "cpu", bucket_bytes=25 << 20, world_size=1)
    assert red.flat.numel() >= sum(s.numel() for s in shapes.values())
    assert len(red.buckets) == 5                      # 24.44 M fp32 parameters (97.8 MB) in buckets of at most 25 MiB
    for n in order:                                   # every slot 16-byte aligned, disjoint, inside its bucket
        assert red.flat_offset(n) % 4 == 0
    red.begin()
    for n in order:
        red.view(n).fill_(1.0)
        red.mark(n)
    assert red.launched == [0, 1, 2, 3, 4]            # buckets complete in backward order
    out = red.finish()
    assert all(float(out[n].sum()) == shapes[n].numel() for n in order)
    red.begin()
    red.mark(order[0])
    with pytest.raises(RuntimeError):
        red.finish()


def _reducer_worker(rank, world, port):
    from deadtrees_b200.parallel import GradBucketReducer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes = [("a", torch.Size([300, 7])), ("b", torch.Size([5])), ("c", torch.Size([1000, 3, 3])), ("d", torch.Size([64]))]
    red = GradBucketReducer(shapes, "cpu", bucket_bytes=12000, world_size=world)
    assert len(red.buckets) == 3
    for step in range(2):
        red.begin()
        for i, (n, s) in enumerate(shapes):
            red.add(n, torch.full(tuple(s), float((rank + 1) * (i + 1) + step)))
        out = red.finish()
        for i, (n, s) in enumerate(shapes):
            want = sum((r + 1) * (i + 1) + step for r in range(world)) / world      # mean over ranks
            assert out[n].shape == s and torch.allclose(out[n], torch.full(tuple(s), want)), (n, rank)
        assert red.launched == [0, 1, 2]
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_gradient_bucket_allreduce_gloo(world):
    """the N>1 training path on CPU: every rank ends with the mean gradient, buckets launched as they complete."""
    port = 31500 + (os.getpid() % 2000) + world
    mp.spawn(_reducer_worker, args=(world, port), nprocs=world, join=True)


def test_fold_upsample_weights_identity():
    """nearest-x2 up-sampling folded into per-class 2 x 2 weights (engine.fold_upsample_weights, DT_CONV_UPS_FOLDED; restated
    in oracle/ref_unet.py) is the same convolution: float64, borders included."""
    import torch.nn.functional as F
    from deadtrees_b200.engine import fold_upsample_weights, pack_weight_folded
    from oracle import ref_unet
    g = torch.Generator().manual_seed(0)
    w = torch.randn(5, 7, 3, 3, generator=g, dtype=torch.float64)
    x = torch.randn(2, 7, 6, 9, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, None, 1, 1)
    wf = ref_unet.fold_upsample_weights(w)                       # (C_out, C_in, a, b, ey, ex)
    N, _, H, W = x.shape
    xp = F.pad(x, (1, 1, 1, 1))
    y = torch.zeros_like(ref)
    for a in range(2):
        for b in range(2):
            y[:, :, a::2, b::2] = F.conv2d(xp[:, :, a: a + H + 1, b: b + W + 1], wf[:, :, a, b])
    assert (y - ref).abs().max().item() < 1e-12
    # the engine's packing holds the same sums: [C_out][(class * 4 + e) * C_in + ci]
    eng = fold_upsample_weights(w.float())                       # (C_out, class, e, C_in)
    assert torch.equal(eng, ref_unet.fold_upsample_weights(w.float()).permute(0, 2, 3, 4, 5, 1).reshape(5, 4, 4, 7))
    packed = pack_weight_folded(torch.randn(16, 96, 3, 3, generator=g), "cpu", 64)
    assert packed.shape == (16, 16 * 64 + 9 * 32) and packed.dtype == torch.bfloat16


def test_train_transform_draws_follow_the_reference_distributions():
    """draw_train_params: OneOf(HFlip, VFlip, p=.5), RandomRotate90(p=.5), RandomBrightnessContrast(p=.5, 0.2 / 0.15)
    (deadtreedata.py:132-146) - frequencies and ranges of the draws."""
    from deadtrees_b200.data.deadtreedata import draw_train_params
    geom, bc = draw_train_params(np.random.default_rng(1), 20000)
    flips = np.bincount(geom[:, 0], minlength=3) / 20000
    rots = np.bincount(geom[:, 1], minlength=4) / 20000
    np.testing.assert_allclose(flips, [0.5, 0.25, 0.25], atol=0.015)
    np.testing.assert_allclose(rots, [0.625, 0.125, 0.125, 0.125], atol=0.015)
    on = bc[:, 0] != 1.0
    assert abs(on.mean() - 0.5) < 0.015 and (bc[~on, 1] == 0).all()
    assert bc[on, 0].min() >= 0.85 and bc[on, 0].max() <= 1.15 and np.abs(bc[on, 1]).max() <= 0.2
    assert abs(bc[on, 0].mean() - 1.0) < 0.005 and abs(bc[on, 1].mean()) < 0.005
    g2, b2 = draw_train_params(np.random.default_rng(1), 20000, beta_times_alpha=True)
    np.testing.assert_allclose(b2[:, 1], bc[:, 1] * bc[:, 0]) and np.testing.assert_array_equal(g2, geom)


def test_gwdl_constructor_contract(capsys):
    """GeneralizedWassersteinDiceLoss(dist_matrix, weighting_mode, reduction) as deadtrees/loss/gwdl.py:44-82: the matrix is
    normalised to a maximum of 1 (with the reference's message), unknown weighting modes are rejected by the same assert."""
    from deadtrees.loss.gwdl import GeneralizedWassersteinDiceLoss as ShimGW
    from deadtrees_b200.loss.gwdl import GeneralizedWassersteinDiceLoss
    assert ShimGW is GeneralizedWassersteinDiceLoss
    gw = GeneralizedWassersteinDiceLoss(dist_matrix=np.array([[0.0, 2.0], [2.0, 0.0]]))
    assert "Normalize the maximum of the distance matrix" in capsys.readouterr().out
    assert gw.matrix() == [[0.0, 1.0], [1.0, 0.0]] and gw.num_classes == 2
    gw = GeneralizedWassersteinDiceLoss(dist_matrix=torch.tensor([[0.0, 1.0], [0.5, 0.0]]))
    assert capsys.readouterr().out == "" and gw.matrix() == [[0.0, 1.0], [0.5, 0.0]]
    with pytest.raises(AssertionError):
        GeneralizedWassersteinDiceLoss(dist_matrix=np.eye(2), weighting_mode="banana")
    with pytest.raises(NotImplementedError):
        GeneralizedWassersteinDiceLoss(dist_matrix=1 - np.eye(2), weighting_mode="GDL")


def test_data_shim_exports():
    import deadtrees.data.deadtreedata as d
    for name in ("DeadtreeDatasetConfig", "val_transform", "train_transform", "transform", "BatchTrainTransform"):
        assert hasattr(d, name), name
    import deadtrees.loss.losses as l
    for name in ("DiceLoss", "FocalLoss", "BoundaryLoss", "class2one_hot", "one_hot2dist"):
        assert hasattr(l, name), name


def test_rebucketing_keeps_the_flat_gradient_buffer():
    """set_process_group() after configure_optimizers() must not leave the optimizer on a stale buffer - the
    reducer re-buckets the SAME flat buffer (a gradient's offset does not depend on the bucket size)."""
    from deadtrees_b200.parallel import GradBucketReducer, backward_param_order
    shapes = _param_shapes()
    order = backward_param_order([n for n, _ in shapes])
    named = [(n, dict(shapes)[n]) for n in order]
    a = GradBucketReducer(named, "cpu", bucket_bytes=25 << 20, world_size=1)
    b = GradBucketReducer(named, "cpu", bucket_bytes=1 << 20, world_size=1, flat=a.flat)
    assert b.flat.data_ptr() == a.flat.data_ptr() and len(b.buckets) > len(a.buckets)
    for n, _ in named:
        assert a.flat_offset(n) == b.flat_offset(n)
        assert a.view(n).data_ptr() == b.view(n).data_ptr()
    with pytest.raises(ValueError):
        GradBucketReducer(named[:-1], "cpu", world_size=1, flat=a.flat)


def test_state_generation_counter():
    from deadtrees_b200 import ops
    g = ops.STATE_GENERATION
    ops.bump_state_generation()
    assert ops.STATE_GENERATION == g + 1


# ---- tile-range shards (deadtrees_b200/sharding.py::ShardPlan) ------------------------------------------------------------
@pytest.mark.parametrize("gy,gx,world,ov", [(45, 45, 8, 32), (45, 45, 4, 32), (45, 45, 2, 32), (7, 5, 3, 8), (9, 4, 2, 16), (40, 40, 8, 0)])
def test_shard_plan_partition(gy, gx, world, ov):
    from deadtrees_b200.sharding import ShardPlan
    ps = [ShardPlan(gy, gx, world, r, ov) for r in range(world)]
    assert ps[0].t0 == 0 and ps[-1].t1 == gy * gx and ps[0].R0 == 0 and ps[-1].R1 == gy
    sizes = [p.t1 - p.t0 for p in ps]
    if ov > 0:
        assert max(sizes) - min(sizes) <= 1                       # balanced to one tile (8 GPUs, cfg2: 253 / 254 tiles)
    for a, b in zip(ps, ps[1:]):
        assert a.t1 == b.t0 and a.R1 == b.R0
        assert a.recv_tail == b.send_head and a.send_halo == b.recv_halo
    rows = [p.mask_rows(10 ** 9, 256) for p in ps]                # mask rows: a partition of the mosaic rows
    assert rows[0][0] == 0 and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    for p in ps:       # everything the stitch of rows [R0, R1) reads is computed locally or received
        have = set(range(p.t0, p.t1))
        if p.recv_tail:
            have |= set(range(*p.recv_tail))
        assert set(range(p.R0 * gx, p.R1 * gx)) <= have
        if p.R0 > 0 and ov > 0:
            halo = set(range(*p.recv_halo)) if p.recv_halo else set()
            assert set(range((p.R0 - 1) * gx, p.R0 * gx)) <= (halo | set(range(p.t0, p.t1)))
        assert p.B0 <= min(p.t0, p.ty_base * gx) and p.B1 >= max(p.t1, p.R1 * gx) and p.B0 % gx == 0


def _shard_exchange_worker(rank, world, port):
    from deadtrees_b200.sharding import ShardPlan, exchange_logits, gather_mask_rows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gy, gx, T, ov, K = 7, 5, 16, 4, 3
    H = (gy - 1) * (T - ov) + T
    g = torch.Generator().manual_seed(5)
    full = torch.randn(gy * gx, T, T, K, generator=g)             # what a single GPU would hold
    plans = [ShardPlan(gy, gx, world, r, ov) for r in range(world)]
    p = plans[rank]
    local = torch.full((p.B1 - p.B0, T, T, K), float("nan"))
    local[p.t0 - p.B0: p.t1 - p.B0] = full[p.t0: p.t1]            # "computed" tiles
    exchange_logits(p, local, T)
    # after the exchange: every tile of the stitched rows is complete, the halo row has its bottom strips
    assert torch.equal(local[p.R0 * gx - p.B0: p.R1 * gx - p.B0], full[p.R0 * gx: p.R1 * gx]), rank
    if p.R0 > 0:
        a = (p.R0 - 1) * gx
        assert torch.equal(local[a - p.B0: a - p.B0 + gx, T - ov:], full[a: a + gx, T - ov:]), rank
    # mask rows -> rank 0
    mask = torch.zeros((H, 9), dtype=torch.uint8)
    y0, y1 = p.mask_rows(H, T)
    mask[y0:y1] = rank + 1
    gather_mask_rows(plans, rank, mask, H, T)
    if rank == 0:
        for k, q in enumerate(plans):
            a, b = q.mask_rows(H, T)
            assert bool((mask[a:b] == k + 1).all()), k
        assert int((mask == 0).sum()) == 0
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shard_exchange_gloo(world):
    """the N > 1 inference path on the CPU: head tiles go to the previous rank, boundary strips to the next one, mask rows
    to rank 0 (gloo; NCCL over NVLink on the GPUs)."""
    port = 31500 + (os.getpid() % 2000) + world
    mp.spawn(_shard_exchange_worker, args=(world, port), nprocs=world, join=True)


def test_batch_plan_equal_batches_and_lead():
    """deployment/inference.py batch_plan: equal batches under the bound, optional short first batch, explicit sizes"""
    from deadtrees_b200.deployment.inference import batch_plan
    assert batch_plan(0, 2025, 405) == [(i * 405, 405) for i in range(5)]
    p = batch_plan(10, 506, 405)                         # 506 tiles: two batches of 253, not 405 + 101
    assert p == [(10, 253), (263, 253)]
    p = batch_plan(0, 254, 405, lead=45)
    assert p == [(0, 45), (45, 209)]
    p = batch_plan(0, 1012, 405, lead=45)                # 967 left: three batches of 323 / 323 / 321
    assert [n for _, n in p] == [45, 323, 323, 321] and p[-1][0] + p[-1][1] == 1012
    assert batch_plan(0, 30, 405, lead=45) == [(0, 30)]  # shard smaller than the lead batch
    assert batch_plan(5, 100, [10, 40]) == [(5, 10), (15, 40), (55, 40), (95, 10)]
    assert batch_plan(0, 0, 8) == []
    for bad in (0, [], [4, 0]):
        with pytest.raises(ValueError):
            batch_plan(0, 10, bad)
    for total in (1, 44, 45, 46, 253, 2025, 7777):
        for lead in (0, 45):
            p = batch_plan(3, total, 405, lead=lead)
            assert p[0][0] == 3 and sum(n for _, n in p) == total
            assert all(a[0] + a[1] == b[0] for a, b in zip(p, p[1:])) and max(n for _, n in p) <= 405
