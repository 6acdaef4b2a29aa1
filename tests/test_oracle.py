"""CPU tests: the oracle against the reference's golden vectors and fixtures generated from the reference."""
import numpy as np
import pytest
import torch

from oracle import ref_fscore, ref_losses, ref_normalize, ref_tiler, ref_unet


# ---- tiler ----------------------------------------------------------------------------------------
class TestBlocksVectorizedGolden:
    """the reference's own known-answer test (tests/test_tiler.py:56-77)."""
    source = np.array([np.arange(16).reshape(4, 4)] * 3)
    target = np.array([
        [[[0, 1], [4, 5]]] * 3, [[[2, 3], [6, 7]]] * 3, [[[8, 9], [12, 13]]] * 3, [[[10, 11], [14, 15]]] * 3])

    def test_make(self):
        np.testing.assert_array_equal(ref_tiler.make_blocks(self.source, 2), self.target)

    def test_unmake(self):
        np.testing.assert_array_equal(ref_tiler.unmake_blocks(self.target[:, 0], 2, 4, 4), self.source[0])


def test_tiler_against_reference_outputs(golden_dir):
    g = np.load(golden_dir / "tiler_blocks.npz")
    for i in range(int(g["ncases"])):
        d = int(g[f"d{i}"])
        np.testing.assert_array_equal(ref_tiler.make_blocks(g[f"x{i}"], d), g[f"blocks{i}"])
        m, n = g[f"x{i}"].shape[1:]
        np.testing.assert_array_equal(ref_tiler.unmake_blocks(g[f"pred{i}"], d, m, n), g[f"merged{i}"])


@pytest.mark.parametrize("size,expect", [((8192, 8192), (16, 16)), ((8192, 7433), (16, 15)), ((2649, 8192), (6, 16))])
def test_inspect_shape_edge_tiles(size, expect):
    """the reference's three real tiles (tests/test_tiler.py:30-46): ceil grid for ragged sizes."""
    assert ref_tiler.inspect_shape(size, (8192, 8192), (512, 512)) == expect


def test_inspect_shape_unaligned():
    with pytest.raises(ValueError):
        ref_tiler.inspect_shape((100, 100), (8192, 8192), (512, 211))


def test_tiler_oracle_roundtrip_ragged():
    rng = np.random.default_rng(0)
    sv = rng.integers(0, 256, size=(4, 83, 120), dtype=np.uint8)
    t = ref_tiler.TilerOracle((128, 128), (32, 32))
    t.load_array(sv)
    assert t.subtiles == (3, 4) and t.subtiles_to_use.sum() == 12
    b = t.get_batches()
    assert b.shape == (12, 4, 32, 32)
    out = t.put_batches(b[:, 0].astype(np.int64))  # identity "prediction" = band 0
    np.testing.assert_array_equal(out, sv[0])


def test_overlap_zero_degenerates_to_blocks():
    rng = np.random.default_rng(1)
    m = rng.integers(0, 256, size=(64, 96, 3), dtype=np.uint8)
    tiles = ref_tiler.extract_tiles(m, 32, 0)
    np.testing.assert_array_equal(tiles.transpose(0, 3, 1, 2), ref_tiler.make_blocks(m.transpose(2, 0, 1), 32))
    logits = rng.standard_normal((tiles.shape[0], 32, 32, 3)).astype(np.float32)
    _, mask = ref_tiler.stitch_blend(logits, 64, 96, 32, 0)
    np.testing.assert_array_equal(mask, ref_tiler.unmake_blocks(logits.argmax(-1), 32, 64, 96))


def test_overlap_grid_and_window():
    assert ref_tiler.overlap_grid(10000, 10000, 256, 32) == (45, 45, 10112, 10112)
    assert ref_tiler.overlap_grid(10000, 10000, 256, 0) == (40, 40, 10240, 10240)
    w = ref_tiler.blend_window(256, 32)
    np.testing.assert_allclose(w[224:] + w[:32], 1.0, rtol=0, atol=1e-6)  # partition of unity across an overlap


# ---- losses / metric ------------------------------------------------------------------------------
def test_losses_against_reference_outputs(golden_dir):
    g = np.load(golden_dir / "losses.npz")
    for i in range(int(g["ncases"])):
        logits, mask = torch.from_numpy(g[f"logits{i}"]), torch.from_numpy(g[f"mask{i}"])
        K = logits.shape[1]
        probs = logits.softmax(dim=1)
        onehot = ref_losses.class2one_hot(mask, K)
        np.testing.assert_array_equal(onehot.numpy(), g[f"onehot{i}"])
        fg = list(range(1, K))
        np.testing.assert_allclose(ref_losses.dice_loss(probs, onehot, fg), g[f"dice{i}"], rtol=1e-6)
        np.testing.assert_allclose(ref_losses.focal_loss(probs, onehot, list(range(K))), g[f"focal{i}"], rtol=1e-6)
        np.testing.assert_allclose(ref_losses.generalized_dice_loss(probs, onehot), g[f"gdl{i}"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(ref_losses.surface_loss(probs, torch.from_numpy(g[f"dist{i}"]), fg),
                                   g[f"surface{i}"], rtol=1e-5, atol=1e-7)
        # Generalized Wasserstein Dice loss: on the probabilities (SemSegment's call) and on raw scores, value + gradient
        D = [[0.0, 1.0, 1.0], [1.0, 0.0, 0.5], [1.0, 0.5, 0.0]]
        D = [r[:K] for r in D[:K]]
        for name, pre in (("gwdl_probs", lambda z: z.softmax(dim=1)), ("gwdl_scores", lambda z: z)):
            z = logits.clone().requires_grad_(True)
            val = ref_losses.gwdl_loss(pre(z), mask, D)
            val.backward()
            np.testing.assert_allclose(val.item(), g[f"{name}{i}"], rtol=1e-12)
            np.testing.assert_allclose(z.grad.numpy(), g[f"grad_{name}{i}"], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(ref_losses.calculate_loss(probs, onehot, ["GWDICE"])["total_loss"].item(), g[f"gwdl_probs{i}"], rtol=1e-12)
    np.testing.assert_allclose(ref_losses.gwdl_loss(torch.from_numpy(g["logits0"]), torch.from_numpy(g["mask0"]),
                                                    [[0.0, 2.0, 4.0], [2.0, 0.0, 1.0], [4.0, 1.0, 0.0]]).item(), g["gwdl_unnorm0"], rtol=1e-12)


def test_class2one_hot_rejects_out_of_range():
    with pytest.raises(AssertionError):
        ref_losses.class2one_hot(torch.tensor([[[0, 3]]]), 3)


@pytest.mark.parametrize("inc,res", [(2, 1.0), (3, 0.6154), (4, 0.2)])
def test_fscore_without_background_known_answers(inc, res):
    """tests/test_dice_metric.py:16,38-52 — equal to smp Fscore(ignore_channels=[0]) on the same inputs."""
    n = 5
    sample = torch.zeros((1, 2, n, n)); sample[:, 0] = 1; sample[:, 0, 2:, 2:] = 0; sample[:, 1, 2:, 2:] = 1
    pred = torch.zeros((1, 2, n, n)); pred[:, 0] = 1; pred[:, 0, inc:, inc:] = 0; pred[:, 1, inc:, inc:] = 1
    assert abs(float(ref_fscore.fscore(pred, sample, ignore_channels=[0])) - res) < 1e-4


def test_normalize_formula():
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, size=(8, 8, 4), dtype=np.uint8)
    out = ref_normalize.val_transform(img)
    ref = ((img.astype(np.float64) / 255.0 - ref_normalize.MEAN) / ref_normalize.STD).transpose(2, 0, 1)
    np.testing.assert_allclose(out, ref, rtol=0, atol=2e-6)
    assert out.dtype == np.float32 and out.shape == (4, 8, 8)


# ---- U-Net restatement ----------------------------------------------------------------------------
@pytest.mark.parametrize("cin,k,count", [(3, 3, 24436659), (4, 3, 24439795)])
def test_unet_parameter_count(cin, k, count):
    m = ref_unet.Unet(in_channels=cin, classes=k)
    assert sum(p.numel() for p in m.parameters()) == count
    assert list(m.parameters())[0].shape == (64, cin, 7, 7)  # PyTorchInference reads channels from it
    assert len(m.state_dict()) == 278


def test_encoder_matches_torchvision_resnet34():
    torchvision = pytest.importorskip("torchvision")
    tv = torchvision.models.resnet34(weights=None).eval()
    enc = ref_unet.ResNet34Encoder(3).eval()
    missing = enc.load_state_dict({k: v for k, v in tv.state_dict().items() if not k.startswith("fc.")}, strict=True)
    x = torch.randn(2, 3, 64, 64, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        feats = enc(x)
        y = tv.maxpool(tv.relu(tv.bn1(tv.conv1(x))))
        y = tv.layer4(tv.layer3(tv.layer2(tv.layer1(y))))
    assert [f.shape[1] for f in feats] == [3, 64, 64, 128, 256, 512]
    assert [f.shape[-1] for f in feats] == [64, 32, 16, 8, 4, 2]
    torch.testing.assert_close(feats[-1], y, rtol=0, atol=0)


def test_unet_forward_shape_and_decoder_wiring():
    m = ref_unet.build_reference_unet(3, 3, seed=0)
    x = torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        y = m(x)
    assert y.shape == (1, 3, 64, 64)
    assert m.decoder.blocks[0].conv1[0].weight.shape == (256, 768, 3, 3)
    assert m.decoder.blocks[3].conv1[0].weight.shape == (32, 128, 3, 3)
    assert m.decoder.blocks[4].conv1[0].weight.shape == (16, 32, 3, 3)
    # PyTorchInference.run semantics: 3-d in -> 2-d out, rgb model on rgbn data slices channels
    out = ref_unet.run_inference(m, torch.randn(4, 64, 64), channels=3)
    assert out.shape == (64, 64) and out.dtype == torch.int64


def test_distance_maps_against_reference_outputs(golden_dir):
    """oracle/ref_dist.one_hot2dist == the reference's own one_hot2dist (losses.py:159-178) on the committed fixtures:
    the truncating form the dataloader uses (int32 one-hot in) and the exact float32 form; includes a class that covers the
    whole tile (scipy's no-background convention) and absent classes (left at 0)."""
    from oracle import ref_dist
    g = np.load(golden_dir / "dist.npz")
    for i in range(int(g["ncases"])):
        lab, K = g[f"labels{i}"], int(g[f"K{i}"])
        onehot = (lab[None] == np.arange(K)[:, None, None]).astype(np.int32)
        got_int = ref_dist.one_hot2dist(onehot, resolution=[1, 1])
        assert got_int.dtype == np.int32
        np.testing.assert_array_equal(got_int, g[f"dist_int{i}"])
        np.testing.assert_array_equal(ref_dist.one_hot2dist(onehot, resolution=[1, 1], dtype=np.float32), g[f"dist_f32{i}"])
        np.testing.assert_array_equal(ref_dist.labels_to_dist(lab[None], K, truncate=True)[0], g[f"dist_int{i}"].astype(np.float32))
        np.testing.assert_array_equal(ref_dist.labels_to_dist(lab[None], K, truncate=False)[0], g[f"dist_f32{i}"])


def test_confusion_matrices_against_sklearn():
    """oracle/ref_confusion.py (torchmetrics' published definition) against sklearn's independent implementation,
    including an absent class (NaN row -> 0) and an empty forest mask."""
    from sklearn.metrics import confusion_matrix as sk_cm
    from oracle import ref_confusion
    rng = np.random.default_rng(5)
    for K, n, absent in ((3, 5000, None), (3, 777, 2), (2, 100, None), (4, 4096, 0)):
        target, pred = rng.integers(0, K, n), rng.integers(0, K, n)
        if absent is not None:
            target[target == absent] = (absent + 1) % K
        lu = rng.integers(0, 3, n)
        m = ref_confusion.epoch_matrices(pred, target, lu, K)
        np.testing.assert_array_equal(m["cm_px"], sk_cm(target, pred, labels=list(range(K))))
        np.testing.assert_allclose(m["cm_norm"], sk_cm(target, pred, labels=list(range(K)), normalize="true"), rtol=1e-15)
        np.testing.assert_array_equal(m["cm_px_masked"], sk_cm(target[lu == 1], pred[lu == 1], labels=list(range(K))))
        np.testing.assert_allclose(m["cm_norm_masked"], sk_cm(target[lu == 1], pred[lu == 1], labels=list(range(K)), normalize="true"), rtol=1e-15)
        assert m["cm_px"].sum() == n and np.isfinite(m["cm_norm"]).all()
    empty = ref_confusion.epoch_matrices(pred, target, np.zeros(n, np.int64), K)
    assert empty["cm_px_masked"].sum() == 0 and (empty["cm_norm_masked"] == 0).all()


def test_train_transform_oracle_properties():
    """oracle/ref_augment.py: identity draws reduce to val_transform, four quarter turns are the identity, a flip is an
    involution, the table saturates, and the brightness term uses the mean over every channel of the sample."""
    from oracle import ref_augment, ref_normalize
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (32, 32, 4), dtype=np.uint8)
    mask, lu = rng.integers(0, 3, (32, 32), dtype=np.uint8), rng.integers(0, 2, (32, 32), dtype=np.uint8)
    out, m, l = ref_augment.train_transform(img, mask, lu, 0, 0, 1.0, 0.0)
    np.testing.assert_array_equal(out, ref_normalize.val_transform(img))
    np.testing.assert_array_equal(m, mask) and np.testing.assert_array_equal(l, lu)
    a = img
    for _ in range(4):
        a = ref_augment.geometric(a, 0, 1)
    np.testing.assert_array_equal(a, img)
    for f in (1, 2):
        np.testing.assert_array_equal(ref_augment.geometric(ref_augment.geometric(img, f, 0), f, 0), img)
    # np.rot90 turns counter-clockwise: the top-right pixel moves to the top-left corner
    np.testing.assert_array_equal(ref_augment.geometric(img, 0, 1)[0, 0], img[0, -1])
    bc = ref_augment.brightness_contrast(img, 1.15, 0.2)
    lut = np.clip(np.arange(256, dtype=np.float32) * np.float32(1.15) + np.float32(0.2 * img.mean()), 0, 255).astype(np.uint8)
    np.testing.assert_array_equal(bc, lut[img])
    assert bc.max() == 255 and (bc >= img).all()
    out2, m2, _ = ref_augment.train_transform(img, mask, lu, 2, 3, 0.9, -0.1, in_channels=3, classes=2)
    assert out2.shape == (3, 32, 32) and m2.max() == 1 and m2.dtype == np.int64


def test_gwdl_oracle_properties():
    """oracle Wasserstein Dice: one sample = the paper's per-sample formula; a confident correct prediction -> ~0; a target
    without foreground -> true positives 0 and a loss of ~1; samples of a batch are coupled only through the true-positive
    term (the reference's broadcast), so a batch of identical samples equals B times the per-sample true positives."""
    g = torch.Generator().manual_seed(3)
    D = [[0.0, 1.0, 1.0], [1.0, 0.0, 0.5], [1.0, 0.5, 0.0]]
    M = torch.tensor(D, dtype=torch.float64)
    z = torch.randn(1, 3, 12, 10, generator=g)
    t = torch.randint(0, 3, (1, 12, 10), generator=g)
    q = z.reshape(1, 3, -1).softmax(1).double()
    w = (M[t.reshape(1, -1)].permute(0, 2, 1) * q).sum(1)
    tp = ((t.reshape(1, -1) != 0) * (1 - w)).sum()
    expect = 1 - (2 * tp + 2.0 ** -52) / (2 * tp + w.sum() + 2.0 ** -52)
    np.testing.assert_allclose(ref_losses.gwdl_loss(z, t, D).item(), expect.item(), rtol=1e-12)
    sure = torch.nn.functional.one_hot(t, 3).permute(0, 3, 1, 2).float() * 60.0
    assert ref_losses.gwdl_loss(sure, t, D).item() < 1e-6
    assert ref_losses.gwdl_loss(z, torch.zeros_like(t), D).item() > 1 - 1e-9
    zb, tb = z.repeat(3, 1, 1, 1), t.repeat(3, 1, 1)
    batch = 1 - (2 * 3 * tp + 2.0 ** -52) / (2 * 3 * tp + w.sum() + 2.0 ** -52)
    np.testing.assert_allclose(ref_losses.gwdl_loss(zb, tb, D).item(), batch.item(), rtol=1e-7)
