"""GPU parity: tiler kernels against the oracle / reference golden vectors — bit-exact."""
import numpy as np
import pytest
import torch

from deadtrees_b200 import ops
from deadtrees_b200.data.deadtreedata import normalize_constants, val_transform
from deadtrees_b200.deployment.inference import blend_window, overlap_grid
from deadtrees_b200.deployment.tiler import Tiler
from deadtrees_b200.utils.data_handling import make_blocks_vectorized, unmake_blocks_vectorized
from oracle import ref_normalize, ref_tiler

pytestmark = pytest.mark.gpu


def test_blocks_reference_known_answer():
    """tests/test_tiler.py:56-77 through the drop-in functions."""
    source = np.array([np.arange(16).reshape(4, 4)] * 3)
    target = np.array([[[[0, 1], [4, 5]]] * 3, [[[2, 3], [6, 7]]] * 3, [[[8, 9], [12, 13]]] * 3,
                       [[[10, 11], [14, 15]]] * 3])
    np.testing.assert_array_equal(make_blocks_vectorized(source, 2), target)
    np.testing.assert_array_equal(unmake_blocks_vectorized(target[:, 0, :, :], 2, 4, 4), source[0])


def test_blocks_golden(golden_dir):
    g = np.load(golden_dir / "tiler_blocks.npz")
    for i in range(int(g["ncases"])):
        d = int(g[f"d{i}"])
        np.testing.assert_array_equal(make_blocks_vectorized(g[f"x{i}"], d), g[f"blocks{i}"])
        m, n = g[f"x{i}"].shape[1:]
        out = unmake_blocks_vectorized(g[f"pred{i}"], d, m, n)
        assert out.dtype == np.int64
        np.testing.assert_array_equal(out, g[f"merged{i}"])


@pytest.mark.parametrize("dtype", [np.uint8, np.int16, np.float32, np.int64])
@pytest.mark.parametrize("p,m,n,d", [(4, 2048, 2048, 256), (3, 96, 160, 32), (1, 7, 21, 7), (2, 30, 12, 6)])
def test_blocks_vs_oracle(dtype, p, m, n, d):
    rng = np.random.default_rng(7)
    x = (rng.integers(0, 200, size=(p, m, n))).astype(dtype)
    np.testing.assert_array_equal(make_blocks_vectorized(x, d), ref_tiler.make_blocks(x, d))
    b = rng.integers(0, 200, size=((m // d) * (n // d), d, d)).astype(dtype)
    np.testing.assert_array_equal(unmake_blocks_vectorized(b, d, m, n), ref_tiler.unmake_blocks(b, d, m, n))


def test_blocks_errors():
    with pytest.raises(ValueError):
        make_blocks_vectorized(np.zeros((3, 10, 10), np.uint8), 4)


@pytest.mark.parametrize("size", [(2048, 2048), (1900, 2048), (700, 513)])
def test_tiler_roundtrip(size):
    """edge tiles: ceil grid, zero padding, crop on the way back (tiler.py:105-170)."""
    rng = np.random.default_rng(3)
    sv = rng.integers(0, 256, size=(4, *size), dtype=np.uint8)
    t, o = Tiler(), ref_tiler.TilerOracle()
    t.load_array(sv); o.load_array(sv)
    b = t.get_batches()
    np.testing.assert_array_equal(b, o.get_batches())
    pred = (b[:, 0] % 3).astype(np.int64)
    t.put_batches(pred)
    np.testing.assert_array_equal(t.result, o.put_batches(pred))
    assert t.result.shape == size and t.result.dtype == np.uint8


@pytest.mark.parametrize("layout", ["hwc", "chw"])
@pytest.mark.parametrize("C,Cm", [(3, 3), (4, 4), (4, 3)])
@pytest.mark.parametrize("H,W,T,ov", [(256, 256, 64, 0), (300, 203, 64, 16), (130, 77, 32, 8)])
def test_gather_normalize_bit_exact(layout, C, Cm, H, W, T, ov):
    rng = np.random.default_rng(11)
    m = rng.integers(0, 256, size=(H, W, C), dtype=np.uint8)
    tiles = ref_tiler.extract_tiles(m, T, ov)                      # raw uint8 zero padding
    ref = np.zeros(tiles.shape[:3] + (4,), dtype=np.float32)
    off, sc = ref_normalize.normalize_constants(Cm)
    ref[..., :Cm] = (tiles[..., :Cm].astype(np.float32) - off) * sc
    src = torch.from_numpy(m if layout == "hwc" else np.ascontiguousarray(m.transpose(2, 0, 1))).cuda()
    gy, gx = overlap_grid(H, W, T, ov)
    o2, s2 = normalize_constants(Cm)
    np.testing.assert_array_equal(o2, off); np.testing.assert_array_equal(s2, sc)
    got = ops.tile_gather_normalize(src, layout, Cm, T, ov, (gy, gx), 0, gy * gx, off, sc, dtype=torch.float32)
    np.testing.assert_array_equal(got.cpu().numpy(), ref)
    got16 = ops.tile_gather_normalize(src, layout, Cm, T, ov, (gy, gx), 0, gy * gx, off, sc, dtype=torch.bfloat16)
    assert torch.equal(got16.cpu(), torch.from_numpy(ref).to(torch.bfloat16))
    # a sub-range of tiles (sharding) equals the slice
    if gy * gx > 2:
        part = ops.tile_gather_normalize(src, layout, Cm, T, ov, (gy, gx), 1, gy * gx - 2, off, sc, dtype=torch.float32)
        np.testing.assert_array_equal(part.cpu().numpy(), ref[1:-1])


@pytest.mark.parametrize("layout,C", [("hwc", 3), ("hwc", 4), ("chw", 3)])
@pytest.mark.parametrize("H,W,T,ov", [(150, 130, 64, 16), (600, 520, 256, 32), (70, 90, 32, 4), (100, 40, 36, 0)])
def test_gather_normalize_padded_frame(layout, C, H, W, T, ov):
    """pad=3: tiles land at offset (3, 3) of a zero-bordered (T+6, T+8) frame (input layout of the TMA stem)."""
    rng = np.random.default_rng(12)
    m = rng.integers(0, 256, size=(H, W, C), dtype=np.uint8)
    gy, gx = overlap_grid(H, W, T, ov)
    off, sc = normalize_constants(C)
    src = torch.from_numpy(m if layout == "hwc" else np.ascontiguousarray(m.transpose(2, 0, 1))).cuda()
    dense = ops.tile_gather_normalize(src, layout, C, T, ov, (gy, gx), 0, gy * gx, off, sc, dtype=torch.bfloat16)
    ref32 = ops.tile_gather_normalize(src, layout, C, T, ov, (gy, gx), 0, gy * gx, off, sc, dtype=torch.float32)
    assert torch.equal(dense, ref32.to(torch.bfloat16))        # fast bf16 kernel == generic kernel, rounded once
    frame = ops.tile_gather_normalize(src, layout, C, T, ov, (gy, gx), 0, gy * gx, off, sc, dtype=torch.bfloat16, pad=3)
    assert frame.shape == (gy * gx, T + 6, T + 8, 4)
    assert torch.equal(frame[:, 3:3 + T, 3:3 + T], dense)
    border = frame.clone(); border[:, 3:3 + T, 3:3 + T] = 0
    assert float(border.float().abs().max()) == 0.0


def test_val_transform_matches_oracle():
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, size=(256, 256, 4), dtype=np.uint8)
    out = val_transform(image=img)["image"]
    assert out.shape == (4, 256, 256) and out.dtype == torch.float32
    np.testing.assert_array_equal(out.cpu().numpy(), ref_normalize.val_transform(img))


@pytest.mark.parametrize("H,W,T", [(256, 256, 64), (300, 203, 64), (100, 50, 32)])
def test_stitch_mask_bit_exact(H, W, T):
    rng = np.random.default_rng(13)
    gy, gx = overlap_grid(H, W, T, 0)
    tm = rng.integers(0, 3, size=(gy * gx, T, T), dtype=np.uint8)
    ref = ref_tiler.unmake_blocks(tm, T, gy * T, gx * T)[:H, :W]
    out = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
    ops.stitch_mask(torch.from_numpy(tm).cuda(), gx, 0, out)
    np.testing.assert_array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("H,W,T,ov,K", [(300, 203, 64, 16, 3), (256, 256, 64, 32, 2), (100, 90, 32, 0, 3), (64, 64, 64, 8, 3),
                                        (128, 100, 64, 12, 4), (96, 83, 32, 8, 1)])
def test_stitch_blend_bit_exact(dtype, H, W, T, ov, K):
    g = torch.Generator().manual_seed(17)
    gy, gx = overlap_grid(H, W, T, ov)
    logits = (torch.randn(gy * gx, T, T, K, generator=g) * 3).to(dtype)
    ref_blend, ref_mask = ref_tiler.stitch_blend(logits.float().numpy(), H, W, T, ov)
    mask = torch.empty((H, W), dtype=torch.uint8, device="cuda")
    blended = torch.empty((H, W, K), dtype=torch.float32, device="cuda")
    ops.stitch_blend_argmax(logits.cuda(), ov, (gy, gx), blend_window(T, ov, "cuda"), mask, blended)
    np.testing.assert_array_equal(blended.cpu().numpy(), ref_blend)
    np.testing.assert_array_equal(mask.cpu().numpy(), ref_mask)
    # mask-only form (what MosaicInference runs): class chosen from the un-normalised sums, divisions only on near-ties
    mask2 = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
    ops.stitch_blend_argmax(logits.cuda(), ov, (gy, gx), blend_window(T, ov, "cuda"), mask2, None)
    np.testing.assert_array_equal(mask2.cpu().numpy(), ref_mask)


@pytest.mark.parametrize("T,ov", [(64, 32), (64, 12), (32, 8)])
def test_stitch_blend_mask_only_with_ties(T, ov):
    """logits drawn from three values: exact ties between classes are everywhere (first maximum must win) and the
    sums of different classes are often within a few ulps of each other."""
    H, W, K = 150, 170, 3
    g = torch.Generator().manual_seed(T + ov)
    gy, gx = overlap_grid(H, W, T, ov)
    vals = torch.tensor([1.0, 1.0078125, -0.5])
    logits = vals[torch.randint(0, 3, (gy * gx, T, T, K), generator=g)].to(torch.bfloat16)
    _, ref_mask = ref_tiler.stitch_blend(logits.float().numpy(), H, W, T, ov)
    mask = torch.full((H, W), 255, dtype=torch.uint8, device="cuda")
    ops.stitch_blend_argmax(logits.cuda(), ov, (gy, gx), blend_window(T, ov, "cuda"), mask, None)
    np.testing.assert_array_equal(mask.cpu().numpy(), ref_mask)


def test_full_size_mosaic_roundtrip_property():
    """BASELINE cfg2 size: gather (overlap 0) then stitch of a per-pixel function reproduces it on 10k x 10k."""
    H = W = 10000
    T = 256
    g = torch.Generator(device="cuda").manual_seed(1)
    m = torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, device="cuda", generator=g)
    gy, gx = overlap_grid(H, W, T, 0)
    out = torch.zeros((H, W), dtype=torch.uint8, device="cuda")
    for t0 in range(0, gy * gx, 400):
        n = min(400, gy * gx - t0)
        tiles = ops.tile_gather_normalize(m, "hwc", 3, T, 0, (gy, gx), t0, n, [0, 0, 0], [1, 1, 1], dtype=torch.float32)
        cls = (tiles[..., 0].to(torch.int32) % 3).to(torch.uint8)     # identity-normalised red channel mod 3
        ops.stitch_mask(cls, gx, t0, out)
    assert torch.equal(out, (m[..., 0] % 3))


@pytest.mark.parametrize("N,T,C,cin,classes", [(8, 64, 4, 4, 3), (5, 256, 4, 3, 2), (3, 37, 3, 3, 3), (2, 16, 1, 1, 3)])
def test_train_transform_vs_oracle(N, T, C, cin, classes):
    """dt_train_transform (train_transform + transform(), deadtreedata.py:132-146, 156-189) against the oracle for the
    same draws: every flip x rotation, contrast / brightness tables that saturate at both ends, the channel slice and the
    two-class merge.  Byte and index work: bit-exact, the normalised floats included."""
    from oracle import ref_augment
    from deadtrees_b200.data.deadtreedata import normalize_constants
    rng = np.random.default_rng(N * 100 + T)
    images = rng.integers(0, 256, (N, T, T, C), dtype=np.uint8)
    masks = rng.integers(0, 3, (N, T, T), dtype=np.uint8)
    lus = rng.integers(0, 4, (N, T, T), dtype=np.uint8)
    geom = np.array([[i % 3, (i // 3 + i) % 4] for i in range(N)], dtype=np.int32)
    bc = np.array([[1.0, 0.0] if i % 4 == 0 else [rng.uniform(0.85, 1.15), rng.uniform(-0.2, 0.2)] for i in range(N)])
    bc[-1] = [1.15, 0.2]
    offset, scale = normalize_constants(C)
    img, m, l = ops.train_transform(torch.from_numpy(images).cuda(), torch.from_numpy(masks).cuda(), torch.from_numpy(lus).cuda(),
                                    geom, bc, offset, scale, cin, merge_classes=classes == 2)
    assert img.shape == (N, cin, T, T) and m.dtype == torch.int64 and l.dtype == torch.int64
    for i in range(N):
        ri, rm, rl = ref_augment.train_transform(images[i], masks[i], lus[i], int(geom[i, 0]), int(geom[i, 1]),
                                                 float(bc[i, 0]), float(bc[i, 1]), in_channels=cin, classes=classes)
        np.testing.assert_array_equal(img[i].cpu().numpy(), ri, err_msg=f"sample {i} geom {geom[i]} bc {bc[i]}")
        np.testing.assert_array_equal(m[i].cpu().numpy(), rm)
        np.testing.assert_array_equal(l[i].cpu().numpy(), rl)
    # image only (no mask / lu)
    img2, m2, l2 = ops.train_transform(torch.from_numpy(images).cuda(), None, None, geom, bc, offset, scale, cin)
    assert m2 is None and l2 is None and torch.equal(img2, img)


def test_batch_train_transform_feeds_training_step():
    """BatchTrainTransform: uint8 tiles -> the tensors SemSegment.training_step takes, incl. the boundary-loss distance maps
    computed from the TRANSFORMED mask (deadtreedata.py:182-185); identity draws equal val_transform."""
    from deadtrees_b200.data.deadtreedata import BatchTrainTransform, train_transform, val_transform
    from oracle import ref_augment, ref_dist
    rng = np.random.default_rng(4)
    images = rng.integers(0, 256, (4, 64, 64, 4), dtype=np.uint8)
    masks = (rng.random((4, 64, 64)) < 0.2).astype(np.uint8) * rng.integers(1, 3, (4, 64, 64), dtype=np.uint8)
    lus = rng.integers(0, 2, (4, 64, 64), dtype=np.uint8)
    tf = BatchTrainTransform(in_channels=4, classes=3, distmap=True, seed=11)
    geom = np.array([[1, 1], [2, 3], [0, 2], [0, 0]], dtype=np.int32)
    bc = np.array([[1.1, 0.1], [0.9, -0.15], [1.0, 0.0], [1.0, 0.0]])
    img, mask, dist, lu = tf(images, masks, lus, params=(geom, bc))
    assert img.shape == (4, 4, 64, 64) and dist.shape == (4, 3, 64, 64) and img.is_cuda
    for i in range(4):
        ri, rm, rl = ref_augment.train_transform(images[i], masks[i], lus[i], *map(int, geom[i]), *map(float, bc[i]))
        np.testing.assert_array_equal(img[i].cpu().numpy(), ri)
        np.testing.assert_array_equal(mask[i].cpu().numpy(), rm)
        np.testing.assert_array_equal(dist[i].cpu().numpy(), ref_dist.labels_to_dist(rm[None], 3)[0])
    np.testing.assert_array_equal(img[3].cpu().numpy(), val_transform(image=images[3])["image"].cpu().numpy())
    # random draws: same seed, same batch
    a = BatchTrainTransform(seed=5)(images, masks, lus)
    b = BatchTrainTransform(seed=5)(images, masks, lus)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and a[2] is None
    one = train_transform(image=images[0], mask=masks[0], lu=lus[0])
    assert one["image"].shape == (4, 64, 64) and one["mask"].shape == (64, 64) and one["lu"].dtype == torch.int64
