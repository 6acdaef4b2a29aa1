"""GeoTIFF tile I/O without GDAL (deadtrees/deployment/tiler.py:82-140) and the double-buffered file pipeline."""
import numpy as np
import pytest

from deadtrees_b200.deployment import geotiff
from deadtrees_b200.deployment.tiler import Tiler, inspect_tile

TAGS = {33550: (0.2, 0.2, 0.0), 33922: (0.0, 0.0, 0.0, 500000.0, 5400000.0, 0.0),
        34735: (1, 1, 0, 3, 1024, 0, 1, 1, 1025, 0, 1, 1, 3072, 0, 1, 25832)}


def make_tile(path, H, W, bands, seed=0, compress="tiff_lzw"):
    from PIL import Image, TiffImagePlugin
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, size=(H, W, bands), dtype=np.uint8)
    info = TiffImagePlugin.ImageFileDirectory_v2()
    for t, v in TAGS.items():
        info[t] = v
        info.tagtype[t] = 3 if t == 34735 else 12
    Image.fromarray(a, {3: "RGB", 4: "RGBA"}[bands]).save(path, format="TIFF", compression=compress, tiffinfo=info)
    return a


@pytest.mark.parametrize("bands,compress", [(4, "tiff_lzw"), (3, None), (4, "tiff_adobe_deflate")])
def test_read_write_roundtrip(tmp_path, bands, compress):
    f = tmp_path / "ortho_ms_2019_500-5400.tif"
    a = make_tile(f, 83, 120, bands, compress=compress)
    vals, tags = geotiff.read_geotiff(f)
    assert vals.shape == (bands, 83, 120) and np.array_equal(vals, a.transpose(2, 0, 1))
    assert {k: tuple(v) for k, v in tags.items()} == TAGS
    mask = (a[..., 0] % 3).astype(np.uint8)
    out = tmp_path / "mask.tif"
    geotiff.write_geotiff(out, mask, tags)
    back, tags2 = geotiff.read_geotiff(out)
    assert back.shape == (1, 83, 120) and np.array_equal(back[0], mask)
    assert {k: tuple(v) for k, v in tags2.items()} == TAGS            # geo-referenced like the input
    from PIL import Image
    with Image.open(out) as im:
        assert im.tag_v2[259] == 5                                    # LZW, as the reference's to_raster(compress="LZW")


def test_tiler_load_file_write_file_without_gdal(tmp_path):
    """Tiler.load_file / write_file (tiler.py:82-140) on a ragged RGB+NIR tile: same arrays as load_array, mask written
    cropped to the true size with the input's geo tags"""
    f = tmp_path / "tile.tif"
    a = make_tile(f, 83, 120, 4, seed=3)
    assert inspect_tile(f, (128, 128), (32, 32)).subtiles == (3, 4)
    t = Tiler(tile_shape=(128, 128), subtile_shape=(32, 32))
    t.load_file(f)
    ref = Tiler(tile_shape=(128, 128), subtile_shape=(32, 32))
    ref.load_array(a.transpose(2, 0, 1))
    assert np.array_equal(t._indata, ref._indata) and np.array_equal(t._subtiles_to_use, ref._subtiles_to_use)
    t._outdata = (t._indata[0] % 3).astype(np.uint8)                # "prediction" = band 0 mod 3 (put_batches runs on the GPU)
    out = tmp_path / "pred.tif"
    t.write_file(out)
    back, tags = geotiff.read_geotiff(out)
    assert np.array_equal(back[0], a[..., 0] % 3) and {k: tuple(v) for k, v in tags.items()} == TAGS


@pytest.mark.gpu
def test_segment_files_pipeline(tmp_path):
    """decode of file i + 1 and encode of mask i - 1 behind the GPU: every written mask equals the one-file-at-a-time result"""
    import torch
    from deadtrees_b200.deployment.inference import MosaicInference
    from deadtrees_b200.engine import UnetEngine
    from gpu_util import pattern_mosaic, trained_model
    model = trained_model(3, 3)
    eng = UnetEngine(model.state_dict(), 3, 3, precision="bf16")
    mi = MosaicInference(eng, tile=64, overlap=16, batch_tiles=6)
    files = []
    from PIL import Image, TiffImagePlugin
    for i, (H, W) in enumerate([(200, 150), (200, 150), (130, 190), (64, 64)]):
        rgb = pattern_mosaic(H, W, 3, seed=50 + i)
        if i == 3:
            rgb[...] = 255                                            # an empty tile: skipped by the validity filter
        info = TiffImagePlugin.ImageFileDirectory_v2()
        for t, v in TAGS.items():
            info[t] = v
            info.tagtype[t] = 3 if t == 34735 else 12
        f = tmp_path / f"ortho_{i}.tif"
        Image.fromarray(rgb, "RGB").save(f, format="TIFF", compression="tiff_lzw", tiffinfo=info)
        files.append((f, rgb))
    valid = lambda band1: not np.isin(band1, [0, 255]).all()          # scripts/inference.py:63-65
    written = geotiff.segment_files(mi, [f for f, _ in files], tmp_path / "out", is_valid=valid)
    assert [w.name for w in written] == ["ortho_0.tif", "ortho_1.tif", "ortho_2.tif"]
    for (f, rgb), w in zip(files, written):
        want = MosaicInference(eng, tile=64, overlap=16, batch_tiles=6).run_host(rgb, "hwc")
        got, tags = geotiff.read_geotiff(w)
        assert np.array_equal(got[0], want) and {k: tuple(v) for k, v in tags.items()} == TAGS
