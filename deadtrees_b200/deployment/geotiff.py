"""GeoTIFF tile I/O without GDAL, and the double-buffered file pipeline around ``MosaicInference``.

The reference reads a tile with ``rioxarray.open_rasterio`` and writes the mask with ``rio.to_raster(compress="LZW",
tiled=True)`` (``deadtrees/deployment/tiler.py:82-140``); rioxarray / GDAL are absent from this image.  The pixel arithmetic
of the hot path only needs the raster values and, for the output, the geo-referencing of the input: ``read_geotiff`` decodes
the bands with Pillow (strips or tiles, LZW / deflate / uncompressed, uint8) and keeps the GeoTIFF tags verbatim
(ModelPixelScale 33550, ModelTiepoint 33922, ModelTransformation 34264, GeoKeyDirectory 34735, GeoDoubleParams 34736,
GeoAsciiParams 34737, GDAL_NODATA 42113); ``write_geotiff`` stores a single-band uint8 mask LZW-compressed with those tags.

``segment_files`` is the I/O overlap the survey asks for: while the GPU segments file *i* (``MosaicInference.run`` with
pinned host buffers: row bands go up and mask bands come back behind the batches), a reader thread decodes file *i + 1*
into the other pinned buffer and a writer thread LZW-encodes the mask of file *i - 1*.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Callable, Dict, Iterable, List, Optional, Tuple, Union

import numpy as np

GEO_TAGS = (33550, 33922, 34264, 34735, 34736, 34737, 42113)
_TAG_TYPES = {33550: 12, 33922: 12, 34264: 12, 34735: 3, 34736: 12, 34737: 2, 42113: 2}   # DOUBLE / SHORT / ASCII


def read_geotiff(path: Union[str, Path]) -> Tuple[np.ndarray, Dict[int, object]]:
    """-> (raster (bands, H, W) uint8 - what ``rioxarray.open_rasterio(f).values`` holds at ``tiler.py:106`` -, geo tags)"""
    from PIL import Image

    Image.MAX_IMAGE_PIXELS = None
    with Image.open(path) as im:
        tags = {t: im.tag_v2[t] for t in GEO_TAGS if t in getattr(im, "tag_v2", {})}
        arr = np.asarray(im)
    if arr.dtype != np.uint8:
        raise ValueError(f"{path}: {arr.dtype} raster; the deadtrees tiles are 8-bit")
    if arr.ndim == 2:
        arr = arr[:, :, None]
    return np.ascontiguousarray(arr.transpose(2, 0, 1)), tags


def read_geotiff_hwc(path: Union[str, Path], out: Optional[np.ndarray] = None) -> Tuple[np.ndarray, Dict[int, object]]:
    """interleaved (H, W, bands) form - the layout the gather kernel streams from; ``out``: a (pinned) buffer to decode into"""
    from PIL import Image

    Image.MAX_IMAGE_PIXELS = None
    with Image.open(path) as im:
        tags = {t: im.tag_v2[t] for t in GEO_TAGS if t in getattr(im, "tag_v2", {})}
        arr = np.asarray(im)
    if arr.ndim == 2:
        arr = arr[:, :, None]
    if out is not None:
        out[: arr.shape[0], : arr.shape[1], : arr.shape[2]] = arr
        return out[: arr.shape[0], : arr.shape[1], : arr.shape[2]], tags
    return np.ascontiguousarray(arr), tags


def write_geotiff(path: Union[str, Path], mask: np.ndarray, tags: Optional[Dict[int, object]] = None,
                  compress: str = "LZW") -> None:
    """single-band uint8 class-id raster, LZW-compressed (``tiler.py:134-140``), geo-referenced like the input"""
    from PIL import Image, TiffImagePlugin

    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    if mask.ndim != 2:
        raise ValueError("write_geotiff stores one band (H, W)")
    info = TiffImagePlugin.ImageFileDirectory_v2()
    for t, v in (tags or {}).items():
        info[t] = v
        info.tagtype[t] = _TAG_TYPES.get(t, info.tagtype.get(t, 12))
    comp = {"LZW": "tiff_lzw", "DEFLATE": "tiff_adobe_deflate", "NONE": None}[compress.upper()]
    Image.fromarray(mask, "L").save(path, format="TIFF", compression=comp, tiffinfo=info)


def segment_files(inference, infiles: Iterable[Union[str, Path]], outdir: Union[str, Path],
                  is_valid: Optional[Callable[[np.ndarray], bool]] = None, workers: int = 2) -> List[Path]:
    """``inference``: a ``MosaicInference``.  Segments every GeoTIFF of ``infiles`` into ``outdir`` (same file names) with the
    decode of the next file and the encode of the previous mask running behind the GPU; -> the written paths.
    ``is_valid(band1)``: the reference's all-0 / all-255 filter (``scripts/inference.py:63-65``); invalid tiles are skipped."""
    import torch

    infiles = [Path(f) for f in infiles]
    outdir = Path(outdir)
    outdir.mkdir(parents=True, exist_ok=True)
    chans = inference.engine.in_channels
    written: List[Path] = []
    bufs: List[dict] = [{}, {}]                 # two sets of pinned staging buffers, sized on first use / growth

    def decode(i: int):
        # pinned (H, W, C) staging: decoded straight into it, so the upload needs no extra host copy
        from PIL import Image
        Image.MAX_IMAGE_PIXELS = None
        with Image.open(infiles[i]) as im:
            tags = {t: im.tag_v2[t] for t in GEO_TAGS if t in getattr(im, "tag_v2", {})}
            arr = np.asarray(im)
        if arr.ndim == 2:
            arr = arr[:, :, None]
        if arr.shape[2] < chans:
            raise ValueError(f"{infiles[i]}: {arr.shape[2]} bands, the model wants {chans}")
        b = bufs[i & 1]
        H, W, C = arr.shape
        if b.get("src") is None or b["src"].shape[0] < H or b["src"].shape[1] < W or b["src"].shape[2] != C:
            b["src"] = torch.empty((H, W, C), dtype=torch.uint8).pin_memory()
            b["mask"] = torch.empty((H, W), dtype=torch.uint8).pin_memory()
        src = b["src"][:H, :W]
        src.numpy()[...] = arr
        return src, b["mask"][:H, :W], tags

    with ThreadPoolExecutor(max_workers=max(2, workers)) as pool:
        nxt = pool.submit(decode, 0) if infiles else None
        pending_write = None
        for i, infile in enumerate(infiles):
            src, mask_host, tags = nxt.result()
            nxt = pool.submit(decode, i + 1) if i + 1 < len(infiles) else None        # decode the next file behind the GPU
            if is_valid is not None and not is_valid(src[..., 0].numpy()):
                continue
            H, W, C = src.shape
            if not src.is_contiguous():       # a staging buffer larger than this file: rows are strided views
                src = src.contiguous().pin_memory()
                mask_host = torch.empty((H, W), dtype=torch.uint8).pin_memory()
            dev = inference._buf("file_mosaic", (H, W, C), torch.uint8)
            inference.run(dev, "hwc", host_src=src, host_out=mask_host)               # bands up / mask bands down, pipelined
            torch.cuda.current_stream().synchronize()
            if pending_write is not None:
                pending_write.result()
            out = outdir / infile.name
            pending_write = pool.submit(write_geotiff, out, mask_host.numpy().copy(), tags)   # encode behind the next file
            written.append(out)
        if pending_write is not None:
            pending_write.result()
    return written
