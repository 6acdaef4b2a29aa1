"""``Tiler`` — drop-in for ``deadtrees.deployment.tiler.Tiler`` with the block arithmetic on the GPU.

Mirrors ``deadtrees/deployment/tiler.py``: ``TileInfo`` (:22-25), ``divisible_without_remainder`` (:28-31),
``inspect_tile`` (:34-56), ``Tiler`` (:59-170).  GeoTIFF I/O (rioxarray) is optional — it is absent from
this image — so ``load_array`` / ``inspect_array`` take the raster as an ndarray of shape
(bands, H, W), exactly what ``rioxarray.open_rasterio(f).values`` returns at ``tiler.py:106``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from pathlib import Path
from typing import Optional, Tuple, Union

import numpy as np

from ..utils.data_handling import make_blocks_vectorized, unmake_blocks_vectorized


@dataclass
class TileInfo:
    size: Tuple[int, int]
    subtiles: Tuple[int, int]


def divisible_without_remainder(a, b):
    if b == 0:
        return False
    return True if a % b == 0 else False


def inspect_array(shape: Tuple[int, int], tile_shape=(8192, 8192), subtile_shape=(512, 512)) -> TileInfo:
    shape = tuple(shape)
    if not divisible_without_remainder(tile_shape[0], subtile_shape[0]):
        raise ValueError(f"Shapes unaligned (v): {tile_shape[0], subtile_shape[0]}")
    if not divisible_without_remainder(tile_shape[1], subtile_shape[1]):
        raise ValueError(f"Shapes unaligned (h): {tile_shape[1], subtile_shape[1]}")
    subtiles = (math.ceil(shape[0] / subtile_shape[0]), math.ceil(shape[1] / subtile_shape[1]))
    return TileInfo(size=shape, subtiles=subtiles)


def inspect_tile(infile, tile_shape=(8192, 8192), subtile_shape=(512, 512)) -> TileInfo:
    """Accepts a path (needs rioxarray), an xarray.DataArray, or an ndarray (bands, H, W) / (H, W)."""
    if isinstance(infile, np.ndarray):
        return inspect_array(infile.shape[-2:], tile_shape, subtile_shape)
    if hasattr(infile, "shape") and not isinstance(infile, (str, Path)):
        return inspect_array(tuple(infile.shape)[-2:], tile_shape, subtile_shape)
    try:
        import rioxarray  # optional dependency
    except ImportError:
        from PIL import Image
        Image.MAX_IMAGE_PIXELS = None
        with Image.open(infile) as im:           # header only: (width, height)
            return inspect_array((im.size[1], im.size[0]), tile_shape, subtile_shape)
    with rioxarray.open_rasterio(infile).sel(band=1, drop=True) as da:
        return inspect_array(tuple(da.shape), tile_shape, subtile_shape)


class Tiler:
    def __init__(self, infile: Optional[Union[str, Path]] = None, tile_shape=(2048, 2048), subtile_shape=(256, 256)):
        self._infile = infile
        self._tile_shape = tuple(tile_shape)
        self._subtile_shape = tuple(subtile_shape)
        if subtile_shape[0] != subtile_shape[1]:
            raise ValueError("Subtile required to have matching x/y dims")
        self._source = None
        self._target = None
        self._indata: Optional[np.ndarray] = None
        self._outdata: Optional[np.ndarray] = None
        self._batch_shape = None
        self._subtiles_to_use: Optional[np.ndarray] = None
        self._tile_info: Optional[TileInfo] = None

    # -- loading ---------------------------------------------------------------------------------
    def _configure(self, tile_shape, subtile_shape):
        self._tile_shape = tuple(tile_shape) if tile_shape else self._tile_shape
        if subtile_shape:
            if subtile_shape[0] != subtile_shape[1]:
                raise ValueError("Subtile required to have matching x/y dims")
        self._subtile_shape = tuple(subtile_shape) if subtile_shape else self._subtile_shape

    def load_array(self, sv: np.ndarray, tile_shape=None, subtile_shape=None) -> None:
        """``sv``: (bands, H, W) raster values; zero-padded to the tile shape (tiler.py:105-132)."""
        self._configure(tile_shape, subtile_shape)
        self._tile_info = inspect_array(sv.shape[1:], self._tile_shape, self._subtile_shape)
        if sv.shape[1] > self._tile_shape[0] or sv.shape[2] > self._tile_shape[1]:
            raise ValueError(f"raster {sv.shape[1:]} larger than tile_shape {self._tile_shape}")
        if self._tile_shape != self._tile_info.size:
            self._indata = np.zeros((sv.shape[0], *self._tile_shape), dtype=sv.dtype)
            self._indata[:, 0: sv.shape[1], 0: sv.shape[2]] = sv
        else:
            self._indata = sv
        self._outdata = np.zeros(self._tile_shape, dtype="uint8")
        mask = np.zeros((self._tile_shape[0] // self._subtile_shape[0], self._tile_shape[1] // self._subtile_shape[1]),
                        dtype=bool)
        mask[0: self._tile_info.subtiles[0], 0: self._tile_info.subtiles[1]] = 1
        self._subtiles_to_use = mask.ravel()

    def load_file(self, infile, tile_shape=None, subtile_shape=None) -> None:
        """``tiler.py:82-132``.  With rioxarray installed this is the reference's call sequence; without it (this image) the
        raster values and the geo-referencing tags come from ``deployment.geotiff.read_geotiff`` (Pillow)."""
        self._infile = infile
        try:
            import rioxarray  # optional dependency (absent in the build image)
        except ImportError:
            from .geotiff import read_geotiff
            values, self._geo_tags = read_geotiff(infile)
            self._source = None
            self.load_array(values, tile_shape, subtile_shape)
            self._target = "geotiff"          # marks: there is a geo-referenced target to write
            return
        self._source = rioxarray.open_rasterio(self._infile, chunks={"band": 4, "x": 256, "y": 256})
        self.load_array(self._source.values, tile_shape, subtile_shape)
        self._target = self._source.sel(band=1, drop=True).astype("uint8").copy(deep=True)

    def write_file(self, outfile) -> None:
        """LZW-compressed single-band mask, geo-referenced like the input (``tiler.py:134-140``)"""
        if isinstance(self._target, str):
            from .geotiff import write_geotiff
            write_geotiff(outfile, self.result, getattr(self, "_geo_tags", None))
        elif self._target is not None:
            self._target[:] = self.result
            self._target.rio.to_raster(outfile, compress="LZW", tiled=True)

    # -- batches ---------------------------------------------------------------------------------
    def get_batches(self) -> np.ndarray:
        subtiles = make_blocks_vectorized(self._indata, self._subtile_shape[0])
        self._batch_shape = self._batch_shape or subtiles.shape
        return subtiles[self._subtiles_to_use]

    def put_batches(self, batches: np.ndarray) -> None:
        batches = np.asarray(batches)
        d = self._subtile_shape[0]
        expanded = np.zeros((self._subtiles_to_use.size, d, d), dtype=np.uint8)
        expanded[self._subtiles_to_use] = batches.astype(np.uint8)  # uint8 on assignment (tiler.py:168)
        self._outdata = unmake_blocks_vectorized(expanded, d, self._tile_shape[0], self._tile_shape[1])
        if self._target is not None and not isinstance(self._target, str):
            self._target = self._target.load()
            self._target.loc[:] = self.result

    @property
    def result(self) -> np.ndarray:
        """stitched class ids cropped to the true raster size (what ``write_file`` stores)."""
        return self._outdata[0: self._tile_info.size[0], 0: self._tile_info.size[1]]
