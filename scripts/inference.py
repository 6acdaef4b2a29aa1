"""Mosaic inference CLI — B200 drop-in for the reference's ``scripts/inference.py`` (same flags).

Per input raster: load -> zero-pad to the tile shape -> sub-tiles -> normalise -> Unet -> argmax -> stitch ->
crop -> write, as ``scripts/inference.py:80-111`` of the reference; the per-batch Python loop there
(:93-105) is one device pipeline here (``MosaicInference``).  GeoTIFF I/O: rioxarray when installed, else the GDAL-free
reader / writer of ``deadtrees_b200.deployment.geotiff``; with ``--overlap N --pipeline`` a directory of GeoTIFF tiles runs
through ``geotiff.segment_files`` (whole rasters on the device, the next file decoded and the previous mask encoded behind the
GPU).  ``.npy`` rasters of shape (bands, H, W) uint8 are read and written as they are.
"""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from deadtrees_b200.deployment.inference import MosaicInference, PyTorchEnsembleInference, PyTorchInference  # noqa: E402
from deadtrees_b200.deployment.tiler import Tiler  # noqa: E402


def is_valid_tile(band1: np.ndarray) -> bool:
    return False if np.isin(band1, [0, 255]).all() else True  # scripts/inference.py:63-65


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("infile", type=Path)
    parser.add_argument("-m", "--model", dest="model", action="append", type=Path, default=[], help="model artefact")
    parser.add_argument("-o", dest="outpath", type=Path, default=Path("."), help="output directory")
    parser.add_argument("--all", action="store_true", dest="all", default=False, help="process complete directory")
    parser.add_argument("--nopreview", action="store_false", dest="preview", default=True, help="produce preview images")
    parser.add_argument("--overlap", type=int, default=0, help="extension: overlapping sub-tiles with blended stitching")
    parser.add_argument("--pipeline", action="store_true", default=False,
                        help="extension: GeoTIFF inputs through the double-buffered file pipeline (no padding to the tile shape)")
    args = parser.parse_args()

    if len(args.model) == 0:
        args.model = [Path("checkpoints/bestmodel.ckpt")]
    if len(args.model) == 1:
        print("Default inference: single model")
        inference = PyTorchInference(args.model[0])
    else:
        print(f"Ensemble inference: {len(args.model)} models")
        inference = PyTorchEnsembleInference(*args.model)

    if args.all:
        infiles = sorted(args.infile.glob("ortho*.tif")) + sorted(args.infile.glob("ortho*.npy"))
    else:
        infiles = [args.infile]

    if args.pipeline and isinstance(inference, PyTorchInference) and all(f.suffix != ".npy" for f in infiles):
        from deadtrees_b200.deployment import geotiff
        pipe = MosaicInference(inference._model.cuda(), tile=256, overlap=args.overlap, batch_tiles=64)
        written = geotiff.segment_files(pipe, infiles, args.outpath, is_valid=is_valid_tile)
        print(f"{len(written)} of {len(infiles)} tiles segmented -> {args.outpath}")
        return

    pipe = None
    for infile in infiles:
        tiler = Tiler()
        if infile.suffix == ".npy":
            tiler.load_array(np.load(infile))
        else:
            tiler.load_file(infile)
        if not is_valid_tile(tiler._indata[0, : tiler._tile_info.size[0], : tiler._tile_info.size[1]]):
            continue
        args.outpath.mkdir(parents=True, exist_ok=True)
        outfile = args.outpath / infile.name
        if isinstance(inference, PyTorchInference):
            if pipe is None:
                pipe = MosaicInference(inference._model.cuda(), tile=tiler._subtile_shape[0], overlap=args.overlap,
                                       batch_tiles=64)
            # whole padded tile on the device: gather + normalise + Unet + argmax + stitch, then crop as write_file
            tiler._outdata = pipe.run_host(np.ascontiguousarray(tiler._indata), "chw")
        else:  # ensemble: the reference's batch loop with the per-pixel mode over models
            import torch
            from deadtrees_b200.data.deadtreedata import val_transform
            out = []
            batches = tiler.get_batches()
            for b in np.array_split(batches, max(1, int(np.ceil(len(batches) / 64))), axis=0):
                x = torch.stack([val_transform(image=i.transpose(1, 2, 0))["image"] for i in b])
                out.append(inference.run(x, device="cuda").cpu().numpy())
            tiler.put_batches(np.concatenate(out, axis=0))
        if infile.suffix == ".npy":
            np.save(outfile, tiler.result)
        else:
            tiler.write_file(outfile)
        if args.preview:
            from PIL import Image
            prev = Path(str(args.outpath) + "_preview")
            prev.mkdir(parents=True, exist_ok=True)
            Image.fromarray(np.uint8(tiler.result * 255), "L").save(prev / (infile.stem + ".png"))


if __name__ == "__main__":
    main()
