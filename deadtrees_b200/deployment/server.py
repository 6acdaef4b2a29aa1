"""REST entry point on the B200 path - drop-in for ``deadtrees/deployment/server.py``.

``POST /segmentation`` (``server.py:87-128``): multipart ``file`` (any image PIL opens), optional query
``model_type`` in {pytorch, onnx} -> ``image/png`` (class ids * 255 as an 8-bit grey image) with the headers
``fraction, model_name, model_type, elapsed`` (``models.py:6-14``).  The request runs
``val_transform`` (uint8 -> normalised fp32, ``dt_tile_gather_normalize``) and ``PyTorchInference.run(device="cuda")``
(tcgen05 forward + ``dt_argmax_nchw``) - no CPU model.  Differences from the reference, all forced by this build's scope:

* the reference loads ``checkpoints/bestmodel.ckpt`` (+ ``.onnx``) at import time (``server.py:18-22``); here the
  checkpoint is loaded on the first request (or by ``create_app(model_file)``), from ``$DEADTREES_CHECKPOINT`` or the
  reference's path, so importing the module needs neither a GPU nor a file;
* ``model_type=onnx`` answers 501: ONNX export / onnxruntime is outside the B200 hot path (SURVEY.md D1 / DESIGN.md 6);
* images whose sides are not equal multiples of 32 are zero-padded to the next multiple and the mask is cropped back
  (the reference would fail inside smp's shape check).
"""
from __future__ import annotations

import io
import os
import threading
from enum import Enum
from pathlib import Path
from typing import Optional

import numpy as np
import torch
from fastapi import FastAPI, File, HTTPException
from PIL import Image
from starlette.responses import HTMLResponse, Response

from ..utils.timer import record_execution_time
from .models import PredictionStats, predictionstats_to_str

MODEL = "bestmodel"


class ModelTypes(Enum):
    """allowed model types"""

    PYTORCH = "pytorch"
    ONNX = "onnx"


def segment_image(model, image: Image.Image) -> np.ndarray:
    """RGB image -> uint8 class-id map of the same size through val_transform + ``model.run`` on the GPU."""
    from ..data.deadtreedata import val_transform
    arr = np.array(image.convert("RGB"))
    H, W = arr.shape[:2]
    side = max(32, -(-max(H, W) // 32) * 32)
    if (H, W) != (side, side):
        padded = np.zeros((side, side, 3), dtype=np.uint8)
        padded[:H, :W] = arr
        arr = padded
    input_tensor = val_transform(image=arr)["image"]
    out = model.run(input_tensor, device="cuda")
    if isinstance(out, torch.Tensor):
        out = out.detach().cpu().numpy()
    return np.asarray(out)[:H, :W]


def create_app(model_file: Optional[os.PathLike] = None, model=None) -> FastAPI:
    """``model_file``: Lightning ``.ckpt``; ``model``: an object with ``run(tensor, device=)`` (tests).  Default: the file
    named by ``$DEADTREES_CHECKPOINT``, else ``checkpoints/bestmodel.ckpt`` as in the reference."""
    app = FastAPI(
        title="DeadTrees image segmentation",
        description="Obtain semantic segmentation maps of the image in input via our UNet (B200 tcgen05 path).",
        version="0.1.0",
    )
    state = {"model": model, "lock": threading.Lock()}

    def get_model():
        with state["lock"]:
            if state["model"] is None:
                from .inference import PyTorchInference
                f = Path(model_file or os.environ.get("DEADTREES_CHECKPOINT", f"checkpoints/{MODEL}.ckpt"))
                if not f.exists():
                    raise HTTPException(status_code=503, detail=f"checkpoint {f} not found")
                state["model"] = PyTorchInference(f)
            return state["model"]

    @app.get("/", response_class=HTMLResponse, include_in_schema=False)
    async def root():
        return ("<!doctype html><html lang='en'><head><meta charset='utf-8'><title>DeadTrees Inference API</title>"
                "<meta http-equiv='refresh' content='7; URL=./docs' /></head><body><h1>DeadTrees Inference API</h1>"
                "<p>REST API for semantic segmentation of dead trees from ortho photos. "
                "<a href='./docs'>OpenAPI documentation</a></p></body></html>")

    @app.post("/segmentation")
    def get_segmentation_map(file: bytes = File(...), model_type: Optional[ModelTypes] = None):
        """Get segmentation maps from image file"""
        model_type = model_type or ModelTypes.PYTORCH
        if model_type == ModelTypes.ONNX:
            raise HTTPException(status_code=501, detail="only pytorch models are served by the B200 build")
        try:
            image = Image.open(io.BytesIO(file)).convert("RGB")
        except Exception as e:
            raise HTTPException(status_code=400, detail=f"cannot decode the image: {e}")
        pytorch_model = get_model()
        # call prediction and measure execution time
        with state["lock"], record_execution_time() as elapsed:      # one request at a time on the device
            out = segment_image(pytorch_model, image)
            seconds = elapsed()
        png = Image.fromarray(np.uint8(out * 255), "L")
        dead_tree_fraction = float(out.sum() / out.size)
        name = getattr(pytorch_model, "model_file", MODEL)
        stats = PredictionStats(fraction=dead_tree_fraction, model_name=Path(str(name)).stem or MODEL,
                                model_type=model_type.value, elapsed=seconds)
        bytes_io = io.BytesIO()
        png.save(bytes_io, format="PNG")
        return Response(bytes_io.getvalue(), headers=predictionstats_to_str(stats), media_type="image/png")

    return app


app = create_app()
