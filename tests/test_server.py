"""REST entry point (deadtrees/deployment/server.py:87-128): request / response contract on the CPU with a stub model,
and the real GPU path behind the same route."""
import io

import numpy as np
import pytest
import torch
from PIL import Image
from starlette.testclient import TestClient

from deadtrees_b200.deployment import server
from deadtrees_b200.deployment.models import PredictionStats, predictionstats_to_str


def png_bytes(arr):
    b = io.BytesIO()
    Image.fromarray(arr).save(b, format="PNG")
    return b.getvalue()


def test_predictionstats_headers_are_strings():
    s = PredictionStats(fraction=0.25, model_name="bestmodel", model_type="pytorch", elapsed=0.5)
    assert predictionstats_to_str(s) == {"fraction": "0.25", "model_name": "bestmodel", "model_type": "pytorch", "elapsed": "0.5"}


def test_shim_paths():
    import deadtrees.deployment.models as m
    import deadtrees.deployment.server as s
    import deadtrees.utils.timer as t
    assert s.app is server.app and m.PredictionStats is PredictionStats and callable(t.record_execution_time)
    assert {r.path for r in s.app.routes} >= {"/", "/segmentation"}


def test_route_contract_without_a_checkpoint(tmp_path, monkeypatch):
    monkeypatch.setenv("DEADTREES_CHECKPOINT", str(tmp_path / "missing.ckpt"))
    c = TestClient(server.create_app())
    img = png_bytes(np.zeros((32, 32, 3), np.uint8))
    assert c.get("/").status_code == 200
    assert c.post("/segmentation", files={"file": ("a.png", img, "image/png")}).status_code == 503      # no checkpoint
    assert c.post("/segmentation?model_type=onnx", files={"file": ("a.png", img, "image/png")}).status_code == 501
    assert c.post("/segmentation?model_type=tf", files={"file": ("a.png", img, "image/png")}).status_code == 422
    assert c.post("/segmentation", files={"file": ("a.png", b"not an image", "image/png")}).status_code == 400
    assert c.post("/segmentation").status_code == 422                                                    # file is required


@pytest.mark.gpu
def test_segmentation_route_on_the_gpu_path(tmp_path):
    """POST an RGB PNG -> grey PNG of class ids * 255 + the reference's four headers; the mask equals the fp32 oracle's
    argmax on >= 99.9 % of the pixels (bf16 default precision), also for a size that is not a multiple of 32."""
    from deadtrees_b200.network.segmodel import SemSegment
    from gpu_util import pattern_mosaic, trained_model
    from oracle import ref_normalize, ref_unet
    model = trained_model(3, 3)
    net = dict(architecture="unet", encoder_name="resnet34", encoder_depth=5, encoder_weights=None,
               decoder_channels=[256, 128, 64, 32, 16], losses=["DICE", "FOCAL"], classes=["bg", "conifer", "broadleaf"],
               in_channels=3)
    seg = SemSegment(net, dict(learning_rate=3e-4, cosineannealing_tmax=10))
    seg.model.load_state_dict(model.state_dict())
    ckpt = tmp_path / "bestmodel.ckpt"
    seg.save_checkpoint(ckpt)
    c = TestClient(server.create_app(ckpt))
    for H, W in ((256, 256), (200, 150)):
        rgb = pattern_mosaic(H, W, 3, seed=H)
        r = c.post("/segmentation", files={"file": ("tile.png", png_bytes(rgb), "image/png")})
        assert r.status_code == 200 and r.headers["content-type"] == "image/png"
        assert r.headers["model_name"] == "bestmodel" and r.headers["model_type"] == "pytorch"
        assert float(r.headers["elapsed"]) > 0
        got = np.array(Image.open(io.BytesIO(r.content)))
        assert got.shape == (H, W) and got.dtype == np.uint8
        side = -(-max(H, W) // 32) * 32
        padded = np.zeros((side, side, 3), np.uint8)
        padded[:H, :W] = rgb
        x = torch.from_numpy(ref_normalize.val_transform(padded))[None]
        ref = ref_unet.run_inference(model, x, 3).numpy()[:H, :W]
        want = np.uint8(ref * 255)                      # the reference's np.uint8(out * 255): class 2 wraps to 254
        agree = float((got == want).mean())
        print(f"/segmentation {H}x{W}: agreement {agree:.5f}, fraction header {r.headers['fraction']}")
        assert agree >= 0.999
        assert abs(float(r.headers["fraction"]) - float(ref.sum() / ref.size)) < 5e-3
