"""Where the end-to-end gap of cfg2 comes from: device-resident vs upload only vs download only vs both (closed jobs and
pipelined), 5 steps each."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import torch
import bench
from deadtrees_b200.deployment.inference import MosaicInference
from deadtrees_b200.network.segmodel import SemSegment

dev = torch.device("cuda:0")
net = dict(architecture="unet", encoder_name="resnet34", encoder_depth=5, encoder_weights=None, decoder_channels=[256, 128, 64, 32, 16],
           losses=["DICE", "FOCAL"], classes=["bg", "a", "b"], in_channels=3, precision="bf16")
torch.manual_seed(0)
seg = SemSegment(net, dict(learning_rate=3e-4, cosineannealing_tmax=10)).eval().cuda()
mi = MosaicInference(seg.model.engine(), tile=256, overlap=32, batch_tiles=405)
mosaic = bench.synthetic_mosaic(10000, dev)
hsrc = torch.empty(mosaic.shape, dtype=torch.uint8, pin_memory=True); hsrc.copy_(mosaic)
hout = torch.empty((10000, 10000), dtype=torch.uint8, pin_memory=True)
mask = torch.zeros((10000, 10000), dtype=torch.uint8, device=dev)


def timed(fn, n=5):
    for _ in range(2):
        fn()
    mi.finish(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    mi.finish()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, kw in [("device", {}), ("device banded", dict(banded=True)), ("upload only", dict(host_src=hsrc)), ("download only", dict(host_out=hout)),
                 ("both closed", dict(host_src=hsrc, host_out=hout)), ("both pipelined", dict(host_src=hsrc, host_out=hout, pipelined=True)),
                 ("upload pipelined", dict(host_src=hsrc, pipelined=True)), ("device", {})]:
    print(f"{name:18s} {timed(lambda: mi.run(mosaic, 'hwc', out=mask, **kw)):8.3f} ms/step", flush=True)
for nbytes in (20 << 20, 100 << 20):
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev); h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    for direction in ("h2d", "d2h"):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
        e1.record(); torch.cuda.synchronize()
        print(f"{direction} {nbytes >> 20} MB: {5 * nbytes / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
