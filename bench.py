#!/usr/bin/env python
"""Benchmark of the deadtrees hot path on B200: whole-mosaic Unet-resnet34 inference (BASELINE cfg 2/3).

    python bench.py --gpus N --steps K --warmup W            # this framework (hand-written CUDA)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port) on host cores

One step = one pass of the hot path over the synthetic 10 000 x 10 000 RGB uint8 mosaic (tile 256,
overlap 32 -> 45 x 45 = 2025 tiles): gather+normalise -> Unet -> head -> blended stitch -> uint8 mask.
N > 1 (torchrun): the same mosaic sharded by tile rows, one boundary-logits send per shard boundary,
masks gathered on rank 0; timing = max over ranks of the CUDA-event time, barrier on both sides.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "unet_mosaic_inference_tiles_per_s"
UNIT = "tiles/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="infer", choices=["infer", "train"],
                    help="infer: cfg2/cfg3 whole-mosaic inference (the headline metric); train: cfg4 training step, "
                         "64 RGB+NIR tiles of 256x256 per GPU, Dice+Focal, clip 0.5, Adam, data parallel")
    ap.add_argument("--train-batch", type=int, default=64, help="tiles per GPU per training step (cfg4)")
    ap.add_argument("--no-graph", action="store_true", help="train workload: launch the ~400 kernels of a step one by one "
                                                            "from Python instead of replaying the captured CUDA graph")
    ap.add_argument("--size", type=int, default=10000, help="mosaic side in pixels")
    ap.add_argument("--tile", type=int, default=256)
    ap.add_argument("--overlap", type=int, default=32)
    ap.add_argument("--batch-tiles", type=int, default=1013,
                    help="upper bound of the tiles per Unet launch sequence (the batches of a mosaic are equal): 1013 = two "
                         "batches of the cfg2 grid.  Measured on B200 with the final round-2 kernels, ms per mosaic device-"
                         "resident / end to end pipelined / end to end as closed jobs: 338 -> 34.0 / 34.4 / -, 405 -> 33.0-33.3 / "
                         "33.3-33.8 / 35.9, 675 -> 32.7-33.2 / 33.2-33.8 / 35.5, 1013 -> 31.7-32.6 / 32.2-32.9 / 35.5-36.1, 2025 -> "
                         "32.1-32.5 / 32.9-33.2 / 37.7-38.4 (boxes differ by their power cap; round 1 kernels: 405 was best)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-sample-tiles", type=int, default=36)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--layer-table", default=None, help="write per-layer conv timings (CUDA events) to this file")
    ap.add_argument("--no-profile", action="store_true", help="skip the per-kernel CUDA-event instrumentation")
    ap.add_argument("--no-extra", action="store_true",
                    help="infer workload: skip the cfg5 (1024x1024 tiles, batch 32 per GPU) and cfg4 (training step) "
                         "measurements that the default run appends to the JSON line as `cfg5` / `cfg4`")
    ap.add_argument("--extra-steps", type=int, default=20, help="timed steps of the appended cfg4 / cfg5 measurements")
    ap.add_argument("--e2e-no-overlap", action="store_true",
                    help="end to end: every mosaic is a closed job (no overlap of successive mosaics' copies with compute)")
    ap.add_argument("--e2e-batches", default=None,
                    help="N > 1: batch sizes of the end-to-end shard pipeline, comma separated (the last one repeats)")
    ap.add_argument("--e2e-sweep", default=None,
                    help="N > 1 experiment: ';'-separated batch plans, each timed end to end and reported on stderr")
    return ap.parse_args()


def workload_config(a, n_tiles):
    return {"workload": f"cfg2: sliding-window Unet-resnet34 inference over a synthetic {a.size}x{a.size} RGB uint8 "
                        f"mosaic, tile {a.tile}, overlap {a.overlap} ({n_tiles} tiles), blended stitch",
            "mosaic": [a.size, a.size, 3], "tile": a.tile, "overlap": a.overlap, "tiles": n_tiles,
            "batch_tiles": a.batch_tiles, "classes": 3, "in_channels": 3,
            "l2_policy": "inputs larger than L2 (300 MB mosaic, >1 GB of activations per step); no explicit flush",
            "parallelism": f"tile-range shards x{a.gpus} (row-aligned stitch, neighbour exchange, mask rows on rank 0)"}


def synthetic_mosaic(size: int, device, seed: int = 1234) -> torch.Tensor:
    """low-frequency pattern + noise so the predicted masks are not constant."""
    g = torch.Generator(device=device).manual_seed(seed)
    yy = torch.arange(size, device=device, dtype=torch.float32)[:, None]
    xx = torch.arange(size, device=device, dtype=torch.float32)[None, :]
    out = torch.empty((size, size, 3), dtype=torch.uint8, device=device)
    for c, (fy, fx, amp) in enumerate([(97.0, 131.0, 80.0), (61.0, 173.0, 70.0), (149.0, 83.0, 90.0)]):
        base = 127.0 + amp * torch.sin(yy / fy) * torch.cos(xx / fx)
        noise = torch.randn((size, size), device=device, generator=g) * 20.0
        out[..., c] = (base + noise).clamp_(0, 255).to(torch.uint8)
    return out


def ncu_conv_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the conv launches of one 135-tile batch, from the committed
    `ncu --set full` capture (profiles/r01_convs_ncu_full_v6.txt, scripts/gpu_profile2.sh) - not measured in this run."""
    import re
    cands = sorted((ROOT / "profiles").glob("r0*_convs_ncu_full*.txt"))
    for f in reversed(cands):
        m = re.search(r"= ([0-9.]+) GB per (\d+)-tile batch", f.read_text())
        n = sum(1 for l in f.read_text().splitlines() if "_kernel" in l)
        if m and n:
            return float(m.group(1)) * 1e9, int(m.group(2)), n, f.name
    return None


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "MEASURED_PEAKS.json (sustained bf16, copy HBM)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons during the timed region (pynvml; nvidia-smi as a fallback)."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.backend = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.backend = "nvml"
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "backend": self.backend}


# ---------------------------------------------------------------------------------------------------
def load_reference_tiler():
    """the reference's own make/unmake_blocks_vectorized from baseline/_ref (pip-installed copy of the
    unmodified reference; loaded by file path because the repo's `deadtrees` shim has the same package
    name).  None when the install is absent."""
    import importlib.util
    f = ROOT / "baseline" / "_ref" / "deadtrees" / "utils" / "data_handling.py"
    if not f.exists():
        return None
    spec = importlib.util.spec_from_file_location("_reference_data_handling", f)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_reference_flow(model, block_chw: np.ndarray, tile: int, ref_dh):
    """the reference's per-file flow on the CPU (scripts/inference.py:85-111) for one (C, m, n) uint8 block:
    make_blocks -> per-tile normalise -> Unet forward -> argmax -> unmake_blocks.  Tiling uses the reference's
    own functions when baseline/_ref exists; normalise / Unet are the oracle port (smp, albumentations absent)."""
    from oracle import ref_normalize, ref_tiler
    mk = ref_dh.make_blocks_vectorized if ref_dh else ref_tiler.make_blocks
    unmk = ref_dh.unmake_blocks_vectorized if ref_dh else ref_tiler.unmake_blocks
    tiles = mk(block_chw, tile)
    x = torch.stack([torch.from_numpy(ref_normalize.val_transform(t.transpose(1, 2, 0))) for t in tiles])
    with torch.no_grad():
        out = model(x).argmax(dim=1).numpy()
    return unmk(out, tile, block_chw.shape[1], block_chw.shape[2])


def run_cpu_sample(a, n_tiles_sample: int, repeats: int, warm: int):
    """times the reference flow (oracle port of the Unet, fp32 torch CPU, all host threads) on a bounded
    sample of the workload: one square block of ~n_tiles_sample tiles."""
    from oracle import ref_unet
    torch.set_num_threads(os.cpu_count() or 1)
    model = ref_unet.build_reference_unet(3, 3, seed=0)
    ref_dh = load_reference_tiler()
    rng = np.random.default_rng(1234)
    side = max(1, int(round(np.sqrt(n_tiles_sample))))
    block = rng.integers(0, 256, size=(3, side * a.tile, side * a.tile), dtype=np.uint8)
    times = []
    for i in range(warm + repeats):
        t0 = time.perf_counter()
        cpu_reference_flow(model, block, a.tile, ref_dh)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return side * side / min(times), times, side * side, ref_dh is not None


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from deadtrees_b200.deployment.inference import overlap_grid
    gy, gx = overlap_grid(a.size, a.size, a.tile, a.overlap)
    n_sample = a.cpu_sample_tiles
    tps, times, n_sample, used_ref = run_cpu_sample(a, n_sample, repeats=a.steps, warm=a.warmup)
    ms = 1000.0 * statistics.mean(times)
    sample = f"one block of {n_sample} tiles of {a.tile}x{a.tile} per step through the reference flow " \
             f"(make_blocks -> normalise -> Unet fp32 -> argmax -> unmake_blocks; tiler = " \
             f"{'reference code from baseline/_ref' if used_ref else 'oracle port'}, Unet/normalise = oracle port), " \
             f"torch CPU, {torch.get_num_threads()} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(a, gy * gx),
        "mpixel_per_s": tps * a.tile * a.tile / 1e6,
        "cpu_baseline": {"value": tps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------------
def main_b200(a):
    import torch.distributed as dist
    from deadtrees_b200 import ops
    from deadtrees_b200.deployment.inference import MosaicInference, overlap_grid
    from deadtrees_b200.engine import UnetEngine, conv_flops_per_tile
    from deadtrees_b200.sharding import ShardPlan, exchange_logits, gather_mask_rows

    ctx = dist_context()
    rank, world, local, dev = ctx

    # random-init weights of the reference architecture (seeded; the oracle is only the INITIALISER here,
    # so that bench, tests and the CPU baseline share one state-dict) -- built on CPU, then handed over
    from deadtrees_b200.network.segmodel import SemSegment
    net = dict(architecture="unet", encoder_name="resnet34", encoder_depth=5, encoder_weights=None,
               decoder_channels=[256, 128, 64, 32, 16], losses=["DICE", "FOCAL"], classes=["bg", "a", "b"],
               in_channels=3, precision=a.precision)
    torch.manual_seed(0)
    seg = SemSegment(net, dict(learning_rate=3e-4, cosineannealing_tmax=10)).eval()
    g = torch.Generator().manual_seed(0)
    for mod in seg.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
    seg.cuda()
    engine = seg.model.engine()

    H = W = a.size
    T, ov = a.tile, a.overlap
    gy, gx = overlap_grid(H, W, T, ov)
    n_tiles = gy * gx
    # N > 1: contiguous tile-index ranges (253 / 254 tiles per rank at N = 8), row-aligned stitching, one grouped
    # neighbour exchange of head tiles / boundary strips, mask rows gathered on rank 0 (deadtrees_b200/sharding.py)
    plans = [ShardPlan(gy, gx, world, r, ov) for r in range(world)]
    plan = plans[rank]
    r0, r1 = plan.R0, plan.R1
    my_tiles = plan.t1 - plan.t0
    mi = MosaicInference(engine, tile=T, overlap=ov, batch_tiles=a.batch_tiles)
    # end to end the mosaic rows of a batch are uploaded behind the previous batch.  The first upload cannot be hidden, so
    # the pipeline (deployment/inference.py batch_plan) starts with ONE tile row and sends the rest in equal batches of
    # at most batch_tiles, which keeps the deep layers' grids full (measured at N = 8, ms per step: [45, 209] 6.30,
    # one batch 6.38, three equal batches 7.41)
    bt_e2e = None
    if a.e2e_batches:
        bt_e2e = [int(v) for v in a.e2e_batches.split(",")]
    mosaic = synthetic_mosaic(a.size, dev)
    host_mosaic = torch.empty(mosaic.shape, dtype=torch.uint8, pin_memory=True)
    host_mosaic.copy_(mosaic)
    host_mask = torch.empty((H, W), dtype=torch.uint8, pin_memory=True)
    mask = torch.zeros((H, W), dtype=torch.uint8, device=dev)
    y0, y1 = plan.mask_rows(H, T)
    exchange = (lambda lg: exchange_logits(plan, lg, T)) if world > 1 else None

    def step_device():
        if world == 1:
            mi.run(mosaic, "hwc", out=mask)
        else:
            mi.run_shard(mosaic, plan, mask, exchange=exchange)
            gather_mask_rows(plans, rank, mask, H, T)

    def step_e2e():
        # host buffers in and out: H2D of the uint8 mosaic rows this shard's tiles read, D2H of its mask rows
        if world == 1:
            # MosaicInference's host pipeline: row bands go up on a copy stream while earlier batches compute, finished
            # mask bands come back behind the compute
            mi.run(mosaic, "hwc", out=mask, host_src=host_mosaic, host_out=host_mask, pipelined=pipelined)
            return H * W * 3, H * W
        mi.run_shard(mosaic, plan, mask, exchange=exchange, host_src=host_mosaic, host_out=host_mask, batch_tiles=bt_e2e,
                     pipelined=pipelined)
        ya, yb = plan.input_rows(H, T)
        return (yb - ya) * W * 3, (y1 - y0) * W

    # back-to-back mosaics overlap like a production loop over files (scripts/inference.py --pipeline): the next mosaic's rows go
    # up as soon as the current one's last gather has read the staging buffer, the last mask band comes back behind the next
    # mosaic's first batch; mi.finish() joins the copy streams INSIDE the timed region
    pipelined = not a.e2e_no_overlap

    def timed(fn, steps, profile=False, after=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ops.PROFILE = {} if profile else None
        l0 = ops.LAUNCHES
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if after is not None:
            after()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        prof, ops.PROFILE = ops.PROFILE, None
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ops.LAUNCHES - l0, prof

    for _ in range(max(a.warmup, 3)):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    ms_total, launches, _ = timed(step_device, a.steps)
    clocks = sampler.stop()
    ms_step = ms_total / a.steps
    value = n_tiles / (ms_step / 1e3)
    # end-to-end through the public pipeline with host buffers (pinned), same number of steps - timed right after the
    # device-resident steps: under the power cap the SM clock keeps sinking for seconds (scripts/exp/e2e_split.py: the same
    # device step 34.3 ms at the start of a script and 35.0 ms three seconds later), so measurements that are compared
    # should be neighbours in time
    for _ in range(2):
        step_e2e()
    mi.finish()
    ms_e2e_total, _, _ = timed(step_e2e, a.steps, after=mi.finish)
    h2d, d2h = step_e2e()
    mi.finish()
    torch.cuda.synchronize()
    ms_e2e = ms_e2e_total / a.steps
    ms_closed = None
    if pipelined:            # the same pipeline with every mosaic a closed job (nothing of step k+1 starts before step k's mask is back)
        pipelined = False
        for _ in range(2):
            step_e2e()
        ms_closed = timed(step_e2e, a.steps)[0] / a.steps
        pipelined = True
    # per-launch CUDA events (the roofline numbers) are taken in a separate pass so that the ~1500 event records of a
    # step do not sit inside the headline's timed region
    prof, ms_prof_total = None, ms_total
    if not a.no_profile:
        ms_prof_total, _, prof = timed(step_device, a.steps, profile=True)
    if a.e2e_sweep and world > 1:
        keep = bt_e2e
        for spec in a.e2e_sweep.split(";"):
            bt_e2e = [int(v) for v in spec.split(",")]
            for _ in range(2):
                step_e2e()
            mi.finish()
            ms_sw, _, _ = timed(step_e2e, a.steps, after=mi.finish)
            if rank == 0:
                print(f"[e2e sweep] batches {spec}: {ms_sw / a.steps:.3f} ms/step", file=sys.stderr, flush=True)
        bt_e2e = keep

    pk = peaks()
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": a.precision if a.precision == "bf16" else "f32", "data": "synthetic",
        "config": workload_config(a, n_tiles),
        "mpixel_per_s": value * T * T / 1e6, "mosaic_mpixel_per_s": H * W / 1e6 / (ms_step / 1e3),
        "e2e": {"value": n_tiles / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": ms_e2e,
                "pipeline": ("successive mosaics overlap (next upload behind the current mosaic's last batch, last mask band "
                             "behind the next mosaic's first batch); all copies and MosaicInference.finish() inside the timed "
                             "region" if pipelined else "every mosaic a closed job"),
                "closed_job": (None if ms_closed is None else
                               {"value": n_tiles / (ms_closed / 1e3), "unit": UNIT, "ms_per_step": ms_closed})},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if prof:
        def agg(kind):
            evs = prof.get(kind, [])
            t = sum(ev[0].elapsed_time(ev[1]) for ev in evs) / 1e3
            w = sum(ev[2] for ev in evs)
            return t, w, len(evs)
        tc, wc, nc = agg("conv")
        tg, wg, ng = agg("gather")
        ts, ws, ns = agg("stitch")
        if tc > 0:
            ach = wc / tc / 1e12
            out["roofline"] = {"bound": "tensor",
                               "kernel": "all conv launches of the Unet (tcgen05 implicit GEMM: conv_stem_rows, conv_row, "
                                         "conv_res, conv_halo, conv_halo_quad, conv_pair, conv_tc kernels)",
                               "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"],
                               "traffic": None, "launches": nc, "share_of_step": tc * 1e3 / ms_prof_total,
                               "flops_per_tile": conv_flops_per_tile(T, 3, 3), "peak_source": pk["source"]}
            tr = ncu_conv_traffic()
            if tr is not None and T == 256:
                bt = min(a.batch_tiles, max(my_tiles, 1))
                out["roofline"]["traffic"] = tr[0] / tr[1] * bt / tr[2]
                out["roofline"]["traffic_source"] = (
                    f"NOT measured in this run: committed ncu --set full capture of the {tr[2]} conv launches of one "
                    f"{tr[1]}-tile batch (profiles/{tr[3]}): {tr[0] / 1e9:.3f} GB DRAM read + write per batch = "
                    f"{tr[0] / tr[1] / 1e6:.1f} MB per tile; scaled to this run's {bt}-tile batches and averaged per "
                    f"launch; bf16 activations in + out of all convs, unfused: 44.6 MB per tile")
        hb = {}
        if tg > 0:
            # algorithmic bytes: T^2 * C * (1 B in + 2 B out) per tile; the bytes the kernel really moves are
            # T^2 * (3 B in + 8 B out): the stem's TMA im2col map needs 4-channel (8-byte) pixels
            moved = wg * (3 + 8) / 9.0
            hb["gather_normalize"] = {"achieved": wg / tg / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                      "frac": wg / tg / 1e9 / pk["hbm_gbs"], "launches": ng,
                                      "moved_gbs": moved / tg / 1e9, "frac_moved": moved / tg / 1e9 / pk["hbm_gbs"]}
        if ts > 0:
            # algorithmic bytes: every covering logit once + 1 B of mask per pixel (ops.stitch_blend_argmax)
            hb["stitch"] = {"achieved": ws / ts / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                            "frac": ws / ts / 1e9 / pk["hbm_gbs"], "launches": ns}
        if world == 1 and n_tiles * (T + 6) * (T + 8) * 8 < (8 << 30):
            # the same gather kernel over the WHOLE mosaic in one launch (the pipeline launches it per batch of
            # batch_tiles tiles, ~80 MB, where launch ramp and tail are a third of the kernel's 20 us)
            frame = torch.zeros((n_tiles, T + 6, T + 8, 4), dtype=torch.bfloat16, device=dev)
            from deadtrees_b200.data.deadtreedata import normalize_constants
            off, sc = normalize_constants(3, None, None)
            best = None
            for i in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.tile_gather_normalize(mosaic, "hwc", 3, T, ov, (gy, gx), 0, n_tiles, off, sc, out=frame, pad=3)
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    best = min(best, e0.elapsed_time(e1)) if best else e0.elapsed_time(e1)
            gb = n_tiles * T * T * 3 * (1 + 2) / 1e9
            hb["gather_normalize_whole_mosaic"] = {"achieved": gb / (best / 1e3), "peak": pk["hbm_gbs"], "unit": "GB/s",
                                                   "frac": gb / (best / 1e3) / pk["hbm_gbs"], "launches": 1,
                                                   "us": 1e3 * best, "bytes": gb * 1e9,
                                                   "frac_moved": gb * 11.0 / 9.0 / (best / 1e3) / pk["hbm_gbs"]}
            del frame
        out["roofline_hbm"] = hb
        if a.layer_table and rank == 0:
            per = {}
            for e0, e1, wk, tag in prof.get("conv", []):
                d = per.setdefault(tag, [0.0, 0.0, 0])
                d[0] += e0.elapsed_time(e1); d[1] += wk; d[2] += 1 if wk > 0 else 0   # (helper launches of a layer carry no work)
            with open(a.layer_table, "w") as fh:
                fh.write(f"{'layer':34s} {'launches':>8s} {'avg_us':>9s} {'TFLOP/s':>9s} {'share%':>7s}\n")
                for tag, (ms, wk, n) in per.items():
                    fh.write(f"{tag:34s} {n:8d} {1e3 * ms / n:9.1f} {wk / (ms / 1e3) / 1e12:9.1f} {100 * ms / ms_prof_total:7.2f}\n")
                fh.write(f"conv total {1e3 * tc:.1f} ms of {ms_prof_total:.1f} ms; gather {1e3 * tg:.2f} ms; stitch {1e3 * ts:.2f} ms\n")
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        tps, times, n_s, used_ref = run_cpu_sample(a, a.cpu_sample_tiles, repeats=3, warm=1)
        out["cpu_baseline"] = {"value": tps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"one block of {n_s} tiles of {T}x{T} through the reference flow (make_blocks -> "
                                         f"normalise -> Unet fp32 -> argmax -> unmake_blocks; Unet = oracle port, tiler = "
                                         f"{'reference code (baseline/_ref)' if used_ref else 'oracle port'}), best of 3"}
    gstep = None
    if not a.no_extra and a.precision == "bf16":
        # BASELINE configs[4] and configs[3] behind the same default invocation, at every N: extra keys of the one line
        del mi, mosaic, mask, host_mosaic, host_mask
        engine._ws.clear()
        torch.cuda.empty_cache()
        out["cfg5"] = cfg5_measure(a, ctx, engine, a.extra_steps, 3)
        out["cfg4"], gstep = train_measure(a, ctx, a.extra_steps, 3, with_profile=True, with_cpu=not a.no_cpu_baseline)
    if rank == 0:
        print(json.dumps(out), flush=True)
    finish(world, gstep)


# ---------------------------------------------------------------------------------------------------
def train_flops_per_tile(T: int, cin: int, classes: int) -> float:
    """algorithmic conv FLOPs of one training step per tile: forward + data gradient + weight gradient of every conv
    (3 x forward) minus the stem's data gradient, which is never needed."""
    from deadtrees_b200.engine import conv_flops_per_tile
    return 3.0 * conv_flops_per_tile(T, cin, classes) - 2.0 * (T // 2) ** 2 * 64 * cin * 49


def synthetic_train_batch(B: int, T: int, cin: int, K: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    u8 = torch.randint(0, 256, (B, cin, T, T), generator=g, dtype=torch.uint8)
    from deadtrees_b200.data.deadtreedata import normalize_constants
    off, sc = normalize_constants(cin, None, None)
    img = (u8.float() - torch.tensor(off[:cin]).view(1, cin, 1, 1)) * torch.tensor(sc[:cin]).view(1, cin, 1, 1)
    yy, xx = torch.meshgrid(torch.arange(T), torch.arange(T), indexing="ij")
    mask = torch.stack([(((yy + 7 * i) // 23 + (xx + 3 * i) // 17) % K) for i in range(B)]).long()
    return img.contiguous(), mask.contiguous()


def main_train_reference(a):
    """cfg4 on the host cores: the oracle's training step (torch CPU autograd + the reference's loss terms + clip + Adam)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import ref_train, ref_unet
    torch.set_num_threads(os.cpu_count() or 1)
    B, T, cin, K = 8, a.tile, 4, 3
    model = ref_unet.build_reference_unet(cin, K, seed=0)
    img, mask = synthetic_train_batch(B, T, cin, K, 1234)
    opt = torch.optim.Adam(model.parameters(), lr=3e-4)
    times = []
    for i in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        ref_train.train_step(model, img, mask, lr=3e-4, clip=0.5, optimizer=opt)
        if i >= a.warmup:
            times.append(time.perf_counter() - t0)
    tps = B / statistics.mean(times)
    sample = f"one training step on {B} RGB+NIR tiles of {T}x{T} per timed step (oracle Unet autograd fp32 + reference loss " \
             f"terms + clip_grad_norm_(0.5) + torch.optim.Adam), torch CPU, {torch.get_num_threads()} threads"
    print(json.dumps({
        "impl": "reference", "metric": "unet_train_tiles_per_s", "value": tps, "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * statistics.mean(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"cfg4: Unet-resnet34 training step, {a.train_batch} RGB+NIR {T}x{T} tiles per GPU, Dice+Focal, "
                               f"clip 0.5, Adam lr 3e-4", "tile": T, "batch_per_gpu": a.train_batch, "in_channels": 4, "classes": 3},
        "cpu_baseline": {"value": tps, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def dist_context():
    """(rank, world, local, device); initialises NCCL under torchrun."""
    import torch.distributed as dist
    from deadtrees_b200._lib import require_device
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    require_device()
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    return rank, world, local, dev


def train_measure(a, ctx, steps: int, warmup: int, with_profile: bool, with_cpu: bool):
    """cfg4: one data-parallel training step per timed step -> (JSON dict, GraphedTrainStep or None)."""
    import torch.distributed as dist
    from deadtrees_b200 import ops
    from deadtrees_b200.network.segmodel import SemSegment

    rank, world, local, dev = ctx
    B, T, cin, K = a.train_batch, 256, 4, 3
    net = dict(architecture="unet", encoder_name="resnet34", encoder_depth=5, encoder_weights=None,
               decoder_channels=[256, 128, 64, 32, 16], losses=["DICE", "FOCAL"], classes=["bg", "a", "b"],
               in_channels=cin, precision=a.precision)
    torch.manual_seed(0)                       # identical replicas on every rank
    seg = SemSegment(net, dict(learning_rate=3e-4, cosineannealing_tmax=10, gradient_clip_val=0.5)).cuda().train()
    eng = seg.model.train_engine()
    if world > 1:
        eng.set_process_group(None, world_size=world)
    (opt,), _ = seg.configure_optimizers()
    img_h, mask_h = synthetic_train_batch(B, T, cin, K, 1234 + rank)       # per-rank batch
    img_h, mask_h = img_h.pin_memory(), mask_h.pin_memory()
    img_d, mask_d = img_h.to(dev), mask_h.to(dev)
    lu = torch.zeros(B)
    stats = [{"file": f"r{rank}t{i}"} for i in range(B)]
    loss_h = torch.zeros((), pin_memory=True)

    use_graph = not a.no_graph
    launches_per_replay = 0
    if use_graph:
        from deadtrees_b200.train_graph import GraphedTrainStep
        l0 = ops.LAUNCHES
        gstep = GraphedTrainStep(seg, opt, B, T)              # world > 1: the NCCL bucket all-reduces are graph nodes too
        launches_per_replay = (ops.LAUNCHES - l0) // 3        # two warm-up passes + the captured one
        gstep.img.copy_(img_d)
        gstep.mask.copy_(mask_d)

    def eager_step():
        loss = seg.training_step({"main": (img_d, mask_d, None, lu, stats)}, 0)
        loss.backward()
        opt.step()
        return loss

    def step_device():
        if use_graph:                        # inputs already in the graph's static buffers
            gstep.replay()
            return gstep.terms.total_loss
        return eager_step()

    def step_e2e():
        if use_graph:
            # double-buffered input: this step's batch was uploaded (pinned host -> staging) behind the previous step;
            # hand it over, replay, and start the upload of the next step's batch - one upload per step
            if not getattr(gstep, "_staged", False):
                gstep.prefetch(img_h, mask_h)
            loss = gstep.step_prefetched()
            gstep.prefetch(img_h, mask_h)
        else:
            img_d.copy_(img_h, non_blocking=True)
            mask_d.copy_(mask_h, non_blocking=True)
            loss = eager_step()
        loss_h.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(loss_h)

    # back-to-back mosaics overlap like a production loop over files (scripts/inference.py --pipeline): the next mosaic's rows go
    # up as soon as the current one's last gather has read the staging buffer, the last mask band comes back behind the next
    # mosaic's first batch; mi.finish() joins the copy streams INSIDE the timed region
    pipelined = not a.e2e_no_overlap

    def timed(fn, steps, profile=False, after=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ops.PROFILE = {} if profile else None
        l0 = ops.LAUNCHES
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if after is not None:
            after()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        prof, ops.PROFILE = ops.PROFILE, None
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ops.LAUNCHES - l0, prof

    for _ in range(max(warmup, 3)):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    ms_total, launches, _ = timed(step_device, steps)
    clocks = sampler.stop()
    if use_graph:
        launches = launches_per_replay * steps          # kernels replayed by the graph (no per-kernel Python call)
    ms_step = ms_total / steps
    # per-kernel CUDA events need individually launched kernels: a separate, eager, instrumented pass (single GPU: the
    # eager step would re-bucket nothing, but its per-kernel NCCL launches are not what the graph replays)
    prof = None
    psteps = min(steps, 5)
    if with_profile and world == 1:
        eager_step()
        _, _, prof = timed(eager_step, psteps, profile=True)
    for _ in range(2):
        step_e2e()
    ms_e2e_total, _, _ = timed(step_e2e, steps)
    ms_e2e = ms_e2e_total / steps
    last_loss = step_e2e()
    pk = peaks()
    fl = train_flops_per_tile(T, cin, K)
    out = {
        "metric": "unet_train_tiles_per_s", "value": world * B / (ms_step / 1e3), "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": a.precision if a.precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"cfg4: Unet-resnet34 training step (train-mode BN fwd + Dice/Focal + bwd + clip 0.5 + Adam), "
                               f"{B} RGB+NIR {T}x{T} tiles per GPU, data parallel x{world}",
                   "tile": T, "batch_per_gpu": B, "in_channels": cin, "classes": K,
                   "l2_policy": "per-step working set (activations + gradients, > 5 GB) far larger than L2; no explicit flush",
                   "parallelism": f"dp{world}, bucketed NCCL gradient all-reduce overlapped with backward",
                   "launch": "whole step replayed from one CUDA graph" if use_graph else "kernels launched one by one"},
        "mpixel_per_s": world * B * T * T / 1e6 / (ms_step / 1e3),
        "e2e": {"value": world * B / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(img_h.numel() * 4 + mask_h.numel() * 8),
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e},
        "gpu_launches": int(launches), "clocks": clocks, "final_loss": last_loss,
        "step_tflops": B * fl / (ms_step / 1e3) / 1e12, "flops_per_tile": fl,
    }
    if prof:
        def agg(evs):
            return sum(e[0].elapsed_time(e[1]) for e in evs) / 1e3, sum(e[2] for e in evs), len(evs)
        conv = prof.get("conv", [])
        tf, wf, nf = agg([e for e in conv if e[3].startswith("train.")])
        td, wd, nd = agg([e for e in conv if e[3].startswith("dgrad.")])
        tw, ww, nw = agg(prof.get("wgrad", []))
        t_all, w_all = tf + td + tw, wf + wd + ww
        if t_all > 0:
            ach = w_all / t_all / 1e12
            out["roofline"] = {"bound": "tensor", "kernel": "tcgen05 conv forward + dgrad (forward kernels on transposed weights) + MN-major wgrad",
                               "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"],
                               "traffic": None, "launches": nf + nd + nw, "share_of_step": (t_all * 1e3 / psteps) / ms_step,
                               "peak_source": pk["source"],
                               "breakdown": {"forward": {"ms_per_step": 1e3 * tf / psteps, "tflops": wf / max(tf, 1e-12) / 1e12, "launches": nf},
                                             "dgrad_tc": {"ms_per_step": 1e3 * td / psteps, "tflops": wd / max(td, 1e-12) / 1e12, "launches": nd},
                                             "wgrad_tc": {"ms_per_step": 1e3 * tw / psteps, "tflops": ww / max(tw, 1e-12) / 1e12, "launches": nw}}}
        if a.layer_table and rank == 0:
            per = {}
            for e0, e1, wk, tag in conv + prof.get("wgrad", []):
                d = per.setdefault(tag, [0.0, 0.0, 0])
                d[0] += e0.elapsed_time(e1); d[1] += wk; d[2] += 1
            # appended to the default invocation (cfg4 behind cfg2) the training table gets its own file
            path = a.layer_table if a.workload == "train" else a.layer_table + ".train"
            with open(path, "w") as fh:
                fh.write(f"{'op.layer':58s} {'launches':>8s} {'avg_us':>9s} {'TFLOP/s':>9s} {'share%':>7s}\n")
                for tag, (ms, wk, n) in per.items():
                    fh.write(f"{tag:58s} {n:8d} {1e3 * ms / n:9.1f} {wk / (ms / 1e3) / 1e12:9.1f} {100 * (ms / psteps) / ms_step:7.2f}\n")
                fh.write(f"tensor-core kernels {1e3 * t_all / psteps:.2f} ms of {ms_step:.2f} ms per step\n")
    if rank == 0 and world == 1 and with_cpu:
        from oracle import ref_train, ref_unet
        torch.set_num_threads(os.cpu_count() or 1)
        model = ref_unet.build_reference_unet(cin, K, seed=0)
        ci, cm = synthetic_train_batch(8, T, cin, K, 1234)
        copt = torch.optim.Adam(model.parameters(), lr=3e-4)
        ts = []
        for i in range(3):
            t0 = time.perf_counter()
            ref_train.train_step(model, ci, cm, lr=3e-4, clip=0.5, optimizer=copt)
            ts.append(time.perf_counter() - t0)
        out["cpu_baseline"] = {"value": 8 / min(ts[1:]), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"one oracle training step on 8 RGB+NIR tiles of {T}x{T} (torch CPU autograd fp32 + reference "
                                         f"loss terms + clip + Adam), best of 2 after 1 warm-up"}
    return out, (gstep if use_graph else None)


def finish(world: int, gstep) -> None:
    """process teardown.  A CUDA graph that holds NCCL nodes must be gone before the communicator is torn down
    (destroy_process_group waited forever with the graph alive): leave the process without the teardown once every rank
    is done."""
    import torch.distributed as dist
    if world <= 1:
        return
    torch.cuda.synchronize()
    dist.barrier()
    if gstep is not None:
        gstep.graph.reset()
        del gstep
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    dist.destroy_process_group()


def main_train(a):
    ctx = dist_context()
    out, gstep = train_measure(a, ctx, a.steps, a.warmup, not a.no_profile, not a.no_cpu_baseline)
    if ctx[0] == 0:
        print(json.dumps(out), flush=True)
    finish(ctx[1], gstep)


def cfg5_measure(a, ctx, engine, steps: int, warmup: int):
    """cfg5: large-tile inference, 32 RGB tiles of 1024 x 1024 per GPU (a 4096 x 8192 uint8 block, overlap 0) through the
    same pipeline (gather+normalise -> Unet -> head+argmax -> mask stitch); every GPU of the box runs its own block
    (weak scaling): whole-box tiles/s and Mpixel/s, convs against the tensor roofline."""
    import torch.distributed as dist
    from deadtrees_b200 import ops
    from deadtrees_b200.deployment.inference import MosaicInference
    from deadtrees_b200.engine import conv_flops_per_tile
    rank, world, local, dev = ctx
    T, gy, gx = 1024, 4, 8
    n = gy * gx
    bt5 = int(os.environ.get("DT_CFG5_BATCH", "32"))      # one batch per block (16: 8.6-8.7 ms, 32: 8.3-8.4 ms per 32 tiles)
    mi = MosaicInference(engine, tile=T, overlap=0, batch_tiles=bt5)
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    block = torch.randint(0, 256, (gy * T, gx * T, 3), dtype=torch.uint8, device=dev, generator=g)
    host_block = torch.empty(block.shape, dtype=torch.uint8, pin_memory=True)
    host_block.copy_(block)
    host_mask = torch.empty((gy * T, gx * T), dtype=torch.uint8, pin_memory=True)
    mask = torch.zeros((gy * T, gx * T), dtype=torch.uint8, device=dev)

    def timed(fn, k, profile=False, after=None):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ops.PROFILE = {} if profile else None
        l0 = ops.LAUNCHES
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        if after is not None:
            after()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        prof, ops.PROFILE = ops.PROFILE, None
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ops.LAUNCHES - l0, prof

    dev_step = lambda: mi.run(block, "hwc", out=mask)
    pipelined = not a.e2e_no_overlap       # successive blocks overlap their copies with the neighbours' compute (as in cfg2)
    e2e_step = lambda: mi.run(block, "hwc", out=mask, host_src=host_block, host_out=host_mask, pipelined=pipelined)
    for _ in range(max(warmup, 3)):
        dev_step()
    sampler = ClockSampler(local)
    sampler.start()
    ms_total, launches, _ = timed(dev_step, steps)
    clocks = sampler.stop()
    ms = ms_total / steps
    ms_prof, _, prof = timed(dev_step, min(steps, 5), profile=True)
    for _ in range(2):
        e2e_step()
    mi.finish()
    ms_e2e = timed(e2e_step, steps, after=mi.finish)[0] / steps
    pk = peaks()
    evs = prof.get("conv", [])
    tc = sum(ev[0].elapsed_time(ev[1]) for ev in evs) / 1e3
    wc = sum(ev[2] for ev in evs)
    out = {"metric": "unet_large_tile_inference_tiles_per_s", "value": world * n / (ms / 1e3), "unit": "tiles/s",
           "mpixel_per_s": world * n * T * T / 1e6 / (ms / 1e3), "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
           "ms_per_step": ms, "scaling": "weak", "dtype": "bf16", "gpu_launches": int(launches), "clocks": clocks,
           "config": {"workload": "cfg5: Unet-resnet34 inference on 32 RGB uint8 tiles of 1024x1024 per GPU (4096x8192 block, "
                                  f"overlap 0), batches of {bt5} tiles", "tile": T, "tiles_per_gpu": n, "batch_tiles": bt5,
                      "l2_policy": "100 MB block and > 8 GB of activations per step; no explicit flush"},
           "e2e": {"value": world * n / (ms_e2e / 1e3), "unit": "tiles/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": int(block.numel()), "d2h_bytes_per_step": int(mask.numel()),
                   "pipeline": "successive blocks overlap; finish() inside the timed region" if pipelined else "closed jobs"}}
    if tc > 0:
        out["roofline"] = {"bound": "tensor", "kernel": "all conv launches (tcgen05 implicit GEMM)", "achieved": wc / tc / 1e12,
                           "peak": pk["tflops"], "unit": "TFLOP/s", "frac": wc / tc / 1e12 / pk["tflops"], "traffic": None,
                           "launches": len(evs), "share_of_step": tc * 1e3 / ms_prof, "flops_per_tile": conv_flops_per_tile(T, 3, 3),
                           "peak_source": pk["source"]}
    del mi, block, mask
    engine._ws.clear()
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    args = parse()
    if args.workload == "train":
        (main_train_reference if args.impl == "reference" else main_train)(args)
    elif args.impl == "reference":
        main_reference(args)
    else:
        main_b200(args)
