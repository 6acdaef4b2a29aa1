"""Inference back-ends — drop-in for ``deadtrees.deployment.inference`` plus the whole-mosaic pipeline.

Mirrors ``deadtrees/deployment/inference.py``: ``Inference`` (:14-27), ``PyTorchInference`` (:30-62),
``PyTorchEnsembleInference`` (:65-116).  ``MosaicInference`` is the B200 form of the hot loop in
``scripts/inference.py:85-111`` (tile -> normalise -> forward -> argmax -> stitch) with every stage on the
device: uint8 mosaic in, uint8 class-id mosaic out.
"""
from __future__ import annotations

import math
from abc import ABC, abstractmethod
from pathlib import Path
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .. import ops
from .._lib import require_device
from ..data.deadtreedata import normalize_constants
from ..engine import UnetEngine
from ..network.segmodel import SemSegment


class Inference(ABC):
    def __init__(self, model_file: Union[str, Path]) -> None:
        self._model_file = model_file if isinstance(model_file, Path) else Path(model_file)
        super().__init__()

    @property
    def model_file(self) -> str:
        return self._model_file.name

    @abstractmethod
    def run(self, input_tensor: torch.Tensor):
        pass


class PyTorchInference(Inference):
    def __init__(self, model_file) -> None:
        super().__init__(model_file)
        if self._model_file.suffix != ".ckpt":
            raise ValueError(f"ckpt file expected, but {self._model_file.suffix} received")
        model = SemSegment.load_from_checkpoint(self._model_file)
        model.eval()
        self._channels = list(model.parameters())[0].shape[1]
        self._model = model.model

    def run(self, input_tensor, device: str = "cuda"):
        """(N, C, H, W) or (C, H, W) float tensor -> int64 class ids, squeezed (inference.py:47-62).

        The reference defaults ``device="cpu"``; this build has no CPU path, so a CPU device raises."""
        if not isinstance(input_tensor, torch.Tensor):
            raise TypeError("no pytorch tensor provided")
        require_device()
        if torch.device(device).type != "cuda":
            raise RuntimeError("deadtrees_b200 runs on CUDA (B200) devices only: pass device='cuda'")
        self._model.to(device)
        if input_tensor.dim() == 3:
            input_tensor.unsqueeze_(0)
        with torch.no_grad():
            # rgb model but rgbn data: the input packer keeps the first `_channels` channels
            out = self._model(input_tensor.to(device))
        return ops.argmax_nchw(out).long().squeeze()


class PyTorchEnsembleInference:
    def __init__(self, *model_files: Path):
        self._models = []
        self._channels = None
        if len(model_files) % 2 == 0:
            raise ValueError("PyTorchEnsembleInference requires an uneven number of models")
        for model_file in model_files:
            model_file = Path(model_file)
            if model_file.suffix != ".ckpt":
                raise ValueError(f"Ckpt file expected, but {model_file.suffix} received")
            model = SemSegment.load_from_checkpoint(model_file)
            model.eval()
            channels = list(model.parameters())[0].shape[1]
            if not self._channels:
                self._channels = channels
            if channels != self._channels:
                raise ValueError("Models are not compatible since they were trained for different channel configs")
            self._models.append(model.model)

    def run(self, input_tensor, device: str = "cuda"):
        if not isinstance(input_tensor, torch.Tensor):
            raise TypeError("No PyTorch tensor provided")
        require_device()
        if input_tensor.dim() == 3:
            input_tensor.unsqueeze_(0)
        if len(self._models) > 15:
            raise ValueError("the majority-vote kernel takes at most 15 models")
        x = input_tensor.to(device)
        masks = None
        for i, model in enumerate(self._models):
            model.to(device)
            with torch.no_grad():
                m = ops.argmax_nchw(model(x))             # uint8 (N, H, W)
            if masks is None:
                masks = torch.empty((len(self._models),) + tuple(m.shape), dtype=torch.uint8, device=m.device)
            masks[i].copy_(m)
        # torch.mode(torch.stack(outs, dim=1), axis=1)[0] of the reference, one fused pass: int64 (N, H, W) / (H, W)
        return ops.mode_vote(masks).squeeze()


# --------------------------------------------------------------------------------------------------

def overlap_grid(H: int, W: int, T: int, overlap: int) -> Tuple[int, int]:
    """number of tiles (gy, gx) at stride T - overlap covering H x W (zero padding beyond the edge)."""
    if not 0 <= overlap < T:
        raise ValueError("overlap must be in [0, T)")
    s = T - overlap
    return max(0, math.ceil((H - T) / s)) + 1, max(0, math.ceil((W - T) / s)) + 1


def blend_window(T: int, overlap: int, device) -> torch.Tensor:
    """1-D blending weights: linear ramp across the overlap, 1 inside (all ones when overlap == 0)."""
    i = np.arange(T, dtype=np.float32)
    w = np.minimum(np.minimum(i + 1, T - i), np.float32(overlap + 1)) / np.float32(overlap + 1)
    return torch.from_numpy(w.astype(np.float32)).to(device)


def batch_plan(first_tile: int, n_tiles: int, batch_tiles, lead: int = 0):
    """[(t0, n), ...] covering ``n_tiles`` tiles from ``first_tile``.  ``batch_tiles``: an upper bound (the batches are then
    EQUAL - a short last batch leaves the deep layers' grids mostly empty) or an explicit sequence of sizes (the last one
    repeats).  ``lead`` > 0 puts a short batch of that many tiles first: with host buffers its upload is the only one
    nothing hides."""
    if isinstance(batch_tiles, (list, tuple)):
        sizes = [int(v) for v in batch_tiles]
        if not sizes or min(sizes) < 1:
            raise ValueError("batch sizes must be positive")
    else:
        if int(batch_tiles) < 1:
            raise ValueError("batch_tiles must be positive")
        sizes = []
        rest = n_tiles
        if 0 < lead < n_tiles:
            sizes.append(lead)
            rest -= lead
        nb = max(1, -(-rest // int(batch_tiles)))
        sizes.append(max(1, -(-rest // nb)))
    plan, t, end = [], first_tile, first_tile + n_tiles
    while t < end:
        n = min(sizes[min(len(plan), len(sizes) - 1)], end - t)
        plan.append((t, n))
        t += n
    return plan


class MosaicInference:
    """Sliding-window segmentation of a whole uint8 mosaic on one GPU (or one tile-row shard of it)."""

    def __init__(self, model, tile: int = 256, overlap: int = 0, batch_tiles: int = 128,
                 precision: Optional[str] = None, mean=None, std=None):
        require_device()
        if isinstance(model, SemSegment):
            model = model.model
        if isinstance(model, UnetEngine):
            self.engine = model
        else:
            if precision is not None:
                model.set_precision(precision)
            self.engine = model.eval().engine()
        if tile % 32:
            raise ValueError("tile must be a multiple of 32 (five stride-2 stages)")
        self.tile, self.overlap, self.batch_tiles = tile, overlap, batch_tiles
        self.offset, self.scale = normalize_constants(self.engine.in_channels, mean, std)
        self.win = blend_window(tile, overlap, self.engine.device)
        self._bufs = {}

    def _buf(self, key, shape, dtype):
        t = self._bufs.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.engine.device)
            self._bufs[key] = t
        return t

    def run(self, mosaic: torch.Tensor, layout: str = "hwc", tile_rows: Optional[Tuple[int, int]] = None,
            out: Optional[torch.Tensor] = None, halo_hook=None, host_src: Optional[torch.Tensor] = None,
            host_out: Optional[torch.Tensor] = None, banded: Optional[bool] = None,
            pipelined: bool = False) -> torch.Tensor:
        """mosaic: CUDA uint8 (H, W, C) ["hwc"] or (C, H, W) ["chw"] -> uint8 class ids (H, W).

        ``tile_rows=(r0, r1)`` restricts the work to that range of tile rows (multi-GPU sharding); the
        returned mask is then only valid on the mosaic rows this shard owns, ``owned_rows(...)``.

        Host pipeline ("hwc" only): with ``host_src`` (pinned uint8 (H, W, C)) the rows a batch of tiles needs are copied
        into ``mosaic`` on a copy stream while the previous batches compute; with ``host_out`` (pinned uint8 (H, W)) and
        no ``halo_hook`` the mask is stitched in bands as soon as their tile rows are done and each band goes back to the
        host behind the compute.  Only the first band's upload and the last band's download are exposed.

        ``pipelined`` (host pipeline only): successive calls overlap like a loop over many mosaics does - ``mosaic`` and
        ``out`` are then staging buffers that only this object touches; the next call's upload starts as soon as this call's
        last gather has read ``mosaic``, and this call's last mask band goes to the host while the next call computes.
        ``host_out`` is complete after :meth:`finish` (or the next ``finish``), not after this call.

        ``banded``: stitch the blended mask band by band behind the batches (True) or in one launch over the shard
        (False); default: bands while a batch's logits fit in L2.  A shard that starts below the first tile row blends
        its first ``overlap`` rows with the previous shard's logits, which only ``halo_hook`` can provide."""
        H, W = (mosaic.shape[0], mosaic.shape[1]) if layout == "hwc" else (mosaic.shape[1], mosaic.shape[2])
        T, ov, eng = self.tile, self.overlap, self.engine
        gy, gx = overlap_grid(H, W, T, ov)
        r0, r1 = (0, gy) if tile_rows is None else tile_rows
        mask = out if out is not None else self._buf("mask", (H, W), torch.uint8)
        ntiles = (r1 - r0) * gx
        # equal batches; with a host source a first batch of one tile row (the only upload nothing hides)
        # (a pipelined call whose predecessor left its events behind has its first rows uploaded behind that call)
        prefetched = pipelined and host_src is not None and getattr(self, "_gather_done", None) is not None
        batches = batch_plan(r0 * gx, ntiles, self.batch_tiles, lead=gx if host_src is not None and not prefetched else 0)
        bt = max([n for _, n in batches] + [1])
        pad = 3 if eng.stem_padded(T) else 0
        x = self._bufs.get(("x", bt, T))
        if x is None:
            x = self._bufs[("x", bt, T)] = eng.alloc_input(bt, T)   # padded frame: borders stay zero
        if ov == 0:
            tmask = self._buf("tmask", (bt, T, T), torch.uint8)
        else:
            halo = 1 if r0 > 0 else 0  # room for the neighbour's last tile row (only its bottom rows are read)
            logits = self._buf("logits", ((r1 - r0 + halo) * gx, T, T, eng.classes), eng.act_dtype)
        step = T - ov
        piped_in = host_src is not None
        piped_out = host_out is not None and halo_hook is None
        # single GPU: the mask is stitched band by band as soon as the tile rows of a band are done - the logits a band
        # needs were written by the last batches and are still in L2 (a batch of 135 tiles holds 53 MB of logits)
        # (only while a batch's logits fit in L2; with larger batches one launch over the whole shard is faster)
        if ov > 0 and halo and halo_hook is None:
            raise ValueError("a shard with tile_rows[0] > 0 and overlap > 0 needs the previous shard's boundary logits: "
                             "pass halo_hook (deadtrees_b200.sharding.make_halo_hook)")
        if banded is None or halo_hook is not None or piped_out or ov == 0:
            banded = halo_hook is None and (piped_out or ov == 0 or
                                            bt * T * T * eng.classes * logits.element_size() <= (96 << 20))
        if (piped_in or host_out is not None) and layout != "hwc":
            raise ValueError("the host pipeline takes interleaved (H, W, C) mosaics")
        main = torch.cuda.current_stream()
        if not pipelined and getattr(self, "_out_done", None) is not None:
            # a closed job after pipelined ones without finish(): their last downloads still read `mask`
            main.wait_event(self._out_done)
            self._out_done = self._gather_done = None
        if piped_in or piped_out:
            if getattr(self, "_copy_streams", None) is None:
                self._copy_streams = (torch.cuda.Stream(device=eng.device), torch.cuda.Stream(device=eng.device))
            cs_in, cs_out = self._copy_streams    # uploads never queue behind a download that waits for compute
            if prefetched:
                cs_in.wait_event(self._gather_done)   # the previous call's gathers have read `mosaic`
            else:
                cs_in.wait_stream(main)           # earlier work on `mosaic` / `mask` is done before they are overwritten
            cs_out.wait_stream(main)
        out_done = getattr(self, "_out_done", None) if pipelined else None   # previous call's downloads still read `mask`
        y_own0, y_own1 = self.owned_rows(H, T, ov, gy, r0, r1)
        copied = min(H, r0 * step)                # mosaic rows [r0 * step, copied) are on the device
        stitched = y_own0                         # mask rows [y_own0, stitched) are final
        for t0, n in batches:
            xb = x[:n]
            if piped_in:
                need = min(H, ((t0 + n - 1) // gx) * step + T)
                if need > copied:
                    with torch.cuda.stream(cs_in):
                        mosaic[copied:need].copy_(host_src[copied:need], non_blocking=True)
                    copied = need
                    main.wait_stream(cs_in)
            ops.tile_gather_normalize(mosaic, layout, eng.in_channels, T, ov, (gy, gx), t0, n, self.offset,
                                      self.scale, out=xb, pad=pad)
            if pipelined and piped_in and t0 + n >= r1 * gx:
                self._gather_done = torch.cuda.Event()
                self._gather_done.record(main)
            if ov == 0:
                eng.forward(xb, mask_out=tmask[:n])
                if out_done is not None:
                    main.wait_event(out_done)
                    out_done = None
                ops.stitch_mask(tmask[:n], gx, t0, mask)
            else:
                lo = t0 - (r0 - halo) * gx
                eng.forward(xb, logits_nhwc_out=logits[lo: lo + n])
            if banded:
                rows_done = (t0 + n) // gx           # complete tile rows so far: mask rows below rows_done * step are final
                y_end = y_own1 if t0 + n >= r1 * gx else min(y_own1, rows_done * step)
                if y_end > stitched:
                    if ov > 0:
                        if out_done is not None:
                            main.wait_event(out_done)
                            out_done = None
                        ops.stitch_blend_argmax(logits, ov, (gy, gx), self.win, mask, row0=stitched, nrows=y_end - stitched,
                                                ty_base=r0 - halo)
                    if piped_out:
                        cs_out.wait_stream(main)
                        with torch.cuda.stream(cs_out):
                            host_out[stitched:y_end].copy_(mask[stitched:y_end], non_blocking=True)
                    stitched = y_end
        if ov > 0 and not banded:
            if halo_hook is not None:
                halo_hook(logits, gx, halo)  # multi-GPU: exchange boundary logits rows with the neighbours
            if out_done is not None:
                main.wait_event(out_done)
            ops.stitch_blend_argmax(logits, ov, (gy, gx), self.win, mask, row0=y_own0, nrows=y_own1 - y_own0, ty_base=r0 - halo)
        if host_out is not None and not piped_out:
            host_out[y_own0:y_own1].copy_(mask[y_own0:y_own1], non_blocking=True)
        if piped_in or piped_out:
            main.wait_stream(cs_in)               # (every upload was already waited for by its gather)
            if pipelined and piped_out:
                self._out_done = torch.cuda.Event()
                self._out_done.record(cs_out)     # the next call's first stitch waits for this; finish() joins the stream
            else:
                main.wait_stream(cs_out)          # callers synchronise the current stream only
        return mask

    def finish(self) -> None:
        """join the copy streams of ``pipelined`` calls into the current stream: after this (and a synchronise of the current
        stream) every ``host_out`` handed to an earlier call is complete, and ``mosaic`` / ``out`` may be touched again."""
        streams = getattr(self, "_copy_streams", None)
        if streams is not None:
            main = torch.cuda.current_stream()
            main.wait_stream(streams[0])
            main.wait_stream(streams[1])
        self._gather_done = None
        self._out_done = None

    def run_shard(self, mosaic: torch.Tensor, plan, out: torch.Tensor, exchange=None, host_src: Optional[torch.Tensor] = None,
                  host_out: Optional[torch.Tensor] = None, batch_tiles=None, pipelined: bool = False) -> torch.Tensor:
        """one rank of a multi-GPU run over tile-RANGE shards (``deadtrees_b200.sharding.ShardPlan``): the tiles
        ``[plan.t0, plan.t1)`` go through the network in batches, ``exchange(logits)`` swaps the few boundary tiles /
        strips with the two neighbours (``sharding.exchange_logits``), and the mask rows ``plan.mask_rows(H, T)`` are
        stitched into ``out``.  "hwc" mosaics; ``host_src`` / ``host_out`` as in :meth:`run` (row bands of the mosaic are
        uploaded behind the batches; the shard's mask rows go back after the stitch).  ``batch_tiles``: one size, or a
        sequence of batch sizes (the last one repeats) - with host buffers a short first batch shortens the upload nothing
        can hide, and large later batches keep the deep layers' grids full.  ``pipelined`` (with ``host_src``): as in
        :meth:`run`, the next call's upload starts once this call's last gather has read ``mosaic`` (a staging buffer then)."""
        H, W = mosaic.shape[0], mosaic.shape[1]
        T, ov, eng = self.tile, self.overlap, self.engine
        gy, gx = overlap_grid(H, W, T, ov)
        if (gy, gx, ov) != (plan.gy, plan.gx, plan.overlap):
            raise ValueError("shard plan does not belong to this mosaic / tiling")
        step = T - ov
        prefetched = pipelined and host_src is not None and getattr(self, "_gather_done", None) is not None
        starts = batch_plan(plan.t0, plan.t1 - plan.t0, batch_tiles or self.batch_tiles,
                            lead=gx if host_src is not None and not prefetched and not isinstance(batch_tiles, (list, tuple)) else 0)
        bt = max([n for _, n in starts] + [1])
        pad = 3 if eng.stem_padded(T) else 0
        x = self._bufs.get(("x", bt, T))
        if x is None:
            x = self._bufs[("x", bt, T)] = eng.alloc_input(bt, T)
        y0, y1 = plan.mask_rows(H, T)
        main = torch.cuda.current_stream()
        if host_src is not None:
            if getattr(self, "_copy_streams", None) is None:
                self._copy_streams = (torch.cuda.Stream(device=eng.device), torch.cuda.Stream(device=eng.device))
            cs_in = self._copy_streams[0]
            if prefetched:
                cs_in.wait_event(self._gather_done)
            else:
                cs_in.wait_stream(main)
        if ov == 0:
            tmask = self._buf("tmask", (bt, T, T), torch.uint8)
        else:
            logits = self._buf("logits", (plan.B1 - plan.B0, T, T, eng.classes), eng.act_dtype)
        copied = plan.input_rows(H, T)[0]
        # with a host mask the interior rows are stitched and sent back as soon as their tile rows are complete; the first
        # stitched tile row waits for the previous rank's boundary strips, the last one for the next rank's head tiles
        early = host_out is not None and ov > 0
        if early:
            cs_out = self._copy_streams[1] if host_src is not None else None
            if cs_out is None:
                if getattr(self, "_copy_streams", None) is None:
                    self._copy_streams = (torch.cuda.Stream(device=eng.device), torch.cuda.Stream(device=eng.device))
                cs_out = self._copy_streams[1]
            cs_out.wait_stream(main)
        first_end = min(y1, (plan.R0 + 1) * step) if plan.ty_base < plan.R0 else y0       # rows that need the halo strips
        stitched = first_end
        last_row = plan.R1 - 1 if plan.recv_tail is not None else plan.R1                 # tile rows complete without the tail
        for t0, n in starts:
            if host_src is not None:
                need = min(H, ((t0 + n - 1) // gx) * step + T)
                if need > copied:
                    with torch.cuda.stream(cs_in):
                        mosaic[copied:need].copy_(host_src[copied:need], non_blocking=True)
                    copied = need
                    main.wait_stream(cs_in)
            ops.tile_gather_normalize(mosaic, "hwc", eng.in_channels, T, ov, (gy, gx), t0, n, self.offset, self.scale,
                                      out=x[:n], pad=pad)
            if pipelined and host_src is not None and t0 + n >= plan.t1:
                self._gather_done = torch.cuda.Event()
                self._gather_done.record(main)
            if ov == 0:
                eng.forward(x[:n], mask_out=tmask[:n])
                ops.stitch_mask(tmask[:n], gx, t0, out)       # overlap 0: a tile's pixels belong to whoever computed it
            else:
                eng.forward(x[:n], logits_nhwc_out=logits[t0 - plan.B0: t0 - plan.B0 + n])
            if early:
                rows_done = min((t0 + n) // gx, last_row)     # mask rows below rows_done * step are final
                y_end = min(y1, rows_done * step)
                if y_end > stitched:
                    ops.stitch_blend_argmax(logits, ov, (gy, gx), self.win, out, row0=stitched, nrows=y_end - stitched,
                                            ty_base=plan.ty_base)
                    cs_out.wait_stream(main)
                    with torch.cuda.stream(cs_out):
                        host_out[stitched:y_end].copy_(out[stitched:y_end], non_blocking=True)
                    stitched = y_end
        if ov > 0:
            if exchange is not None:
                exchange(logits)
            if not early:
                if y1 > y0:
                    ops.stitch_blend_argmax(logits, ov, (gy, gx), self.win, out, row0=y0, nrows=y1 - y0, ty_base=plan.ty_base)
            else:
                for a, b in ((y0, first_end), (stitched, y1)):
                    if b > a:
                        ops.stitch_blend_argmax(logits, ov, (gy, gx), self.win, out, row0=a, nrows=b - a, ty_base=plan.ty_base)
                        host_out[a:b].copy_(out[a:b], non_blocking=True)
                main.wait_stream(cs_out)
        if host_out is not None and y1 > y0 and not early:
            host_out[y0:y1].copy_(out[y0:y1], non_blocking=True)
        return out

    @staticmethod
    def owned_rows(H: int, T: int, overlap: int, gy: int, r0: int, r1: int) -> Tuple[int, int]:
        """mosaic rows whose output a shard with tile rows [r0, r1) writes."""
        s = T - overlap
        return min(H, r0 * s), (H if r1 >= gy else min(H, r1 * s))

    def run_host(self, mosaic: np.ndarray, layout: str = "hwc") -> np.ndarray:
        """host ndarray in, host ndarray out (pinned staging; the end-to-end path of ``scripts/inference.py``)."""
        src = torch.from_numpy(np.ascontiguousarray(mosaic))
        src = src if src.is_pinned() else src.pin_memory()
        dev = self._buf("mosaic_dev", tuple(src.shape), torch.uint8)
        if layout != "hwc":
            dev.copy_(src, non_blocking=True)
            return self.run(dev, layout).cpu().numpy()
        H, W = src.shape[0], src.shape[1]
        host_mask = torch.empty((H, W), dtype=torch.uint8, pin_memory=True)
        self.run(dev, layout, host_src=src, host_out=host_mask)
        torch.cuda.current_stream().synchronize()
        return host_mask.numpy()
