"""One whole training step (forward, loss, backward, clip + Adam) captured in a CUDA graph.

At the reference's batch sizes the step is ~400 short kernels; launched one by one from Python the host, not the GPU,
sets the step time.  ``GraphedTrainStep`` records the kernel sequence of ``SemSegment.training_step`` +
``loss.backward()`` + ``optimizer.step()`` (``deadtrees/network/segmodel.py:210-229``, ``:420-429``) once and replays it;
everything that varies from step to step lives on the device (inputs in static buffers, Adam's step count and learning
rate, the clip factor, the NaN-skip decision), so no host synchronisation remains inside a step.  The label-range check
of ``class2one_hot`` (``losses.py:129``) and the NaN warning are evaluated from device flags AFTER the replay
(``check()``), instead of stalling the pipeline in the middle of the step as the reference does.
"""
from __future__ import annotations

from typing import Dict

import torch

from .loss.fused import SegLossTerms
from . import ops
from .optim import FusedAdam


class GraphedTrainStep:
    def __init__(self, seg, optimizer: FusedAdam, batch: int, tile: int):
        if seg.boundary_loss is not None:
            raise NotImplementedError("the graphed step takes (img, mask) batches: run boundary-loss training through training_step()")
        self.seg, self.opt = seg, optimizer
        self.engine = seg.model.train_engine()
        if optimizer._flat is None:
            optimizer.attach_engine(self.engine)
        # data parallel: the bucketed NCCL all-reduces run on the reducer's side stream, which forks from and joins the
        # capturing stream through events, so they become nodes of the same graph (every rank captures the same sequence)
        dev = self.engine.device
        self.dice_mode = seg._dice_mode()
        self.gwdl_matrix = seg.dice_loss.matrix() if self.dice_mode == 0 and seg.dice_loss is not None else None
        self.use_focal = seg.focal_loss is not None
        cin = seg.model.in_channels
        self.img = torch.zeros((batch, cin, tile, tile), dtype=torch.float32, device=dev)
        self.mask = torch.zeros((batch, tile, tile), dtype=torch.int64, device=dev)
        self.terms: SegLossTerms = None
        # warm-up on a side stream (lazy allocations, cudaFuncSetAttribute calls, tensor-map encodes), then capture;
        # the warm-up steps must not train: snapshot and restore every piece of state they touch
        snap = self._snapshot()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()
        self._restore(snap)

    def _state_tensors(self) -> Dict[str, torch.Tensor]:
        f = self.opt._flat
        st = {"p": f["p"], "m": f["m"], "v": f["v"], "state": f["state"]}
        for n, b in self.seg.model.named_buffers():
            st["buf." + n] = b
        return st

    def _snapshot(self):
        return {k: v.clone() for k, v in self._state_tensors().items()}

    def _restore(self, snap) -> None:
        with torch.no_grad():
            for k, v in self._state_tensors().items():
                v.copy_(snap[k])

    def _body(self) -> None:
        with torch.no_grad():
            logits, tape = self.engine.forward(self.img)
            self.terms = SegLossTerms(logits, self.mask, self.dice_mode, self.use_focal)
            g_logits = self.terms.grad_logits(1.0)
            self.total = self.terms.out[2:3]
            if self.gwdl_matrix is not None:           # GWDICE: Wasserstein Dice term next to the focal term
                self.gwdl, coef = ops.gwdl_loss(self.terms.logits, self.terms.labels, self.gwdl_matrix, softmax_twice=True)
                ops.gwdl_loss_backward(self.terms.logits, self.terms.labels, self.gwdl_matrix, coef, 1.0, g_logits, softmax_twice=True)
                self.total = self.total + self.gwdl
            grads = self.engine.backward(tape, g_logits)
            self.opt.step(loss=self.total)                  # skipped on the device when the loss is not finite
        for n, p in self.engine.params.items():
            p.grad = grads[n]

    def replay(self) -> None:
        """one step on whatever the static buffers hold: pushes a learning rate changed by a scheduler to the device
        (outside the graph, same stream), replays, and invalidates the caches of folded weights (``Unet.engine()``) -
        a replay changes parameters and running statistics without touching any tensor version counter."""
        self.opt.push_lr()
        self.graph.replay()
        ops.bump_state_generation()

    def __call__(self, img: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        """copies the batch into the static buffers, replays the step, returns the (device) total loss."""
        self.img.copy_(img, non_blocking=True)
        self.mask.copy_(mask, non_blocking=True)
        self.replay()
        return self.total[0]

    def prefetch(self, img: torch.Tensor, mask: torch.Tensor) -> None:
        """starts the upload of the NEXT batch (pinned host tensors) into staging buffers on a copy stream; it runs
        behind the step that is being replayed.  ``step_prefetched()`` then consumes it."""
        if getattr(self, "_stage", None) is None:
            dev = self.img.device
            self._stage = (torch.empty_like(self.img), torch.empty_like(self.mask))
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._stage_free = torch.cuda.Event()
            self._stage_free.record()
        cs = self._copy_stream
        cs.wait_event(self._stage_free)                # the previous hand-over out of the staging buffers is done
        with torch.cuda.stream(cs):
            self._stage[0].copy_(img, non_blocking=True)
            self._stage[1].copy_(mask, non_blocking=True)
        self._staged = True

    def step_prefetched(self) -> torch.Tensor:
        """hands the prefetched batch over to the graph's static buffers (device-to-device) and replays the step."""
        if not getattr(self, "_staged", False):
            raise RuntimeError("step_prefetched() needs a prefetch() first")
        main = torch.cuda.current_stream()
        main.wait_stream(self._copy_stream)
        self.img.copy_(self._stage[0], non_blocking=True)
        self.mask.copy_(self._stage[1], non_blocking=True)
        self._stage_free.record(main)
        self._staged = False
        self.replay()
        return self.total[0]

    def check(self) -> float:
        """host-side checks of the last replayed step (one synchronisation): label range, finite loss."""
        self.terms.check_labels()
        loss = float(self.total[0])
        if loss != loss or loss in (float("inf"), float("-inf")):
            import logging
            logging.getLogger(__name__).warning("Train loss is NaN! What is going on? (optimizer step skipped)")
        return loss

    def log_terms(self) -> Dict[str, torch.Tensor]:
        t = self.terms
        return {"train/dice_loss": self.gwdl if self.gwdl_matrix is not None else t.dice_loss,
                "train/focal_loss": t.focal_loss, "train/total_loss": self.total[0],
                "train/dice": t.fscore, "train/dice_with_bg": t.fscore_with_bg}
