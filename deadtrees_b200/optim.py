"""Fused global-norm clip + Adam step (``configure_optimizers`` at ``deadtrees/network/segmodel.py:420-429``
with ``gradient_clip_val: 0.5`` from ``configs/trainer/default.yaml:18``)."""
from __future__ import annotations

import torch

from . import ops


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam semantics (no weight decay, no amsgrad) with the global-norm clip folded into the update.

    Two paths: per-parameter launches (any parameter list), or - after :meth:`attach_engine` - ONE ``dt_sumsq`` and ONE
    ``dt_adam_step`` over the train engine's flat parameter / gradient buffers."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, max_grad_norm: float = 0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, max_grad_norm=max_grad_norm))
        self._flat = None

    def attach_engine(self, engine) -> "FusedAdam":
        """use the flat buffers of a ``UnetTrainEngine`` (all of its parameters must be in this optimizer)."""
        mine = {id(p) for g in self.param_groups for p in g["params"]}
        if len(self.param_groups) != 1 or any(id(p) not in mine for p in engine.params.values()):
            raise ValueError("attach_engine needs one param group holding every parameter of the engine")
        fp = engine.flatten_parameters()
        dev = fp.device
        self._engine = engine
        self._flat = dict(p=fp, g=engine.reducer.flat, m=torch.zeros_like(fp), v=torch.zeros_like(fp),
                          state=torch.tensor([0.0, self.param_groups[0]["lr"]], dtype=torch.float32, device=dev),
                          scratch=torch.zeros(4, dtype=torch.float32, device=dev),
                          acc=torch.zeros(1, dtype=torch.float64, device=dev), lr=self.param_groups[0]["lr"])
        return self

    @torch.no_grad()
    def push_lr(self) -> None:
        """flat path: copies a learning rate changed by a scheduler (``CosineAnnealingLR`` of ``configure_optimizers``,
        segmodel.py:420-429) into the device-side step state.  ``GraphedTrainStep`` calls this before every replay - the
        captured graph reads the rate from the device, so the copy stays outside the graph."""
        if self._flat is None:
            return
        f, lr = self._flat, self.param_groups[0]["lr"]
        if lr != f["lr"]:
            f["lr"] = lr
            f["state"][1:2].fill_(lr)

    @torch.no_grad()
    def step(self, closure=None, loss: torch.Tensor = None):
        """``loss`` (flat path only): a device scalar; the update is skipped on the device when it is not finite."""
        if self._flat is not None:
            # every piece of step state lives on the device (step count, lr, clip factor): the launches below can be
            # captured in a CUDA graph and replayed; a changed learning rate is pushed with one tiny copy
            f, group = self._flat, self.param_groups[0]
            eng = self._engine
            if (eng.reducer.flat.data_ptr() != f["g"].data_ptr() or eng.flat_params is None
                    or eng.flat_params.data_ptr() != f["p"].data_ptr() or getattr(eng.model, "_train_engine", eng) is not eng):
                raise RuntimeError("the train engine this optimizer was attached to has been rebuilt (set_precision / device "
                                   "change) or its buffers replaced: call configure_optimizers() / attach_engine() again")
            self.push_lr()
            if loss is not None and group["max_grad_norm"] <= 0 and eng.reducer.world > 1:
                # data parallel without clipping: the non-finite-loss skip must be the same decision on every rank (with
                # clipping the all-reduced gradient norm already carries a NaN / Inf to all of them)
                import torch.distributed as dist
                loss = loss.detach().clone().reshape(1)
                dist.all_reduce(loss, group=eng.reducer.group)
            acc = None
            if group["max_grad_norm"] > 0:
                acc = f["acc"]
                acc.zero_()
                ops.sumsq(f["g"], acc)
            ops.adam_step_dev(f["p"], f["g"], f["m"], f["v"], f["state"], beta1=group["betas"][0], beta2=group["betas"][1],
                              eps=group["eps"], sumsq_acc=acc, max_norm=group["max_grad_norm"], loss=loss, scratch=f["scratch"])
            return
        for group in self.param_groups:
            params = [p for p in group["params"] if p.grad is not None]
            acc = None
            if group["max_grad_norm"] > 0:
                acc = torch.zeros((1,), dtype=torch.float64, device=params[0].device)
                for p in params:
                    ops.sumsq(p.grad.contiguous(), acc)
            for p in params:
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
                ops.adam_step(p, p.grad.contiguous(), st["exp_avg"], st["exp_avg_sq"], lr=group["lr"],
                              beta1=group["betas"][0], beta2=group["betas"][1], eps=group["eps"], step=st["step"],
                              sumsq_acc=acc, max_norm=group["max_grad_norm"])
