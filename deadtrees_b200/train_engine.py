"""Training-step engine for Unet(resnet34): train-mode forward (batch-statistics BatchNorm) and the backward pass,
sequenced over the CUDA library.

Follows autograd of ``smp.Unet.forward`` as ``SemSegment.training_step`` drives it in the reference
(``deadtrees/network/segmodel.py:210-229``; layer list SURVEY.md Appendix A; formulas SURVEY.md Appendix D).
Precision modes: ``"fp32"`` (CUDA-core check mode, every tensor fp32) and ``"bf16"`` (activations and activation
gradients in bf16, fp32 accumulation; convolutions, data gradients and weight gradients on the tcgen05 kernels where
the layer shape allows, generic CUDA-core kernels otherwise).  Master weights, BatchNorm parameters and all parameter
gradients are fp32.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch

from . import ops
from ._lib import CONV_PAIR, CONV_TRANSPOSED, CONV_X_PAD3, require_device
from .engine import RESNET34_LAYERS, RESNET34_PLANES
from .parallel import GradBucketReducer, backward_param_order

BN_EPS, BN_MOMENTUM = 1e-5, 0.1


class _ConvBN:
    """tape entry of one conv (+ BatchNorm (+ residual) (+ ReLU))."""
    __slots__ = ("conv", "bn", "x", "y", "a", "relu", "stride", "pad", "mean", "invstd", "scale", "shift", "residual", "frame")

    def __init__(self, **kw):
        for k in self.__slots__:
            setattr(self, k, kw.get(k))


class UnetTrainEngine:
    def __init__(self, model, precision: str = "bf16", wgrad_tc: bool = True, dgrad_tc: bool = True):
        require_device()
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.model = model
        self.precision = precision
        self.act_dtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self.in_channels, self.classes = model.in_channels, model.classes
        self.wgrad_tc, self.dgrad_tc = wgrad_tc, dgrad_tc
        self.params = dict(model.named_parameters())
        self.buffers = dict(model.named_buffers())
        self.param_names = list(self.params)
        dev = next(model.parameters()).device
        self.device = dev
        self._ones = torch.ones(1024, dtype=torch.float32, device=dev)
        self._zeros = torch.zeros(1024, dtype=torch.float32, device=dev)
        # when a list, every kernel-level operation appends (kind, name, inputs..., outputs...) so tests can check
        # each op of the real pipeline against the oracle on the SAME inputs (layer-wise, teacher-forced parity)
        self.trace: Optional[List] = None
        # every parameter gradient is written straight into one flat fp32 buffer laid out in backward order; with a
        # process group its buckets are all-reduced on a side stream while the backward pass continues
        self.reducer: Optional[GradBucketReducer] = None
        self.set_process_group(None, world_size=1)
        # kernel layouts (bf16 forward / data-gradient packings) of every conv weight: one batched repack per step
        self.packer = ops.WeightPacker(dev)
        self.flat_params = None
        self.head_tc = precision == "bf16" and os.environ.get("DT_TRAIN_HEAD_TC", "1") == "1"
        self._head_b16 = None
        # BatchNorm backward of relu(bn(y)) layers can recompute the ReLU mask from y instead of reading the activation
        # (dt_bn_train_bwd_relu); measured on B200 it is not faster (11.67 vs 11.60 ms per step: these passes are not
        # bound by the bytes of that one tensor), so it stays opt-in
        self.bn_mask_from_y = os.environ.get("DT_BN_MASK_FROM_Y", "0") == "1"
        self.dgrad_s2_zero_insert = os.environ.get("DT_DGRAD_S2_ZERO_INSERT", "1") != "0"
        # wide 3x3/s1 convolutions (forward and the data gradients run as forward convs) on CTA pairs where the shape fits
        self.pair_flag = CONV_PAIR if (precision == "bf16" and os.environ.get("DT_CONV_PAIR", "1") != "0") else 0

    def set_process_group(self, group, world_size: Optional[int] = None, bucket_bytes: int = 25 << 20) -> None:
        """(re)buckets the flat gradient buffer for ``group``.  The buffer itself is kept: the position of a gradient does
        not depend on the bucket size, so an optimizer that already holds ``reducer.flat`` keeps stepping on live data."""
        order = backward_param_order(self.param_names)
        old = self.reducer.flat if self.reducer is not None else None
        self.reducer = GradBucketReducer([(n, self.params[n].shape) for n in order], self.device, bucket_bytes=bucket_bytes,
                                         group=group, world_size=world_size, flat=old)

    def _rec(self, kind: str, name: str, **tensors) -> None:
        if self.trace is not None:
            self.trace.append((kind, name, tensors))

    def flatten_parameters(self) -> torch.Tensor:
        """re-points every parameter into ONE flat fp32 buffer with the layout of the flat gradient buffer, so the
        optimizer is a single fused clip + Adam launch (``FusedAdam.attach_engine``).  Idempotent."""
        if self.flat_params is None or self.flat_params.numel() != self.reducer.flat.numel():
            flat = torch.zeros_like(self.reducer.flat)
            with torch.no_grad():
                for n, p in self.params.items():
                    off = self.reducer.flat_offset(n)
                    v = flat[off: off + p.numel()].view(p.shape)
                    v.copy_(p.data)
                    p.data = v
            self.flat_params = flat
            self.packer.clear()           # the job table of the batched repack holds the OLD addresses of the weights
        return self.flat_params

    # ---- forward pieces --------------------------------------------------------------------------------
    def _conv_raw(self, x: torch.Tensor, wname: str, stride: int, pad: int, frame: Optional[torch.Tensor] = None) -> torch.Tensor:
        """raw convolution output (no BN / activation) of x (N, H, W, C_x) with the master weights `wname`; `frame`: the
        stem's zero-bordered input frame (x is then its interior view)."""
        w = self.params[wname]
        C_out, C_in, R, S = w.shape
        N, H, W, Cx = x.shape
        stem = R == 7
        if frame is not None:
            wp = self.packer.get((wname, 2, None), w, 2)
            return ops.conv2d(frame, wp, self._ones, self._zeros, N=N, H=H, W=W, C_in=Cx, C_x=Cx, C_out=C_out, R=R, S=S,
                              stride=stride, pad=pad, relu=False, algo_cin=C_in, flags=CONV_X_PAD3, tag="train." + wname)
        mode = 0 if self.precision == "fp32" else (2 if stem else 1)
        wp = self.packer.get((wname, mode, None), w, mode)
        return ops.conv2d(x, wp, self._ones, self._zeros, N=N, H=H, W=W, C_in=Cx, C_x=Cx, C_out=C_out, R=R, S=S,
                          stride=stride, pad=pad, relu=False, algo_cin=C_in, flags=self.pair_flag, tag="train." + wname)

    def _conv_bn(self, tape: List, x: torch.Tensor, conv: str, bn: str, stride: int, pad: int, relu: bool = True,
                 residual: Optional[torch.Tensor] = None, frame: Optional[torch.Tensor] = None) -> torch.Tensor:
        y = self._conv_raw(x, conv + ".weight", stride, pad, frame=frame)
        scale, shift, mean, invstd = ops.bn_train_stats(
            y, self.params[bn + ".weight"], self.params[bn + ".bias"], self.buffers.get(bn + ".running_mean"),
            self.buffers.get(bn + ".running_var"), BN_EPS, BN_MOMENTUM)
        nbt = self.buffers.get(bn + ".num_batches_tracked")
        if nbt is not None:
            self._nbt.append(nbt)       # bumped together at the end of the forward pass (one multi-tensor launch)
        a = ops.bn_apply(y, scale, shift, residual=residual, relu=relu)
        self._rec("conv_bn_fwd", conv, x=x, y=y, a=a, residual=residual, mean=mean, invstd=invstd, scale=scale, shift=shift,
                  relu=relu, stride=stride, pad=pad)
        tape.append(_ConvBN(conv=conv, bn=bn, x=x, y=y, a=a, relu=relu, stride=stride, pad=pad, mean=mean,
                            invstd=invstd, scale=scale, shift=shift, residual=residual, frame=frame))
        return a

    def forward(self, x_nchw: torch.Tensor):
        """(N, C, T, T) float -> (logits (N, K, T, T) fp32, tape)."""
        N, _, T, _ = x_nchw.shape
        if T % 32:
            raise ValueError("tile size must be a multiple of 32")
        tape: List = []
        self._nbt: List[torch.Tensor] = []
        self.packer.refresh()       # the master weights changed in the optimizer step: repack every layout, one launch
        if self.precision == "bf16" and self.wgrad_tc and ops.stem_wgrad_tc_supported(N, T, T):
            # zero-bordered stem frame: the TMA im2col map over it feeds the forward conv AND the weight gradient
            frame = ops.pack_input_nchw_frame(x_nchw, self.in_channels)
            x4 = frame[:, 3:3 + T, 3:3 + T, :]
        else:
            frame, x4 = None, ops.pack_input_nchw(x_nchw, self.in_channels, self.act_dtype)
        f = {1: self._conv_bn(tape, x4, "encoder.conv1", "encoder.bn1", 2, 3, frame=frame)}
        pool, pool_idx = ops.maxpool3x3s2_idx(f[1])
        self._rec("maxpool_fwd", "pool", x=f[1], y=pool)
        tape.append(("pool", (f[1], pool_idx)))
        cur = pool
        for li, (planes, nblk) in enumerate(zip(RESNET34_PLANES, RESNET34_LAYERS), start=1):
            for b in range(nblk):
                p = f"encoder.layer{li}.{b}"
                stride = 2 if (b == 0 and li > 1) else 1
                a1 = self._conv_bn(tape, cur, p + ".conv1", p + ".bn1", stride, 1)
                if stride == 2:
                    idn = self._conv_bn(tape, cur, p + ".downsample.0", p + ".downsample.1", 2, 0, relu=False)
                else:
                    idn = cur
                cur = self._conv_bn(tape, a1, p + ".conv2", p + ".bn2", 1, 1, residual=idn)
            f[li + 1] = cur
        xcur = f[5]
        skips = [f[4], f[3], f[2], f[1], None]
        for i in range(5):
            p = f"decoder.blocks.{i}"
            cat = ops.upsample_concat(xcur, skips[i])
            self._rec("upcat_fwd", p, x=xcur, skip=skips[i], y=cat)
            tape.append(("cat", xcur.shape[-1]))
            a1 = self._conv_bn(tape, cat, p + ".conv1.0", p + ".conv1.1", 1, 1)
            xcur = self._conv_bn(tape, a1, p + ".conv2.0", p + ".conv2.1", 1, 1)
        hw = self.params["segmentation_head.0.weight"]
        logits = torch.empty((N, self.classes, T, T), dtype=torch.float32, device=self.device)
        if self.head_tc and xcur.shape[-1] == 16 and T % 16 == 0:
            # bf16 path: the head on the tensor cores like every other layer (weights rounded to bf16, 16 padded rows)
            if self._head_b16 is None:
                self._head_b16 = torch.zeros(16, dtype=torch.float32, device=self.device)
            self._head_b16[: self.classes].copy_(self.params["segmentation_head.0.bias"].detach())
            ops.head_tc(xcur, self.packer.get(("segmentation_head.0.weight", 1, 16), hw, 1, rows_pad=16), self._head_b16,
                        self.classes, logits_nchw=logits)
        else:
            ops.head(xcur, self.packer.get(("segmentation_head.0.weight", 0, None), hw, 0),
                     self.params["segmentation_head.0.bias"], logits_nchw=logits)
        self._rec("head_fwd", "segmentation_head.0", x=xcur, y=logits)
        tape.append(("head", xcur))
        if self._nbt:
            torch._foreach_add_(self._nbt, 1)
        return logits, tape

    # ---- backward pieces -------------------------------------------------------------------------------
    def _dgrad(self, gy: torch.Tensor, wname: str, x_shape, stride: int, pad: int,
               addend: Optional[torch.Tensor] = None, fp32_weights: bool = False) -> torch.Tensor:
        """data gradient: gy (N, Ho, Wo, Cg >= C_out) -> gx of shape x_shape (+ addend)."""
        w = self.params[wname]
        C_out, C_in, R, S = w.shape
        N, H, W, Cx = x_shape
        Cg = gy.shape[-1]
        tc = self.precision == "bf16" and self.dgrad_tc and not fp32_weights and Cx == C_in and C_in % 16 == 0
        if tc and R == 3 and S == 3 and stride == 1 and pad == 1 and Cg % 16 == 0:
            # the data gradient of a stride-1 conv is a stride-1 conv of gy with the flipped, transposed weights
            wp = self.packer.get((wname, 3, Cg), w, 3, cout_pad=Cg)
            gx = ops.conv2d(gy, wp, self._ones, self._zeros, N=N, H=H, W=W, C_in=Cg, C_x=Cg, C_out=C_in, R=3, S=3, stride=1,
                            pad=1, relu=False, residual=addend, flags=self.pair_flag, tag="dgrad." + wname)
        elif (tc and self.dgrad_s2_zero_insert and stride == 2 and ((R == 3 and pad == 1) or (R == 1 and pad == 0))
              and Cg % 64 == 0 and H == 2 * gy.shape[1] and W == 2 * gy.shape[2]):
            # stride-2 conv: zero-insert gy to the input resolution, then the stride-1 form on the fast halo / TMA kernels
            G = ops.zero_insert2x(gy)
            wp = self.packer.get((wname, 3, Cg), w, 3, cout_pad=Cg)
            gx = ops.conv2d(G, wp, self._ones, self._zeros, N=N, H=H, W=W, C_in=Cg, C_x=Cg, C_out=C_in, R=R, S=S, stride=1,
                            pad=pad, relu=False, residual=addend, algo_cin=Cg // 4, flags=self.pair_flag,
                            tag="dgrad." + wname)   # algo_cin: the algorithmic FLOPs
        elif tc and stride == 2 and ((R == 3 and pad == 1) or (R == 1 and pad == 0)) and Cg % 64 == 0 and Cg == C_out:
            # stride-2 conv: every output pixel gathers the taps whose source coordinate is even (gather producer)
            wp = self.packer.get((wname, 4, Cg), w, 4, cout_pad=Cg)
            gx = ops.conv2d(gy, wp, self._ones, self._zeros, N=N, H=H, W=W, C_in=Cg, C_x=Cg, C_out=C_in, R=R, S=S, stride=2,
                            pad=pad, relu=False, residual=addend, flags=CONV_TRANSPOSED, tag="dgrad." + wname)
        else:
            gx = ops.conv2d_dgrad_direct(gy, w, x_shape, stride, pad, addend=addend,
                                         round_weights=self.precision == "bf16" and not fp32_weights)
        self._rec("dgrad", wname, gy=gy, addend=addend, gx=gx, stride=stride, pad=pad, fp32_weights=fp32_weights)
        return gx

    def _wgrad(self, x: torch.Tensor, gy: torch.Tensor, wname: str, stride: int, pad: int, want_bias: bool = False,
               frame: Optional[torch.Tensor] = None):
        w = self.params[wname]
        out = self.reducer.view(wname)
        bname = wname[:-len("weight")] + "bias"
        if frame is not None:
            dw, db = ops.stem_wgrad_tc(frame, gy, w.shape, out=out, tag="wgrad." + wname), None
        elif self.precision == "bf16" and self.wgrad_tc and ops.wgrad_tc_supported(x, gy, w.shape, stride, pad):
            dw = ops.conv2d_wgrad_tc(x, gy, w.shape, stride, out=out, tag="wgrad." + wname)
            db = ops.channel_sum(gy, w.shape[0], out=self.reducer.view(bname)) if want_bias else None
        else:
            dw, db = ops.conv2d_wgrad_direct(x, gy, w.shape, stride, pad, want_bias=want_bias, out=out,
                                             bias_out=self.reducer.view(bname) if want_bias else None)
        self.reducer.mark(wname)
        if want_bias:
            self.reducer.mark(bname)
        self._rec("wgrad", wname, x=x, gy=gy, dw=dw, db=db, stride=stride, pad=pad)
        return dw, db

    def _conv_bn_bwd(self, e: _ConvBN, g: torch.Tensor, grads: Dict[str, torch.Tensor], want_gz: bool = False,
                     need_dx: bool = True, addend: Optional[torch.Tensor] = None):
        """g = dL/d(a) -> (dL/d(x) (+ addend), gz); fills grads for the conv weight and the BN affine pair."""
        # relu(bn(y)) without a residual: the mask comes from y itself, the activation is not read again
        from_y = e.relu and e.residual is None and self.bn_mask_from_y
        gy, gz, dgamma, dbeta = ops.bn_train_bwd(g, e.a if (e.relu and not from_y) else None, e.y, e.mean, e.invstd, e.scale,
                                                 want_gz=want_gz, dgamma=self.reducer.view(e.bn + ".weight"),
                                                 dbeta=self.reducer.view(e.bn + ".bias"),
                                                 relu_shift=e.shift if from_y else None)
        self.reducer.mark(e.bn + ".weight")
        self.reducer.mark(e.bn + ".bias")
        grads[e.bn + ".weight"], grads[e.bn + ".bias"] = dgamma, dbeta
        self._rec("bn_bwd", e.bn, g=g, a=e.a if e.relu else None, y=e.y, mean=e.mean, invstd=e.invstd, scale=e.scale,
                  gy=gy, gz=gz, dgamma=dgamma, dbeta=dbeta)
        grads[e.conv + ".weight"], _ = self._wgrad(e.x, gy, e.conv + ".weight", e.stride, e.pad, frame=e.frame)
        gx = self._dgrad(gy, e.conv + ".weight", e.x.shape, e.stride, e.pad, addend=addend) if need_dx else None
        return gx, gz

    def backward(self, tape: List, grad_logits: torch.Tensor) -> Dict[str, torch.Tensor]:
        grads: Dict[str, torch.Tensor] = {}
        self.reducer.begin()
        it = list(tape)
        kind, d4 = it.pop()
        assert kind == "head"
        hw = self.params["segmentation_head.0.weight"]
        K = hw.shape[0]
        # bf16 mode: the logits gradient is stored with 16 channels (3 real) so the head runs on the tensor-core kernels
        tc_head = self.precision == "bf16" and self.wgrad_tc and self.dgrad_tc
        g = ops.nchw_to_nhwc(grad_logits, 16 if tc_head else K, self.act_dtype)
        dw, db = self._wgrad(d4, g, "segmentation_head.0.weight", 1, 1, want_bias=True)
        grads["segmentation_head.0.weight"], grads["segmentation_head.0.bias"] = dw, db
        g = self._dgrad(g, "segmentation_head.0.weight", d4.shape, 1, 1, fp32_weights=not tc_head)
        # decoder, last block first
        g_skip: Dict[int, torch.Tensor] = {}
        for i in reversed(range(5)):
            e2, e1 = it.pop(), it.pop()
            kind, cx = it.pop()
            assert kind == "cat"
            g, _ = self._conv_bn_bwd(e2, g, grads)
            g_cat, _ = self._conv_bn_bwd(e1, g, grads)
            g, gs = ops.upsample_concat_bwd(g_cat, cx)
            self._rec("upcat_bwd", f"decoder.blocks.{i}", g_cat=g_cat, g_low=g, g_skip=gs, cx=cx)
            if gs is not None:
                g_skip[4 - i] = gs          # skips = [f4, f3, f2, f1, None]
        # encoder, deepest block first; g = dL/d(f5)
        for li in reversed(range(1, 5)):
            nblk = RESNET34_LAYERS[li - 1]
            for b in reversed(range(nblk)):
                strided = b == 0 and li > 1
                e2 = it.pop()
                ed = it.pop() if strided else None
                e1 = it.pop()
                # out = relu(bn2(conv2(a1)) + idn): gz is the gradient of the sum, shared by both branches
                g_a1, gz = self._conv_bn_bwd(e2, g, grads, want_gz=True)
                if strided:
                    # the block input is f_li (a decoder skip for li >= 2): merge that gradient here
                    gx, _ = self._conv_bn_bwd(e1, g_a1, grads, addend=g_skip.get(li))
                    g, _ = self._conv_bn_bwd(ed, gz, grads, addend=gx)
                else:
                    g, _ = self._conv_bn_bwd(e1, g_a1, grads, addend=gz)
        kind, (f1, pool_idx) = it.pop()
        assert kind == "pool"
        gp = g
        g = ops.maxpool3x3s2_bwd_idx(pool_idx, gp, f1.shape, addend=g_skip.get(1))
        self._rec("maxpool_bwd", "pool", x=f1, gout=gp, addend=g_skip.get(1), gx=g)
        e0 = it.pop()
        self._conv_bn_bwd(e0, g, grads, need_dx=False)
        assert not it
        return self.reducer.finish()      # joins the all-reduce stream; views into the flat gradient buffer


class _UnetTrainFn(torch.autograd.Function):
    """autograd node of the whole train-mode Unet: forward / backward run in the CUDA library."""

    @staticmethod
    def forward(ctx, engine: UnetTrainEngine, x: torch.Tensor, *params):
        logits, tape = engine.forward(x)
        ctx.engine, ctx.tape = engine, tape
        return logits

    @staticmethod
    def backward(ctx, grad_logits):
        """runs the CUDA backward pass; parameter gradients are views into the engine's flat gradient buffer and are
        assigned to ``.grad`` here (each backward OVERWRITES them - zero_grad / accumulation is not needed or supported)."""
        eng = ctx.engine
        grads = eng.backward(ctx.tape, grad_logits.contiguous().float())
        ctx.tape = None
        for n in eng.param_names:
            eng.params[n].grad = grads[n]
        return (None, None) + (None,) * len(eng.param_names)


def unet_train_forward(engine: UnetTrainEngine, x: torch.Tensor) -> torch.Tensor:
    return _UnetTrainFn.apply(engine, x, *[engine.params[n] for n in engine.param_names])
