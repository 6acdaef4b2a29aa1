"""Data-parallel training across the GPUs of one box: bucketed gradient all-reduce overlapped with
the backward pass.

The reference trains on one GPU (``configs/trainer/default.yaml:3``); plain data parallelism of that step means
replicated weights, per-rank batches and BatchNorm statistics (no SyncBN), and ONE collective: the mean of the
parameter gradients.  ``GradBucketReducer`` receives every parameter gradient the moment the backward engine has
launched its kernel (reverse-topological order: head, decoder, encoder), packs it into a flat fp32 bucket and, when a
bucket is full, launches ``all_reduce`` on a side stream (NCCL over NVLink on GPUs, gloo in the CPU tests) while
the remaining data/weight-gradient kernels keep running on the compute stream.  ``finish()`` joins the streams and hands
back views into the buckets, already averaged.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class GradBucketReducer:
    def __init__(self, named_shapes: Sequence[Tuple[str, torch.Size]], device, bucket_bytes: int = 25 << 20,
                 group=None, world_size: Optional[int] = None, flat: Optional[torch.Tensor] = None):
        """``named_shapes``: (name, shape) of every parameter in the order the backward pass produces the gradients.
        ``flat``: an existing flat buffer of the same layout to re-bucket (offsets do not depend on ``bucket_bytes``).

        BatchNorm running statistics are per-rank, exactly as under plain ``DistributedDataParallel`` without SyncBN
        (every rank normalises with its own batch): they are NOT averaged here; rank 0's are the ones a checkpoint keeps."""
        self.group = group
        self.world = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        self.slots: Dict[str, Tuple[int, int, torch.Size]] = {}   # name -> (bucket, offset inside the bucket, shape)
        sizes: List[int] = []
        cur, off = 0, 0
        for name, shape in named_shapes:
            n = int(torch.Size(shape).numel())
            n_pad = (n + 3) // 4 * 4                               # keep every gradient 16-byte aligned
            if off > 0 and (off + n_pad) * 4 > bucket_bytes:
                sizes.append(off)
                cur, off = cur + 1, 0
            self.slots[name] = (cur, off, torch.Size(shape))
            off += n_pad
        sizes.append(off)
        # ONE flat fp32 buffer holds every gradient in backward order; a bucket is a contiguous slice of it, so the
        # optimizer can run a single fused clip + Adam launch over `flat`
        if flat is not None and (flat.numel() != sum(sizes) or flat.device != self.device or flat.dtype != torch.float32):
            raise ValueError("flat gradient buffer does not match the parameter layout")
        self.flat = flat if flat is not None else torch.zeros(sum(sizes), dtype=torch.float32, device=self.device)
        self.bucket_offsets = [sum(sizes[:i]) for i in range(len(sizes))]
        self.buckets = [self.flat[o: o + n] for o, n in zip(self.bucket_offsets, sizes)]
        self.expect = [0] * len(sizes)
        for b, _, _ in self.slots.values():
            self.expect[b] += 1
        self.comm_stream = torch.cuda.Stream(device=self.device) if (self.cuda and self.world > 1) else None
        self._pending = [0] * len(sizes)
        self._works: List = []
        self.launched: List[int] = []          # bucket indices in launch order (observable by tests)

    def flat_offset(self, name: str) -> int:
        b, off, _ = self.slots[name]
        return self.bucket_offsets[b] + off

    def begin(self) -> None:
        self._pending = list(self.expect)
        self._works, self.launched = [], []

    def view(self, name: str) -> torch.Tensor:
        b, off, shape = self.slots[name]
        return self.buckets[b][off: off + shape.numel()].view(shape)

    def add(self, name: str, grad: torch.Tensor) -> None:
        """copy a finished gradient into its slot (callers that cannot write into ``view(name)`` directly)."""
        self.view(name).copy_(grad)
        self.mark(name)

    def mark(self, name: str) -> None:
        """called by the backward engine right after the kernel writing ``view(name)`` has been launched."""
        b, _, _ = self.slots[name]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch(b)

    def _launch(self, b: int) -> None:
        self.launched.append(b)
        if self.world == 1:
            return
        buf = self.buckets[b]
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record()                                  # gradients of this bucket are complete at this point
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ready)
                w = dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                self._works.append((w, buf))
        else:
            self._works.append((dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True), buf))

    def finish(self) -> Dict[str, torch.Tensor]:
        """waits for the collectives, averages, and returns name -> gradient views into the buckets."""
        if any(self._pending):
            missing = [n for n, (b, _, _) in self.slots.items() if self._pending[b] > 0]
            raise RuntimeError(f"gradient buckets incomplete; parameters without a gradient in: {missing[:4]} ...")
        for w, buf in self._works:
            if self.cuda:
                with torch.cuda.stream(self.comm_stream):
                    w.wait()
                    buf.mul_(1.0 / self.world)
            else:
                w.wait()
                buf.mul_(1.0 / self.world)
        if self.cuda and self._works:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        return {n: self.view(n) for n in self.slots}


def backward_param_order(param_names: Sequence[str]) -> List[str]:
    """order in which ``UnetTrainEngine.backward`` finishes the parameter gradients: head, decoder blocks 4..0,
    encoder layer4..layer1 (blocks in reverse), stem."""
    def key(n: str):
        if n.startswith("segmentation_head"):
            return (0, 0, 0)
        if n.startswith("decoder.blocks."):
            i = int(n.split(".")[2])
            return (1, -i, 0 if ".conv2." in n else 1)
        if n.startswith("encoder.layer"):
            li, b = int(n.split(".")[1][5:]), int(n.split(".")[2])
            sub = 0 if (".conv2" in n or ".bn2" in n) else (1 if (".conv1" in n or ".bn1" in n) else 2)
            return (2, -li, -b * 4 + sub)
        return (3, 0, 0)
    return sorted(param_names, key=key)
