"""Data-parallel gradient averaging for the training step (the reference wraps the model
in ``torch.nn.parallel.DistributedDataParallel``, train.py:153, and its only cross-GPU traffic is the gradient
all-reduce fired by ``loss.backward()``, train.py:259).

``FlatGradSync`` keeps the same semantics -- parameters broadcast from rank 0 at construction, gradients AVERAGED over
ranks every step -- but is built so that the step can be replayed as CUDA graphs: every ``p.grad`` is a view into ONE
flat fp32 buffer (autograd accumulates into existing ``.grad`` tensors in place), so

    graph A: flat.zero_() ; forward ; backward        (no collective inside the capture)
    eager  : ONE NCCL all-reduce of the flat buffer over NVLink / NVSwitch (592 MiB for DiffMa-XL), pre-scaled by 1/N
    graph B: fused AdamW step

With ~5 000 kernel launches per DiffMa-XL/4 step the eager DDP step is host-bound (96 ms on 2 GPUs against 39 ms for
the single-GPU graph); torch's DDP reducer hooks could not be captured on this stack (they deadlock), a plain
all-reduce between two graph replays needs no capture at all.  The all-reduce is not overlapped with the backward:
at NVLink-5 bus bandwidth it is ~2-3 ms of a ~40 ms step.
"""
from __future__ import annotations

from typing import Iterable

import torch
import torch.distributed as dist


class FlatGradSync:
    def __init__(self, params: Iterable[torch.nn.Parameter], world_size: int, broadcast: bool = True):
        self.params = [p for p in params if p.requires_grad]
        self.world = world_size
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            if p.dtype != torch.float32:
                raise TypeError("FlatGradSync: master parameters are expected in fp32 (autocast handles the compute dtype)")
            p.grad = self.flat[off:off + n].view_as(p)
            off += n
        if broadcast and world_size > 1:
            for p in self.params:                       # what DDP's constructor does (train.py:153)
                dist.broadcast(p.data, src=0)

    def zero(self) -> None:
        """Zero every gradient with one kernel (do NOT call ``zero_grad(set_to_none=True)``: it would detach the views)."""
        self.flat.zero_()

    def check_views(self) -> None:
        base = self.flat.untyped_storage().data_ptr()
        for p in self.params:
            if p.grad is None or p.grad.untyped_storage().data_ptr() != base:
                raise RuntimeError("FlatGradSync: a parameter's .grad no longer aliases the flat buffer "
                                   "(something called zero_grad(set_to_none=True) or replaced .grad)")

    def allreduce(self) -> None:
        """Average the gradients over all ranks (sum of pre-scaled buffers: one collective, no second pass)."""
        if self.world > 1:
            self.flat.mul_(1.0 / self.world)
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
