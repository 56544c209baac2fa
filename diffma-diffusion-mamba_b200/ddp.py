"""Data-parallel training state for the DiffMa step (SURVEY.md section 8a row a13 and the loop around it).

The reference wraps the model in ``torch.nn.parallel.DistributedDataParallel`` (train.py:153): parameters are broadcast
from rank 0, and ``loss.backward()`` (train.py:259) all-reduces fp32 gradients in 25 MB buckets overlapped with the
rest of the backward; then ``opt.step()`` (AdamW, lr 1e-4, weight decay 0: train.py:201,262) and ``update_ema``
(train.py:34-43,264) walk ~1 900 parameter tensors one by one.

``FlatTrainState`` keeps the same semantics -- broadcast at construction, gradients AVERAGED over ranks, AdamW, EMA --
but lays the state out so that the whole step is ONE CUDA graph per rank:

* parameters, gradients, both Adam moments and the EMA copy are five flat fp32 buffers, offsets 256-byte aligned.
  ``p.data`` is a view into the parameter buffer (or, for the ``lowp`` GEMM weights, a bf16 leaf aliasing a flat bf16
  shadow the optimizer kernel refreshes).  ``.grad`` is cleared at ``begin_step`` so autograd hands its result tensors
  over as they are; one multi-tensor copy per bucket moves them into the flat gradient buffer (no accumulate / cast
  kernel per parameter), and after ``finish_backward`` every fp32 ``.grad`` is the view of its reduced slot again;
* the gradient buffer is cut into buckets from its END (the backward produces the last blocks' gradients first).  A
  post-accumulate hook per parameter counts a bucket down; when it is complete its all-reduce (NCCL over NVLink /
  NVSwitch, SUM) is enqueued on a communication stream while the backward of the earlier blocks keeps running on the
  compute stream -- inside the capture this becomes a fork in the graph, so replays overlap the same way;
* the 1/N of the average is folded into the optimizer kernel's ``grad_scale``: no extra pass over the buffer;
* ``dm_adamw_ema_step`` (csrc/dm_optim.cu) does AdamW + EMA in one pass after the streams join.

``FlatGradSync`` (round 1: one blocking all-reduce between two graphs) is kept for comparison runs.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import ctypes as C

import torch
import torch.distributed as dist

_ALIGN = 64                      # elements (256 B): every parameter view stays 16-byte aligned for the kernels


class FlatTrainState:
    def __init__(self, params: Iterable[torch.nn.Parameter], world_size: int = 1, *, lr: float = 1e-4,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, ema_decay: Optional[float] = 0.9999,
                 bucket_mib: float = 48.0, broadcast: bool = True, process_group=None, overlap: bool = True,
                 lowp: Optional[Iterable[torch.nn.Parameter]] = None):
        """``lowp``: parameters that autocast consumes in bf16 only (``autocast_leaf_params(model)``).  Their fp32 masters
        stay in the flat buffer, but the ``nn.Parameter`` the modules see becomes a bf16 LEAF aliasing a flat bf16 shadow
        that the optimizer kernel refreshes: the forward needs no fp32 -> bf16 cast per weight, autograd hands back bf16
        gradients with no bf16 -> fp32 cast per weight, and one multi-tensor copy per bucket moves them into the flat fp32
        gradient buffer.  Same arithmetic as autocast over fp32 parameters (the forward sees bf16(master) either way)."""
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatTrainState: no trainable parameters")
        self.world, self.group, self.overlap = world_size, process_group, overlap
        self.lr, self.betas, self.eps, self.weight_decay, self.ema_decay = lr, betas, eps, weight_decay, ema_decay
        dev = self.params[0].device
        offs, off = [], 0
        for p in self.params:
            if p.dtype != torch.float32:
                raise TypeError("FlatTrainState: master parameters are expected in fp32 (autocast handles the compute dtype)")
            offs.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.offsets, self.total = offs, off
        self.flat_p = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros_like(self.flat_p)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        self.step_t = torch.zeros((), dtype=torch.float32, device=dev)        # device-side step counter (graph replays)
        lowp_ids = {id(p) for p in (lowp or ())}
        self.is_lowp = [id(p) in lowp_ids for p in self.params]
        self.flat_s = torch.zeros(off, dtype=torch.bfloat16, device=dev) if any(self.is_lowp) else None
        self._g_views = [self.flat_g[o:o + p.numel()].view_as(p) for p, o in zip(self.params, offs)]
        with torch.no_grad():
            for p, o, lp in zip(self.params, offs, self.is_lowp):
                view = self.flat_p[o:o + p.numel()].view_as(p)
                view.copy_(p.data)
                if lp:
                    p.data = self.flat_s[o:o + p.numel()].view_as(p)          # bf16 leaf; filled below
                    p.grad = None
                else:
                    p.data = view
                    p.grad = self.flat_g[o:o + p.numel()].view_as(p)
        if broadcast and world_size > 1:
            dist.broadcast(self.flat_p, src=0, group=process_group)           # what DDP's constructor does (train.py:153)
        if self.flat_s is not None:
            self.flat_s.copy_(self.flat_p)
        self.ema = self.flat_p.clone() if ema_decay is not None else None     # deepcopy(model) of train.py:156
        # ---- buckets: contiguous ranges of the flat gradient buffer, built from the END -----------------------
        want = int(bucket_mib * (1 << 20) / 4)
        self.buckets = []                                                     # [lo, hi, n_params]
        hi, n = off, 0
        self.bucket_of = [0] * len(self.params)
        for i in range(len(self.params) - 1, -1, -1):
            n += 1
            self.bucket_of[i] = len(self.buckets)
            if hi - offs[i] >= want or i == 0:
                self.buckets.append([offs[i], hi, n])
                hi, n = offs[i], 0
        self._pending = [b[2] for b in self.buckets]
        self._fired = [False] * len(self.buckets)
        self._flushed = [False] * len(self.buckets)
        self._of_bucket = [[i for i in range(len(self.params)) if self.bucket_of[i] == b] for b in range(len(self.buckets))]
        self.comm_stream = torch.cuda.Stream(device=dev) if (dev.type == "cuda" and world_size > 1) else None
        self._hooks = []
        if world_size > 1 and overlap:
            for i, p in enumerate(self.params):
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(self.bucket_of[i])))

    # ---- gradient synchronisation ---------------------------------------------------------------------------
    def _make_hook(self, b: int):
        def hook(_param):
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._reduce_bucket(b)
        return hook

    def _flush(self, b: int) -> None:
        """The gradients autograd produced for bucket ``b`` (fp32, or bf16 for the lowp leaves) -> their fp32 slots in the
        flat gradient buffer: one multi-tensor copy instead of one accumulate kernel per parameter."""
        if self._flushed[b]:
            return
        self._flushed[b] = True
        idx = [i for i in self._of_bucket[b] if self.params[i].grad is not None]
        if idx:
            torch._foreach_copy_([self._g_views[i] for i in idx], [self.params[i].grad for i in idx])

    def _reduce_bucket(self, b: int) -> None:
        if self._fired[b]:
            return
        self._fired[b] = True
        self._flush(b)
        lo, hi, _ = self.buckets[b]
        chunk = self.flat_g[lo:hi]
        if self.comm_stream is None:                                           # CPU (gloo) path of the tests
            dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)
            return
        self.comm_stream.wait_stream(torch.cuda.current_stream())              # the bucket's gradients are complete
        with torch.cuda.stream(self.comm_stream):
            dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group)

    def begin_step(self) -> None:
        """Zero the flat gradient buffer with one kernel, re-arm the buckets and clear every ``.grad``: autograd then hands
        its result tensors over as they are (no accumulate kernel per parameter); ``_flush`` copies them into the flat
        buffer bucket by bucket and ``finish_backward`` points every ``.grad`` at its (reduced) slot again.  One ``backward()`` per
        ``begin_step()``: gradient accumulation over several backward passes is not supported by this state (the reference's
        loop, train.py:252-264, does one backward per optimizer step as well)."""
        self.flat_g.zero_()
        self._pending = [b[2] for b in self.buckets]
        self._fired = [False] * len(self.buckets)
        self._flushed = [False] * len(self.buckets)
        for p in self.params:
            p.grad = None

    def finish_backward(self, reduce: bool = True) -> None:
        """After ``loss.backward()``: reduce whatever the hooks have not (unused parameters, overlap off) and make the
        compute stream wait for the communication stream.  Gradients hold the SUM over ranks afterwards."""
        if self.world == 1 or not self.overlap or not reduce:
            for b in range(len(self.buckets)):
                self._flush(b)
        if self.world > 1 and reduce:
            if self.overlap:
                for b in range(len(self.buckets)):
                    self._reduce_bucket(b)
            else:
                dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=self.group)   # one blocking collective
            if self.comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
        for p, lp, v in zip(self.params, self.is_lowp, self._g_views):
            if not lp:
                p.grad = v               # fp32 leaves: ``.grad`` is the (summed) slot of the flat buffer again

    def check_views(self) -> None:
        gb, pb = self.flat_g.untyped_storage().data_ptr(), self.flat_p.untyped_storage().data_ptr()
        sb = None if self.flat_s is None else self.flat_s.untyped_storage().data_ptr()
        for p, lp in zip(self.params, self.is_lowp):
            if lp:
                if p.data.untyped_storage().data_ptr() != sb or p.dtype != torch.bfloat16:
                    raise RuntimeError("FlatTrainState: a bf16 leaf parameter no longer aliases the flat shadow buffer")
                continue
            if (p.grad is not None and p.grad.untyped_storage().data_ptr() != gb) or p.data.untyped_storage().data_ptr() != pb:
                raise RuntimeError("FlatTrainState: a parameter or its .grad no longer aliases the flat buffers "
                                   "(something called zero_grad(set_to_none=True), .to(), or replaced .data / .grad)")

    # ---- optimizer + EMA --------------------------------------------------------------------------------------
    def optimizer_step(self) -> None:
        """AdamW + EMA over the whole flat state: one ``dm_adamw_ema_step`` launch (CUDA only)."""
        self.step_t.add_(1.0)
        scale = 1.0 / self.world
        b1, b2 = self.betas
        if self.flat_p.is_cuda:
            from . import _cabi, ops
            args = _cabi.AdamwArgs(
                self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                None if self.ema is None else self.ema.data_ptr(), self.step_t.data_ptr(),
                None if self.flat_s is None else self.flat_s.data_ptr(), self.total, self.lr, b1, b2, self.eps,
                self.weight_decay, 0.0 if self.ema_decay is None else self.ema_decay, scale)
            st = _cabi.lib().dm_adamw_ema_step_ex(C.byref(args), torch.cuda.current_stream(self.flat_p.device).cuda_stream)
            _cabi.check(st, "dm_adamw_ema_step_ex")
            ops.LAUNCH_COUNTER["kernels"] += 1
            ops.invalidate_weight_caches()      # (a graph REPLAY of this call does not run this line: see replayed())
            return
        raise RuntimeError("FlatTrainState.optimizer_step: diffma_b200 has no CPU path (state on "
                           f"{self.flat_p.device}); oracle/ref_optim.py holds the reference update for tests")

    @staticmethod
    def replayed() -> None:
        """Call after replaying a CUDA graph that contains ``optimizer_step``: the weights changed behind autograd's
        back, so the inference-path weight caches (ops.weights_key) must be dropped before the next no-grad forward."""
        from . import ops
        ops.invalidate_weight_caches()

    def master_state(self, named_params) -> dict:
        """name -> fp32 master tensor (views into the flat parameter buffer).  With ``lowp`` leaves ``model.state_dict()``
        holds their bf16 shadows; checkpoints (train.py:293-300 saves model.module.state_dict()) take the masters."""
        by_id = {id(p): o for p, o in zip(self.params, self.offsets)}
        return {n: self.flat_p[by_id[id(p)]:by_id[id(p)] + p.numel()].view_as(p) for n, p in named_params if id(p) in by_id}

    @torch.no_grad()
    def load_master_state(self, named_params, state_dict: dict, strict: bool = True) -> None:
        """Load fp32 weights (a checkpoint's ``model`` entry) into an EXISTING state: into the flat masters, then refresh the
        bf16 shadows.  ``net.load_state_dict`` after construction would write the ``lowp`` leaves' bf16 shadows only and
        leave their masters stale -- either build the state after loading, or load through this method."""
        by_id = {id(p): o for p, o in zip(self.params, self.offsets)}
        missing = []
        for n, p in named_params:
            if id(p) not in by_id:
                continue
            if n not in state_dict:
                missing.append(n)
                continue
            o = by_id[id(p)]
            self.flat_p[o:o + p.numel()].view_as(p).copy_(state_dict[n].to(self.flat_p.device, torch.float32))
        if missing and strict:
            raise KeyError(f"load_master_state: missing {len(missing)} entries, e.g. {missing[:3]}")
        if self.flat_s is not None:
            self.flat_s.copy_(self.flat_p)

    def ema_state(self, named_params) -> dict:
        """name -> EMA tensor (views into the flat EMA buffer), for checkpoints (train.py:293-300 saves ema.state_dict())."""
        if self.ema is None:
            raise RuntimeError("FlatTrainState was built without an EMA copy")
        by_id = {id(p): o for p, o in zip(self.params, self.offsets)}
        return {n: self.ema[by_id[id(p)]:by_id[id(p)] + p.numel()].view_as(p) for n, p in named_params if id(p) in by_id}


_LOWP_SUFFIXES = ("in_proj.weight", "in_proj.bias", "out_proj.weight", "out_proj.bias", "adaLN_modulation.1.weight",
                  "adaLN_modulation.1.bias", "attention_network.1.weight", "attention_network.1.bias", "x_proj.weight",
                  "dt_proj.weight")


def autocast_leaf_params(model: torch.nn.Module) -> List[torch.nn.Parameter]:
    """The big GEMM operands of every block (in/out projections, adaLN modulation, attention Linear): consumed only through
    ``.to(bf16)`` / autocast ``F.linear`` -- candidates for ``FlatTrainState(lowp=...)``.  LayerNorm, conv1d, A_log, D and
    dt_proj.bias stay fp32 leaves (the kernels read them in fp32)."""
    return [p for n, p in model.named_parameters() if p.requires_grad and n.endswith(_LOWP_SUFFIXES)]


class FlatGradSync:
    """Round-1 variant: every ``.grad`` a view into one flat buffer, ONE blocking all-reduce between two graphs."""

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
