"""Torch-facing operators over the C-ABI (``include/diffma_b200.h``) -- device memory and streams only.

Two levels:

* ``mamba1_scan`` / ``mamba2_ssd``: the B200-native ops.  They take tokens-major activations for one or
  several mixers ("groups") and a ``ScanPlan`` (directions + gather table resident on the device) and run
  ALL directions of ALL groups in one C-ABI call.  Used by ``mixer.py``.
* ``mamba_inner_fn`` / ``mamba_split_conv1d_scan_combined`` / ``RMSNormGated`` ...: the upstream
  ``mamba_ssm`` signatures the reference imports (block/mamba.py:11, block/mamba2.py:17-21), implemented on
  top of the ops above.  ``shims/`` re-exports them under the upstream module paths.

There is no CPU path: every op raises on non-CUDA tensors (the oracle lives in ``oracle/`` and is test-only).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F

from . import _cabi

__all__ = ["ScanPlan", "mamba1_scan", "mamba2_ssd", "mamba_inner_fn", "selective_scan_fn", "causal_conv1d_fn",
           "mamba_split_conv1d_scan_combined", "mamba_chunk_scan_combined", "RMSNormGated"]


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return _cabi.DM_F32
    if t.dtype == torch.bfloat16:
        return _cabi.DM_BF16
    raise TypeError(f"diffma_b200 kernels take fp32 or bf16 activations, got {t.dtype} "
                    "(the reference's fp16 autocast maps to bf16 here, see DESIGN.md)")


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: diffma_b200 has no CPU path (tensor on {t.device}); "
                           "use oracle/ for CPU reference values in tests")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream_handle(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


# --------------------------------------------------------------------------------------------------
# direction plans
# --------------------------------------------------------------------------------------------------
@dataclass
class ScanPlan:
    """How one mixer call scans its tokens (reference semantics: SURVEY.md App. A.4).

    ``table``  int32 device tensor (n_dir, seqlen): scanned token j of direction k is source token
               ``table[k, j]``; a row starting with -1 is the identity; ``None`` = all identity.
    ``layout`` where the per-direction outputs go:
               "concat"   token-order rows, directions side by side: (B, L_src, n_dir, D) -> the merge of
                          CrossMerge (block/mamba.py:60-82) becomes a sum over the direction axis, which the
                          out-projection absorbs;
               "disjoint" token-order rows, directions cover disjoint token sets (EfficientVMamba): (B, L_src, D);
               "stacked"  scan-order rows per direction: (B, n_dir, seqlen, D) (what mamba_inner_fn returns).
    """
    n_dir: int
    seqlen: int
    src_len: int
    table: Optional[torch.Tensor]
    layout: str
    table_host: Optional[torch.Tensor] = None      # CPU copy of ``table`` (identity markers are read on the host)

    @property
    def out_order(self) -> int:
        return _cabi.DM_OUT_SCAN_ORDER if self.layout == "stacked" else _cabi.DM_OUT_TOKEN_ORDER

    def inverse_table(self):
        """Per direction the inverse permutation (source token -> scanned position) as device int64 tensors (None for an
        identity direction), or None altogether if some direction does not cover every source token exactly once."""
        if self.seqlen != self.src_len:
            return None
        if getattr(self, "_inv", None) is None:
            inv = []
            for k in range(self.n_dir):
                if self.table is None or int(self.table_host[k][0]) < 0:
                    inv.append(None)
                else:
                    row = self.table_host[k].long()
                    r = torch.empty_like(row)
                    r[row] = torch.arange(row.numel())
                    inv.append(r.to(self.table.device))
            self._inv = inv
        return self._inv

    def out_shape(self, batch: int, d: int):
        if self.layout == "concat":
            return (batch, self.src_len, self.n_dir, d)
        if self.layout == "disjoint":
            return (batch, self.src_len, d)
        return (batch, self.n_dir, self.seqlen, d)

    def out_strides(self, d: int):
        """(batch, dir, token) strides in elements."""
        if self.layout == "concat":
            return (self.src_len * self.n_dir * d, d, self.n_dir * d)
        if self.layout == "disjoint":
            return (self.src_len * d, 0, d)
        return (self.n_dir * self.seqlen * d, self.seqlen * d, d)

    @staticmethod
    def build(orders: Sequence[Optional[Sequence[int]]], src_len: int, layout: str, device) -> "ScanPlan":
        """``orders``: one gather list per direction (``None`` = identity)."""
        n_dir = len(orders)
        lens = {len(o) for o in orders if o is not None}
        assert len(lens) <= 1, "all directions of a call must have the same length"
        seqlen = lens.pop() if lens else src_len
        if all(o is None for o in orders):
            table = None
        else:
            rows = []
            for o in orders:
                if o is None:
                    assert seqlen == src_len
                    rows.append(torch.full((seqlen,), -1, dtype=torch.int32))
                else:
                    r = torch.as_tensor(list(o), dtype=torch.int64)
                    assert r.numel() == seqlen and int(r.min()) >= 0 and int(r.max()) < src_len
                    rows.append(r.to(torch.int32))
            host = torch.stack(rows).contiguous()
            table = host.to(device)
            plan = ScanPlan(n_dir, seqlen, src_len, table, layout)
            plan.table_host = host
            return plan
        return ScanPlan(n_dir, seqlen, src_len, table, layout)


# --------------------------------------------------------------------------------------------------
# Mamba-1
# --------------------------------------------------------------------------------------------------
@dataclass
class Mamba1Weights:
    conv_weight: torch.Tensor       # (D, W) fp32
    conv_bias: Optional[torch.Tensor]
    x_proj_weight: torch.Tensor     # (R+2N, D) act dtype
    dt_proj_weight: torch.Tensor    # (D, R) act dtype
    dt_bias: Optional[torch.Tensor]  # (D) fp32
    A: torch.Tensor                 # (D, N) fp32 = -exp(A_log)
    D: Optional[torch.Tensor]       # (D) fp32


def _chk(t: Optional[torch.Tensor], dtype, shape, name):
    if t is None:
        return None
    if t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous() or not t.is_cuda:
        raise RuntimeError(f"{name}: expected contiguous CUDA {dtype} {tuple(shape)}, got {t.dtype} "
                           f"{tuple(t.shape)} contiguous={t.is_contiguous()} on {t.device}")
    return t


# ---- inference weight caches ------------------------------------------------------------------------------
# The no-grad fast path caches act-dtype copies / fused forms of the weights (blocks._fused_weights, DiffMa._cached,
# Mamba.scan_weights).  Keys hold (epoch, data_ptr, _version) of EVERY parameter consumed: ``_version`` catches
# ordinary in-place updates, ``data_ptr`` catches ``load_state_dict(assign=True)`` / re-flattened parameters, and the
# epoch is for updates autograd cannot see -- a CUDA-graph replay of a captured optimizer step, or a kernel writing the
# flat parameter buffer (ddp.FlatTrainState bumps it; loops that only REPLAY a captured step must call
# ``invalidate_weight_caches()`` themselves before the next no-grad forward).  A captured GraphedSampler bakes in pointers
# to the cached copies: it checks the epoch on every step and asks to be rebuilt when the weights have changed.
_WEIGHTS_EPOCH = [0]


def invalidate_weight_caches() -> None:
    _WEIGHTS_EPOCH[0] += 1


def weights_epoch() -> int:
    return _WEIGHTS_EPOCH[0]


def weights_key(params, *extra):
    return (_WEIGHTS_EPOCH[0], tuple(extra), tuple((p.data_ptr(), p._version) for p in params))


# kernels of OURS launched so far (bench.py reports it as gpu_launches; CUDA-graph replays are counted by the
# sampler as captured launches x replays)
LAUNCH_COUNTER = {"kernels": 0}


def _strides(x: torch.Tensor):
    """(batch, token) strides in elements, normalised for size-1 dimensions (whose torch strides are arbitrary)."""
    ts = x.stride(1) if x.shape[1] > 1 else x.shape[2]
    bs = x.stride(0) if x.shape[0] > 1 else x.shape[1] * ts
    return bs, ts


# scratch of the scan kernel's dynamic schedule: one zero-initialised buffer per (device, launch geometry), kept alive
# for the life of the process (CUDA graphs hold its address); every launch leaves it re-armed (include/diffma_b200.h)
_SCHED_WS = {}
USE_DYNAMIC_SCHEDULE = True
# A small kernel between conv + x_proj and the scan hands the scan delta = softplus(dt_proj(dt_low) + bias) as fp16
# (inference, bf16, d_inner 1024): the MUFU-bound scan sheds the softplus and the dt_proj MMA (131 -> 109 us at the
# headline shape).  DIFFMA_DELTA=0: the scan evaluates them itself (round-1 behaviour; A/B runs).
import os as _os
USE_DELTA_HANDOVER = _os.environ.get("DIFFMA_DELTA", "1") != "0"


def _sched_workspace(device, batch: int, n_dir: int, d_inner: int, groups: int):
    """One workspace per (device, STREAM, launch geometry): two same-shape scans running concurrently on different
    streams (two samplers, a prefetch stream) must not share tickets and hand-over states.  Launches on one stream are
    ordered, and a captured graph keeps using the workspace of the stream it was captured on."""
    stream = torch.cuda.current_stream(device)
    key = (str(device), stream.cuda_stream, batch, n_dir, d_inner, groups)
    ws = _SCHED_WS.get(key)
    if ws is None:
        need = int(_cabi.lib().dm_mamba1_sched_workspace_bytes(batch, n_dir, d_inner, groups))
        if need <= 0:
            return None
        if torch.cuda.is_current_stream_capturing():
            # a memset recorded into a graph would not have run before an eager launch that finds the cached buffer
            raise RuntimeError("mamba1_scan: the ready-queue workspace for this shape must be allocated (and zeroed) "
                               "before CUDA-graph capture: run the step once eagerly on the capture stream first")
        ws = torch.zeros(need, dtype=torch.uint8, device=device)
        _SCHED_WS[key] = ws
    return ws


def mamba1_args(xz: List[torch.Tensor], weights: List[Mamba1Weights], plan: ScanPlan, bufs=None, dynamic=None,
                chunk_states=None, z_gated: bool = False, delta=None, out_strides=None):
    """Fill a ``dm_mamba1_args`` for the given groups.  Returns (args, (out, u, x_dbl)); the tensors own the memory
    the struct points at and must outlive the launch.  ``bufs`` = existing (out-shaped, u, x_dbl) tensors to point at
    instead of allocating (the backward passes dout / the saved intermediates)."""
    G = len(xz)
    assert 1 <= G <= _cabi.DM_MAX_GROUPS and len(weights) == G
    x0 = xz[0]
    _require_cuda(x0, "mamba1_scan")
    B, Lsrc, D2 = x0.shape
    D = D2 // 2
    assert Lsrc == plan.src_len, f"plan built for {plan.src_len} source tokens, got {Lsrc}"
    N = weights[0].A.shape[1]
    E = weights[0].x_proj_weight.shape[0]
    R = E - 2 * N
    a = _cabi.Mamba1Args()
    a.batch, a.n_dir, a.seqlen = B, plan.n_dir, plan.seqlen
    a.d_inner, a.d_state, a.dt_rank, a.d_conv = D, N, R, weights[0].conv_weight.shape[1]
    a.act_dtype, a.out_order, a.n_groups = _dtype_code(x0), plan.out_order, G
    a.order = _ptr(plan.table)
    a.z_is_gated = int(bool(z_gated))
    if USE_DYNAMIC_SCHEDULE if dynamic is None else dynamic:
        ws = _sched_workspace(x0.device, B, plan.n_dir, D, G)
        if ws is not None:
            a.sched_workspace, a.sched_workspace_bytes = ws.data_ptr(), ws.numel()
    # (batch, direction, token) strides of the out-shaped buffer; ``out_strides`` overrides them for a strided view (the
    # backward reads an upstream gradient that may be broadcast over the directions: direction stride 0)
    obs, ods, ots = plan.out_strides(D) if out_strides is None else out_strides
    # one allocation per kind so groups are adjacent (lets callers view them as a batch)
    if bufs is None:
        out_all = torch.empty((G,) + plan.out_shape(B, D), dtype=x0.dtype, device=x0.device)
        u_all = torch.empty((G, B, plan.n_dir, plan.seqlen, D), dtype=x0.dtype, device=x0.device)
        xd_all = torch.empty((G, B, plan.n_dir, plan.seqlen, E), dtype=torch.float32, device=x0.device)
    else:
        out_all, u_all, xd_all = bufs
        assert (out_strides is not None or out_all.is_contiguous()) and u_all.is_contiguous() and xd_all.is_contiguous()
        assert tuple(out_all.shape) == (G,) + plan.out_shape(B, D) and out_all.dtype == x0.dtype
    for g in range(G):
        x, w = xz[g], weights[g]
        if x.shape != x0.shape or x.dtype != x0.dtype or x.stride(2) != 1:
            raise RuntimeError("mamba1_scan: every group needs the same (B, L, 2D) shape/dtype, channel stride 1")
        gs = a.group[g]
        gs.xz = x.data_ptr()
        gs.xz_batch_stride, gs.xz_token_stride = _strides(x)
        gs.out, gs.out_batch_stride, gs.out_dir_stride, gs.out_token_stride = out_all[g].data_ptr(), obs, ods, ots
        gs.u, gs.x_dbl = u_all[g].data_ptr(), xd_all[g].data_ptr()
        gs.conv_weight = _chk(w.conv_weight, torch.float32, (D, a.d_conv), "conv_weight").data_ptr()
        gs.conv_bias = _ptr(_chk(w.conv_bias, torch.float32, (D,), "conv_bias"))
        gs.x_proj_weight = _chk(w.x_proj_weight, x0.dtype, (E, D), "x_proj_weight").data_ptr()
        gs.dt_proj_weight = _chk(w.dt_proj_weight, x0.dtype, (D, R), "dt_proj_weight").data_ptr()
        gs.dt_bias = _ptr(_chk(w.dt_bias, torch.float32, (D,), "dt_bias"))
        gs.A = _chk(w.A, torch.float32, (D, N), "A").data_ptr()
        gs.D = _ptr(_chk(w.D, torch.float32, (D,), "D"))
        if delta is not None:                   # (G, B, n_dir, seqlen, D) fp16: softplus'ed delta handed to the scan kernel
            gs.delta = _chk(delta[g], torch.float16, (B, plan.n_dir, plan.seqlen, D), "delta").data_ptr()
        if chunk_states is not None:            # training: recurrence checkpoints for the backward (fp32, contiguous)
            gs.chunk_states = _chk(chunk_states[g], torch.float32, chunk_states.shape[1:], "chunk_states").data_ptr()
    return a, (out_all, u_all, xd_all)


def mamba1_state_shape(G: int, B: int, plan: ScanPlan, D: int, N: int):
    """Shape of the recurrence checkpoints the training forward hands to the backward (include/diffma_b200.h)."""
    ct = int(_cabi.lib().dm_mamba1_bwd_chunk_tokens())
    return (G, B, plan.n_dir, (plan.seqlen + ct - 1) // ct, D, N)


def mamba1_scan_raw(xz: List[torch.Tensor], weights: List[Mamba1Weights], plan: ScanPlan, chunk_states=None,
                    z_gated: bool = False):
    """One C-ABI call: conv1d+SiLU -> x_proj -> dt_proj -> softplus -> scan -> D skip -> SiLU(z) gate.

    xz[g]: (B, L_src, 2D) tokens-major (last-dim stride 1).  Returns (out, u, x_dbl), each with a leading group
    axis: ``out[g]`` has ``plan.out_shape``; u (scan order) and x_dbl are the intermediates the backward reads.
    """
    delta = None
    x0 = xz[0]
    if (USE_DELTA_HANDOVER and chunk_states is None and not z_gated and x0.dtype == torch.bfloat16
            and x0.shape[-1] == 2048):
        delta = torch.empty((len(xz), x0.shape[0], plan.n_dir, plan.seqlen, 1024), dtype=torch.float16, device=x0.device)
    a, bufs = mamba1_args(xz, weights, plan, chunk_states=chunk_states, z_gated=z_gated, delta=delta)
    st = _cabi.lib().dm_mamba1_scan_fwd(C.byref(a), C.c_void_p(_stream_handle(xz[0].device)))
    _cabi.check(st, "dm_mamba1_scan_fwd")
    LAUNCH_COUNTER["kernels"] += 3 if delta is not None else 2        # conv + x_proj, [delta,] scan
    return bufs


def mamba1_scan(xz: List[torch.Tensor], weights: List[Mamba1Weights], plan: ScanPlan, z_gated: bool = False) -> torch.Tensor:
    """-> (G,) + plan.out_shape.  ``z_gated``: the z half of ``xz`` already holds silu(z) (inference path whose
    in-projection epilogue applied it, ``gemm_bf16_tn(..., silu_from=d_inner)``); not differentiable."""
    if torch.is_grad_enabled() and any(t.requires_grad for t in xz):
        if z_gated:
            raise RuntimeError("mamba1_scan: z_gated inputs are inference-only (the backward needs the raw z)")
        from . import autograd_ops
        return autograd_ops.Mamba1ScanFn.apply(plan, len(xz), *xz, *autograd_ops.flatten_weights(weights))
    return mamba1_scan_raw(xz, weights, plan, z_gated=z_gated)[0]


# --------------------------------------------------------------------------------------------------
# Mamba-2
# --------------------------------------------------------------------------------------------------
@dataclass
class Mamba2Weights:
    conv_weight: torch.Tensor       # (D + 2N, W) fp32
    conv_bias: Optional[torch.Tensor]
    dt_bias: Optional[torch.Tensor]  # (H) fp32
    A: torch.Tensor                 # (H) fp32 = -exp(A_log)
    D: Optional[torch.Tensor]       # (H) fp32


def mamba2_ssd_raw(zxbcdt: List[torch.Tensor], weights: List[Mamba2Weights], plan: ScanPlan, d_inner: int,
                   d_state: int, nheads: int, gate: bool = True, want_sumsq: bool = True):
    """One C-ABI call: conv1d+SiLU over x|B|C -> softplus(dt) -> SSD recurrence -> D skip -> SiLU(z) gate.

    Returns (v, sumsq): v (G,)+plan.out_shape ; sumsq (G, B, n_dir, rows) fp32 = sum_c v^2 addressed like v's
    rows (rows = L_src for token-order layouts, seqlen for "stacked"), or None.
    """
    G = len(zxbcdt)
    assert 1 <= G <= _cabi.DM_MAX_GROUPS and len(weights) == G
    x0 = zxbcdt[0]
    _require_cuda(x0, "mamba2_ssd")
    B, Lsrc, Cin = x0.shape
    assert Lsrc == plan.src_len and Cin == 2 * d_inner + 2 * d_state + nheads
    a = _cabi.Mamba2Args()
    a.batch, a.n_dir, a.seqlen = B, plan.n_dir, plan.seqlen
    a.d_inner, a.d_state, a.nheads, a.d_conv = d_inner, d_state, nheads, weights[0].conv_weight.shape[1]
    a.act_dtype, a.out_order, a.n_groups, a.gate = _dtype_code(x0), plan.out_order, G, int(gate)
    a.order = _ptr(plan.table)
    obs, ods, ots = plan.out_strides(d_inner)
    v_all = torch.empty((G,) + plan.out_shape(B, d_inner), dtype=x0.dtype, device=x0.device)
    rows = plan.seqlen if plan.layout == "stacked" else plan.src_len
    ss_all = torch.zeros((G, B, plan.n_dir, rows), dtype=torch.float32, device=x0.device) if want_sumsq else None
    Cc = d_inner + 2 * d_state
    for g in range(G):
        x, w = zxbcdt[g], weights[g]
        if x.shape != x0.shape or x.dtype != x0.dtype or x.stride(2) != 1:
            raise RuntimeError("mamba2_ssd: every group needs the same (B, L, C) shape/dtype, channel stride 1")
        gs = a.group[g]
        gs.zxbcdt = x.data_ptr()
        gs.in_batch_stride, gs.in_token_stride = _strides(x)
        gs.out, gs.out_batch_stride, gs.out_dir_stride, gs.out_token_stride = v_all[g].data_ptr(), obs, ods, ots
        if want_sumsq:
            gs.sumsq, gs.sumsq_batch_stride, gs.sumsq_dir_stride = ss_all[g].data_ptr(), plan.n_dir * rows, rows
        gs.conv_weight = _chk(w.conv_weight, torch.float32, (Cc, a.d_conv), "conv_weight").data_ptr()
        gs.conv_bias = _ptr(_chk(w.conv_bias, torch.float32, (Cc,), "conv_bias"))
        gs.dt_bias = _ptr(_chk(w.dt_bias, torch.float32, (nheads,), "dt_bias"))
        gs.A = _chk(w.A, torch.float32, (nheads,), "A").data_ptr()
        gs.D = _ptr(_chk(w.D, torch.float32, (nheads,), "D"))
    st = _cabi.lib().dm_mamba2_ssd_fwd(C.byref(a), C.c_void_p(_stream_handle(x0.device)))
    _cabi.check(st, "dm_mamba2_ssd_fwd")
    LAUNCH_COUNTER["kernels"] += 1
    return v_all, ss_all


def mamba2_ssd(zxbcdt, weights, plan, d_inner, d_state, nheads, gate=True, want_sumsq=True):
    if torch.is_grad_enabled() and any(t.requires_grad for t in zxbcdt):
        from . import autograd_ops
        v, ss = autograd_ops.Mamba2SsdFn.apply(plan, len(zxbcdt), d_inner, d_state, nheads, gate, want_sumsq, *zxbcdt,
                                               *autograd_ops.flatten_weights(weights))
        return v, (ss if want_sumsq else None)
    return mamba2_ssd_raw(zxbcdt, weights, plan, d_inner, d_state, nheads, gate, want_sumsq)


# --------------------------------------------------------------------------------------------------
# upstream-signature operators (what block/mamba.py and block/mamba2.py import)
# --------------------------------------------------------------------------------------------------
def _autocast_dtype(x: torch.Tensor):
    if torch.is_autocast_enabled():
        dt = torch.get_autocast_dtype("cuda")
        if dt == torch.float16:
            raise TypeError("diffma_b200: fp16 autocast is not supported, use torch.autocast('cuda', torch.bfloat16)")
        return dt
    return x.dtype


def mamba_inner_fn(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight,
                   out_proj_bias, A, B=None, C=None, D=None, delta_bias=None, B_proj_bias=None,
                   C_proj_bias=None, delta_softplus=True):
    """[upstream mamba_ssm.ops.selective_scan_interface.mamba_inner_fn]; call sites block/mamba.py:346-393.

    xz (B, 2D, L) in the reference's channel-major layout -> (B, L, d_model).  Same mixed-precision
    rule as upstream's ``custom_fwd``: under autocast the projection weights and activations run in the
    autocast dtype; conv weights, A, D, delta_bias stay fp32.
    """
    if B is not None or C is not None or B_proj_bias is not None or C_proj_bias is not None:
        raise NotImplementedError("diffma_b200.mamba_inner_fn: only input-dependent B and C without projection "
                                  "biases are supported (all DiffMa uses)")
    if not delta_softplus:
        raise NotImplementedError("diffma_b200.mamba_inner_fn: delta_softplus=False is never used by DiffMa")
    _require_cuda(xz, "mamba_inner_fn")
    if xz.stride(-1) != 1 and xz.stride(1) != 1:
        raise RuntimeError("mamba_inner_fn: xz must have unit stride along L or along channels")
    act = _autocast_dtype(xz)
    Bsz, D2, L = xz.shape
    xz_t = xz.to(act).transpose(1, 2)
    if xz_t.stride(2) != 1:
        xz_t = xz_t.contiguous()                       # channel-major -> tokens-major (one copy)
    w = Mamba1Weights(
        conv_weight=conv1d_weight.reshape(conv1d_weight.shape[0], -1).float().contiguous(),
        conv_bias=None if conv1d_bias is None else conv1d_bias.float().contiguous(),
        x_proj_weight=x_proj_weight.to(act).contiguous(), dt_proj_weight=delta_proj_weight.to(act).contiguous(),
        dt_bias=None if delta_bias is None else delta_bias.float().contiguous(),
        A=A.float().contiguous(), D=None if D is None else D.float().contiguous())
    plan = ScanPlan(1, L, L, None, "stacked")
    with torch.autocast("cuda", enabled=False):
        y = mamba1_scan([xz_t], [w], plan)[0][:, 0]     # (B, L, D)
        return F.linear(y, out_proj_weight.to(act), None if out_proj_bias is None else out_proj_bias.to(act))


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False):
    """[upstream selective_scan_fn] -- imported at block/mamba.py:11 but only reached with
    ``use_fast_path=False``, which no DiffMa block ever sets."""
    raise NotImplementedError("diffma_b200: selective_scan_fn with a precomputed delta is a dead path in DiffMa "
                              "(use_fast_path is always True); use mamba_inner_fn")


def causal_conv1d_fn(x, weight, bias=None, seq_idx=None, initial_states=None, return_final_states=False,
                     final_states_out=None, activation=None):
    """[upstream causal_conv1d.causal_conv1d_fn] -- reached only on the non-fused paths DiffMa never takes."""
    raise NotImplementedError("diffma_b200: standalone causal_conv1d_fn is a dead path in DiffMa; the conv is "
                              "fused into mamba_inner_fn / mamba_split_conv1d_scan_combined")


def causal_conv1d_update(*args, **kwargs):
    raise NotImplementedError("diffma_b200: single-token decode (step()) is never used by a diffusion model")


def mamba_split_conv1d_scan_combined(zxbcdt, conv1d_weight, conv1d_bias, dt_bias, A, D, chunk_size,
                                     initial_states=None, seq_idx=None, dt_limit=(0.0, float("inf")),
                                     return_final_states=False, activation="silu", rmsnorm_weight=None,
                                     rmsnorm_eps=1e-6, outproj_weight=None, outproj_bias=None, headdim=None,
                                     ngroups=1, norm_before_gate=True):
    """[upstream ssd_combined.mamba_split_conv1d_scan_combined]; call sites block/mamba2.py:392-696.

    zxbcdt (B, L, 2*d_in + 2*N + H) -> (B, L, d_model).  ``chunk_size`` only tiles the upstream kernels; any
    chunking yields the same recurrence, so it is accepted and ignored.
    """
    if initial_states is not None or seq_idx is not None or return_final_states:
        raise NotImplementedError("diffma_b200: initial_states / seq_idx / final states are never used by DiffMa")
    if tuple(dt_limit) != (0.0, float("inf")):
        raise NotImplementedError("diffma_b200: dt_limit clamping is never used by DiffMa")
    if activation not in ("silu", "swish") or ngroups != 1:
        raise NotImplementedError("diffma_b200: only activation=silu, ngroups=1")
    if rmsnorm_weight is not None and norm_before_gate:
        raise NotImplementedError("diffma_b200: norm_before_gate=True is never used by DiffMa")
    if D.dim() != 1:
        raise NotImplementedError("diffma_b200: D must be per-head (H,)")
    _require_cuda(zxbcdt, "mamba_split_conv1d_scan_combined")
    act = _autocast_dtype(zxbcdt)
    H = D.shape[0]
    assert headdim is not None
    d_in = H * headdim
    Bsz, L, Cin = zxbcdt.shape
    N = (Cin - 2 * d_in - H) // 2
    z_in = zxbcdt.to(act)
    if z_in.stride(2) != 1:
        z_in = z_in.contiguous()
    w = Mamba2Weights(conv_weight=conv1d_weight.reshape(conv1d_weight.shape[0], -1).float().contiguous(),
                      conv_bias=None if conv1d_bias is None else conv1d_bias.float().contiguous(),
                      dt_bias=None if dt_bias is None else dt_bias.float().contiguous(),
                      A=A.float().contiguous(), D=D.float().contiguous())
    plan = ScanPlan(1, L, L, None, "stacked")
    with torch.autocast("cuda", enabled=False):
        v, ss = mamba2_ssd([z_in], [w], plan, d_in, N, H, gate=True, want_sumsq=rmsnorm_weight is not None)
        v = v[0][:, 0]                                                     # (B, L, d_in)
        if rmsnorm_weight is not None:
            rstd = torch.rsqrt(ss[0][:, 0] / d_in + rmsnorm_eps)           # (B, L)
            v = (v.float() * rstd.unsqueeze(-1) * rmsnorm_weight.float()).to(act)
        if outproj_weight is not None:
            v = F.linear(v, outproj_weight.to(act), None if outproj_bias is None else outproj_bias.to(act))
        return v


def mamba_chunk_scan_combined(*args, **kwargs):
    """[upstream mamba_chunk_scan_combined] -- imported at block/mamba2.py:20, reached only with
    ``use_mem_eff_path=False``, which DiffMa never sets."""
    raise NotImplementedError("diffma_b200: mamba_chunk_scan_combined is a dead path in DiffMa "
                              "(use_mem_eff_path is always True)")


class RMSNormGated(torch.nn.Module):
    """[upstream layernorm_gated.RMSNorm] parameter holder used at block/mamba2.py:347-350; its weight and
    eps are handed to ``mamba_split_conv1d_scan_combined``.  ``forward`` (never called on DiffMa's fused
    path) computes rmsnorm(x * silu(z)) * weight with torch ops on the device."""

    def __init__(self, hidden_size, eps=1e-5, norm_before_gate=True, group_size=None, device=None, dtype=None):
        super().__init__()
        self.eps = eps
        self.weight = torch.nn.Parameter(torch.ones(hidden_size, device=device, dtype=dtype))
        self.register_parameter("bias", None)
        self.group_size = group_size
        self.norm_before_gate = norm_before_gate

    def forward(self, x, z=None):
        _require_cuda(x, "RMSNormGated")
        xf = x.float()
        if z is not None and not self.norm_before_gate:
            xf = xf * F.silu(z.float())
        gs = self.group_size or xf.shape[-1]
        g = xf.reshape(*xf.shape[:-1], -1, gs)
        out = (g * torch.rsqrt(g.square().mean(-1, keepdim=True) + self.eps)).reshape(xf.shape) * self.weight.float()
        if z is not None and self.norm_before_gate:
            out = out * F.silu(z.float())
        return out.to(x.dtype)


# --------------------------------------------------------------------------------------------------
# row-wise glue of Spiral_MambaBlock.forward (reference block/mamba_block.py:100-115)
# --------------------------------------------------------------------------------------------------
def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda:
        raise RuntimeError(f"{what}: expected a contiguous CUDA fp32 tensor, got {t.dtype} on {t.device}")
    return t


def spiral_pre(x, skip, ln_weight, ln_bias, mod, w, act_dtype, eps: float = 1e-5) -> torch.Tensor:
    """x (B,L,D) fp32 [+ skip] -> (2, B*L, D) act dtype: [modulate(LN(x)) ; modulate(LN(x)) * w]."""
    B, L, D = x.shape
    out2 = torch.empty((2, B * L, D), dtype=act_dtype, device=x.device)
    st = _cabi.lib().dm_spiral_pre(_f32c(x, "x").data_ptr(), None if skip is None else _f32c(skip, "skip").data_ptr(),
                                   _f32c(ln_weight, "ln_weight").data_ptr(), _f32c(ln_bias, "ln_bias").data_ptr(),
                                   _mod2d(mod).data_ptr(), mod.stride(0),
                                   None if w is None else _f32c(w, "w").data_ptr(), out2.data_ptr(), B, L, D, eps,
                                   _dtype_code(out2), _stream_handle(x.device))
    _cabi.check(st, "dm_spiral_pre")
    LAUNCH_COUNTER["kernels"] += 1
    return out2


def spiral_post_ln(ab, ln_weight, ln_bias) -> torch.Tensor:
    """ab (2, rows, D) -> LayerNorm(cat(ab[0], ab[1])) (rows, 2D), same dtype."""
    _, rows, D = ab.shape
    out = torch.empty((rows, 2 * D), dtype=ab.dtype, device=ab.device)
    st = _cabi.lib().dm_spiral_post_ln(ab.data_ptr(), _f32c(ln_weight, "ln_weight").data_ptr(),
                                       _f32c(ln_bias, "ln_bias").data_ptr(), out.data_ptr(), rows, D, 1e-5,
                                       _dtype_code(ab), _stream_handle(ab.device))
    _cabi.check(st, "dm_spiral_post_ln")
    LAUNCH_COUNTER["kernels"] += 1
    return out


def spiral_post_mix(x, skip, ab, hidden, w3, b3, mod) -> torch.Tensor:
    """alpha = sigmoid(w3 . silu(hidden) + b3); (x + skip) + gate * (alpha*ab[0] + (1-alpha)*ab[1]) -> fp32 (B,L,D)."""
    B, L, D = x.shape
    out = torch.empty_like(x)
    if not (ab.is_contiguous() and hidden.is_contiguous() and hidden.dtype == ab.dtype):
        raise RuntimeError("spiral_post_mix: ab / hidden must be contiguous and of the same dtype")
    st = _cabi.lib().dm_spiral_post_mix(_f32c(x, "x").data_ptr(), None if skip is None else _f32c(skip, "skip").data_ptr(),
                                        ab.data_ptr(), hidden.data_ptr(), _f32c(w3, "w3").data_ptr(),
                                        _f32c(b3, "b3").data_ptr(), _mod2d(mod).data_ptr(), mod.stride(0),
                                        out.data_ptr(), B, L, D, _dtype_code(ab), _stream_handle(x.device))
    _cabi.check(st, "dm_spiral_post_mix")
    LAUNCH_COUNTER["kernels"] += 1
    return out


def spiral_post_mix_pre(x, skip, ab, hidden, w3, b3, mod, skip_next, ln_weight, ln_bias, mod_next, w, eps: float = 1e-5):
    """``spiral_post_mix`` of one block and ``spiral_pre`` of the next (or of the final layer) in one launch.
    Returns (x_new fp32 (B,L,D), out2 (2, B*L, D) in ab's dtype)."""
    B, L, D = x.shape
    x_out = torch.empty_like(x)
    out2 = torch.empty((2, B * L, D), dtype=ab.dtype, device=x.device)
    if not (ab.is_contiguous() and hidden.is_contiguous() and hidden.dtype == ab.dtype):
        raise RuntimeError("spiral_post_mix_pre: ab / hidden must be contiguous and of the same dtype")
    st = _cabi.lib().dm_spiral_post_mix_pre(
        _f32c(x, "x").data_ptr(), None if skip is None else _f32c(skip, "skip").data_ptr(), ab.data_ptr(),
        hidden.data_ptr(), _f32c(w3, "w3").data_ptr(), _f32c(b3, "b3").data_ptr(), _mod2d(mod).data_ptr(), mod.stride(0),
        x_out.data_ptr(), None if skip_next is None else _f32c(skip_next, "skip_next").data_ptr(),
        _f32c(ln_weight, "ln_weight").data_ptr(), _f32c(ln_bias, "ln_bias").data_ptr(), _mod2d(mod_next).data_ptr(),
        mod_next.stride(0), None if w is None else _f32c(w, "w").data_ptr(), out2.data_ptr(), B, L, D, eps,
        _dtype_code(ab), _stream_handle(x.device))
    _cabi.check(st, "dm_spiral_post_mix_pre")
    LAUNCH_COUNTER["kernels"] += 1
    return x_out, out2


def spiral_post_mix_fold(x, skip, ab, g2, colsum, cvec, ln2_eps, w3, b3, mod, pre=None):
    """``spiral_post_mix`` / ``spiral_post_mix_pre`` with the attention network's LayerNorm folded around its Linear
    (``dm_spiral_post_mix_fold``): ``g2`` (2, B*L, D) holds the raw products a W'_a^T, b W'_b^T (fp32 or ab's dtype), the
    kernel derives mean / rstd of cat(a, b) from the rows it reads anyway.  ``pre`` = None, or (skip_next, ln_weight,
    ln_bias, mod_next, w, eps) to open the next block in the same launch.  Returns x_new, or (x_new, out2)."""
    B, L, D = x.shape
    if not (ab.is_contiguous() and g2.is_contiguous() and tuple(g2.shape) == tuple(ab.shape) == (2, B * L, D)):
        raise RuntimeError("spiral_post_mix_fold: ab / g2 must be contiguous (2, B*L, D)")
    if g2.dtype not in (torch.float32, ab.dtype):
        raise TypeError("spiral_post_mix_fold: g2 must be fp32 or of ab's dtype")
    x_out = torch.empty_like(x)
    a = _cabi.SpiralFoldArgs()
    a.x, a.skip = _f32c(x, "x").data_ptr(), None if skip is None else _f32c(skip, "skip").data_ptr()
    a.ab, a.g2, a.g2_dtype, a.act_dtype = ab.data_ptr(), g2.data_ptr(), _dtype_code(g2), _dtype_code(ab)
    a.colsum, a.cvec = _f32c(colsum, "colsum").data_ptr(), _f32c(cvec, "cvec").data_ptr()
    a.w3, a.b3 = _f32c(w3, "w3").data_ptr(), _f32c(b3, "b3").data_ptr()
    a.mod, a.mod_batch_stride, a.x_out = _mod2d(mod).data_ptr(), mod.stride(0), x_out.data_ptr()
    a.batch, a.seqlen, a.d_model, a.ln2_eps = B, L, D, ln2_eps
    out2 = None
    if pre is not None:
        skip_next, ln_weight, ln_bias, mod_next, w, eps = pre
        out2 = torch.empty((2, B * L, D), dtype=ab.dtype, device=x.device)
        a.skip_next = None if skip_next is None else _f32c(skip_next, "skip_next").data_ptr()
        a.ln_weight, a.ln_bias = _f32c(ln_weight, "ln_weight").data_ptr(), _f32c(ln_bias, "ln_bias").data_ptr()
        a.mod_next, a.mod_next_batch_stride = _mod2d(mod_next).data_ptr(), mod_next.stride(0)
        a.w, a.out2, a.eps = None if w is None else _f32c(w, "w").data_ptr(), out2.data_ptr(), eps
    st = _cabi.lib().dm_spiral_post_mix_fold(C.byref(a), _stream_handle(x.device))
    _cabi.check(st, "dm_spiral_post_mix_fold")
    LAUNCH_COUNTER["kernels"] += 1
    return x_out if pre is None else (x_out, out2)


def step_head(x, patch_weight, pos_bias, patch: int, t, t_table, y, y2_mean, act):
    """Head of DiffMa.forward in one launch (``dm_step_head``): returns (h (B, L, D) fp32 = PatchEmbed(x) + pos_embed,
    silu_c (B, 2D) in ``act`` = silu(cat(t_table[t] + y, t_table[t] + mean_T(y2)))); ``y2_mean`` may be the pooled (B, D)
    tensor or the un-pooled (B, T, D) one."""
    _require_cuda(x, "step_head")
    B, Cc, H, Wd = x.shape
    D = pos_bias.shape[-1]
    if H != Wd or H % patch or t.dtype != torch.int64 or tuple(patch_weight.shape) != (Cc * patch * patch, D):
        raise RuntimeError("step_head: square images divisible by the patch, int64 timesteps and a (C*p*p, D) weight expected")
    L = (H // patch) ** 2
    y2_tokens = y2_mean.shape[1] if y2_mean.dim() == 3 else 1             # (B, T, D): the kernel takes the token mean itself
    if (tuple(pos_bias.shape) != (L, D) or tuple(y.shape) != (B, D) or y2_mean.shape[0] != B or y2_mean.shape[-1] != D
            or y2_mean.dim() not in (2, 3)):
        raise RuntimeError("step_head: pos_bias (L, D), y (B, D), y2 (B, D) or (B, T, D) expected")
    h = torch.empty((B, L, D), dtype=torch.float32, device=x.device)
    sc = torch.empty((B, 2 * D), dtype=act, device=x.device)
    st = _cabi.lib().dm_step_head(
        _f32c(x, "x").data_ptr(), _f32c(patch_weight, "patch_weight").data_ptr(), _f32c(pos_bias, "pos_bias").data_ptr(),
        h.data_ptr(), B, Cc, H, patch, t.contiguous().data_ptr(), _f32c(t_table, "t_table").data_ptr(), t_table.shape[0],
        _f32c(y, "y").data_ptr(), _f32c(y2_mean, "y2_mean").data_ptr(), y2_tokens, sc.data_ptr(), D, _dtype_code(sc),
        _stream_handle(x.device))
    _cabi.check(st, "dm_step_head")
    LAUNCH_COUNTER["kernels"] += 1
    return h, sc


def final_linear_unpatchify(hn, weight, bias, batch: int, patch: int, out_channels: int):
    """Tail of DiffMa.forward in one launch (``dm_final_linear_unpatchify``): hn (B*L, 512) bf16 -> (B, C_out, S, S) bf16.
    Returns None when the shape is outside the kernel's range (the caller then runs the GEMM and the permute)."""
    N = patch * patch * out_channels
    rows, D = hn.shape
    if (hn.dtype != torch.bfloat16 or D != 512 or N > 128 or not hn.is_contiguous() or weight.dtype != torch.bfloat16
            or not weight.is_contiguous() or bias.dtype != torch.bfloat16 or tuple(weight.shape) != (N, D) or rows % batch):
        return None
    L = rows // batch
    g = int(round(L ** 0.5))
    if g * g != L:
        return None
    out = torch.empty((batch, out_channels, g * patch, g * patch), dtype=torch.bfloat16, device=hn.device)
    st = _cabi.lib().dm_final_linear_unpatchify(hn.data_ptr(), weight.data_ptr(), bias.contiguous().data_ptr(), out.data_ptr(),
                                                batch, g, patch, out_channels, D, _cabi.DM_BF16, _stream_handle(hn.device))
    _cabi.check(st, "dm_final_linear_unpatchify")
    LAUNCH_COUNTER["kernels"] += 1
    return out


def _mod2d(mod: torch.Tensor) -> torch.Tensor:
    """adaLN output (B, 3D) fp32, rows may be strided (a slice of the all-blocks GEMM), channels contiguous."""
    if mod.dtype != torch.float32 or mod.dim() != 2 or mod.stride(1) != 1 or not mod.is_cuda or mod.stride(0) % 4:
        raise RuntimeError("mod: expected CUDA fp32 (B, 3D) with unit channel stride and 16-byte aligned rows")
    return mod


# --------------------------------------------------------------------------------------------------
# tcgen05 GEMM
# --------------------------------------------------------------------------------------------------
def gemm_bf16_tn(a: torch.Tensor, b: torch.Tensor, row_scale: Optional[torch.Tensor] = None,
                 bias: Optional[torch.Tensor] = None, silu_from: Optional[int] = None) -> torch.Tensor:
    """C[g] = silu_{cols >= silu_from}(row_scale[g][:, None] * (A[g] @ b[g].T) + bias[g]) on the hand-written tcgen05
    kernel (``dm_gemm_bf16_tn_ex``).  b (G, N, K) bf16, K contiguous.  a is (G, M, K) -- or (G, M, S, K) with S = 3:
    then A[g] = a[g].sum(-2) is formed in shared memory in front of the MMA (the CrossMerge direction sum as the
    out-projection's A producer).  row_scale (G, M) fp32, bias (G, N) fp32.  Returns (G, M, N) bf16."""
    _require_cuda(a, "gemm_bf16_tn")
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16 or a.dim() not in (3, 4) or b.dim() != 3:
        raise TypeError("gemm_bf16_tn: (G, M, [S,] K) x (G, N, K) bf16 operands expected")
    G, M, K = a.shape[0], a.shape[1], a.shape[-1]
    S = a.shape[2] if a.dim() == 4 else 1
    N = b.shape[1]
    if b.shape[0] != G or b.shape[2] != K or a.stride(-1) != 1 or b.stride(2) != 1:
        raise RuntimeError("gemm_bf16_tn: shape / stride mismatch")
    c = torch.empty((G, M, N), dtype=torch.bfloat16, device=a.device)
    g = _cabi.GemmArgs()
    g.A, g.a_group_stride, g.a_row_stride = a.data_ptr(), a.stride(0), a.stride(1)
    g.a_sum_stride, g.n_sum = (a.stride(2) if S > 1 else 0), S
    g.B, g.b_group_stride, g.b_row_stride = b.data_ptr(), b.stride(0), b.stride(1)
    g.C, g.c_group_stride, g.c_row_stride = c.data_ptr(), c.stride(0), c.stride(1)
    keep = []
    if row_scale is not None:
        rs = _f32c(row_scale, "row_scale")
        assert tuple(rs.shape) == (G, M)
        g.row_scale = rs.data_ptr()
        keep.append(rs)
    if bias is not None:
        bs = _f32c(bias, "bias")
        assert tuple(bs.shape) == (G, N)
        g.bias = bs.data_ptr()
        keep.append(bs)
    g.silu_from = N if silu_from is None else int(silu_from)
    g.groups, g.M, g.N, g.K = G, M, N, K
    st = _cabi.lib().dm_gemm_bf16_tn_ex(C.byref(g), C.c_void_p(_stream_handle(a.device)))
    _cabi.check(st, "dm_gemm_bf16_tn_ex")
    LAUNCH_COUNTER["kernels"] += 1
    return c


def p_sample_update(model_out, x, noise, table, t, clip_denoised=False, want_pred_xstart=True):
    """One reverse-diffusion update on ``dm_p_sample_update``.  model_out (N, 2C, H, W), x / noise (N, C, H, W) fp32
    contiguous, table (9, n_steps) fp32 (diffusion._ROWS order), t (N,) int64 -> (sample, pred_xstart | None)."""
    _require_cuda(x, "p_sample_update")
    N = x.shape[0]
    chw = x.numel() // N
    if tuple(model_out.shape) != (N, 2 * x.shape[1]) + tuple(x.shape[2:]) or t.dtype != torch.int64:
        raise RuntimeError("p_sample_update: model_out must be (N, 2C, H, W) and t int64")
    sample = torch.empty_like(x)
    pred = torch.empty_like(x) if want_pred_xstart else None
    st = _cabi.lib().dm_p_sample_update(_f32c(model_out, "model_out").data_ptr(), _f32c(x, "x").data_ptr(),
                                        _f32c(noise, "noise").data_ptr(), _f32c(table, "table").data_ptr(),
                                        t.contiguous().data_ptr(), sample.data_ptr(), _ptr(pred), N, chw, table.shape[1],
                                        int(bool(clip_denoised)), _stream_handle(x.device))
    _cabi.check(st, "dm_p_sample_update")
    LAUNCH_COUNTER["kernels"] += 1
    return sample, pred


# --------------------------------------------------------------------------------------------------
# adjoints of the row kernels (training path) and of the CrossScan gather
# --------------------------------------------------------------------------------------------------
def spiral_pre_bwd(x, skip, ln_weight, ln_bias, mod, w, d_out2, eps: float = 1e-5):
    """Adjoint of ``spiral_pre``: d_out2 (2, B*L, D) -> (dx (B,L,D) fp32, d_mod (B, 3D) fp32 [gate part zero],
    d_ln_weight (D), d_ln_bias (D))."""
    B, L, D = x.shape
    dx = torch.empty_like(x)
    small = torch.zeros(B * 3 * D + 2 * D, dtype=torch.float32, device=x.device)
    d_mod, d_lw, d_lb = small[:B * 3 * D].view(B, 3 * D), small[B * 3 * D:B * 3 * D + D], small[B * 3 * D + D:]
    if not d_out2.is_contiguous():
        d_out2 = d_out2.contiguous()
    st = _cabi.lib().dm_spiral_pre_bwd(
        _f32c(x, "x").data_ptr(), None if skip is None else _f32c(skip, "skip").data_ptr(),
        _f32c(ln_weight, "ln_weight").data_ptr(), _f32c(ln_bias, "ln_bias").data_ptr(), _mod2d(mod).data_ptr(), mod.stride(0),
        None if w is None else _f32c(w, "w").data_ptr(), d_out2.data_ptr(), dx.data_ptr(), d_mod.data_ptr(), 3 * D,
        d_lw.data_ptr(), d_lb.data_ptr(), B, L, D, eps, _dtype_code(d_out2), _stream_handle(x.device))
    _cabi.check(st, "dm_spiral_pre_bwd")
    LAUNCH_COUNTER["kernels"] += 1
    return dx, d_mod, d_lw, d_lb


def spiral_post_bwd(d_x_out, ab, lnab, hidden, att_w, ln2_weight, w3, b3, mod, B: int, L: int):
    """Adjoint of post_ln -> attention_network[1] -> post_mix.  d_x_out (B,L,D) fp32; ab (2, rows, D), lnab (rows, 2D),
    hidden (rows, D), att_w (D, 2D) in the act dtype.  Returns (d_ab (2, rows, D) act, d_mod (B, 3D) fp32 [gate part],
    d_ln2_weight (2D), d_ln2_bias (2D), d_att_w (D, 2D) act dtype, d_att_b (D) fp32, d_w3 (D), d_b3 (1))."""
    rows, D = hidden.shape
    dev = d_x_out.device
    d_ab = torch.empty_like(ab)
    d_hidden = torch.empty_like(hidden)
    small = torch.zeros(B * 3 * D + D + 4 + 4 * D, dtype=torch.float32, device=dev)
    o = B * 3 * D
    d_mod, d_w3, d_b3 = small[:o].view(B, 3 * D), small[o:o + D], small[o + D:o + D + 1]
    d_l2w, d_l2b = small[o + D + 4:o + D + 4 + 2 * D], small[o + D + 4 + 2 * D:]
    lib, st = _cabi.lib(), _stream_handle(dev)
    code = _dtype_code(ab)
    if not d_x_out.is_contiguous():
        d_x_out = d_x_out.contiguous()
    s = lib.dm_spiral_post_mix_bwd(_f32c(d_x_out, "d_x_out").data_ptr(), ab.data_ptr(), hidden.data_ptr(),
                                   _f32c(w3, "w3").data_ptr(), _f32c(b3, "b3").data_ptr(), _mod2d(mod).data_ptr(),
                                   mod.stride(0), d_ab.data_ptr(), d_hidden.data_ptr(), d_mod.data_ptr(), 3 * D,
                                   d_w3.data_ptr(), d_b3.data_ptr(), B, L, D, code, st)
    _cabi.check(s, "dm_spiral_post_mix_bwd")
    d_lnab = torch.mm(d_hidden, att_w)                                        # (rows, 2D) act dtype
    d_att_w = torch.mm(d_hidden.t(), lnab)                                    # (D, 2D) act dtype: the caller casts to the parameter's
    d_att_b = torch.sum(d_hidden, 0, dtype=torch.float32)
    s = lib.dm_spiral_post_ln_bwd(ab.data_ptr(), _f32c(ln2_weight, "ln2_weight").data_ptr(), d_lnab.data_ptr(),
                                  d_ab.data_ptr(), d_l2w.data_ptr(), d_l2b.data_ptr(), B, L, D, 1e-5, code, st)
    _cabi.check(s, "dm_spiral_post_ln_bwd")
    LAUNCH_COUNTER["kernels"] += 2
    return d_ab, d_mod, d_l2w, d_l2b, d_att_w, d_att_b, d_w3, d_b3


def merge_directions(g_scan: torch.Tensor, plan: ScanPlan, out_dtype) -> Optional[torch.Tensor]:
    """(G, B, K, L, C) fp32 gradients in scan order -> (G, B, L_src, C) in ``out_dtype``, summed over the K directions
    (adjoint of the CrossScan gather) in one kernel.  None if some direction is not a full permutation (caller falls
    back to index_add)."""
    inv = plan.inverse_table()
    if inv is None or not g_scan.is_cuda or g_scan.dtype != torch.float32 or g_scan.shape[-1] % 8:
        return None
    G, B, K, L, Cc = g_scan.shape
    idx = getattr(plan, "_flat_inv32", None)
    if idx is None:
        dev = g_scan.device
        cols = [(torch.arange(L, device=dev) if inv[k] is None else inv[k]) + k * L for k in range(K)]
        idx = torch.stack(cols, 1).reshape(-1).to(torch.int32).contiguous()    # (L_src * K): [l][k]
        plan._flat_inv32 = idx
    out = torch.empty((G, B, plan.src_len, Cc), dtype=out_dtype, device=g_scan.device)
    st = _cabi.lib().dm_merge_directions(g_scan.contiguous().data_ptr(), idx.data_ptr(), out.data_ptr(), G * B, plan.src_len,
                                         K, K * L, Cc, _dtype_code(out), _stream_handle(g_scan.device))
    _cabi.check(st, "dm_merge_directions")
    LAUNCH_COUNTER["kernels"] += 1
    return out
