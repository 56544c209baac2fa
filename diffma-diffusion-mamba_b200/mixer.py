"""``Mamba`` / ``Mamba2`` mixers with the reference's constructor, parameter names and ``forward(h, scan_type)``.

Second-level drop-in boundary of SURVEY.md section 8b: same ctor kwargs and state-dict keys as the reference's
``block/mamba.py:226-315`` and ``block/mamba2.py:234-357`` (so its checkpoints load), same result as
``Mamba.forward`` (block/mamba.py:317-403) / ``Mamba2.forward`` (block/mamba2.py:359-712), but B200-first:

* the scan-order tables are int32 device tensors built once per (scan_type, device) -- the reference
  re-uploads Python lists on every gather (block/mamba.py:27,30);
* no ``xs (B,K,2D,L)`` / ``out_m (B,K,L,d)`` is materialised: the kernels gather rows by the table while
  loading and write each direction's gated output straight to its un-permuted row;
* every direction of every mixer handed to ``mix_groups`` runs in ONE C-ABI call (a Spiral block's two
  mixers = one launch pair instead of ~90 launches);
* CrossMerge's sum over directions is absorbed by the out-projection: out = [y_0|y_1|y_2] . [W;W;W]^T.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


def _eff_orders(L: int):
    """Token lists of the four EfficientVMamba sub-scans (semantics of block/mamba.py:170-183)."""
    s = int(round(math.sqrt(L)))
    if s * s != L or s % 2:
        raise ValueError("EfficientVMamba cross-scan needs an even square token grid")
    g = torch.arange(L).reshape(s, s)
    return [g[::2, ::2].reshape(-1).tolist(), g.t()[::2, 1::2].reshape(-1).tolist(),
            g[::2, 1::2].reshape(-1).tolist(), g.t()[1::2, 1::2].reshape(-1).tolist()]


class _MixerBase(nn.Module):
    """Direction plans shared by both mixers."""

    def _init_orders(self, token_list, token_list_reversal, origina_list, origina_list_reversal):
        # plain attributes like the reference (not buffers, not in the state dict)
        self.token_list = token_list
        self.token_list_reversal = token_list_reversal
        self.origina_list = origina_list
        self.origina_list_reversal = origina_list_reversal
        self._plans = {}

    def plan(self, scan_type: str, L: int, device) -> ops.ScanPlan:
        key = (scan_type, L, str(device))
        p = self._plans.get(key)
        if p is None:
            if scan_type == "spiral":
                p = ops.ScanPlan.build([None, self.token_list, self.token_list_reversal], L, "concat", device)
            elif scan_type == "zigma":
                p = ops.ScanPlan.build([self.token_list], L, "concat", device)
            elif scan_type == "vmamba":
                p = ops.ScanPlan.build(list(self.token_list), L, "concat", device)
            elif scan_type == "eff":
                p = ops.ScanPlan.build(_eff_orders(L), L, "disjoint", device)
            elif scan_type == "vim":
                p = ops.ScanPlan.build([None, list(range(L - 1, -1, -1))], L, "stacked", device)
            else:
                raise ValueError(f"unknown scan_type {scan_type!r}")
            self._plans[key] = p
        return p


class Mamba(_MixerBase):
    """Mamba-1 mixer; ctor mirrors reference block/mamba.py:227-249."""

    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1,
                 dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True, bias=False,
                 use_fast_path=True, layer_idx=None, device=None, dtype=None, token_list=[],
                 token_list_reversal=[], origina_list=[], origina_list_reversal=[]):
        fk = {"device": device, "dtype": dtype}
        super().__init__()
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.use_fast_path, self.layer_idx = use_fast_path, layer_idx
        self.in_proj = nn.Linear(d_model, self.d_inner * 2, bias=bias, **fk)
        self.conv1d = nn.Conv1d(self.d_inner, self.d_inner, bias=conv_bias, kernel_size=d_conv,
                                groups=self.d_inner, padding=d_conv - 1, **fk)
        self._init_orders(token_list, token_list_reversal, origina_list, origina_list_reversal)
        self.activation = "silu"
        self.x_proj = nn.Linear(self.d_inner, self.dt_rank + d_state * 2, bias=False, **fk)
        self.dt_proj = nn.Linear(self.dt_rank, self.d_inner, bias=True, **fk)
        std = self.dt_rank ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(self.dt_proj.weight, std)
        elif dt_init == "random":
            nn.init.uniform_(self.dt_proj.weight, -std, std)
        else:
            raise NotImplementedError
        dt = torch.exp(torch.rand(self.d_inner, **fk) * (math.log(dt_max) - math.log(dt_min))
                       + math.log(dt_min)).clamp(min=dt_init_floor)
        with torch.no_grad():
            self.dt_proj.bias.copy_(dt + torch.log(-torch.expm1(-dt)))
        self.dt_proj.bias._no_reinit = True
        A = torch.arange(1, d_state + 1, dtype=torch.float32, device=device).repeat(self.d_inner, 1).contiguous()
        self.A_log = nn.Parameter(torch.log(A))
        self.A_log._no_weight_decay = True
        self.D = nn.Parameter(torch.ones(self.d_inner, device=device))
        self.D._no_weight_decay = True
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=bias, **fk)
        self._wcache = {}

    # ---- weights in the form the C-ABI wants ------------------------------------------------------
    def scan_weights(self, act_dtype) -> ops.Mamba1Weights:
        frozen = not torch.is_grad_enabled()
        key = ops.weights_key([p for p in (self.conv1d.weight, self.conv1d.bias, self.x_proj.weight, self.dt_proj.weight,
                                           self.dt_proj.bias, self.A_log, self.D) if p is not None], act_dtype)
        if frozen and self._wcache.get("key") == key:
            return self._wcache["w"]
        w = ops.Mamba1Weights(
            conv_weight=self.conv1d.weight.reshape(self.d_inner, self.d_conv).float().contiguous(),
            conv_bias=None if self.conv1d.bias is None else self.conv1d.bias.float().contiguous(),
            x_proj_weight=self.x_proj.weight.to(act_dtype).contiguous(),
            dt_proj_weight=self.dt_proj.weight.to(act_dtype).contiguous(),
            dt_bias=self.dt_proj.bias.float().contiguous(),
            A=(-torch.exp(self.A_log.float())).contiguous(), D=self.D.float().contiguous())
        if frozen:
            self._wcache = {"key": key, "w": w}
        return w

    def forward(self, hidden_states, scan_type, inference_params=None):
        """hidden_states (B, L, d_model) -> (B, L, d_model)   [reference block/mamba.py:317-403]"""
        if inference_params is not None:
            raise NotImplementedError("diffma_b200: single-token decode / inference cache is never used by DiffMa")
        return mix_groups([self], [hidden_states], scan_type)[0]

    def step(self, *a, **k):
        raise NotImplementedError("diffma_b200: Mamba.step (autoregressive decode) is out of scope for diffusion")


class Mamba2(_MixerBase):
    """Mamba-2 mixer; ctor mirrors reference block/mamba2.py:235-267 (ngroups=1, rmsnorm, no TP/SP)."""

    def __init__(self, d_model, d_state=128, d_conv=4, conv_init=None, expand=2, headdim=64, d_ssm=None, ngroups=1,
                 A_init_range=(1, 16), D_has_hdim=False, rmsnorm=True, norm_before_gate=False, dt_min=0.001,
                 dt_max=0.1, dt_init_floor=1e-4, dt_limit=(0.0, float("inf")), bias=False, conv_bias=True,
                 chunk_size=256, use_mem_eff_path=True, layer_idx=None, process_group=None,
                 sequence_parallel=True, device=None, dtype=None, token_list=[], token_list_reversal=[],
                 origina_list=[], origina_list_reversal=[]):
        fk = {"device": device, "dtype": dtype}
        super().__init__()
        if process_group is not None:
            raise NotImplementedError("diffma_b200: tensor/sequence parallel Mamba-2 is dead code in DiffMa")
        if ngroups != 1 or D_has_hdim or not rmsnorm or norm_before_gate or (d_ssm not in (None, int(expand * d_model))):
            raise NotImplementedError("diffma_b200: Mamba2 supports the configuration DiffMa uses "
                                      "(ngroups=1, per-head D, gated RMSNorm with norm_before_gate=False)")
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(expand * d_model)
        self.headdim, self.ngroups = headdim, ngroups
        self.d_ssm = self.d_inner
        assert self.d_ssm % headdim == 0
        self.nheads = self.d_ssm // headdim
        self.rmsnorm, self.norm_before_gate = rmsnorm, norm_before_gate
        self.dt_limit, self.chunk_size = dt_limit, chunk_size
        self.use_mem_eff_path, self.layer_idx = use_mem_eff_path, layer_idx
        self.activation = "silu"
        d_in_proj = 2 * self.d_inner + 2 * ngroups * d_state + self.nheads
        self.in_proj = nn.Linear(d_model, d_in_proj, bias=bias, **fk)
        conv_dim = self.d_ssm + 2 * ngroups * d_state
        self.conv1d = nn.Conv1d(conv_dim, conv_dim, bias=conv_bias, kernel_size=d_conv, groups=conv_dim,
                                padding=d_conv - 1, **fk)
        if conv_init is not None:
            nn.init.uniform_(self.conv1d.weight, -conv_init, conv_init)
        self._init_orders(token_list, token_list_reversal, origina_list, origina_list_reversal)
        dt = torch.exp(torch.rand(self.nheads, **fk) * (math.log(dt_max) - math.log(dt_min)) + math.log(dt_min))
        dt = torch.clamp(dt, min=dt_init_floor)
        self.dt_bias = nn.Parameter(dt + torch.log(-torch.expm1(-dt)))
        self.dt_bias._no_weight_decay = True
        A = torch.empty(self.nheads, dtype=torch.float32, device=device).uniform_(*A_init_range)
        self.A_log = nn.Parameter(torch.log(A).to(dtype=dtype))
        self.A_log._no_weight_decay = True
        self.D = nn.Parameter(torch.ones(self.nheads, device=device))
        self.D._no_weight_decay = True
        self.norm = ops.RMSNormGated(self.d_ssm, eps=1e-5, norm_before_gate=norm_before_gate,
                                     group_size=self.d_ssm // ngroups, **fk)
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=bias, **fk)
        self._wcache = {}

    def scan_weights(self) -> ops.Mamba2Weights:
        frozen = not torch.is_grad_enabled()
        key = ops.weights_key([p for p in (self.conv1d.weight, self.conv1d.bias, self.dt_bias, self.A_log, self.D)
                               if p is not None])
        if frozen and self._wcache.get("key") == key:
            return self._wcache["w"]
        cd = self.conv1d.weight.shape[0]
        w = ops.Mamba2Weights(
            conv_weight=self.conv1d.weight.reshape(cd, self.d_conv).float().contiguous(),
            conv_bias=None if self.conv1d.bias is None else self.conv1d.bias.float().contiguous(),
            dt_bias=self.dt_bias.float().contiguous(), A=(-torch.exp(self.A_log.float())).contiguous(),
            D=self.D.float().contiguous())
        if frozen:
            self._wcache = {"key": key, "w": w}
        return w

    def forward(self, u, scan_type, seqlen=None, seq_idx=None, inference_params=None):
        """u (B, L, d_model) -> (B, L, d_model)   [reference block/mamba2.py:359-712]"""
        if inference_params is not None or seq_idx is not None or seqlen is not None:
            raise NotImplementedError("diffma_b200: decode cache / seq_idx / flattened batches are never used by DiffMa")
        return mix_groups([self], [u], scan_type)[0]

    def step(self, *a, **k):
        raise NotImplementedError("diffma_b200: Mamba2.step (autoregressive decode) is out of scope for diffusion")


# ------------------------------------------------------------------------------------------------------
# the fused forward of one or several mixers of the same kind
# ------------------------------------------------------------------------------------------------------
_SUM_FIRST = os.environ.get("DIFFMA_SUM_FIRST", "1") != "0"     # training out-projection: sum the directions before the GEMM


def _act_dtype(x: torch.Tensor):
    if torch.is_autocast_enabled():
        dt = torch.get_autocast_dtype("cuda")
        if dt == torch.float16:
            raise TypeError("diffma_b200: fp16 autocast is not supported; use bfloat16 (DESIGN.md, precision)")
        return dt
    return x.dtype


def _merge_out_proj(y: torch.Tensor, plan: ops.ScanPlan, weight: torch.Tensor, bias, scan_type: str, mamba2: bool):
    """y: plan.out_shape of one group -> (B, L, d_model).  The direction sum rides on the GEMM's K axis."""
    B = y.shape[0]
    D = y.shape[-1]
    if plan.layout == "concat":
        K = plan.n_dir
        w = weight if K == 1 else weight.repeat(1, K)
        return F.linear(y.reshape(B, plan.src_len, K * D), w, bias if bias is None else bias * K)
    if plan.layout == "disjoint":
        return F.linear(y, weight, bias)
    # "stacked": only ViM -> (out_fwd + flip(out_bwd)) / 2
    o = F.linear(y, weight, bias)                                     # (B, 2, L, d_model)
    # Mamba-1 flips the FEATURE axis of the backward branch (reference block/mamba.py:366);
    # Mamba-2 flips the token axis (block/mamba2.py:522).  Bug-compatible on purpose.
    return (o[:, 0] + torch.flip(o[:, 1], [1 if mamba2 else 2])) / 2


def mix_groups(mixers: Sequence[nn.Module], inputs: Sequence[torch.Tensor], scan_type: str) -> List[torch.Tensor]:
    """Run ``mixers[g](inputs[g], scan_type)`` for all g with one scan launch.  All mixers must be the same class
    and size (a Spiral block's ``mamba1``/``mamba2`` pair: reference block/mamba_block.py:107-108)."""
    m0 = mixers[0]
    x0 = inputs[0]
    B, L, _ = x0.shape
    act = _act_dtype(x0)
    is_m2 = isinstance(m0, Mamba2)
    if is_m2 and scan_type == "eff":
        raise TypeError("EfficientVMamba + Mamba-2 is broken in the reference itself (block/mamba2.py:704, "
                        "SURVEY App. D#4); refusing to invent semantics")
    plan = m0.plan(scan_type, L, x0.device)
    with torch.autocast("cuda", enabled=False):
        proj = [F.linear(x.to(act), m.in_proj.weight.to(act), None if m.in_proj.bias is None else m.in_proj.bias.to(act))
                for m, x in zip(mixers, inputs)]
        outs = []
        if not is_m2:
            y = ops.mamba1_scan(proj, [m.scan_weights(act) for m in mixers], plan)
            if _SUM_FIRST and torch.is_grad_enabled() and plan.layout == "concat" and plan.n_dir > 1 and y.requires_grad:
                # training: sum the directions first (one kernel for all groups; its adjoint is an expanded view the
                # reverse-scan kernel reads with direction stride 0), then a K = d_inner projection per group -- instead of
                # a repeated (d_model, K * d_inner) weight per step, its summed gradient, and per-group slices of y whose
                # adjoints each zero-fill and add a full (G, B, L, K, D) tensor
                ys = y.sum(3).unbind(0)
                for g, m in enumerate(mixers):
                    bias = None if m.out_proj.bias is None else m.out_proj.bias.to(act) * plan.n_dir
                    outs.append(F.linear(ys[g], m.out_proj.weight.to(act), bias))
                return outs
            yg = y.unbind(0)
            for g, m in enumerate(mixers):
                bias = None if m.out_proj.bias is None else m.out_proj.bias.to(act)
                outs.append(_merge_out_proj(yg[g], plan, m.out_proj.weight.to(act), bias, scan_type, False))
            return outs
        v, ss = ops.mamba2_ssd(proj, [m.scan_weights() for m in mixers], plan, m0.d_inner, m0.d_state, m0.nheads,
                               gate=True, want_sumsq=True)
        for g, m in enumerate(mixers):
            # gated RMSNorm (block/mamba2.py:347-350,402): rstd per (token, direction) row, weight per channel
            rstd = torch.rsqrt(ss[g] / m.d_inner + m.norm.eps)                    # (B, K, rows)
            if plan.layout == "concat" and m.out_proj.bias is None:
                # rstd is a per-(token, direction) scalar and the norm weight a per-channel one, so
                #   sum_k rstd_k (v_k * w_norm) W^T = sum_k rstd_k * (v_k (W * w_norm)^T):
                # the GEMM consumes v as the kernel wrote it, the norm weight is folded into the (512 x 1024) projection
                # weight and rstd scales the 512-wide product -- 4x less elementwise traffic than normalising the
                # (B, L, K, 1024) tensor in fp32, forward and backward (profiles/r02_notes.md).
                Wn = (m.out_proj.weight.float() * m.norm.weight.float()[None, :]).to(act)
                o = F.linear(v[g], Wn)                                            # (B, L, K, d_model)
                outs.append((o * rstd.transpose(1, 2).unsqueeze(-1)).sum(2).to(act))
                continue
            if plan.layout == "concat":
                scale = rstd.transpose(1, 2).unsqueeze(-1)                        # (B, L, K, 1)
            elif plan.layout == "disjoint":
                scale = rstd.sum(1).unsqueeze(-1)
            else:
                scale = rstd.unsqueeze(-1)                                        # (B, K, L, 1)
            vn = (v[g].float() * scale * m.norm.weight.float()).to(act)
            bias = None if m.out_proj.bias is None else m.out_proj.bias.to(act)
            outs.append(_merge_out_proj(vn, plan, m.out_proj.weight.to(act), bias, scan_type, True))
        return outs
