"""CPU restatement of the external ops DiffMa's mixers call (TEST INFRASTRUCTURE).

PARITY UNPINNED against the reference's own wheels (second-source pin below): the originals live in the un-vendored wheels ``mamba-ssm==2.0.4`` and
``causal-conv1d==1.2.2.post1`` (reference ``environment.yml:67,36``), imported by the
reference at ``block/mamba.py:11-23`` and ``block/mamba2.py:9-21``.  Each function
below restates the *published reference algorithm* of the named upstream function
 in plain torch on CPU; the reference call sites that fix the
argument meaning are cited per function.  Second source: ``tests/test_oracle_vllm_pin.py`` checks these
functions against outputs of vLLM 0.22's ports of the same upstream kernels recorded on a B200
(``tests/golden/vllm_*.npz``, generator ``tests/golden/make_golden_vllm.py``).

All functions compute in ``compute_dtype`` (fp32 by default, fp64 for tight checks)
and return tensors in the input dtype, like the upstream ``*_ref`` functions do.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F


def _softplus(x: torch.Tensor) -> torch.Tensor:
    # torch/upstream convention: identity above threshold 20
    return torch.where(x > 20.0, x, torch.log1p(torch.exp(torch.clamp(x, max=20.0))))


def causal_conv1d_ref(x, weight, bias=None, activation: Optional[str] = None,
                      compute_dtype=torch.float32):
    """[upstream causal_conv1d_interface.causal_conv1d_ref]  x (B,C,L), weight (C,W), bias (C).

    u[b,c,l] = act(bias[c] + sum_k weight[c,k] * x[b,c,l-(W-1)+k]),  x[...,j<0] = 0.
    Reached from the reference through ``mamba_inner_fn`` (block/mamba.py:346) and
    ``mamba_split_conv1d_scan_combined`` (block/mamba2.py:392).
    """
    assert activation in (None, "silu", "swish")
    dtype_in = x.dtype
    xc = x.to(compute_dtype)
    C, W = weight.shape
    L = x.shape[-1]
    out = F.conv1d(xc, weight.to(compute_dtype).unsqueeze(1),
                   None if bias is None else bias.to(compute_dtype), padding=W - 1, groups=C)
    out = out[..., :L]
    if activation is not None:
        out = F.silu(out)
    return out.to(dtype_in)


def selective_scan_ref(u, delta, A, B, C, D=None, z=None, delta_bias=None,
                       delta_softplus=False, return_last_state=False,
                       compute_dtype=torch.float32):
    """[upstream selective_scan_interface.selective_scan_ref]

    u, delta, z: (B,D,L); A: (D,N); B, C: (B,N,L) (input dependent, one group); D: (D,).
    h_l = exp(delta_l A) h_{l-1} + delta_l B_l u_l ;  y_l = <h_l, C_l> + D u_l ; out = y silu(z).
    Sequential over L on purpose (it is the definition, not an optimisation).
    """
    dtype_in = u.dtype
    u_ = u.to(compute_dtype)
    delta_ = delta.to(compute_dtype)
    if delta_bias is not None:
        delta_ = delta_ + delta_bias.to(compute_dtype)[None, :, None]
    if delta_softplus:
        delta_ = _softplus(delta_)
    A_ = A.to(compute_dtype)
    B_ = B.to(compute_dtype)
    C_ = C.to(compute_dtype)
    bsz, dim, L = u_.shape
    N = A_.shape[1]
    h = torch.zeros(bsz, dim, N, dtype=compute_dtype)
    ys = []
    dA = torch.exp(delta_[..., None] * A_[None, :, None, :])            # (B,D,L,N)
    dBu = (delta_ * u_)[..., None] * B_.transpose(1, 2)[:, None, :, :]  # (B,D,L,N)
    for l in range(L):
        h = dA[:, :, l] * h + dBu[:, :, l]
        ys.append(torch.einsum("bdn,bn->bd", h, C_[:, :, l]))
    y = torch.stack(ys, dim=2)
    out = y if D is None else y + u_ * D.to(compute_dtype)[None, :, None]
    if z is not None:
        out = out * F.silu(z.to(compute_dtype))
    out = out.to(dtype_in)
    return (out, h) if return_last_state else out


def mamba_inner_ref(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                    out_proj_weight, out_proj_bias, A, B=None, C=None, D=None,
                    delta_bias=None, B_proj_bias=None, C_proj_bias=None,
                    delta_softplus=True, compute_dtype=torch.float32):
    """[upstream selective_scan_interface.mamba_inner_ref]; call sites block/mamba.py:346-393.

    xz (B,2D,L); conv1d_weight (D,1,W); x_proj_weight (R+2N,D); delta_proj_weight (D,R);
    out_proj_weight (d_model,D); A (D,N).  Returns (B,L,d_model).
    """
    assert B is None and C is None and B_proj_bias is None and C_proj_bias is None, \
        "DiffMa always uses input-dependent B, C without projection biases"
    dtype_in = xz.dtype
    L = xz.shape[-1]
    R = delta_proj_weight.shape[1]
    N = A.shape[-1]
    x, z = xz.to(compute_dtype).chunk(2, dim=1)
    u = causal_conv1d_ref(x, conv1d_weight.reshape(conv1d_weight.shape[0], -1), conv1d_bias,
                          activation="silu", compute_dtype=compute_dtype)
    x_dbl = torch.einsum("bdl,ed->ble", u, x_proj_weight.to(compute_dtype))        # (B,L,R+2N)
    delta = torch.einsum("blr,dr->bdl", x_dbl[..., :R], delta_proj_weight.to(compute_dtype))
    Bm = x_dbl[..., R:R + N].transpose(1, 2)
    Cm = x_dbl[..., R + N:].transpose(1, 2)
    y = selective_scan_ref(u, delta, A, Bm, Cm, D, z=z, delta_bias=delta_bias,
                           delta_softplus=delta_softplus, compute_dtype=compute_dtype)
    out = torch.einsum("bdl,ed->ble", y, out_proj_weight.to(compute_dtype))
    if out_proj_bias is not None:
        out = out + out_proj_bias.to(compute_dtype)
    return out.to(dtype_in)


def rmsnorm_gated_ref(x, weight, z=None, eps=1e-5, group_size=None, norm_before_gate=False,
                      compute_dtype=torch.float32):
    """[upstream layernorm_gated.rms_norm_ref]; used as block/mamba2.py:349,402 (norm_before_gate=False)."""
    dtype_in = x.dtype
    x_ = x.to(compute_dtype)
    if z is not None and not norm_before_gate:
        x_ = x_ * F.silu(z.to(compute_dtype))
    if group_size is None or group_size == x_.shape[-1]:
        rstd = torch.rsqrt(x_.square().mean(dim=-1, keepdim=True) + eps)
        out = x_ * rstd * weight.to(compute_dtype)
    else:
        g = x_.reshape(*x_.shape[:-1], -1, group_size)
        rstd = torch.rsqrt(g.square().mean(dim=-1, keepdim=True) + eps)
        out = (g * rstd).reshape(x_.shape) * weight.to(compute_dtype)
    if z is not None and norm_before_gate:
        out = out * F.silu(z.to(compute_dtype))
    return out.to(dtype_in)


def ssd_sequential_ref(x, dt, A, Bm, Cm, D=None, compute_dtype=torch.float32):
    """Mamba-2 state recurrence, one token at a time (the definition).

    x (B,L,H,P); dt (B,L,H) already softplus'ed; A (H,); Bm, Cm (B,L,G,N) with H % G == 0; D (H,) or (H,P).
    S_t = exp(dt_t A_h) S_{t-1} + dt_t x_t (x) B_t ;  y_t = S_t C_t + D_h x_t.
    """
    bsz, L, H, P = x.shape
    G, N = Bm.shape[2], Bm.shape[3]
    x_ = x.to(compute_dtype)
    dt_ = dt.to(compute_dtype)
    A_ = A.to(compute_dtype)
    rep = H // G
    B_ = Bm.to(compute_dtype).repeat_interleave(rep, dim=2)   # (B,L,H,N)
    C_ = Cm.to(compute_dtype).repeat_interleave(rep, dim=2)
    S = torch.zeros(bsz, H, P, N, dtype=compute_dtype)
    ys = []
    for t in range(L):
        decay = torch.exp(dt_[:, t] * A_[None, :])                               # (B,H)
        S = decay[..., None, None] * S + (dt_[:, t, :, None] * x_[:, t])[..., None] * B_[:, t, :, None, :]
        ys.append(torch.einsum("bhpn,bhn->bhp", S, C_[:, t]))
    y = torch.stack(ys, dim=1)
    if D is not None:
        D_ = D.to(compute_dtype)
        y = y + x_ * (D_[None, None, :, None] if D_.dim() == 1 else D_[None, None])
    return y, S


def ssd_chunked_ref(x, dt, A, Bm, Cm, chunk_size, D=None, compute_dtype=torch.float32):
    """Chunked (SSD / 'ssd_minimal_discrete'-style) evaluation of the same recurrence.

    Used only as an independent formulation for self-consistency tests of ``ssd_sequential_ref``
    and as the tiling the CUDA kernel follows (intra-chunk masked-decay product + inter-chunk
    state passing).  Must equal ``ssd_sequential_ref`` up to rounding for any chunk size.
    """
    bsz, L, H, P = x.shape
    G, N = Bm.shape[2], Bm.shape[3]
    rep = H // G
    pad = (-L) % chunk_size
    x_ = F.pad(x.to(compute_dtype), (0, 0, 0, 0, 0, pad))
    dt_ = F.pad(dt.to(compute_dtype), (0, 0, 0, pad))
    B_ = F.pad(Bm.to(compute_dtype).repeat_interleave(rep, dim=2), (0, 0, 0, 0, 0, pad))
    C_ = F.pad(Cm.to(compute_dtype).repeat_interleave(rep, dim=2), (0, 0, 0, 0, 0, pad))
    A_ = A.to(compute_dtype)
    nC = (L + pad) // chunk_size
    Q = chunk_size
    x_ = x_.reshape(bsz, nC, Q, H, P)
    dt_ = dt_.reshape(bsz, nC, Q, H)
    B_ = B_.reshape(bsz, nC, Q, H, N)
    C_ = C_.reshape(bsz, nC, Q, H, N)
    a = dt_ * A_[None, None, None, :]                       # log-decay per step
    acs = torch.cumsum(a, dim=2)                            # (B,nC,Q,H) inclusive
    # intra-chunk: y_i += sum_{j<=i} C_i.B_j exp(acs_i-acs_j) dt_j x_j
    CB = torch.einsum("bcihn,bcjhn->bchij", C_, B_)
    seg = acs.permute(0, 1, 3, 2)[..., :, None] - acs.permute(0, 1, 3, 2)[..., None, :]   # (B,nC,H,i,j)
    mask = torch.tril(torch.ones(Q, Q, dtype=torch.bool))
    Lmat = torch.where(mask, torch.exp(torch.where(mask, seg, torch.zeros_like(seg))), torch.zeros_like(seg))
    y_intra = torch.einsum("bchij,bcjh,bcjhp->bcihp", CB * Lmat, dt_, x_)
    # chunk states (state at end of chunk from that chunk's inputs only)
    decay_to_end = torch.exp(acs[:, :, -1:, :] - acs)       # (B,nC,Q,H)
    states = torch.einsum("bcjh,bcjh,bcjhp,bcjhn->bchpn", decay_to_end, dt_, x_, B_)
    # inter-chunk passing
    S = torch.zeros(bsz, H, P, N, dtype=compute_dtype)
    y_inter = []
    for c in range(nC):
        y_inter.append(torch.einsum("bihn,bhpn,bih->bihp", C_[:, c], S, torch.exp(acs[:, c])))
        S = torch.exp(acs[:, c, -1])[..., None, None] * S + states[:, c]
    y = (y_intra + torch.stack(y_inter, dim=1)).reshape(bsz, nC * Q, H, P)[:, :L]
    if D is not None:
        D_ = D.to(compute_dtype)
        y = y + x.to(compute_dtype) * (D_[None, None, :, None] if D_.dim() == 1 else D_[None, None])
    return y, S


def mamba_chunk_scan_combined_ref(x, dt, A, B, C, chunk_size, D=None, z=None, dt_bias=None,
                                  dt_softplus=False, dt_limit=(0.0, float("inf")),
                                  compute_dtype=torch.float32):
    """[upstream ssd_combined.mamba_chunk_scan_combined semantics] imported at block/mamba2.py:20.

    x (B,L,H,P); dt (B,L,H); A (H,); B, C (B,L,G,N); z (B,L,H,P) optional gate (y * silu(z)).
    """
    dtype_in = x.dtype
    dt_ = dt.to(compute_dtype)
    if dt_bias is not None:
        dt_ = dt_ + dt_bias.to(compute_dtype)
    if dt_softplus:
        dt_ = _softplus(dt_)
    if dt_limit != (0.0, float("inf")):
        dt_ = dt_.clamp(min=dt_limit[0], max=dt_limit[1])
    y, _ = ssd_sequential_ref(x, dt_, A, B, C, D, compute_dtype=compute_dtype)
    if z is not None:
        y = y * F.silu(z.to(compute_dtype))
    return y.to(dtype_in)


def mamba_split_conv1d_scan_ref(zxbcdt, conv1d_weight, conv1d_bias, dt_bias, A, D, chunk_size,
                                dt_limit=(0.0, float("inf")), activation="silu",
                                rmsnorm_weight=None, rmsnorm_eps=1e-6, outproj_weight=None,
                                outproj_bias=None, headdim=None, ngroups=1, norm_before_gate=True,
                                compute_dtype=torch.float32):
    """[upstream ssd_combined.mamba_split_conv1d_scan_ref]; call sites block/mamba2.py:392-696.

    zxbcdt (B,L,2*d_in + 2*G*N + H) ordered [z | x | B | C | dt] (block/mamba2.py:300-301);
    conv1d_weight (d_in+2GN, W).  Returns (B,L,d_model) (or (B,L,d_in) without out-proj).
    """
    assert activation in ("silu", "swish")
    dtype_in = zxbcdt.dtype
    if D.dim() == 1:
        assert headdim is not None
        H = D.shape[0]
    else:
        H, headdim = D.shape
    bsz, L, _ = zxbcdt.shape
    d_in = H * headdim
    N = (zxbcdt.shape[-1] - 2 * d_in - H) // ngroups // 2
    z, xBC, dt = torch.split(zxbcdt.to(compute_dtype), [d_in, d_in + 2 * ngroups * N, H], dim=-1)
    xBC = causal_conv1d_ref(xBC.transpose(1, 2), conv1d_weight, conv1d_bias, activation="silu",
                            compute_dtype=compute_dtype).transpose(1, 2)
    x, Bm, Cm = torch.split(xBC, [d_in, ngroups * N, ngroups * N], dim=-1)
    x = x.reshape(bsz, L, H, headdim)
    Bm = Bm.reshape(bsz, L, ngroups, N)
    Cm = Cm.reshape(bsz, L, ngroups, N)
    zz = z.reshape(bsz, L, H, headdim)
    y = mamba_chunk_scan_combined_ref(x, dt, A, Bm, Cm, chunk_size, D=D,
                                      z=zz if rmsnorm_weight is None else None,
                                      dt_bias=dt_bias, dt_softplus=True, dt_limit=dt_limit,
                                      compute_dtype=compute_dtype)
    y = y.reshape(bsz, L, d_in)
    if rmsnorm_weight is not None:
        y = rmsnorm_gated_ref(y, rmsnorm_weight, z=z, eps=rmsnorm_eps,
                              group_size=d_in // ngroups, norm_before_gate=norm_before_gate,
                              compute_dtype=compute_dtype)
    if outproj_weight is not None:
        y = y.to(compute_dtype) @ outproj_weight.to(compute_dtype).t()
        if outproj_bias is not None:
            y = y + outproj_bias.to(compute_dtype)
    return y.to(dtype_in)
