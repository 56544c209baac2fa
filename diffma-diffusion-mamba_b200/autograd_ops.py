"""autograd wrappers of the scan ops (training path).

``Mamba1ScanFn`` runs the same C-ABI forward as inference, keeps the intermediates (u, x_dbl) and implements upstream
``MambaInnerFn.backward`` (minus the out-projection, which stays a torch GEMM) on ``dm_mamba1_scan_bwd``: reverse scan
kernel -> four small library GEMMs (through x_proj / dt_proj) -> conv backward kernel -> un-permute + sum directions.
``Mamba2SsdFn`` differentiates ``dm_mamba2_ssd_fwd`` (upstream ``MambaSplitConv1dScanCombinedFn.backward`` minus the
RMSNorm scale and the out-projection, which stay torch ops) through the SAME reverse-scan kernel: the SSD recurrence
is the S6 recurrence with ``A[d, n] = A_head(d)`` and ``delta[d] = dt_head(d)`` (self-consistency (2)),
so the Mamba-2 operands are laid out as a Mamba-1 problem (dt as the dt_low slots + a one-hot dt_proj), the kernel
returns per-channel gradients and the host reduces them per head.  Correct-first: it spends 16 exps per
(token, channel) where a dedicated SSD backward needs one per (token, head).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi, ops

import os

# DIFFMA_M2_BWD=torch: differentiate the Mamba-2 conv / gathers with torch ops around the reverse-scan kernel (round-1
# path, kept for A/B runs and as the checker of the CUDA orchestration)
_M2_BWD_CUDA = os.environ.get("DIFFMA_M2_BWD", "cuda") != "torch"

_W1 = ("conv_weight", "conv_bias", "x_proj_weight", "dt_proj_weight", "dt_bias", "A", "D")
_W2 = ("conv_weight", "conv_bias", "dt_bias", "A", "D")


def flatten_weights(weights):
    fields = _W1 if isinstance(weights[0], ops.Mamba1Weights) else _W2
    return [getattr(w, f) for w in weights for f in fields]


def _det(t):
    return None if t is None else t.detach()


def gather_scan_order(t: torch.Tensor, plan) -> torch.Tensor:
    """(B, L_src, C) -> (B, K, L, C): row j of direction k = source token plan.table[k, j] (the CrossScan gather)."""
    rows = []
    for k in range(plan.n_dir):
        if plan.table is None or int(plan.table_host[k][0]) < 0:
            rows.append(t)
        else:
            rows.append(t.index_select(1, plan.table[k].long()))
    return torch.stack(rows, 1)


def scan_to_token_sum(g_scan: torch.Tensor, plan) -> torch.Tensor:
    """Adjoint of ``gather_scan_order``: (B, K, L, C) gradients in scan order -> (B, L_src, C), summed over directions."""
    B, K = g_scan.shape[:2]
    inv = plan.inverse_table()          # (K, L_src) long: source token -> scanned position, or None (partial cover)
    if inv is not None:                 # every direction is a full permutation: gather, no atomics
        acc = None
        for k in range(K):
            part = g_scan[:, k] if inv[k] is None else g_scan[:, k].index_select(1, inv[k])
            acc = part if acc is None else acc + part
        return acc
    acc = torch.zeros((B, plan.src_len, g_scan.shape[-1]), dtype=g_scan.dtype, device=g_scan.device)
    for k in range(K):
        if plan.table is None or int(plan.table_host[k][0]) < 0:
            acc += g_scan[:, k]
        else:
            acc.index_add_(1, plan.table[k].long(), g_scan[:, k])
    return acc


def scan_to_token_sum_all(g_scan: torch.Tensor, plan) -> torch.Tensor:
    """``scan_to_token_sum`` for all groups at once: (G, B, K, L, C) -> (G, B, L_src, C).  When every direction is a
    full permutation this is ONE gather (index (L_src * K) built once per plan: token l, direction k -> k * L +
    inverse_k[l]) and one sum over K, instead of per-group, per-direction index_selects and adds."""
    G, B, K, L, Cc = g_scan.shape
    inv = plan.inverse_table()
    if inv is None:
        return torch.stack([scan_to_token_sum(g_scan[g], plan) for g in range(G)])
    idx = getattr(plan, "_flat_inv", None)
    if idx is None:
        dev = g_scan.device
        cols = [(torch.arange(L, device=dev) if inv[k] is None else inv[k]) + k * L for k in range(K)]
        idx = torch.stack(cols, 1).reshape(-1)                 # (L_src * K): [l][k]
        plan._flat_inv = idx
    if K == 1:
        return g_scan.view(G * B, L, Cc).index_select(1, idx).view(G, B, L, Cc)
    return g_scan.view(G * B, K * L, Cc).index_select(1, idx).view(G, B, L, K, Cc).sum(3)


class Mamba1ScanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, G, *tensors):
        xz = [t.detach() for t in tensors[:G]]
        flat = tensors[G:]
        n = len(_W1)
        weights = [ops.Mamba1Weights(*[_det(v) for v in flat[g * n:(g + 1) * n]]) for g in range(G)]
        B, _, D2 = xz[0].shape
        # recurrence checkpoints written by the forward: the reverse-scan kernel then skips its own forward sweep
        states = torch.empty(ops.mamba1_state_shape(G, B, plan, D2 // 2, weights[0].A.shape[1]), dtype=torch.float32,
                             device=xz[0].device)
        out, u, x_dbl = ops.mamba1_scan_raw(xz, weights, plan, chunk_states=states)
        ctx.plan, ctx.G = plan, G
        ctx.none_mask = [t is None for t in flat]
        ctx.save_for_backward(*xz, *[t for t in flat if t is not None], u, x_dbl, states)
        return out

    @staticmethod
    def backward(ctx, dout):
        plan, G = ctx.plan, ctx.G
        saved = list(ctx.saved_tensors)
        xz = saved[:G]
        ws, x_dbl, u = saved.pop(), saved.pop(), saved.pop()
        it = iter(saved[G:])
        flat = [None if m else next(it) for m in ctx.none_mask]
        n = len(_W1)
        weights = [ops.Mamba1Weights(*flat[g * n:(g + 1) * n]) for g in range(G)]
        x0 = xz[0]
        dev = x0.device
        B, Lsrc, D2 = x0.shape
        D = D2 // 2
        K, L = plan.n_dir, plan.seqlen
        E = x_dbl.shape[-1]
        N = weights[0].A.shape[1]
        R = E - 2 * N
        dout = dout.to(x0.dtype)
        ostr = None
        if (plan.layout == "concat" and dout.dim() == 5 and dout.stride(4) == 1 and not dout.is_contiguous()
                and all(st % 8 == 0 for st in dout.stride()[:4]) and dout.data_ptr() % 16 == 0
                and (G == 1 or dout.stride(0) > 0)):
            # e.g. the gradient of a sum over the directions: an expanded view (direction stride 0) the kernel reads as is
            ostr = (dout.stride(1), dout.stride(3), dout.stride(2))                          # (batch, direction, token)
        else:
            dout = dout.contiguous()
        a, _ = ops.mamba1_args(xz, weights, plan, bufs=(dout, u, x_dbl), out_strides=ostr)
        f32 = dict(dtype=torch.float32, device=dev)
        W = weights[0].conv_weight.shape[1]
        d_xz_scan = torch.empty((G, B, K, L, 2 * D), **f32)
        du = torch.empty((G, B, K, L, D), **f32)
        ddelta = torch.empty((G, B, K, L, D), **f32)
        # every accumulated buffer in ONE zero-filled allocation (one memset instead of six)
        sizes = [G * B * K * L * E, G * D * N, G * D, G * D, G * D * W, G * D]
        acc = torch.zeros(sum(sizes), **f32)
        d_x_dbl, dA, dD, ddtb, dcw, dcb = (t.view(shape) for t, shape in zip(
            acc.split(sizes), [(G, B, K, L, E), (G, D, N), (G, D), (G, D), (G, D, W), (G, D)]))
        gr = (_cabi.Mamba1BwdGroup * G)()
        for g in range(G):
            w = weights[g]
            gr[g].states_valid = 1                  # ws = the forward's checkpoints
            gr[g].dout = dout[g].data_ptr()
            gr[g].d_xz_scan, gr[g].du, gr[g].ddelta = d_xz_scan[g].data_ptr(), du[g].data_ptr(), ddelta[g].data_ptr()
            gr[g].d_x_dbl, gr[g].dA = d_x_dbl[g].data_ptr(), dA[g].data_ptr()
            gr[g].dD = dD[g].data_ptr() if w.D is not None else None
            gr[g].d_dt_bias = ddtb[g].data_ptr() if w.dt_bias is not None else None
            gr[g].state_workspace = ws[g].data_ptr()
            gr[g].d_conv_weight = dcw[g].data_ptr()
            gr[g].d_conv_bias = dcb[g].data_ptr() if w.conv_bias is not None else None
        lib, st = _cabi.lib(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(lib.dm_mamba1_scan_bwd(C.byref(a), gr, 1, st), "dm_mamba1_scan_bwd(phase 1)")
        T = B * K * L
        # the four GEMM-shaped gradients through x_proj / dt_proj, batched over the groups.  bf16 activations: fp32
        # operands on the TF32 tensor path (10-bit mantissa >= the bf16 the forward used); fp32 activations: exact fp32.
        tf32_prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = x0.dtype != torch.float32
        try:
            Wdt = torch.stack([w.dt_proj_weight for w in weights]).float()                    # (G, D, R)
            Wx = torch.stack([w.x_proj_weight for w in weights]).float()                      # (G, E, D)
            dd = ddelta.view(G, T, D)
            dxd = d_x_dbl.view(G, T, E)
            dxd[:, :, :R] = torch.bmm(dd, Wdt)                                                # d dt_low
            halves = x_dbl.view(G, T, E)[:, :, :R].contiguous().view(torch.bfloat16).float()  # (G, T, 2R): [hi | lo]
            dt_low = halves[:, :, :R] + halves[:, :, R:]
            dWdt = torch.bmm(dd.transpose(1, 2), dt_low)                                      # (G, D, R)
            du.view(G, T, D).baddbmm_(dxd, Wx)                                                # du += d_x_dbl . W_x
            dWx = torch.bmm(dxd.transpose(1, 2), u.view(G, T, D).float())                     # (G, E, D)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32_prev
        _cabi.check(lib.dm_mamba1_scan_bwd(C.byref(a), gr, 2, st), "dm_mamba1_scan_bwd(phase 2)")
        ops.LAUNCH_COUNTER["kernels"] += 2
        dxz_all = ops.merge_directions(d_xz_scan, plan, x0.dtype)                             # (G, B, L_src, 2D), one kernel
        if dxz_all is None:
            dxz_all = scan_to_token_sum_all(d_xz_scan, plan).to(x0.dtype)
        dWx = dWx.to(weights[0].x_proj_weight.dtype)
        dWdt = dWdt.to(weights[0].dt_proj_weight.dtype)
        grads = []
        for g in range(G):
            w = weights[g]
            per = {"conv_weight": dcw[g], "conv_bias": dcb[g] if w.conv_bias is not None else None,
                   "x_proj_weight": dWx[g], "dt_proj_weight": dWdt[g],
                   "dt_bias": ddtb[g] if w.dt_bias is not None else None, "A": dA[g],
                   "D": dD[g] if w.D is not None else None}
            grads += [per[f] for f in _W1]
        return (None, None, *dxz_all.unbind(0), *grads)


def s6_backward_cuda(u, z_src, dt_raw, Bm, Cm, A_h, D_h, dtb_h, dv, plan, nheads):
    """Reverse scan of the SSD recurrence on ``dm_mamba1_scan_bwd`` (phase 1), all groups in one launch.

    u (G,B,K,L,D) act dtype = silu(conv(x)) in scan order; z_src[g] (B,L_src,D) act; dt_raw (G,B,K,L,H) fp32 (before
    bias / softplus); Bm, Cm (G,B,K,L,N) fp32; A_h / D_h / dtb_h [g] (H,) fp32 (D_h, dtb_h may be None); dv (G,) +
    plan.out_shape act dtype = gradient of the gated output.  Returns per-CHANNEL fp32 gradients in scan order:
    dz, du, ddelta (G,B,K,L,D); dB, dC (G,B,K,L,N); dA (G,D,N); dD, ddtb (G,D).
    """
    G, B, K, L, D = u.shape
    H, N = nheads, Bm.shape[-1]
    P, R = D // H, 32
    if H > R or N != 16:
        raise NotImplementedError("diffma_b200: Mamba-2 backward needs nheads <= 32 and d_state == 16 (all DiffMa uses)")
    dev, act = u.device, u.dtype
    f32 = dict(dtype=torch.float32, device=dev)
    # x_dbl rows as the Mamba-1 kernels read them: [dt_low hi: 32 bf16 | dt_low lo: 32 bf16 | B: 16 f32 | C: 16 f32]
    xd_bf = torch.zeros((G, B, K, L, 4 * R), dtype=torch.bfloat16, device=dev)
    hi = dt_raw.to(torch.bfloat16)
    xd_bf[..., :H] = hi
    xd_bf[..., R:R + H] = (dt_raw - hi.float()).to(torch.bfloat16)
    x_dbl = xd_bf.view(torch.float32)                                   # (G,B,K,L,64)
    x_dbl[..., R:R + N] = Bm
    x_dbl[..., R + N:] = Cm
    head = torch.arange(D, device=dev) // P
    onehot = (torch.arange(R, device=dev)[None, :] == head[:, None]).to(torch.float32)      # no host scalar: graph-capturable
    xz, weights = [], []
    for g in range(G):
        t = torch.zeros((B, plan.src_len, 2 * D), dtype=act, device=dev)
        t[..., D:] = z_src[g]
        xz.append(t)
        weights.append(ops.Mamba1Weights(
            conv_weight=torch.zeros((D, 4), **f32), conv_bias=None,
            x_proj_weight=torch.zeros((R + 2 * N, D), dtype=act, device=dev), dt_proj_weight=onehot.to(act),
            dt_bias=None if dtb_h[g] is None else dtb_h[g].float()[head].contiguous(),
            A=A_h[g].float()[head].unsqueeze(1).expand(D, N).contiguous(),
            D=None if D_h[g] is None else D_h[g].float()[head].contiguous()))
    a, _ = ops.mamba1_args(xz, weights, plan, bufs=(dv.contiguous(), u.contiguous(), x_dbl))
    ct = int(_cabi.lib().dm_mamba1_bwd_chunk_tokens())
    nch = (L + ct - 1) // ct
    d_xz_scan = torch.zeros((G, B, K, L, 2 * D), **f32)
    du = torch.empty((G, B, K, L, D), **f32)
    ddelta = torch.empty((G, B, K, L, D), **f32)
    d_x_dbl = torch.zeros((G, B, K, L, R + 2 * N), **f32)
    dA = torch.zeros((G, D, N), **f32)
    dD = torch.zeros((G, D), **f32)
    ddtb = torch.zeros((G, D), **f32)
    ws = torch.empty((G, B, K, nch, D, N), **f32)
    dcw = torch.zeros((G, D, 4), **f32)
    gr = (_cabi.Mamba1BwdGroup * G)()
    dvc = dv.contiguous()
    for g in range(G):
        gr[g].dout = dvc[g].data_ptr()
        gr[g].d_xz_scan, gr[g].du, gr[g].ddelta = d_xz_scan[g].data_ptr(), du[g].data_ptr(), ddelta[g].data_ptr()
        gr[g].d_x_dbl, gr[g].dA = d_x_dbl[g].data_ptr(), dA[g].data_ptr()
        gr[g].dD = dD[g].data_ptr() if D_h[g] is not None else None
        gr[g].d_dt_bias = ddtb[g].data_ptr() if dtb_h[g] is not None else None
        gr[g].state_workspace = ws[g].data_ptr()
        gr[g].d_conv_weight = dcw[g].data_ptr()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _cabi.check(_cabi.lib().dm_mamba1_scan_bwd(C.byref(a), gr, 1, st), "dm_mamba1_scan_bwd(phase 1, SSD operands)")
    ops.LAUNCH_COUNTER["kernels"] += 1
    if any(d is None for d in dtb_h):       # the kernel skipped the accumulation: d(dt_bias) is sum of d(delta_raw)
        ddtb = ddelta.sum(dim=(1, 2, 3))
    return dict(dz=d_xz_scan[..., D:], du=du, ddelta=ddelta, dB=d_x_dbl[..., R:R + N], dC=d_x_dbl[..., R + N:],
                dA=dA, dD=dD, ddtb=ddtb)


def mamba2_backward(zx, weights, plan, d_inner, d_state, nheads, v, gv, gss, s6_backward=s6_backward_cuda):
    """Gradients of ``ops.mamba2_ssd_raw`` w.r.t. zxbcdt and (conv_weight, conv_bias, dt_bias, A, D) of every group.

    Device-agnostic glue (split, conv recomputation and its backward on torch ops, per-head reductions, scatter back
    to source tokens) around ``s6_backward`` -- the CUDA reverse scan in the product; tests substitute the oracle's
    autograd to check the glue on CPU.  gv: gradient of v (plan.out_shape); gss: gradient of sumsq (G,B,K,rows) or None.
    """
    if plan.layout == "disjoint":
        raise NotImplementedError("diffma_b200: Mamba-2 backward for the EfficientVMamba split (broken in the reference "
                                  "itself, block/mamba2.py:704)")
    G = len(zx)
    D, N, H = d_inner, d_state, nheads
    P = D // H
    Cc = D + 2 * N
    act = zx[0].dtype
    # total gradient of v: direct + through sumsq = sum_c v^2 (the RMSNorm statistic the forward hands out)
    dv = gv.float()
    if gss is not None:
        g2 = gss.float()
        g2 = g2.transpose(2, 3).unsqueeze(-1) if plan.layout == "concat" else g2.unsqueeze(-1)   # -> v's row layout
        dv = dv + 2.0 * v.float() * g2
    dv = dv.to(act)
    # recompute the conv + SiLU in scan order with torch ops (fp32), keeping the graph for its backward
    xbc_scan, acts, u, Bm, Cm, dt_raw, cw, cb = [], [], [], [], [], [], [], []
    with torch.enable_grad():
        for g in range(G):
            w = weights[g]
            xs = gather_scan_order(zx[g][..., D:D + Cc], plan).float().requires_grad_(True)     # (B,K,L,Cc)
            wg = w.conv_weight.detach().float().requires_grad_(True)
            bg = None if w.conv_bias is None else w.conv_bias.detach().float().requires_grad_(True)
            B_, K, L, _ = xs.shape
            pre = torch.nn.functional.conv1d(xs.reshape(B_ * K, L, Cc).transpose(1, 2), wg.unsqueeze(1), bg,
                                             padding=wg.shape[1] - 1, groups=Cc)[..., :L]
            a = torch.nn.functional.silu(pre).transpose(1, 2).reshape(B_, K, L, Cc)
            xbc_scan.append(xs); acts.append(a); cw.append(wg); cb.append(bg)
            u.append(a[..., :D].detach().to(act))
            Bm.append(a[..., D:D + N].detach())
            Cm.append(a[..., D + N:].detach())
            dt_raw.append(gather_scan_order(zx[g][..., D + Cc:], plan).float())
    r = s6_backward(torch.stack(u), [zx[g][..., :D] for g in range(G)], torch.stack(dt_raw), torch.stack(Bm),
                    torch.stack(Cm), [w.A for w in weights], [w.D for w in weights], [w.dt_bias for w in weights],
                    dv, plan, H)
    dzx, grads = [], []
    for g in range(G):
        w = weights[g]
        d_act = torch.cat([r["du"][g], r["dB"][g], r["dC"][g]], dim=-1)
        wanted = [xbc_scan[g], cw[g]] + ([cb[g]] if cb[g] is not None else [])
        got = torch.autograd.grad(acts[g], wanted, grad_outputs=d_act)
        B_, K, L, _ = d_act.shape
        ddt = r["ddelta"][g].reshape(B_, K, L, H, P).sum(-1)
        d_scan = torch.cat([r["dz"][g], got[0], ddt], dim=-1)                                   # (B,K,L,2D+2N+H)
        dzx.append(scan_to_token_sum(d_scan, plan).to(act))
        per = {"conv_weight": got[1], "conv_bias": got[2] if cb[g] is not None else None,
               "dt_bias": None if w.dt_bias is None else r["ddtb"][g].reshape(H, P).sum(-1),
               "A": r["dA"][g].reshape(H, P * N).sum(-1),
               "D": None if w.D is None else r["dD"][g].reshape(H, P).sum(-1)}
        grads += [per[f] for f in _W2]
    return dzx, grads


def mamba2_backward_cuda(zx, weights, plan, d_inner, d_state, nheads, v, gv, gss):
    """``mamba2_backward`` on the C-ABI only (no torch conv / gather / cat): ``dm_mamba2_ssd_bwd`` phase 0 (operand
    preparation) -> ``dm_mamba1_scan_bwd`` phases 1 + 2 on the SSD operands (reverse scan, conv backward of x) ->
    ``dm_mamba2_ssd_bwd`` phase 2 (conv backward of B | C) -> per-head reductions -> ``dm_merge_directions_multi``
    ([dz | dx | dB dC | d dt] un-permuted and summed over the directions).  Same return value as ``mamba2_backward``.
    Returns None when the plan is not a set of full permutations (caller falls back to the torch glue)."""
    inv = plan.inverse_table()
    G = len(zx)
    D, N, H = d_inner, d_state, nheads
    if inv is None or H > 32 or N != 16 or D % 128 or plan.layout not in ("concat", "stacked"):
        return None
    P = D // H
    R, E = 32, 64
    Cc = D + 2 * N
    x0 = zx[0]
    dev, act = x0.device, x0.dtype
    B, Lsrc, Cin = x0.shape
    K, L = plan.n_dir, plan.seqlen
    f32 = dict(dtype=torch.float32, device=dev)
    es = x0.element_size()
    # total gradient of v: direct + through sumsq = sum_c v^2 (the RMSNorm statistic the forward hands out)
    dv = gv
    if gss is not None:
        g2s = gss.float()
        g2s = g2s.transpose(2, 3).unsqueeze(-1) if plan.layout == "concat" else g2s.unsqueeze(-1)   # -> v's row layout
        dv = torch.addcmul(gv.float(), v.float(), g2s, value=2.0)
    dv = dv.to(act).contiguous()                                              # plan.out_shape: (G,B,L_src,K,D) | (G,B,K,L,D)
    lib, st = _cabi.lib(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    # ---- phase 0: u, x_dbl ------------------------------------------------------------------------------------
    u = torch.empty((G, B, K, L, D), dtype=act, device=dev)
    x_dbl = torch.empty((G, B, K, L, E), **f32)
    a2 = _cabi.Mamba2Args()
    a2.batch, a2.n_dir, a2.seqlen = B, K, L
    a2.d_inner, a2.d_state, a2.nheads, a2.d_conv = D, N, H, weights[0].conv_weight.shape[1]
    a2.act_dtype, a2.out_order, a2.n_groups, a2.gate = ops._dtype_code(x0), plan.out_order, G, 1
    a2.order = ops._ptr(plan.table)
    ct = int(lib.dm_mamba1_bwd_chunk_tokens())
    nch = (L + ct - 1) // ct
    d_xz_scan = torch.empty((G, B, K, L, 2 * D), **f32)
    du = torch.empty((G, B, K, L, D), **f32)
    ddelta = torch.empty((G, B, K, L, D), **f32)
    ws = torch.empty((G, B, K, nch, D, N), **f32)
    d_bc = torch.empty((G, B, K, L, 2 * N), **f32)
    sizes = [G * B * K * L * E, G * D * N, G * D, G * D, G * Cc * 4, G * Cc]
    acc = torch.zeros(sum(sizes), **f32)                                      # every accumulated buffer: one memset
    d_x_dbl, dA, dD, ddtb, dcw, dcb = (t.view(shape) for t, shape in zip(
        acc.split(sizes), [(G, B, K, L, E), (G, D, N), (G, D), (G, D), (G, Cc, 4), (G, Cc)]))
    g2 = (_cabi.Mamba2BwdGroup * G)()
    bs, ts = ops._strides(x0)
    for g in range(G):
        x, w = zx[g], weights[g]
        if x.shape != x0.shape or x.dtype != act or x.stride(2) != 1 or ops._strides(x) != (bs, ts):
            return None
        gs = a2.group[g]
        gs.zxbcdt, gs.in_batch_stride, gs.in_token_stride = x.data_ptr(), bs, ts
        gs.conv_weight = w.conv_weight.data_ptr()
        gs.conv_bias = ops._ptr(w.conv_bias)
        g2[g].u, g2[g].x_dbl = u[g].data_ptr(), x_dbl[g].data_ptr()
        g2[g].d_x_dbl, g2[g].d_bc = d_x_dbl[g].data_ptr(), d_bc[g].data_ptr()
        g2[g].d_conv_weight = dcw[g].data_ptr()
        g2[g].d_conv_bias = dcb[g].data_ptr() if w.conv_bias is not None else None
    _cabi.check(lib.dm_mamba2_ssd_bwd(C.byref(a2), g2, 0, st), "dm_mamba2_ssd_bwd(phase 0)")
    # ---- the S6 view of the SSD operands: one-hot dt_proj, per-channel A / D / dt_bias (tiny, built on the device) ----
    head = torch.arange(D, device=dev) // P
    onehot = (torch.arange(R, device=dev)[None, :] == head[:, None]).to(act).contiguous()       # (D, 32)
    a1 = _cabi.Mamba1Args()
    a1.batch, a1.n_dir, a1.seqlen = B, K, L
    a1.d_inner, a1.d_state, a1.dt_rank, a1.d_conv = D, N, R, 4
    a1.act_dtype, a1.out_order, a1.n_groups = ops._dtype_code(x0), plan.out_order, G
    a1.order = ops._ptr(plan.table)
    obs, ods, ots = plan.out_strides(D)
    gr = (_cabi.Mamba1BwdGroup * G)()
    keep = []
    for g in range(G):
        w = weights[g]
        A_c = w.A.float()[head].unsqueeze(1).expand(D, N).contiguous()
        D_c = None if w.D is None else w.D.float()[head].contiguous()
        b_c = None if w.dt_bias is None else w.dt_bias.float()[head].contiguous()
        keep += [A_c, D_c, b_c]
        gs = a1.group[g]
        gs.xz = zx[g].data_ptr() - D * es                 # z of [z | x | ...] lands at the S6 layout's offset d_inner
        gs.xz_batch_stride, gs.xz_token_stride = bs, ts
        gs.out, gs.out_batch_stride, gs.out_dir_stride, gs.out_token_stride = dv[g].data_ptr(), obs, ods, ots
        gs.u, gs.x_dbl = u[g].data_ptr(), x_dbl[g].data_ptr()
        gs.conv_weight, gs.conv_bias = w.conv_weight.data_ptr(), ops._ptr(w.conv_bias)
        gs.x_proj_weight = onehot.data_ptr()              # unused by the backward kernels (validated non-null only)
        gs.dt_proj_weight = onehot.data_ptr()
        gs.dt_bias, gs.A, gs.D = ops._ptr(b_c), A_c.data_ptr(), ops._ptr(D_c)
        r = gr[g]
        r.dout = dv[g].data_ptr()
        r.d_xz_scan, r.du, r.ddelta = d_xz_scan[g].data_ptr(), du[g].data_ptr(), ddelta[g].data_ptr()
        r.d_x_dbl, r.dA = d_x_dbl[g].data_ptr(), dA[g].data_ptr()
        r.dD = dD[g].data_ptr() if w.D is not None else None
        r.d_dt_bias = ddtb[g].data_ptr() if w.dt_bias is not None else None
        r.state_workspace, r.states_valid = ws[g].data_ptr(), 0
        r.d_conv_weight = dcw[g].data_ptr()               # rows [0, D) of the (Cc, 4) buffer
        r.d_conv_bias = dcb[g].data_ptr() if w.conv_bias is not None else None
    _cabi.check(lib.dm_mamba1_scan_bwd(C.byref(a1), gr, 1, st), "dm_mamba1_scan_bwd(phase 1, SSD operands)")
    for g in range(G):
        a1.group[g].xz = zx[g].data_ptr() + D * es        # x of [z | x | ...] at offset 0 for the conv backward
    _cabi.check(lib.dm_mamba1_scan_bwd(C.byref(a1), gr, 2, st), "dm_mamba1_scan_bwd(phase 2, SSD operands)")
    _cabi.check(lib.dm_mamba2_ssd_bwd(C.byref(a2), g2, 2, st), "dm_mamba2_ssd_bwd(phase 2)")
    ops.LAUNCH_COUNTER["kernels"] += 4
    # ---- per-head reductions, merge of the directions ---------------------------------------------------------
    ddt = ddelta.view(G, B, K, L, H, P).sum(-1)                                # (G, B, K, L, H) d (raw dt)
    Hp = (H + 7) // 8 * 8
    if Hp != H:
        ddt = torch.nn.functional.pad(ddt, (0, Hp - H))
    ddt = ddt.contiguous()
    idx = getattr(plan, "_flat_inv32", None)
    if idx is None:
        cols = [(torch.arange(L, device=dev) if inv[k] is None else inv[k]) + k * L for k in range(K)]
        idx = torch.stack(cols, 1).reshape(-1).to(torch.int32).contiguous()
        plan._flat_inv32 = idx
    total = 2 * D + 2 * N + Hp
    out = torch.empty((G, B, Lsrc, total), dtype=act, device=dev)
    segs = (_cabi.MergeSegment * 4)()
    base = d_xz_scan.data_ptr()
    for i, (ptr, ch, rs) in enumerate(((base + D * 4, D, 2 * D), (base, D, 2 * D), (d_bc.data_ptr(), 2 * N, 2 * N),
                                       (ddt.data_ptr(), Hp, Hp))):
        segs[i].src, segs[i].channels, segs[i].row_stride = ptr, ch, rs
    _cabi.check(lib.dm_merge_directions_multi(segs, 4, idx.data_ptr(), out.data_ptr(), G * B, Lsrc, K, K * L,
                                              ops._dtype_code(out), st), "dm_merge_directions_multi")
    ops.LAUNCH_COUNTER["kernels"] += 1
    dzx = [out[g][..., :Cin] if total != Cin else out[g] for g in range(G)]
    grads = []
    for g in range(G):
        w = weights[g]
        if w.dt_bias is None:
            db_h = None
        else:
            db_h = ddtb[g].view(H, P).sum(-1)
        per = {"conv_weight": dcw[g], "conv_bias": dcb[g] if w.conv_bias is not None else None, "dt_bias": db_h,
               "A": dA[g].view(H, P * N).sum(-1), "D": None if w.D is None else dD[g].view(H, P).sum(-1)}
        grads += [per[f] for f in _W2]
    return dzx, grads


class Mamba2SsdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, G, d_inner, d_state, nheads, gate, want_sumsq, *tensors):
        zx = [t.detach() for t in tensors[:G]]
        flat = tensors[G:]
        n = len(_W2)
        weights = [ops.Mamba2Weights(*[_det(v) for v in flat[g * n:(g + 1) * n]]) for g in range(G)]
        v, ss = ops.mamba2_ssd_raw(zx, weights, plan, d_inner, d_state, nheads, gate, want_sumsq)
        ctx.plan, ctx.G, ctx.dims, ctx.gate, ctx.want_sumsq = plan, G, (d_inner, d_state, nheads), gate, want_sumsq
        ctx.none_mask = [t is None for t in flat]
        ctx.save_for_backward(*zx, *[t for t in flat if t is not None], v)
        if ss is None:
            ss = v.new_zeros(())
            ctx.mark_non_differentiable(ss)
        return v, ss

    @staticmethod
    def backward(ctx, gv, gss):
        if not ctx.gate:
            raise NotImplementedError("diffma_b200: Mamba-2 backward is built for the gated output (gate=True), the "
                                      "only form DiffMa uses")
        plan, G = ctx.plan, ctx.G
        saved = list(ctx.saved_tensors)
        v = saved.pop()
        zx = saved[:G]
        it = iter(saved[G:])
        flat = [None if m else next(it) for m in ctx.none_mask]
        n = len(_W2)
        weights = [ops.Mamba2Weights(*flat[g * n:(g + 1) * n]) for g in range(G)]
        d_inner, d_state, nheads = ctx.dims
        res = None
        if zx[0].is_cuda and _M2_BWD_CUDA:
            res = mamba2_backward_cuda(zx, weights, plan, d_inner, d_state, nheads, v, gv, gss if ctx.want_sumsq else None)
        if res is None:         # partial-cover plans (EfficientVMamba) and the CPU glue tests: torch ops around the scan kernel
            res = mamba2_backward(zx, weights, plan, d_inner, d_state, nheads, v, gv, gss if ctx.want_sumsq else None)
        dzx, grads = res
        return (None, None, None, None, None, None, None, *dzx, *grads)


# ------------------------------------------------------------------------------------------------------
# fused row kernels of the Spiral block in the training path (forward = the inference kernels of csrc/dm_block.cu,
# backward = their adjoints in csrc/dm_block_bwd.cu)
# ------------------------------------------------------------------------------------------------------
class SpiralPreFn(torch.autograd.Function):
    """x (B,L,D) fp32 [+ skip], norm1 weight / bias, mod (B, 3D) fp32, w (B*L) fp32 or None -> (2, B*L, D) act dtype:
    [modulate(LN(x + skip)) ; the same * w]   (reference block/mamba_block.py:101-105, model.py:290-292)."""

    @staticmethod
    def forward(ctx, x, skip, ln_w, ln_b, mod, w, act_dtype):
        x, mod = x.detach(), mod.detach()
        lw, lb = ln_w.detach().float().contiguous(), ln_b.detach().float().contiguous()
        sk = None if skip is None else skip.detach()
        out2 = ops.spiral_pre(x, sk, lw, lb, mod, w, act_dtype)
        ctx.save_for_backward(x, sk, lw, lb, mod, w)
        ctx.has_skip = skip is not None
        return out2

    @staticmethod
    def backward(ctx, d_out2):
        x, sk, lw, lb, mod, w = ctx.saved_tensors
        dx, d_mod, d_lw, d_lb = ops.spiral_pre_bwd(x, sk, lw, lb, mod, w, d_out2)
        return dx, (dx if ctx.has_skip else None), d_lw, d_lb, d_mod, None, None


class SpiralPostFn(torch.autograd.Function):
    """ab (2, B*L, D) act dtype -> x_out = (x + skip) + gate * (alpha a + (1 - alpha) b), alpha = sigmoid(Linear(SiLU(
    Linear(LN(cat(a, b))))))   (reference block/mamba_block.py:110-114): post_ln kernel, one GEMM, post_mix kernel."""

    @staticmethod
    def forward(ctx, x, skip, ab, ln2_w, ln2_b, att_w, att_b, w3, b3, mod):
        x, ab, mod = x.detach(), ab.detach(), mod.detach()
        sk = None if skip is None else skip.detach()
        act = ab.dtype
        B, L, D = x.shape
        l2w, l2b = ln2_w.detach().float().contiguous(), ln2_b.detach().float().contiguous()
        aw = att_w.detach().to(act).contiguous()
        w3f, b3f = w3.detach().float().reshape(-1).contiguous(), b3.detach().float().reshape(-1).contiguous()
        with torch.autocast("cuda", enabled=False):
            lnab = ops.spiral_post_ln(ab, l2w, l2b)
            hidden = torch.nn.functional.linear(lnab, aw, att_b.detach().to(act))
            x_out = ops.spiral_post_mix(x, sk, ab, hidden, w3f, b3f, mod)
        ctx.save_for_backward(ab, lnab, hidden, aw, l2w, w3f, b3f, mod)
        ctx.has_skip, ctx.BL, ctx.shapes = skip is not None, (B, L), (w3.shape, b3.shape)
        ctx.att_dtypes = (att_w.dtype, att_b.dtype)           # fp32 masters, or bf16 leaves (ddp.FlatTrainState lowp)
        return x_out

    @staticmethod
    def backward(ctx, d_x_out):
        ab, lnab, hidden, aw, l2w, w3f, b3f, mod = ctx.saved_tensors
        B, L = ctx.BL
        with torch.autocast("cuda", enabled=False):
            d_ab, d_mod, d_l2w, d_l2b, d_aw, d_att_b, d_w3, d_b3 = ops.spiral_post_bwd(d_x_out, ab, lnab, hidden, aw, l2w, w3f,
                                                                                     b3f, mod, B, L)
        d_aw, d_att_b = d_aw.to(ctx.att_dtypes[0]), d_att_b.to(ctx.att_dtypes[1])
        return (d_x_out, (d_x_out if ctx.has_skip else None), d_ab, d_l2w, d_l2b, d_aw, d_att_b, d_w3.view(ctx.shapes[0]),
                d_b3.view(ctx.shapes[1]), d_mod)
