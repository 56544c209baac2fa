"""autograd wrappers of the scan ops (training path).

``Mamba1ScanFn`` runs the same C-ABI forward as inference, keeps the intermediates (u, x_dbl) and implements upstream
``MambaInnerFn.backward`` (minus the out-projection, which stays a torch GEMM) on ``dm_mamba1_scan_bwd``: reverse scan
kernel -> four small library GEMMs (through x_proj / dt_proj) -> conv backward kernel -> un-permute + sum directions.
``Mamba2SsdFn`` has no backward kernel yet and raises instead of silently falling back.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi, ops

_W1 = ("conv_weight", "conv_bias", "x_proj_weight", "dt_proj_weight", "dt_bias", "A", "D")
_W2 = ("conv_weight", "conv_bias", "dt_bias", "A", "D")


def flatten_weights(weights):
    fields = _W1 if isinstance(weights[0], ops.Mamba1Weights) else _W2
    return [getattr(w, f) for w in weights for f in fields]


def _det(t):
    return None if t is None else t.detach()


class Mamba1ScanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, G, *tensors):
        xz = [t.detach() for t in tensors[:G]]
        flat = tensors[G:]
        n = len(_W1)
        weights = [ops.Mamba1Weights(*[_det(v) for v in flat[g * n:(g + 1) * n]]) for g in range(G)]
        out, u, x_dbl = ops.mamba1_scan_raw(xz, weights, plan)
        ctx.plan, ctx.G = plan, G
        ctx.none_mask = [t is None for t in flat]
        ctx.save_for_backward(*xz, *[t for t in flat if t is not None], u, x_dbl)
        return out

    @staticmethod
    def backward(ctx, dout):
        plan, G = ctx.plan, ctx.G
        saved = list(ctx.saved_tensors)
        xz = saved[:G]
        x_dbl, u = saved.pop(), saved.pop()
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
        dout = dout.to(x0.dtype).contiguous()
        a, _ = ops.mamba1_args(xz, weights, plan, bufs=(dout, u, x_dbl))
        f32 = dict(dtype=torch.float32, device=dev)
        nch = (L + 7) // 8
        d_xz_scan = torch.empty((G, B, K, L, 2 * D), **f32)
        du = torch.empty((G, B, K, L, D), **f32)
        ddelta = torch.empty((G, B, K, L, D), **f32)
        d_x_dbl = torch.zeros((G, B, K, L, E), **f32)
        dA = torch.zeros((G, D, N), **f32)
        dD = torch.zeros((G, D), **f32)
        ddtb = torch.zeros((G, D), **f32)
        dcw = torch.zeros((G, D, weights[0].conv_weight.shape[1]), **f32)
        dcb = torch.zeros((G, D), **f32)
        ws = torch.empty((G, B, K, nch, D, N), **f32)
        gr = (_cabi.Mamba1BwdGroup * G)()
        for g in range(G):
            w = weights[g]
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
        dWx, dWdt = [], []
        # the four GEMM-shaped gradients through x_proj / dt_proj.  bf16 activations: fp32 operands on the TF32 tensor
        # path (10-bit mantissa >= the bf16 the forward used); fp32 activations: exact fp32.
        tf32_prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = x0.dtype != torch.float32
        try:
            for g in range(G):
                w = weights[g]
                dd = ddelta[g].view(T, D)
                dxd = d_x_dbl[g].view(T, E)
                dxd[:, :R] = dd @ w.dt_proj_weight.float()                                   # d dt_low
                halves = x_dbl[g].view(T, E)[:, :R].contiguous().view(torch.bfloat16)        # (T, 2R): [hi | lo]
                dt_low = halves[:, :R].float() + halves[:, R:].float()
                dWdt.append(dd.t() @ dt_low)
                du[g].view(T, D).addmm_(dxd, w.x_proj_weight.float())                        # du += d_x_dbl . W_x
                dWx.append(dxd.t() @ u[g].view(T, D).float())
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32_prev
        _cabi.check(lib.dm_mamba1_scan_bwd(C.byref(a), gr, 2, st), "dm_mamba1_scan_bwd(phase 2)")
        ops.LAUNCH_COUNTER["kernels"] += 2
        # scan order -> source-token order, summed over directions (adjoint of the CrossScan gather)
        dxz = []
        inv = plan.inverse_table()          # (K, L_src) long: source token -> scanned position, or None (partial cover)
        for g in range(G):
            if inv is not None:             # every direction is a full permutation: gather, no atomics
                acc = None
                for k in range(K):
                    part = d_xz_scan[g][:, k] if inv[k] is None else d_xz_scan[g][:, k].index_select(1, inv[k])
                    acc = part if acc is None else acc + part
            else:
                acc = torch.zeros((B, Lsrc, 2 * D), **f32)
                for k in range(K):
                    if plan.table is None or int(plan.table_host[k][0]) < 0:
                        acc += d_xz_scan[g][:, k]
                    else:
                        acc.index_add_(1, plan.table[k].long(), d_xz_scan[g][:, k])
            dxz.append(acc.to(x0.dtype))
        grads = []
        for g in range(G):
            w = weights[g]
            per = {"conv_weight": dcw[g], "conv_bias": dcb[g] if w.conv_bias is not None else None,
                   "x_proj_weight": dWx[g].to(w.x_proj_weight.dtype), "dt_proj_weight": dWdt[g].to(w.dt_proj_weight.dtype),
                   "dt_bias": ddtb[g] if w.dt_bias is not None else None, "A": dA[g],
                   "D": dD[g] if w.D is not None else None}
            grads += [per[f] for f in _W1]
        return (None, None, *dxz, *grads)


class Mamba2SsdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, G, d_inner, d_state, nheads, gate, want_sumsq, *tensors):
        zx = [t.detach() for t in tensors[:G]]
        flat = tensors[G:]
        n = len(_W2)
        weights = [ops.Mamba2Weights(*[_det(v) for v in flat[g * n:(g + 1) * n]]) for g in range(G)]
        v, ss = ops.mamba2_ssd_raw(zx, weights, plan, d_inner, d_state, nheads, gate, want_sumsq)
        if ss is None:
            ss = v.new_zeros(())
        ctx.mark_non_differentiable(ss)
        return v, ss

    @staticmethod
    def backward(ctx, gv, gss):
        raise NotImplementedError("diffma_b200: the Mamba-2 backward kernel is not built yet; training with "
                                  "--use-mamba2 is a later milestone (DESIGN.md)")
