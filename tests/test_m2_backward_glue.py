"""CPU check of the host-side glue of the Mamba-2 backward (``autograd_ops.mamba2_backward``): split, conv
recomputation + its backward, the Sigma v^2 path of the gated RMSNorm, per-head reductions of the per-channel S6
gradients and the scatter back to source tokens.  The CUDA reverse scan is replaced by the oracle's autograd
(``selective_scan_ref``), so what is verified here is exactly what the GPU test cannot localise."""
import pytest
import torch
import torch.nn.functional as F

from diffma_b200 import autograd_ops, ops
from oracle import ref_ops

D, N, H, W = 32, 16, 2, 4
P = D // H
CC = D + 2 * N


def _dv_scan(dv, plan):
    """gradient of v in plan layout -> (B, K, L, D) scan order."""
    if plan.layout == "stacked":
        return dv
    return torch.stack([autograd_ops.gather_scan_order(dv[:, :, k], plan)[:, k] for k in range(plan.n_dir)], 1)


def oracle_s6_backward(u, z_src, dt_raw, Bm, Cm, A_h, D_h, dtb_h, dv, plan, nheads):
    G, B, K, L, Dd = u.shape
    out = {k: [] for k in ("dz", "du", "ddelta", "dB", "dC", "dA", "dD", "ddtb")}
    head = torch.arange(Dd) // (Dd // nheads)
    for g in range(G):
        f = lambda t: t.detach().double().reshape(B * K, L, -1).transpose(1, 2).contiguous().requires_grad_(True)   # noqa: E731
        uu, bb, cc = f(u[g]), f(Bm[g]), f(Cm[g])
        dl = f(dt_raw[g][..., head])
        zz = f(autograd_ops.gather_scan_order(z_src[g], plan))
        A = A_h[g].double()[head].unsqueeze(1).expand(Dd, N).contiguous().requires_grad_(True)
        Dv = D_h[g].double()[head].contiguous().requires_grad_(True)
        db = dtb_h[g].double()[head].contiguous().requires_grad_(True)
        y = ref_ops.selective_scan_ref(uu, dl, A, bb, cc, Dv, z=zz, delta_bias=db, delta_softplus=True,
                                       compute_dtype=torch.float64)
        go = _dv_scan(dv[g].double(), plan).reshape(B * K, L, Dd).transpose(1, 2)
        gr = torch.autograd.grad(y, [zz, uu, dl, bb, cc, A, Dv, db], grad_outputs=go)
        r = lambda t: t.transpose(1, 2).reshape(B, K, L, -1).float()    # noqa: E731
        for key, val in zip(("dz", "du", "ddelta", "dB", "dC"), gr[:5]):
            out[key].append(r(val))
        out["dA"].append(gr[5].float()); out["dD"].append(gr[6].float()); out["ddtb"].append(gr[7].float())
    return {k: torch.stack(v) for k, v in out.items()}


def forward_ref(zx, w, plan):
    """Differentiable fp64 restatement of dm_mamba2_ssd_fwd in plan layout -> (v, sumsq)."""
    B, Ls, _ = zx.shape
    K, L = plan.n_dir, plan.seqlen
    xs = autograd_ops.gather_scan_order(zx, plan)                                   # (B,K,L,C)
    z, xbc, dt = xs[..., :D], xs[..., D:D + CC], xs[..., D + CC:]
    pre = F.conv1d(xbc.reshape(B * K, L, CC).transpose(1, 2), w["conv_weight"].unsqueeze(1), w["conv_bias"],
                   padding=W - 1, groups=CC)[..., :L]
    a = F.silu(pre).transpose(1, 2).reshape(B * K, L, CC)
    x, Bm, Cm = a[..., :D].reshape(B * K, L, H, P), a[..., D:D + N].reshape(B * K, L, 1, N), a[..., D + N:].reshape(B * K, L, 1, N)
    dtv = ref_ops._softplus(dt.reshape(B * K, L, H) + w["dt_bias"])
    y, _ = ref_ops.ssd_sequential_ref(x, dtv, w["A"], Bm, Cm, w["D"], compute_dtype=torch.float64)
    v = (y.reshape(B, K, L, D) * F.silu(z)).reshape(B, K, L, D)
    ss = v.square().sum(-1)                                                          # (B,K,L) scan order
    if plan.layout == "stacked":
        return v, ss
    inv = plan.inverse_table()
    vt = torch.stack([v[:, k] if inv[k] is None else v[:, k].index_select(1, inv[k]) for k in range(K)], 2)   # (B,Ls,K,D)
    st = torch.stack([ss[:, k] if inv[k] is None else ss[:, k].index_select(1, inv[k]) for k in range(K)], 1)  # (B,K,Ls)
    return vt, st


@pytest.mark.parametrize("layout,K", [("concat", 3), ("stacked", 2), ("concat", 1)])
def test_mamba2_backward_glue_matches_autograd(layout, K):
    with torch.enable_grad():       # other test modules switch autograd off globally at import
        _run_glue_case(layout, K)


def _run_glue_case(layout, K):
    torch.manual_seed(3)
    B, L, G = 2, 9, 2
    orders = [None] + [torch.randperm(L).tolist() for _ in range(K - 1)] if K > 1 else [torch.randperm(L).tolist()]
    plan = ops.ScanPlan.build(orders, L, layout, "cpu")
    zx, ws, refs = [], [], []
    for g in range(G):
        zx.append(torch.randn(B, L, 2 * D + 2 * N + H, dtype=torch.float64))
        ws.append(dict(conv_weight=torch.randn(CC, W, dtype=torch.float64) * 0.4,
                       conv_bias=torch.randn(CC, dtype=torch.float64) * 0.1,
                       dt_bias=torch.randn(H, dtype=torch.float64) * 0.3 - 1.0,
                       A=-torch.exp(torch.randn(H, dtype=torch.float64) * 0.5), D=torch.randn(H, dtype=torch.float64)))
    leaves, vs, sss = [], [], []
    for g in range(G):
        lz = zx[g].clone().requires_grad_(True)
        lw = {k: v.clone().requires_grad_(True) for k, v in ws[g].items()}
        v, ss = forward_ref(lz, lw, plan)
        leaves.append((lz, lw)); vs.append(v); sss.append(ss)
    v_all, ss_all = torch.stack(vs), torch.stack(sss)
    gv, gss = torch.randn_like(v_all), torch.randn_like(ss_all) * 0.1
    flat = [t for lz, lw in leaves for t in [lz] + [lw[k] for k in autograd_ops._W2]]
    want = torch.autograd.grad((v_all * gv).sum() + (ss_all * gss).sum(), flat)

    weights = [ops.Mamba2Weights(**{k: v.float() for k, v in ws[g].items()}) for g in range(G)]
    dzx, grads = autograd_ops.mamba2_backward([t.float() for t in zx], weights, plan, D, N, H, v_all.detach().float(),
                                              gv.float(), gss.float(), s6_backward=oracle_s6_backward)
    it = iter(want)
    for g in range(G):
        torch.testing.assert_close(dzx[g].double(), next(it), rtol=2e-4, atol=2e-4)
        for j, name in enumerate(autograd_ops._W2):
            torch.testing.assert_close(grads[g * len(autograd_ops._W2) + j].double(), next(it), rtol=2e-4, atol=2e-4,
                                       msg=lambda m, name=name: f"{name}: {m}")


@pytest.mark.parametrize("K,with_identity", [(3, True), (4, False), (1, False), (2, True)])
def test_batched_direction_merge_is_the_adjoint_of_the_scan_gather(K, with_identity):
    """autograd_ops.scan_to_token_sum_all (one gather + one sum for all groups, used by the Mamba-1 backward) equals the
    per-group scan_to_token_sum and is the adjoint of gather_scan_order: <gather(x), g> == <x, merge(g)>."""
    torch.manual_seed(K)
    G, B, L, Cc = 2, 3, 12, 5
    orders = [torch.randperm(L).tolist() for _ in range(K)]
    if with_identity:
        orders[0] = None
    plan = ops.ScanPlan.build(orders, L, "concat", "cpu")
    g_scan = torch.randn(G, B, K, L, Cc, dtype=torch.float64)
    merged = autograd_ops.scan_to_token_sum_all(g_scan, plan)
    assert merged.shape == (G, B, L, Cc)
    for g in range(G):
        torch.testing.assert_close(merged[g], autograd_ops.scan_to_token_sum(g_scan[g], plan))
    x = torch.randn(B, L, Cc, dtype=torch.float64)
    lhs = (autograd_ops.gather_scan_order(x, plan) * g_scan[0]).sum()
    rhs = (x * merged[0]).sum()
    torch.testing.assert_close(lhs, rhs)
    # a second call reuses the cached flat index
    torch.testing.assert_close(autograd_ops.scan_to_token_sum_all(g_scan, plan), merged)


def test_batched_direction_merge_partial_cover_falls_back():
    """EfficientVMamba-style plans (each direction covers a quarter of the tokens): no inverse permutation exists, the
    batched merge must fall back to the index_add path and still be the adjoint of the gather."""
    L = 16
    quarters = [list(range(i, L, 4)) for i in range(4)]
    plan = ops.ScanPlan.build(quarters, L, "disjoint", "cpu")
    assert plan.inverse_table() is None
    g_scan = torch.randn(2, 2, 4, L // 4, 3, dtype=torch.float64)
    merged = autograd_ops.scan_to_token_sum_all(g_scan, plan)
    x = torch.randn(2, L, 3, dtype=torch.float64)
    torch.testing.assert_close((autograd_ops.gather_scan_order(x, plan) * g_scan[1]).sum(), (x * merged[1]).sum())
