"""GPU gradient parity (-m gpu): dm_mamba1_scan_bwd through autograd vs autograd of the CPU oracle.

Tolerance: upstream's own backward tests use rtol/atol 1e-3 for weight gradients in fp32 (widened to the activation
tolerance when z is present); here every gradient must agree with the fp32 oracle to 2e-3 of the gradient's max norm.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from diffma_b200 import _cabi
    _cabi.lib()
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


def _params(D, dm, seed, N=16, R=32):
    g = torch.Generator().manual_seed(seed)
    return dict(
        conv_w=torch.randn(D, 1, 4, generator=g) * 0.4, conv_b=torch.randn(D, generator=g) * 0.1,
        x_proj=torch.randn(R + 2 * N, D, generator=g) / D ** 0.5, dt_proj=torch.randn(D, R, generator=g) / R ** 0.5,
        out_proj=torch.randn(dm, D, generator=g) / D ** 0.5,
        A=-torch.exp(torch.log(torch.arange(1, N + 1).float()).expand(D, N) + 0.3 * torch.randn(D, N, generator=g)),
        D=1 + 0.1 * torch.randn(D, generator=g), dt_bias=torch.randn(D, generator=g) - 2.0)


def _relerr(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


@pytest.mark.parametrize("B,L,D", [(2, 21, 128), (1, 8, 256), (2, 40, 128)])
def test_mamba_inner_fn_grads_fp32(dev, B, L, D):
    from diffma_b200 import ops
    from oracle import ref_ops
    torch.set_grad_enabled(True)
    g = torch.Generator().manual_seed(L)
    xz = torch.randn(B, 2 * D, L, generator=g)
    p = _params(D, 64, seed=L + 1)
    gout = torch.randn(B, L, 64, generator=g)
    order = ["conv_w", "conv_b", "x_proj", "dt_proj", "out_proj", "A", "D", "dt_bias"]

    def run(fn, to):
        leaves = {k: to(v).clone().requires_grad_(True) for k, v in p.items()}
        x = to(xz).clone().requires_grad_(True)
        out = fn(x, leaves["conv_w"], leaves["conv_b"], leaves["x_proj"], leaves["dt_proj"], leaves["out_proj"], None,
                 leaves["A"], None, None, leaves["D"], delta_bias=leaves["dt_bias"], delta_softplus=True)
        out.backward(to(gout))
        return out.detach().cpu(), x.grad.cpu(), [leaves[k].grad.cpu() for k in order]

    o_ref, gx_ref, gw_ref = run(ref_ops.mamba_inner_ref, lambda t: t)
    o, gx, gw = run(ops.mamba_inner_fn, lambda t: t.to(dev))
    torch.set_grad_enabled(False)
    assert _relerr(o, o_ref) < 1e-3
    assert _relerr(gx, gx_ref) < 2e-3, ("xz", _relerr(gx, gx_ref))
    for name, a, b in zip(order, gw, gw_ref):
        assert _relerr(a, b) < 2e-3, (name, _relerr(a, b))


def test_spiral_mixer_grads_match_oracle(dev):
    """Three directions + gather/merge adjoints: Mamba(...).forward(h, 'spiral') gradients vs the oracle mixer."""
    from diffma_b200 import mixer, scan_orders, synth
    from oracle import ref_model
    torch.set_grad_enabled(True)
    ml, inv = scan_orders.spiral(4)
    kw = dict(token_list=ml[2], token_list_reversal=ml[3], origina_list=inv[2], origina_list_reversal=inv[3])
    torch.manual_seed(0)
    m = mixer.Mamba(d_model=512, d_state=16, d_conv=4, expand=2, **kw)
    synth.fill_trained_like_(m, seed=5)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    h = torch.randn(2, 16, 512, generator=g)
    gout = torch.randn(2, 16, 512, generator=g)
    h_ref = h.clone().requires_grad_(True)
    ref_model.mamba1_mixer_ref(sd, "", h_ref, "spiral", kw).backward(gout)
    m = m.to(dev)
    h_gpu = h.to(dev).requires_grad_(True)
    m(h_gpu, "spiral").backward(gout.to(dev))
    torch.set_grad_enabled(False)
    assert _relerr(h_gpu.grad.cpu(), h_ref.grad) < 2e-3
    for name, prm in m.named_parameters():
        assert prm.grad is not None, name
        assert _relerr(prm.grad.cpu(), sd[name].grad) < 3e-3, (name, _relerr(prm.grad.cpu(), sd[name].grad))


def test_training_step_bf16_autocast_runs_and_matches_fp32_direction(dev):
    """One DiffMa-S/4 training_losses + backward under bf16 autocast: finite, and the gradient points the same way as the
    fp32 run (cosine > 0.98 on the largest parameter tensors)."""
    from diffma_b200 import create_model_and_diffusion, synth
    torch.set_grad_enabled(True)
    grads = {}
    for mode in ("fp32", "bf16"):
        torch.manual_seed(0)
        net, diffusion = create_model_and_diffusion("DiffMa-S/4", respacing="")
        synth.fill_trained_like_(net, seed=11)
        net = net.to(dev).train()
        b = synth.synthetic_batch(4, tokens=49, seed=9, device=dev)
        t = torch.tensor([10, 200, 500, 900], device=dev)
        noise = torch.randn(4, 4, 28, 28, generator=torch.Generator().manual_seed(1)).to(dev)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16"):
            loss = diffusion.training_losses(net, b["x"], t, dict(y=b["y"], y2=b["y2"], w=b["w"]), noise=noise)["loss"].mean()
        loss.backward()
        assert torch.isfinite(loss)
        grads[mode] = {n: p.grad.detach().float().flatten() for n, p in net.named_parameters() if p.grad is not None}
        assert all(torch.isfinite(v).all() for v in grads[mode].values())
    torch.set_grad_enabled(False)
    for n in ("blocks.1.mamba1.in_proj.weight", "blocks.2.mamba2.out_proj.weight", "blocks.0.mamba1.x_proj.weight",
              "blocks.3.mamba1.A_log", "blocks.1.mamba2.conv1d.weight"):
        cos = torch.nn.functional.cosine_similarity(grads["fp32"][n], grads["bf16"][n], dim=0).item()
        assert cos > 0.98, (n, cos)


# ---------------------------------------------------------------------------------------------------------
# Mamba-2 (training): autograd through dm_mamba2_ssd_fwd + the reverse-scan kernel fed with
# SSD operands (autograd_ops.Mamba2SsdFn) vs autograd of the oracle
# ---------------------------------------------------------------------------------------------------------
def _m2_params(d_model, seed, d_in=1024, N=16, H=16):
    g = torch.Generator().manual_seed(seed)
    cc = d_in + 2 * N
    return dict(conv_w=torch.randn(cc, 4, generator=g) * 0.4, conv_b=torch.randn(cc, generator=g) * 0.1,
                dt_bias=torch.randn(H, generator=g) * 0.5 - 1.5, A=-torch.exp(0.5 * torch.randn(H, generator=g)),
                D=1 + 0.1 * torch.randn(H, generator=g), norm_w=1 + 0.1 * torch.randn(d_in, generator=g),
                out_proj=torch.randn(d_model, d_in, generator=g) / d_in ** 0.5)


@pytest.mark.parametrize("B,L", [(2, 21), (1, 70)])
def test_mamba_split_conv1d_scan_combined_grads_fp32(dev, B, L):
    from diffma_b200 import ops
    from oracle import ref_ops
    torch.set_grad_enabled(True)
    d_in, N, H, dm = 1024, 16, 16, 64
    g = torch.Generator().manual_seed(L)
    zx = torch.randn(B, L, 2 * d_in + 2 * N + H, generator=g)
    p = _m2_params(dm, seed=L + 1)
    gout = torch.randn(B, L, dm, generator=g)
    order = ["conv_w", "conv_b", "dt_bias", "A", "D", "norm_w", "out_proj"]

    def run(fn, to):
        leaves = {k: to(v).clone().requires_grad_(True) for k, v in p.items()}
        x = to(zx).clone().requires_grad_(True)
        out = fn(x, leaves["conv_w"], leaves["conv_b"], leaves["dt_bias"], leaves["A"], leaves["D"], 256,
                 rmsnorm_weight=leaves["norm_w"], rmsnorm_eps=1e-5, outproj_weight=leaves["out_proj"], headdim=64,
                 ngroups=1, norm_before_gate=False)
        out.backward(to(gout))
        return out.detach().cpu(), x.grad.cpu(), [leaves[k].grad.cpu() for k in order]

    o_ref, gx_ref, gw_ref = run(ref_ops.mamba_split_conv1d_scan_ref, lambda t: t)
    o, gx, gw = run(ops.mamba_split_conv1d_scan_combined, lambda t: t.to(dev))
    torch.set_grad_enabled(False)
    assert _relerr(o, o_ref) < 1e-3
    assert _relerr(gx, gx_ref) < 3e-3, ("zxbcdt", _relerr(gx, gx_ref))
    for name, a, b in zip(order, gw, gw_ref):
        assert _relerr(a, b) < 3e-3, (name, _relerr(a, b))


def test_spiral_mamba2_mixer_grads_match_oracle(dev):
    from diffma_b200 import mixer, scan_orders, synth
    from oracle import ref_model
    torch.set_grad_enabled(True)
    ml, inv = scan_orders.spiral(4)
    kw = dict(token_list=ml[2], token_list_reversal=ml[3], origina_list=inv[2], origina_list_reversal=inv[3])
    torch.manual_seed(0)
    m = mixer.Mamba2(d_model=512, d_state=16, d_conv=4, expand=2, **kw)
    synth.fill_trained_like_(m, seed=5)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    h = torch.randn(2, 16, 512, generator=g)
    gout = torch.randn(2, 16, 512, generator=g)
    h_ref = h.clone().requires_grad_(True)
    ref_model.mamba2_mixer_ref(sd, "", h_ref, "spiral", kw).backward(gout)
    m = m.to(dev)
    h_gpu = h.to(dev).requires_grad_(True)
    m(h_gpu, "spiral").backward(gout.to(dev))
    torch.set_grad_enabled(False)
    assert _relerr(h_gpu.grad.cpu(), h_ref.grad) < 3e-3
    for name, prm in m.named_parameters():
        assert prm.grad is not None, name
        assert _relerr(prm.grad.cpu(), sd[name].grad) < 4e-3, (name, _relerr(prm.grad.cpu(), sd[name].grad))


def test_mamba2_training_step_bf16_runs(dev):
    """DiffMa-S/4 --use-mamba2: training_losses + backward under bf16 autocast gives finite gradients for every
    parameter and points the same way as the fp32 run."""
    from diffma_b200 import create_model_and_diffusion, synth
    torch.set_grad_enabled(True)
    grads = {}
    for mode in ("fp32", "bf16"):
        torch.manual_seed(0)
        net, diffusion = create_model_and_diffusion("DiffMa-S/4", use_mamba2=True, respacing="")
        synth.fill_trained_like_(net, seed=11)
        net = net.to(dev).train()
        b = synth.synthetic_batch(4, tokens=49, seed=9, device=dev)
        t = torch.tensor([10, 200, 500, 900], device=dev)
        noise = torch.randn(4, 4, 28, 28, generator=torch.Generator().manual_seed(1)).to(dev)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16"):
            loss = diffusion.training_losses(net, b["x"], t, dict(y=b["y"], y2=b["y2"], w=b["w"]), noise=noise)["loss"].mean()
        loss.backward()
        assert torch.isfinite(loss)
        grads[mode] = {n: p.grad.detach().float().flatten() for n, p in net.named_parameters() if p.grad is not None}
        assert all(torch.isfinite(v).all() for v in grads[mode].values())
    torch.set_grad_enabled(False)
    for n in ("blocks.1.mamba1.in_proj.weight", "blocks.2.mamba2.out_proj.weight", "blocks.3.mamba1.A_log",
              "blocks.1.mamba2.conv1d.weight", "blocks.0.mamba1.dt_bias", "blocks.2.mamba1.norm.weight"):
        cos = torch.nn.functional.cosine_similarity(grads["fp32"][n], grads["bf16"][n], dim=0).item()
        assert cos > 0.97, (n, cos)


# ---------------------------------------------------------------------------------------------------------
# fused training path of the Spiral block (row kernels + hand-written adjoints, csrc/dm_block_bwd.cu) vs torch autograd of
# the op-by-op module path (reference block/mamba_block.py:100-115 differentiated by autograd)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,side,with_skip,with_w", [(3, 7, True, True), (2, 14, False, True), (5, 4, True, True)])
def test_fused_train_block_grads_match_module_path_fp32(dev, monkeypatch, B, side, with_skip, with_w):
    from diffma_b200 import blocks, scan_orders, synth
    L = side * side
    ml, inv = scan_orders.spiral(side)
    torch.manual_seed(0)
    blk = blocks.Spiral_MambaBlock(D_dim=512, E_dim=1024, dt_rank=16, dim_inner=1024, d_state=16, token_list=ml[2],
                                   token_list_reversal=ml[3], origina_list=inv[2], origina_list_reversal=inv[3],
                                   use_mamba2=False)
    synth.fill_trained_like_(blk, seed=3)
    blk = blk.to(dev).train()
    g = torch.Generator().manual_seed(B * 100 + side)
    x0 = torch.randn(B, L, 512, generator=g).to(dev)
    s0 = torch.randn(B, L, 512, generator=g).to(dev) * 0.5
    c0 = torch.randn(B, 1024, generator=g).to(dev)
    w = torch.sigmoid(torch.randn(B, L, 1, generator=g)).to(dev) if with_w else None
    gout = torch.randn(B, L, 512, generator=g).to(dev)
    res = {}
    for fused in (False, True):
        monkeypatch.setattr(blocks, "_FUSED_TRAIN", fused)
        blk.zero_grad(set_to_none=True)
        with torch.enable_grad():
            x = x0.clone().requires_grad_(True)
            sk = s0.clone().requires_grad_(True) if with_skip else None
            c = c0.clone().requires_grad_(True)
            out = blk(x, c, w, skip=sk) if with_skip else blk(x, c, w)
            out.backward(gout)
        res[fused] = dict(out=out.detach(), dx=x.grad, dc=c.grad, dskip=None if sk is None else sk.grad,
                          **{n: p.grad.clone() for n, p in blk.named_parameters()})
    torch.set_grad_enabled(False)
    for k, ref in res[False].items():
        if ref is None:
            continue
        got = res[True][k]
        assert got is not None, k
        assert _relerr(got, ref) < 2e-3, (k, _relerr(got, ref))


def test_fused_train_model_step_matches_module_path(dev, monkeypatch):
    """DiffMa-S/4 training loss + backward: fused training path vs module path, fp32 (tight) and bf16 autocast (the two
    paths round differently: cosine of the big gradient tensors)."""
    from diffma_b200 import blocks, create_model_and_diffusion, synth
    for mode in ("fp32", "bf16"):
        grads, losses = {}, {}
        for fused in (False, True):
            monkeypatch.setattr(blocks, "_FUSED_TRAIN", fused)
            torch.manual_seed(0)
            net, diffusion = create_model_and_diffusion("DiffMa-S/4", respacing="")
            synth.fill_trained_like_(net, seed=11)
            net = net.to(dev).train()
            b = synth.synthetic_batch(4, tokens=49, seed=9, device=dev)
            t = torch.tensor([10, 200, 500, 900], device=dev)
            noise = torch.randn(4, 4, 28, 28, generator=torch.Generator().manual_seed(1)).to(dev)
            from diffma_b200 import ops
            n0 = ops.LAUNCH_COUNTER["kernels"]
            with torch.enable_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=mode == "bf16"):
                loss = diffusion.training_losses(net, b["x"], t, dict(y=b["y"], y2=b["y2"], w=b["w"]), noise=noise)["loss"].mean()
                loss.backward()
            launched = ops.LAUNCH_COUNTER["kernels"] - n0
            # 4 blocks: scan fwd (2) + bwd (2) + merge (1) always; the fused path adds pre, post_ln, post_mix + 3 adjoints
            assert launched >= (4 * 11 if fused else 4 * 5), (fused, launched)
            losses[fused] = float(loss)
            grads[fused] = {n: p.grad.detach().float() for n, p in net.named_parameters() if p.grad is not None}
        torch.set_grad_enabled(False)
        assert set(grads[True]) == set(grads[False])
        assert abs(losses[True] - losses[False]) < (1e-4 if mode == "fp32" else 2e-2) * max(1.0, abs(losses[False]))
        for n, ref in grads[False].items():
            got = grads[True][n]
            if mode == "fp32":
                assert _relerr(got, ref) < 3e-3, (n, _relerr(got, ref))
            elif ref.numel() >= 4096:
                cos = torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0).item()
                assert cos > 0.97, (n, cos)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_mamba2_backward_cuda_orchestration_equals_torch_glue(dev, monkeypatch, dtype):
    """dm_mamba2_ssd_bwd (operand preparation + B/C conv backward kernels, dm_merge_directions_multi) around the reverse-scan
    kernel == the round-1 route (torch conv / gather / cat glue around the same kernel) on a spiral Mamba-2 mixer: every
    gradient, fp32 tight and bf16 within the rounding of the bf16 outputs."""
    from diffma_b200 import autograd_ops, mixer, scan_orders, synth
    ml, inv = scan_orders.spiral(7)
    kw = dict(token_list=ml[4], token_list_reversal=ml[5], origina_list=inv[4], origina_list_reversal=inv[5])
    torch.manual_seed(0)
    m = mixer.Mamba2(d_model=512, d_state=16, d_conv=4, expand=2, **kw)
    synth.fill_trained_like_(m, seed=5)
    m = m.to(dev)
    g = torch.Generator().manual_seed(3)
    h0 = torch.randn(3, 49, 512, generator=g).to(dev)
    gout = torch.randn(3, 49, 512, generator=g).to(dev)
    res = {}
    for cuda_route in (False, True):
        monkeypatch.setattr(autograd_ops, "_M2_BWD_CUDA", cuda_route)
        m.zero_grad(set_to_none=True)
        with torch.enable_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16):
            h = h0.clone().requires_grad_(True)
            out = m(h, "spiral")
            out.backward(gout.to(out.dtype))
        res[cuda_route] = dict(h=h.grad.float(), **{n: p.grad.float().clone() for n, p in m.named_parameters()})
    torch.set_grad_enabled(False)
    tol = 2e-3 if dtype == torch.float32 else 3e-2
    for k, ref in res[False].items():
        assert _relerr(res[True][k], ref) < tol, (k, _relerr(res[True][k], ref))
