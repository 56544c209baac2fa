"""GPU tests (-m gpu) of the training-step tail: dm_adamw_ema_step vs the oracle (reference train.py:34-43,201,262-264) and
the captured single-GPU training step of train_bench.run (flat state + CUDA graph) on a small model."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from diffma_b200 import _cabi
    _cabi.lib()
    return torch.device("cuda:0")


@pytest.mark.parametrize("n,wd,ema_on", [(1 << 20, 0.0, True), (4099, 0.01, True), (257, 0.0, False), (3, 0.0, True)])
def test_adamw_ema_kernel_matches_oracle(dev, n, wd, ema_on):
    """fp32 kernel vs the fp64 oracle over 4 steps: rtol 2e-6 on parameters / EMA, 1e-5 on the moments."""
    from diffma_b200 import _cabi
    from oracle import ref_optim
    lib = _cabi.lib()
    g = torch.Generator().manual_seed(n)
    pad = (n + 3) // 4 * 4
    p = torch.randn(pad, generator=g)
    m, v = torch.zeros(pad), torch.zeros(pad)
    e = p.clone()
    P, M, V, E = (t.clone().to(dev) for t in (p, m, v, e))
    step_t = torch.zeros((), device=dev)
    rp, rm, rv, re_ = p.double(), m.double(), v.double(), e.double()
    for step in range(1, 5):
        grad = torch.randn(pad, generator=g) * (0.1 if step % 2 else 10.0)
        G = grad.to(dev)
        step_t.add_(1.0)
        st = lib.dm_adamw_ema_step(P.data_ptr(), G.data_ptr(), M.data_ptr(), V.data_ptr(), E.data_ptr() if ema_on else None,
                                   step_t.data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, wd, 0.999, 0.5,
                                   torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(st, "dm_adamw_ema_step")
        rp[:n], rm[:n], rv[:n], ne = ref_optim.adamw_ema_ref(rp[:n], grad[:n].double(), rm[:n], rv[:n], re_[:n], step, lr=1e-3,
                                                             weight_decay=wd, ema_decay=0.999, grad_scale=0.5)
        re_[:n] = ne
        torch.testing.assert_close(P.cpu().double()[:n], rp[:n], rtol=2e-6, atol=2e-6)
        torch.testing.assert_close(M.cpu().double()[:n], rm[:n], rtol=1e-5, atol=2e-6)   # fp32 rounding of O(1..10) terms that nearly cancel
        torch.testing.assert_close(V.cpu().double()[:n], rv[:n], rtol=1e-5, atol=1e-9)
        if ema_on:
            torch.testing.assert_close(E.cpu().double()[:n], re_[:n], rtol=2e-6, atol=2e-6)
        else:
            assert torch.equal(E.cpu(), e)
        assert torch.equal(P.cpu()[n:], p[n:]), "elements beyond n must not be touched"


def test_captured_training_step_runs_and_updates_weights(dev):
    """train_bench.run on DiffMa-S/4 (batch 4): the whole step -- forward, backward (dm_mamba1_scan_bwd), AdamW + EMA --
    is ONE CUDA graph over the flat state; the loss is finite and replaying it really moves the weights and the EMA."""
    sys.path.insert(0, ROOT)
    import train_bench
    from diffma_b200 import ops
    e0 = ops.weights_epoch()
    with torch.enable_grad():                              # other test modules switch grad mode off process-wide
        res = train_bench.run(model="DiffMa-S/4", batch=4, steps=3, warmup=3, world=1, rank=0, device=dev)
    assert res["cuda_graph"] is True, res["capture_note"]
    assert res["loss"] == res["loss"] and 0 < res["loss"] < 10
    assert res["value"] > 0 and res["n_gpus"] == 1 and res["exposed_allreduce_ms"] is None
    assert ops.weights_epoch() > e0                        # the loop invalidated the inference weight caches


def test_flat_train_state_matches_torch_adamw_and_ema(dev):
    """One eager step of FlatTrainState on a small module == torch.optim.AdamW + the reference's update_ema loop."""
    import copy
    from diffma_b200.ddp import FlatTrainState
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(33, 17), torch.nn.SiLU(), torch.nn.Linear(17, 5)).to(dev)
    ref = copy.deepcopy(net)
    ema_ref = copy.deepcopy(net)
    opt = torch.optim.AdamW(ref.parameters(), lr=1e-3, weight_decay=0.0)
    state = FlatTrainState(net.parameters(), 1, lr=1e-3, ema_decay=0.99)
    x = torch.randn(8, 33, device=dev)
    for _ in range(3):
        with torch.enable_grad():
            state.begin_step()
            net(x).square().mean().backward()
            state.finish_backward()
            state.optimizer_step()
            opt.zero_grad()
            ref(x).square().mean().backward()
            opt.step()
        with torch.no_grad():
            for pe, pr in zip(ema_ref.parameters(), ref.parameters()):
                pe.mul_(0.99).add_(pr, alpha=0.01)
    ema = state.ema_state(net.named_parameters())
    for (name, p), pr, pe in zip(net.named_parameters(), ref.parameters(), ema_ref.parameters()):
        torch.testing.assert_close(p, pr, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(ema[name], pe, rtol=1e-5, atol=1e-6)
    state.check_views()


def test_bf16_leaf_weights_train_like_autocast_over_fp32_masters(dev):
    """FlatTrainState(lowp=...): the block GEMM weights become bf16 leaves over fp32 masters (no per-weight casts in the
    step).  Same arithmetic as autocast over fp32 parameters: after 3 steps on DiffMa-S/4 the fp32 masters, the EMA and the
    losses agree with the all-fp32-leaf run to rounding (the only difference: gradients of shared-dtype sums)."""
    from diffma_b200 import create_model_and_diffusion, synth
    from diffma_b200.ddp import FlatTrainState, autocast_leaf_params
    results = []
    for use_lowp in (False, True):
        torch.manual_seed(0)
        net, diffusion = create_model_and_diffusion("DiffMa-S/4", respacing="")
        synth.fill_trained_like_(net, seed=11)
        net = net.to(dev).train()
        lowp = autocast_leaf_params(net) if use_lowp else None
        if use_lowp:
            assert len(lowp) >= 12 * len(net.blocks)
        state = FlatTrainState(net.parameters(), 1, lr=1e-3, ema_decay=0.9, lowp=lowp)
        b = synth.synthetic_batch(4, tokens=49, seed=100, device=dev)
        kw = dict(y=b["y"], y2=b["y2"], w=b["w"])
        g = torch.Generator(device=dev).manual_seed(5)
        losses = []
        for _ in range(3):
            t = torch.randint(0, diffusion.num_timesteps, (4,), device=dev, generator=g)
            noise = torch.randn(b["x"].shape, device=dev, generator=g)
            with torch.enable_grad():
                state.begin_step()
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    loss = diffusion.training_losses(net, b["x"], t, kw, noise=noise)["loss"].mean()
                loss.backward()
                state.finish_backward()
                state.optimizer_step()
            losses.append(float(loss.detach()))
        state.check_views()
        if use_lowp:
            for p, lp in zip(state.params, state.is_lowp):
                assert (p.dtype == torch.bfloat16) == lp
            # the shadow the modules see is bf16(master) after every step
            torch.testing.assert_close(state.flat_s.float(), state.flat_p.to(torch.bfloat16).float(), rtol=0, atol=0)
        results.append((losses, state.flat_p.clone(), state.ema.clone(), state.flat_g.clone()))
    (l0, p0, e0, g0), (l1, p1, e1, g1) = results
    assert l0[0] == pytest.approx(l1[0], rel=1e-6)                 # first forward: identical weights
    for a, c in zip(l0, l1):
        assert a == pytest.approx(c, rel=2e-2)
    # gradients of the last step: same values up to bf16 rounding of single terms
    denom = g0.abs().max().clamp_min(1e-12)
    assert float((g0 - g1).abs().max() / denom) < 5e-2
    # Adam normalises the update: compare the moved distance, not the raw values
    # (a near-zero gradient whose sign differs between the runs moves a weight by up to 2 lr per step; the backward's
    #  atomics already make two runs of the SAME configuration differ that way)
    torch.testing.assert_close(p1, p0, rtol=0, atol=2 * 3 * 1e-3 * 1.1)
    assert float((p1 - p0).abs().mean()) < 2e-4
    torch.testing.assert_close(e1, e0, rtol=0, atol=1e-3)
