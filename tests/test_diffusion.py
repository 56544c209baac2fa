"""Row a12: the device-resident diffusion loop reproduces the reference's diffusion/ arithmetic.

Goldens (tests/golden/diffusion.npz) were written by importing /root/reference/diffusion itself."""
import numpy as np
import pytest
import torch

from diffma_b200 import diffusion as D
from helpers import load

G = load("diffusion.npz")


def fake_model(x, t, **kw):
    return torch.cat([0.3 * x + 0.001 * t.view(-1, 1, 1, 1).float(), torch.tanh(x)], dim=1)


@pytest.mark.parametrize("tag,resp", [("s250", "250"), ("full", "")])
def test_tables_and_map_bit_exact(tag, resp):
    d = D.create_diffusion(resp)
    assert np.array_equal(np.array(d.timestep_map), G[f"{tag}_timestep_map"])
    assert np.array_equal(d.betas, G[f"{tag}_betas"])
    for k in ("sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
              "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
              "posterior_mean_coef1", "posterior_mean_coef2"):
        assert np.array_equal(d.tables64[k], G[f"{tag}_{k}"]), k      # float64, bit-exact
    if resp == "250":
        assert d.timestep_map[:5] == [0, 4, 8, 12, 16] and d.timestep_map[-3:] == [991, 995, 999]


def _inputs(d):
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4, 4, 28, 28, generator=g)
    n = torch.randn(4, 4, 28, 28, generator=g)
    t = torch.tensor([0, 1, d.num_timesteps // 2, d.num_timesteps - 1])
    return x, n, t


@pytest.mark.parametrize("tag,resp", [("s250", "250"), ("full", "")])
def test_p_mean_variance_and_losses(tag, resp):
    d = D.create_diffusion(resp)
    x, n, t = _inputs(d)
    pmv = d.p_mean_variance(fake_model, x, t, clip_denoised=False)
    for k in ("mean", "variance", "log_variance", "pred_xstart"):
        np.testing.assert_allclose(pmv[k].numpy(), G[f"{tag}_pmv_{k}"], rtol=1e-6, atol=1e-6)
    tl = d.training_losses(fake_model, x, t, noise=n)
    for k in ("loss", "mse", "vb"):
        np.testing.assert_allclose(tl[k].numpy(), G[f"{tag}_loss_{k}"], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(d.q_sample(x, t, noise=n).numpy(), G[f"{tag}_q_sample"], rtol=1e-6, atol=1e-6)
    s = d.p_sample(fake_model, x, t, clip_denoised=False, noise=n)["sample"]
    np.testing.assert_allclose(s.numpy(), G[f"{tag}_p_sample"], rtol=1e-6, atol=1e-6)


def test_model_sees_original_timesteps_and_loop_runs():
    d = D.create_diffusion("10")
    seen = []

    def model(x, t, **kw):
        seen.append(t.clone())
        return torch.cat([0.1 * x, torch.zeros_like(x)], dim=1)

    out = d.p_sample_loop(model, (2, 4, 6, 6), noise=torch.randn(2, 4, 6, 6), clip_denoised=False)
    assert out.shape == (2, 4, 6, 6) and torch.isfinite(out).all()
    assert [int(s[0]) for s in seen] == d.timestep_map[::-1]           # respace.py:124-129 semantics


def test_unsupported_configs_raise():
    with pytest.raises(NotImplementedError):
        D.create_diffusion("250", noise_schedule="squaredcos_cap_v2")
    with pytest.raises(ValueError):
        D.space_timesteps(10, [20])
