"""Pins ``oracle/ref_ops.py`` to outputs of independent GPU builds of the upstream kernels (vLLM's ports of
mamba_ssm's ``selective_scan_fwd`` CUDA kernel, the Triton SSD kernels, gated RMSNorm and causal conv1d), recorded
on a B200 by ``tests/golden/make_golden_vllm.py``.  CPU-only: reads the committed ``tests/golden/vllm_*.npz``."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_ops
from helpers import GOLDEN, load

torch.set_grad_enabled(False)


def _have(name):
    return os.path.exists(os.path.join(GOLDEN, name))


def _t(g, k):
    return torch.from_numpy(np.asarray(g[k]))


SCAN = [("scan_f32_a", 2e-4, 2e-4), ("scan_f32_b", 2e-4, 2e-4), ("scan_f32_c", 2e-4, 2e-4),
        ("scan_bf16_a", 1.6e-2, 1.6e-2), ("scan_bf16_b", 1.6e-2, 1.6e-2)]


@pytest.mark.parametrize("tag,rtol,atol", SCAN)
def test_selective_scan_ref_matches_upstream_cuda_kernel(tag, rtol, atol):
    """oracle selective_scan_ref == vLLM's port of upstream selective_scan_fwd_kernel.cuh (B200 run).
    fp32: 2e-4 (the kernel uses exp2f fast math); bf16 I/O: one bf16 rounding of the output (2^-8) + slack."""
    name = f"vllm_{tag}.npz"
    if not _have(name):
        pytest.skip(f"{name} not generated yet (run tests/golden/make_golden_vllm.py on the GPU box)")
    g = load(name)
    out, last = ref_ops.selective_scan_ref(_t(g, "u"), _t(g, "delta"), _t(g, "A"), _t(g, "B"), _t(g, "C"), _t(g, "D"),
                                           z=_t(g, "z"), delta_bias=_t(g, "delta_bias"), delta_softplus=True,
                                           return_last_state=True)
    ref = _t(g, "out")
    assert torch.isfinite(ref).all() and ref.abs().max() > 0.1
    torch.testing.assert_close(out, ref, rtol=rtol, atol=atol)
    torch.testing.assert_close(last, _t(g, "last_state"), rtol=max(rtol, 1e-3), atol=max(atol, 1e-3))


# fp32 inputs still run the Triton dots in TF32 (10-bit mantissa) on the GPU side: element-wise tolerances are
# upstream's bf16-class ones; the relative RMS error (2e-3) is the tight criterion.
SSD = [("ssd_f32_a", 1e-2, 2.5e-2), ("ssd_f32_b", 1e-2, 2.5e-2), ("ssd_bf16_a", 3e-2, 5e-2)]


@pytest.mark.parametrize("tag,rtol,atol", SSD)
def test_ssd_ref_matches_upstream_triton_ssd(tag, rtol, atol):
    """oracle mamba_chunk_scan_combined_ref == vLLM's port of upstream's Triton SSD (chunk cumsum / chunk state /
    state passing / bmm / chunk scan).  Tolerances are upstream's own for this op (tests/ops/triton/test_ssd.py):
    tl.dot runs tf32 for fp32 inputs."""
    name = f"vllm_{tag}.npz"
    if not _have(name):
        pytest.skip(f"{name} not generated yet")
    g = load(name)
    x = _t(g, "x")
    y = ref_ops.mamba_chunk_scan_combined_ref(x, _t(g, "dt"), _t(g, "A"), _t(g, "B"), _t(g, "C"), int(g["chunk"]),
                                              D=_t(g, "D"), z=_t(g, "z"), dt_bias=_t(g, "dt_bias"), dt_softplus=True)
    ref = _t(g, "out")
    assert torch.isfinite(ref).all() and ref.abs().max() > 0.1
    torch.testing.assert_close(y, ref, rtol=rtol, atol=atol)
    assert (y - ref).square().mean().sqrt() < 2e-3 * ref.square().mean().sqrt()
    dt = ref_ops._softplus(_t(g, "dt") + _t(g, "dt_bias"))
    _, S = ref_ops.ssd_sequential_ref(x, dt, _t(g, "A"), _t(g, "B"), _t(g, "C"))
    torch.testing.assert_close(S, _t(g, "last_state"), rtol=rtol, atol=atol)


def test_rmsnorm_gated_ref_matches_upstream_triton():
    if not _have("vllm_rmsnorm_gated.npz"):
        pytest.skip("vllm_rmsnorm_gated.npz not generated yet")
    g = load("vllm_rmsnorm_gated.npz")
    y = ref_ops.rmsnorm_gated_ref(_t(g, "x"), _t(g, "w"), z=_t(g, "z"), eps=1e-5, norm_before_gate=False)
    torch.testing.assert_close(y, _t(g, "y"), rtol=1e-5, atol=1e-5)


def test_causal_conv1d_ref_matches_upstream_port():
    if not _have("vllm_conv1d.npz"):
        pytest.skip("vllm_conv1d.npz not generated yet")
    g = load("vllm_conv1d.npz")
    y = ref_ops.causal_conv1d_ref(_t(g, "x"), _t(g, "w"), _t(g, "b"), activation="silu")
    torch.testing.assert_close(y, _t(g, "y"), rtol=1e-5, atol=1e-5)
