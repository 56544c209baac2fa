"""Live on the B200: our C-ABI path against independent builds of the upstream kernels that ship in the image
(vLLM's port of mamba_ssm's selective_scan_fwd CUDA kernel and of the Triton SSD kernels).  vLLM is LIBRARY code
used as a comparator only; the tests skip when it cannot be imported."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
N = 16


def _vllm_scan():
    try:
        from vllm.model_executor.layers.mamba.ops.mamba_ssm import selective_scan_fn
        return selective_scan_fn
    except Exception as e:       # noqa: BLE001
        pytest.skip(f"vLLM comparator not importable: {e!r}")


def _m1_inputs(B, L, D, dev, seed):
    g = torch.Generator().manual_seed(seed)
    R, dm = 32, D // 2
    p = dict(
        xz=torch.randn(B, 2 * D, L, generator=g),
        conv_w=torch.randn(D, 1, 4, generator=g) * 0.4, conv_b=torch.randn(D, generator=g) * 0.1,
        Wx=torch.randn(R + 2 * N, D, generator=g) / D ** 0.5, Wdt=torch.randn(D, R, generator=g) / R ** 0.5,
        Wout=torch.randn(dm, D, generator=g) / D ** 0.5,
        A=-torch.exp(torch.log(torch.arange(1, N + 1).float())[None, :] + 0.3 * torch.randn(D, N, generator=g)),
        D=torch.randn(D, generator=g),
        dtb=torch.log(torch.expm1(torch.exp(torch.empty(D).uniform_(-6.9, -2.3, generator=g)))))
    return {k: v.to(dev) for k, v in p.items()}


@pytest.mark.parametrize("dtype,rtol,atol", [(torch.float32, 2e-3, 1e-3), (torch.bfloat16, 3e-2, 5e-2)])
@pytest.mark.parametrize("L", [50, 196])
def test_mamba_inner_fn_vs_upstream_cuda_scan(dtype, rtol, atol, L):
    """ops.mamba_inner_fn (ours, one fused C-ABI call) == conv1d -> x_proj -> dt_proj -> upstream selective-scan CUDA
    kernel (vLLM port) -> out_proj, i.e. the op sequence of upstream MambaInnerFn.forward."""
    scan = _vllm_scan()
    from diffma_b200 import ops
    dev = torch.device("cuda:0")
    B, D, R = 3, 256, 32
    p = _m1_inputs(B, L, D, dev, 7)
    cast = lambda t: t.to(dtype)         # noqa: E731
    xz = cast(p["xz"])
    ours = ops.mamba_inner_fn(xz, p["conv_w"], p["conv_b"], cast(p["Wx"]), cast(p["Wdt"]), cast(p["Wout"]), None,
                              p["A"], None, None, p["D"], delta_bias=p["dtb"], delta_softplus=True)
    x, z = xz.chunk(2, dim=1)
    u = F.silu(F.conv1d(x.float(), p["conv_w"], p["conv_b"], padding=3, groups=D)[..., :L]).to(dtype)
    x_dbl = torch.einsum("bdl,ed->ble", u, cast(p["Wx"]))
    delta = torch.einsum("blr,dr->bdl", x_dbl[..., :R], cast(p["Wdt"])).contiguous()
    Bm = x_dbl[..., R:R + N].transpose(1, 2).contiguous()
    Cm = x_dbl[..., R + N:].transpose(1, 2).contiguous()
    states = torch.zeros(B, D, N, device=dev, dtype=dtype)
    y = scan(u.contiguous(), states, delta, p["A"], Bm, Cm, p["D"], z=z.contiguous().clone(), delta_bias=p["dtb"],
             delta_softplus=True)
    ref = torch.einsum("bdl,ed->ble", y, cast(p["Wout"]))
    torch.cuda.synchronize()
    assert torch.isfinite(ref).all()
    torch.testing.assert_close(ours.float(), ref.float(), rtol=rtol, atol=atol)
