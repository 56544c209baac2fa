"""Pin ``oracle/ref_ops.py`` against INDEPENDENT GPU builds of the upstream kernels (run on the B200 box).

    gpurun -- 'python tests/golden/make_golden_vllm.py'      # writes gpurun_out/golden_vllm/*.npz + report.json

The reference's arithmetic for this path lives in ``mamba-ssm==2.0.4`` / ``causal-conv1d==1.2.2.post1`` (reference
``environment.yml:67,36``), which are not installable here.  The image does carry vLLM, whose
``csrc/mamba/mamba_ssm/selective_scan_fwd.cu`` (``torch.ops._C.selective_scan_fwd``) is a port of upstream's
``selective_scan_fwd_kernel.cuh`` and whose ``vllm/model_executor/layers/mamba/ops/{ssd_*,layernorm_gated,
causal_conv1d}.py`` are ports of upstream's Triton SSD / gated-RMSNorm kernels and conv1d.  They are LIBRARY code
(never on our product path); here they play the part of "the reference's native kernels run on the box": this
script feeds them seeded inputs and stores inputs + outputs as small fixtures.  ``tests/test_oracle_vllm_pin.py``
(CPU) then checks the oracle restatement against those outputs, and ``tests/test_gpu_vs_vllm.py`` (GPU) checks our
CUDA path against the same kernels live.

It also times the upstream-port scan kernel at the bench shapes (``report.json``: ``scan_us``), which is the
"reference mamba_ssm CUDA build" figure BASELINE.json asks to be reported next to ours.

Every op is wrapped in try/except: a failure is recorded in ``report.json`` and the others still run.
"""
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(ROOT, "gpurun_out", "golden_vllm")
N = 16


def scan_inputs(seed, B, D, L, dtype, dt_scale=1.0):
    """SURVEY 8d kernel-microbench distributions; everything generated on CPU so the CPU test can regenerate them."""
    g = torch.Generator().manual_seed(seed)
    u = torch.randn(B, D, L, generator=g)
    delta = torch.randn(B, D, L, generator=g) * 0.5 * dt_scale
    z = torch.randn(B, D, L, generator=g)
    Bm = torch.randn(B, N, L, generator=g)
    Cm = torch.randn(B, N, L, generator=g)
    A = -torch.exp(torch.log(torch.arange(1, N + 1).float())[None, :] + 0.3 * torch.randn(D, N, generator=g))
    Dv = torch.randn(D, generator=g)
    dtb = torch.log(torch.expm1(torch.exp(torch.empty(D).uniform_(-6.9, -2.3, generator=g))))   # softplus^-1 of [1e-3, 0.1]
    r = lambda t: t.to(dtype).float()       # noqa: E731   (store what the kernel actually sees)
    return dict(u=r(u), delta=r(delta), z=r(z), B=r(Bm), C=r(Cm), A=A, D=Dv, delta_bias=dtb)


def run_scan(inp, dtype, dev):
    from vllm.model_executor.layers.mamba.ops.mamba_ssm import selective_scan_fn
    t = {k: v.to(dev) for k, v in inp.items()}
    u, delta, z = (t[k].to(dtype).contiguous() for k in ("u", "delta", "z"))
    Bm, Cm = t["B"].to(dtype).contiguous(), t["C"].to(dtype).contiguous()
    states = torch.zeros(u.shape[0], u.shape[1], N, device=dev, dtype=dtype)
    out = selective_scan_fn(u, states, delta, t["A"].contiguous(), Bm, Cm, t["D"].contiguous(), z=z,
                            delta_bias=t["delta_bias"].contiguous(), delta_softplus=True)
    torch.cuda.synchronize()
    return out.float().cpu(), states.float().cpu()


def ssd_inputs(seed, B, L, H, P, dtype):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, L, H, P, generator=g)
    z = torch.randn(B, L, H, P, generator=g)
    dt = torch.randn(B, L, H, generator=g) * 0.5
    Bm = torch.randn(B, L, 1, N, generator=g)
    Cm = torch.randn(B, L, 1, N, generator=g)
    A = -torch.exp(0.5 * torch.randn(H, generator=g))
    Dv = torch.randn(H, generator=g)
    dtb = torch.log(torch.expm1(torch.exp(torch.empty(H).uniform_(-4.6, -1.2, generator=g))))
    r = lambda t: t.to(dtype).float()       # noqa: E731
    return dict(x=r(x), z=r(z), dt=r(dt), B=r(Bm), C=r(Cm), A=A, D=Dv, dt_bias=dtb)


def run_ssd(inp, dtype, dev, chunk):
    from vllm.model_executor.layers.mamba.ops.ssd_combined import mamba_chunk_scan_combined_varlen
    B, L, H, P = inp["x"].shape
    t = {k: v.to(dev) for k, v in inp.items()}
    flat = lambda a: a.reshape(B * L, *a.shape[2:]).to(dtype).contiguous()      # noqa: E731
    x, z, dt, Bm, Cm = flat(t["x"]), flat(t["z"]), flat(t["dt"]), flat(t["B"]), flat(t["C"])
    cu = torch.arange(0, (B + 1) * L, L, dtype=torch.int32, device=dev)
    bounds, seq_idx, last = [0], [], []
    for b in range(B):
        pos = 0
        while pos < L:
            pos = min(pos + chunk, L)
            bounds.append(b * L + pos)
            seq_idx.append(b)
        last.append(len(seq_idx) - 1)
    out = torch.empty_like(x)
    states = mamba_chunk_scan_combined_varlen(
        x, dt, t["A"], Bm, Cm, chunk, cu, torch.tensor(bounds, dtype=torch.int32, device=dev),
        torch.tensor(last, dtype=torch.int32, device=dev), torch.tensor(seq_idx, dtype=torch.int32, device=dev), out,
        D=t["D"], z=z, dt_bias=t["dt_bias"], dt_softplus=True, state_dtype=torch.float32)
    torch.cuda.synchronize()
    return out.float().cpu().reshape(B, L, H, P), states.float().cpu()


def run_rmsnorm(dev):
    from vllm.model_executor.layers.mamba.ops.layernorm_gated import rms_norm_gated
    g = torch.Generator().manual_seed(5)
    x = torch.randn(6, 37, 256, generator=g)
    z = torch.randn(6, 37, 256, generator=g)
    w = 1.0 + 0.2 * torch.randn(256, generator=g)
    y = rms_norm_gated(x.to(dev), w.to(dev), None, z=z.to(dev), eps=1e-5, group_size=None, norm_before_gate=False)
    torch.cuda.synchronize()
    return dict(x=x, z=z, w=w, y=y.float().cpu())


def run_conv(dev):
    from vllm.model_executor.layers.mamba.ops.causal_conv1d import causal_conv1d_fn
    g = torch.Generator().manual_seed(6)
    B, C, L, W = 3, 96, 41, 4
    x = torch.randn(B, C, L, generator=g)
    w = torch.randn(C, W, generator=g) * 0.5
    b = torch.randn(C, generator=g) * 0.1
    xf = x.permute(0, 2, 1).reshape(B * L, C).contiguous().to(dev).t()           # (dim, total tokens), channel-last
    conv_states = torch.zeros(B + 1, W - 1, C, device=dev).transpose(1, 2)       # (lines, dim, width-1), dim stride 1
    qsl = torch.arange(0, (B + 1) * L, L, dtype=torch.int32, device=dev)
    y = causal_conv1d_fn(xf, w.to(dev), b.to(dev), conv_states, qsl,
                         cache_indices=torch.arange(1, B + 1, dtype=torch.int32, device=dev),    # line 0 = vLLM's null block
                         has_initial_state=torch.zeros(B, dtype=torch.bool, device=dev), activation="silu")
    torch.cuda.synchronize()
    y = y.float().cpu().reshape(C, B, L).permute(1, 0, 2).contiguous()
    return dict(x=x, w=w, b=b, y=y)


def time_scan(dev, B, D, L, iters=20):
    """Upstream-port scan kernel alone (inputs resident, L2 flushed between launches), microseconds per launch."""
    dtype = torch.bfloat16
    u = torch.randn(B, D, L, device=dev, dtype=dtype)
    delta = (torch.randn(B, D, L, device=dev) * 0.5).to(dtype)
    z0 = torch.randn(B, D, L, device=dev, dtype=dtype)
    Bm = torch.randn(B, 1, N, L, device=dev, dtype=dtype)
    Cm = torch.randn(B, 1, N, L, device=dev, dtype=dtype)
    A = -torch.exp(torch.log(torch.arange(1, N + 1, device=dev).float())[None, :].repeat(D, 1))
    Dv = torch.randn(D, device=dev)
    dtb = torch.full((D,), -4.0, device=dev)
    states = torch.zeros(B, D, N, device=dev, dtype=dtype)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    z = z0.clone()

    from vllm.model_executor.layers.mamba.ops.mamba_ssm import selective_scan_fn

    def launch():       # output is written in place into z
        selective_scan_fn(u, states, delta, A, Bm, Cm, Dv, z=z, delta_bias=dtb, delta_softplus=True)

    launch()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return {"median_us": ts[len(ts) // 2], "min_us": ts[0], "shape": [B, D, L], "dtype": "bf16"}


def main():
    os.makedirs(OUT, exist_ok=True)
    dev = torch.device("cuda:0")
    report = {"errors": {}, "torch": torch.__version__, "gpu": torch.cuda.get_device_name(0)}
    t0 = time.time()
    try:
        import vllm
        report["vllm"] = getattr(vllm, "__version__", "?")
    except Exception:
        report["errors"]["import"] = traceback.format_exc()
    report["import_s"] = time.time() - t0

    cases = [("scan_f32_a", 101, 2, 48, 50, torch.float32, 1.0), ("scan_f32_b", 102, 1, 32, 196, torch.float32, 1.0),
             ("scan_f32_c", 103, 2, 32, 300, torch.float32, 2.0), ("scan_bf16_a", 104, 2, 48, 50, torch.bfloat16, 1.0),
             ("scan_bf16_b", 105, 1, 32, 196, torch.bfloat16, 1.0)]
    for tag, seed, B, D, L, dtype, sc in cases:
        try:
            inp = scan_inputs(seed, B, D, L, dtype, sc)
            out, last = run_scan(inp, dtype, dev)
            np.savez_compressed(os.path.join(OUT, f"vllm_{tag}.npz"), out=out.numpy(), last_state=last.numpy(),
                                **{k: v.numpy() for k, v in inp.items()})
            report[tag] = {"out_abs_max": float(out.abs().max())}
        except Exception:
            report["errors"][tag] = traceback.format_exc()

    for tag, seed, B, L, H, P, dtype, chunk in [("ssd_f32_a", 201, 2, 50, 4, 16, torch.float32, 16),
                                                ("ssd_f32_b", 202, 1, 196, 2, 64, torch.float32, 64),
                                                ("ssd_bf16_a", 203, 2, 196, 2, 64, torch.bfloat16, 256)]:
        try:
            inp = ssd_inputs(seed, B, L, H, P, dtype)
            out, st = run_ssd(inp, dtype, dev, chunk)
            np.savez_compressed(os.path.join(OUT, f"vllm_{tag}.npz"), out=out.numpy(), last_state=st.numpy(),
                                chunk=np.array(chunk), **{k: v.numpy() for k, v in inp.items()})
            report[tag] = {"out_abs_max": float(out.abs().max())}
        except Exception:
            report["errors"][tag] = traceback.format_exc()

    for tag, fn in (("rmsnorm_gated", run_rmsnorm), ("conv1d", run_conv)):
        try:
            r = fn(dev)
            np.savez_compressed(os.path.join(OUT, f"vllm_{tag}.npz"), **{k: v.numpy() for k, v in r.items()})
            report[tag] = {"out_abs_max": float(r["y"].abs().max())}
        except Exception:
            report["errors"][tag] = traceback.format_exc()

    report["scan_us"] = {}
    for tag, B, D, L in (("c2_b16_L196", 2 * 16 * 3, 1024, 196), ("L2_b32_L784", 2 * 32 * 3, 1024, 784),
                         ("c2_b16_L784", 2 * 16 * 3, 1024, 784)):
        try:
            report["scan_us"][tag] = time_scan(dev, B, D, L)
        except Exception:
            report["errors"]["time_" + tag] = traceback.format_exc()
    with open(os.path.join(OUT, "report.json"), "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps({k: v for k, v in report.items() if k != "errors"}, indent=1))
    for k, v in report["errors"].items():
        print("ERROR", k, v.splitlines()[-1])


if __name__ == "__main__":
    sys.exit(main())
