#!/usr/bin/env python
"""Mamba-1 kernel microbench at a block's shape (2 mixers x batch x 3 spiral directions x L tokens x 1024 channels).

    python tools/scan_micro.py [--batch 16] [--side 14] [--iters 20] [--phase 0|1|2] [--check]

Prints one JSON line with the median CUDA-event time of each phase (L2 flushed between launches).  Under ncu:
``ncu --set full --import-source on --clock-control none -k regex:m1_scan -s 3 -c 1 -o gpurun_out/scan python
tools/scan_micro.py --iters 1 --phase 2``.  ``--check`` compares phase 2's output with the output of the same library
run with DM_SCAN_SCHED=static (for validating scheduling variants) -- it needs a fresh process per setting, so it only
prints a checksum.
"""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from diffma_b200 import _cabi, ops, scan_orders  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--side", type=int, default=14)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--phase", type=int, default=0)
ap.add_argument("--fp32", action="store_true")
ap.add_argument("--static", action="store_true", help="static schedule (no scheduler workspace)")
ap.add_argument("--delta", action="store_true", help="experiment: hand the scan a precomputed fp16 delta (torch) instead of "
                                                     "letting it run dt_proj + softplus")
a_ = ap.parse_args()
dev = torch.device("cuda:0")
n = a_.side
L, B, D = n * n, a_.batch, 1024
dt = torch.float32 if a_.fp32 else torch.bfloat16
ml, _ = scan_orders.spiral(n)
plan = ops.ScanPlan.build([None, ml[0], ml[1]], L, "concat", dev)
g = torch.Generator(device="cpu").manual_seed(0)
xz = [torch.randn(B, L, 2 * D, generator=g).to(dev, dt) for _ in range(2)]
w = [ops.Mamba1Weights((torch.randn(D, 4, generator=g) * 0.4).to(dev), torch.zeros(D, device=dev),
                       (torch.randn(64, D, generator=g) / 32).to(dev, dt),
                       (torch.randn(D, 32, generator=g) / 5.6).to(dev, dt),
                       (torch.randn(D, generator=g) - 3).to(dev),
                       -torch.exp(torch.log(torch.arange(1, 17).float()).expand(D, 16)
                                  + 0.3 * torch.randn(D, 16, generator=g)).contiguous().to(dev),
                       torch.ones(D, device=dev)) for _ in range(2)]
# product path (bf16): a small kernel after conv + x_proj hands delta = softplus(dt_proj + bias) to the scan as fp16
use_delta = ops.USE_DELTA_HANDOVER and not a_.fp32 and not a_.delta
delta0 = torch.empty((2, B, 3, L, D), dtype=torch.float16, device=dev) if use_delta else None
a, keep = ops.mamba1_args(xz, w, plan, dynamic=not a_.static, delta=delta0)
lib, st = _cabi.lib(), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
res = {}
_cabi.check(lib.dm_mamba1_scan_phase(C.byref(a), 1, st), "phase 1")
if a_.delta:
    xd = keep[2]                                                    # (G, B, K, L, 64) fp32 rows [dt hi | dt lo | B | C]
    hl = xd[..., :32].contiguous().view(torch.bfloat16).float()     # (..., 64): hi 32, lo 32
    dt_low = hl[..., :32] + hl[..., 32:]
    delta = torch.stack([torch.nn.functional.softplus(dt_low[g] @ w[g].dt_proj_weight.float().t() + w[g].dt_bias)
                         for g in range(2)]).to(torch.float16).contiguous()
    torch.cuda.synchronize()
    _cabi.check(lib.dm_mamba1_scan_phase(C.byref(a), 2, st), "phase 2 (reference)")
    ref_out = keep[0].float().clone()
    a, keep2 = ops.mamba1_args(xz, w, plan, dynamic=not a_.static, bufs=keep, delta=delta)
for phase in ((1, 2) if a_.phase == 0 else (a_.phase,)):
    for _ in range(3):
        _cabi.check(lib.dm_mamba1_scan_phase(C.byref(a), phase, st), "phase")
    ts = []
    for _ in range(a_.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.dm_mamba1_scan_phase(C.byref(a), phase, st)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    res[f"phase{phase}_us"] = round(ts[len(ts) // 2], 2)
    res[f"phase{phase}_min_us"] = round(ts[0], 2)
out = keep[0].float()
if a_.delta:
    res["delta_vs_inkernel_maxabs"] = float((out - ref_out).abs().max())
res.update(batch=B, L=L, schedule='static' if a_.static else 'dynamic', dtype=str(dt), out_sum=float(out.sum()), out_abs=float(out.abs().sum()),
           finite=bool(torch.isfinite(out).all()), env={k: v for k, v in os.environ.items() if k.startswith("DM_")})
print(json.dumps(res))
