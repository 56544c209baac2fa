#!/usr/bin/env python
"""Where do the scan kernel's warp-units run and how long does each live?  (debugging aid)

    python tools/scan_trace.py --build          # here (CPU): nvcc -DDM_SCAN_TRACE -> tools/_bin/libdm_trace.so
    python tools/scan_trace.py [--batch 16]     # on the GPU box: one launch, per-unit {smid, warpid, t0, t1}

Prints a JSON summary: warps per SM sub-partition (warpid % 4) histogram, unit lifetime by sub-partition load,
kernel span.  Raw records go to gpurun_out/scan_trace.npy.
"""
import argparse
import ctypes as C
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "tools", "_bin", "libdm_trace.so")

ap = argparse.ArgumentParser()
ap.add_argument("--build", action="store_true")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--side", type=int, default=14)
ap.add_argument("--extra", default="", help="extra nvcc -D flags for --build")
ap.add_argument("--lib", default=LIB)
ap.add_argument("--static", action="store_true")
a_ = ap.parse_args()

if a_.build:
    os.makedirs(os.path.dirname(a_.lib), exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(ROOT, "diffma-diffusion-mamba_b200", "csrc", "*.cu")))
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
           "-I", os.path.join(ROOT, "include"), "-DDM_SCAN_TRACE", *a_.extra.split(), "-shared", "-o", a_.lib, *srcs]
    subprocess.run(cmd, check=True)
    print("built", a_.lib)
    sys.exit(0)

import numpy as np  # noqa: E402
import torch  # noqa: E402
from diffma_b200 import _cabi, ops, scan_orders  # noqa: E402

dev = torch.device("cuda:0")
n = a_.side
L, B, D = n * n, a_.batch, 1024
ml, _ = scan_orders.spiral(n)
plan = ops.ScanPlan.build([None, ml[0], ml[1]], L, "concat", dev)
g = torch.Generator(device="cpu").manual_seed(0)
bf = torch.bfloat16
xz = [torch.randn(B, L, 2 * D, generator=g).to(dev, bf) for _ in range(2)]
w = [ops.Mamba1Weights((torch.randn(D, 4, generator=g) * 0.4).to(dev), torch.zeros(D, device=dev),
                       (torch.randn(64, D, generator=g) / 32).to(dev, bf), (torch.randn(D, 32, generator=g) / 5.6).to(dev, bf),
                       (torch.randn(D, generator=g) - 3).to(dev),
                       -torch.exp(torch.log(torch.arange(1, 17).float()).expand(D, 16)
                                  + 0.3 * torch.randn(D, 16, generator=g)).contiguous().to(dev),
                       torch.ones(D, device=dev)) for _ in range(2)]
a, keep = ops.mamba1_args(xz, w, plan, dynamic=not a_.static)
lib = C.CDLL(a_.lib)
lib.dm_mamba1_scan_phase.restype = C.c_int
lib.dm_mamba1_scan_phase.argtypes = [C.POINTER(_cabi.Mamba1Args), C.c_int, C.c_void_p]
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for phase in (1, 2, 2, 2):
    assert lib.dm_mamba1_scan_phase(C.byref(a), phase, st) == 0
torch.cuda.synchronize()
units = 2 * B * 3 * (D // 64)
if not a_.static:
    units *= (((L + 7) // 8) + 2) // 5          # work items of the dynamic schedule (default segmentation)
units = min(units, 8192)
buf = (C.c_ulonglong * (4 * units))()
assert lib.dm_debug_scan_trace(buf, units) == 0
r = np.ctypeslib.as_array(buf).reshape(units, 4).astype(np.int64)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", "scan_trace.npy"), r)
sm, wid, t0, t1 = r[:, 0], r[:, 1], r[:, 2], r[:, 3]
t0 = t0 - t0.min()
t1 = t1 - r[:, 2].min()
smsp = wid % 4
key = sm * 4 + smsp
cnt = np.bincount(key, minlength=(sm.max() + 1) * 4)
life = (t1 - t0) / 1e3
by_load = {}
for k in np.unique(cnt[cnt > 0]):
    sel = cnt[key] == k
    by_load[int(k)] = {"smsps": int((cnt == k).sum()), "units": int(sel.sum()), "life_us_mean": round(float(life[sel].mean()), 1),
                       "life_us_max": round(float(life[sel].max()), 1), "end_us_mean": round(float(t1[sel].mean() / 1e3), 1)}
print(json.dumps({"units": int(units), "sms": int(sm.max() + 1), "warps_per_sm_hist": np.bincount(np.bincount(sm)).tolist(),
                  "warps_per_smsp_hist": np.bincount(cnt).tolist(), "by_smsp_load": by_load,
                  "kernel_span_us": round(float(t1.max() / 1e3), 1), "start_spread_us": round(float(t0.max() / 1e3), 1),
                  "warpid_hist": np.bincount(wid).tolist()}))
