#!/usr/bin/env python
"""Where the time of ONE replayed denoising step goes: kernel durations vs the gaps between consecutive kernels.

    python tools/graph_gaps.py [--model DiffMa-B/2 --batch 16]

Replays the captured step (GraphedSampler, as bench.py builds it) under torch.profiler (CUPTI) and reports, for the last
replay: the span from the first kernel's start to the last kernel's end, the sum of kernel durations, the idle time between
consecutive kernels grouped by the (previous -> next) kernel pair.  CUPTI timestamps inside a graph replay are the only
in-graph numbers available here (ncu serialises and cold-starts every kernel)."""
import argparse
import collections
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
import bench  # noqa: E402
from diffma_b200 import synth  # noqa: E402
from diffma_b200.diffusion import GraphedSampler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="DiffMa-B/2")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--input-size", type=int, default=28)
ap.add_argument("--mamba2", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda:0")
net, diffusion = bench.build_model(a, dev)
patch = int(a.model.split("/")[1])
L = (a.input_size // patch) ** 2
b = synth.synthetic_batch(a.batch, input_size=a.input_size, tokens=L, seed=100, device=dev)


def model_fn(x, t, **kw):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return net(x, t, **kw).float()


s = GraphedSampler(diffusion, model_fn, tuple(b["x"].shape), dict(y=b["y"], y2=b["y2"], w=b["w"]), dev,
                   clip_denoised=False, warmup=2, use_graph=True, pool_y2=True)
s.reset(b["x"])
for _ in range(5):
    s.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        s.step()
    torch.cuda.synchronize()


def short(n):
    n = re.sub(r"^void ", "", n)
    n = n.replace("(anonymous namespace)::", "").replace("dm::", "")
    return re.sub(r"[<(].*", "", n)[:44]


ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and "emcpy" not in e.name
             and "emset" not in e.name), key=lambda e: e.time_range.start)
ends = [i for i, e in enumerate(ev) if "p_sample_update" in e.name]
lo, hi = ends[-2] + 1, ends[-1] + 1
step = ev[lo:hi]
span = step[-1].time_range.end - step[0].time_range.start
busy = sum(e.time_range.end - e.time_range.start for e in step)
gaps = collections.Counter()
cnt = collections.Counter()
dur = collections.Counter()
for p, n in zip(step[:-1], step[1:]):
    g = n.time_range.start - p.time_range.end
    k = f"{short(p.name)} -> {short(n.name)}"
    gaps[k] += g
    cnt[k] += 1
for e in step:
    dur[short(e.name)] += e.time_range.end - e.time_range.start
print(json.dumps({
    "kernels": len(step), "span_us": round(span, 1), "sum_kernel_us": round(busy, 1), "idle_us": round(span - busy, 1),
    "idle_frac": round((span - busy) / span, 4),
    "kernel_us": {k: round(v, 1) for k, v in dur.most_common(12)},
    "gaps_us": {k: [cnt[k], round(v, 1)] for k, v in gaps.most_common(14)}}))
