#!/usr/bin/env python
"""Kernel-time table of ONE eager training step (BASELINE config C4 shapes) from torch.profiler (CUPTI).

    python tools/train_profile.py [--model DiffMa-XL/4 --batch 32 --mamba2]

Prints the top kernels by summed device time and the share of the step each takes (eager launch: host-bound wall time is
NOT what is reported, only device kernel time)."""
import argparse
import collections
import json
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
from diffma_b200 import create_model_and_diffusion, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="DiffMa-XL/4")
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--mamba2", action="store_true")
ap.add_argument("--no-lowp", action="store_true", help="fp32 leaf weights (autocast casts every weight every step)")
a = ap.parse_args()
dev = torch.device("cuda:0")
net, diffusion = create_model_and_diffusion(a.model, use_mamba2=a.mamba2, respacing="")
synth.fill_trained_like_(net, seed=11)
net = net.to(dev).train()
from diffma_b200.ddp import FlatTrainState, autocast_leaf_params  # noqa: E402
state = FlatTrainState(net.parameters(), 1, lr=1e-4, weight_decay=0.0, ema_decay=0.9999,
                       lowp=None if a.no_lowp else autocast_leaf_params(net))
patch = int(a.model.split("/")[1])
L = (28 // patch) ** 2
b = synth.synthetic_batch(a.batch, tokens=L, seed=100, device=dev)
kw = dict(y=b["y"], y2=b["y2"], w=b["w"])


def step():
    state.begin_step()
    t = torch.randint(0, diffusion.num_timesteps, (a.batch,), device=dev)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = diffusion.training_losses(net, b["x"], t, kw)["loss"].mean()
    loss.backward()
    state.finish_backward()
    state.optimizer_step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
# device time by the ATen / autograd op that launched the kernels (which part of the eager glue costs what)
ops_rows = []
for e in sorted(prof.key_averages(), key=lambda e: -getattr(e, "self_device_time_total", 0))[:60]:
    t = getattr(e, "self_device_time_total", 0)
    if t > 0:
        ops_rows.append({"op": e.key[:60], "calls": e.count, "self_device_us": round(t, 1)})
agg = collections.Counter()
cnt = collections.Counter()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        n = re.sub(r"<.*", "", e.name)
        n = re.sub(r"^void ", "", n)
        agg[n] += e.device_time if hasattr(e, "device_time") else e.cuda_time
        cnt[n] += 1
total = sum(agg.values())
rows = [{"kernel": k[:70], "count": cnt[k], "us": round(v, 1), "share": round(100 * v / total, 1)} for k, v in agg.most_common(22)]
print(json.dumps({"model": a.model, "batch": a.batch, "mamba2": a.mamba2, "device_time_us": round(total, 1),
                  "launches": sum(cnt.values()), "top": rows, "by_op": ops_rows}, indent=1))
