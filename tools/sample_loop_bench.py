#!/usr/bin/env python
"""BASELINE config C5: full 250-step p_sample_loop of DiffMa-XXL/2 (batch 8 per GPU) through the graphed sampler.
Prints one JSON line: end-to-end images/s of the whole sampling loop (device time, CUDA events)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffma_b200 import create_model_and_diffusion, synth
from diffma_b200.diffusion import GraphedSampler

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="DiffMa-XXL/2")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--mamba2", action="store_true")
ap.add_argument("--loops", type=int, default=2)
a = ap.parse_args()
dev = torch.device("cuda:0")
net, diffusion = create_model_and_diffusion(a.model, use_mamba2=a.mamba2, respacing="250")
synth.fill_trained_like_(net, seed=11)
net = net.to(dev).eval()
patch = int(a.model.split("/")[1])
L = (28 // patch) ** 2
b = synth.synthetic_batch(a.batch, tokens=L, seed=1, device=dev)

def model_fn(x, t, **kw):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return net(x, t, **kw).float()

s = GraphedSampler(diffusion, model_fn, tuple(b["x"].shape), dict(y=b["y"], y2=b["y2"], w=b["w"]), dev, pool_y2=True)
z = torch.randn_like(b["x"])
out = s.run(z)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.loops):
    out = s.run(z)
e1.record()
torch.cuda.synchronize()
sec = e0.elapsed_time(e1) * 1e-3 / a.loops
assert torch.isfinite(out).all()
print(json.dumps({"metric": "sampling_loop_images_per_s", "value": round(a.batch / sec, 3), "unit": "images/s",
                  "config": {"workload": f"{a.model} {'mamba2' if a.mamba2 else 'mamba1'} p_sample_loop, 250 respaced steps, batch {a.batch}, bf16",
                             "kernels_per_step": s.kernels_per_step},
                  "s_per_batch": round(sec, 4), "ms_per_step": round(sec / 250 * 1e3, 4)}))
