#!/usr/bin/env python
"""Headline benchmark: diffusion-step images/s of DiffMa on B200 (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model DiffMa-B/2]
                    [--batch 16] [--input-size 28] [--mamba2]

A "step" is one denoising step (``p_sample``: the whole DiffMa forward + posterior update) over one batch of
synthetic latents, bf16 autocast, captured as a CUDA graph.  N>1 (torchrun): every rank runs the same per-GPU
batch on its own latents (weak scaling, no data-path collective -- sampling shards by batch).

``value``       device-resident images/s over all ranks (CUDA events, barrier + synchronize both sides, max over ranks);
                exactly ``--steps`` steps per timed repetition, repeated until >= 0.5 s are timed, median reported
``configs``     (default run only) the other BASELINE.json configs from the same process: C2 at L=784, C3 (DiffMa-L/2
                --use-mamba2, batch 32), north_star's scan shape (DiffMa-L/2, batch 32, L=784: kernel time, HBM and MUFU
                fractions, upstream CUDA kernel), C5's per-GPU step (DiffMa-XXL/2, batch 8) and a short C4 training leg
                (DiffMa-XL/4, batch 32 per GPU, fwd + bwd + overlapped NCCL all-reduce + AdamW + EMA in one CUDA graph)
``e2e``         same step through the public API with HOST (pinned) inputs: H2D of x/t/y/y2/w, step, D2H of the sample;
                inputs and results double buffered, the host consumes every sample one step behind the device
``roofline``    the dominant kernel (m1_scan_kernel / m2_ssd_kernel) timed alone at the workload's shape with L2 flushes
``cpu_baseline`` the oracle (restatement of the reference's CPU selective_scan_ref path) on the host cores, N=1 only
``--impl reference``  times that CPU path with all host threads on a bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "diffusion_step_images_per_s"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="DiffMa-B/2")
    ap.add_argument("--batch", type=int, default=16, help="per-GPU batch")
    ap.add_argument("--input-size", type=int, default=28, help="latent side; 28 = 224x224 images (L=196), 56 -> L=784")
    ap.add_argument("--mamba2", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the side results for the other BASELINE configs")
    ap.add_argument("--no-train-leg", action="store_true", help="skip the short C4 training leg (XL/4, 5 graph steps)")
    ap.add_argument("--budget-s", type=float, default=150.0, help="wall-clock budget of the --impl reference run")
    return ap.parse_args()


def workload(args):
    from diffma_b200.model import _DEPTH
    fam, rest = args.model.split("-")
    size, patch = rest.split("/")
    L = (args.input_size // int(patch)) ** 2
    return {"workload": f"{args.model} {'mamba2' if args.mamba2 else 'mamba1'} 1 denoise step (p_sample), "
                        f"{args.input_size}x{args.input_size}x4 latents (224x224 images at 28), L={L} tokens, "
                        f"per-GPU batch {args.batch}, bf16 autocast",
            "registry_key": args.model, "depth": _DEPTH[size], "tokens": L, "per_gpu_batch": args.batch,
            "mixer": "mamba2" if args.mamba2 else "mamba1", "respacing": "250",
            "l2_policy": "working set per step (bf16 weights + activations) exceeds the 126 MB L2; "
                         "kernel microbench flushes L2 explicitly between iterations"}


# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.th.join(timeout=2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower() == "active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def dist_setup(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl" if args.impl == "ours" else "gloo", rank=rank, world_size=world)
    return world, rank, local


def barrier(world):
    if world > 1:
        torch.distributed.barrier()


def max_over_ranks(x, world, device):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


# ----------------------------------------------------------------------------------------------------
def build_model(args, device):
    from diffma_b200 import create_model_and_diffusion, synth
    torch.manual_seed(0)
    net, diffusion = create_model_and_diffusion(args.model, input_size=args.input_size, use_mamba2=args.mamba2,
                                                respacing="250")
    synth.fill_trained_like_(net, seed=11)
    return net.to(device).eval(), diffusion


def upstream_cuda_scan_us(n_seq, D, L, device, flush, iters=10):
    """The upstream selective-scan CUDA kernel at the same shape, as a comparator (never on our path): vLLM's
    ``torch.ops._C.selective_scan_fwd`` is a port of mamba_ssm's ``selective_scan_fwd_kernel.cuh`` (the reference's
    wheel itself has no sm_100 build and is not installable offline).  It covers ONLY the scan + gate: dt_proj,
    conv1d and x_proj are separate launches upstream, while our scan kernel includes dt_proj.  None if unavailable."""
    try:
        from vllm.model_executor.layers.mamba.ops.mamba_ssm import selective_scan_fn
        bf = torch.bfloat16
        u = torch.randn(n_seq, D, L, device=device, dtype=bf)
        delta = (torch.randn(n_seq, D, L, device=device) * 0.5).to(bf)
        z = torch.randn(n_seq, D, L, device=device, dtype=bf)
        Bm = torch.randn(n_seq, 1, 16, L, device=device, dtype=bf)
        Cm = torch.randn(n_seq, 1, 16, L, device=device, dtype=bf)
        A = -torch.arange(1, 17, device=device).float().repeat(D, 1)
        Dv, dtb = torch.ones(D, device=device), torch.full((D,), -4.0, device=device)
        states = torch.zeros(n_seq, D, 16, device=device, dtype=bf)
        ts = []
        for i in range(iters + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            selective_scan_fn(u, states, delta, A, Bm, Cm, Dv, z=z, delta_bias=dtb, delta_softplus=True)
            e1.record()
            torch.cuda.synchronize(device)
            if i >= 2:
                ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return {"launch_us": round(ts[len(ts) // 2], 1), "shape": [n_seq, D, L],
                "impl": "vLLM torch.ops._C.selective_scan_fwd (port of mamba_ssm selective_scan_fwd_kernel.cuh), bf16, "
                        "scan + gate only"}
    except Exception as e:      # noqa: BLE001 -- comparator only
        return {"unavailable": repr(e)[:200]}


def kernel_roofline(args, device, peaks, batch=None, input_size=None, mamba2=None, with_upstream=True):
    """Time the dominant kernel alone at the workload's per-block shape (2 mixers x 3 directions x batch)."""
    from diffma_b200 import _cabi, ops, scan_orders
    import ctypes as C
    patch = int(args.model.split("/")[1])
    n = (input_size or args.input_size) // patch
    L, B, D = n * n, batch or args.batch, 1024
    mamba2 = args.mamba2 if mamba2 is None else mamba2
    ml, _ = scan_orders.spiral(n)
    plan = ops.ScanPlan.build([None, ml[0], ml[1]], L, "concat", device)
    g = torch.Generator(device="cpu").manual_seed(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)
    iters = 20
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    res = {}
    if not mamba2:
        xz = [torch.randn(B, L, 2 * D, generator=g).to(device, torch.bfloat16) for _ in range(2)]
        w = [ops.Mamba1Weights((torch.randn(D, 4, generator=g) * 0.4).to(device), torch.zeros(D, device=device),
                               (torch.randn(64, D, generator=g) / 32).to(device, torch.bfloat16),
                               (torch.randn(D, 32, generator=g) / 5.6).to(device, torch.bfloat16),
                               (torch.randn(D, generator=g) - 3).to(device),
                               -torch.exp(torch.log(torch.arange(1, 17).float()).expand(D, 16)
                                          + 0.3 * torch.randn(D, 16, generator=g)).contiguous().to(device),
                               torch.ones(D, device=device)) for _ in range(2)]
        # as the product path launches it: kernel P hands delta = softplus(dt_proj(dt_low) + bias) to the scan as fp16
        delta = torch.empty((2, B, 3, L, D), dtype=torch.float16, device=device) if ops.USE_DELTA_HANDOVER else None
        a, keep = ops.mamba1_args(xz, w, plan, delta=delta)
        lib, st = _cabi.lib(), C.c_void_p(torch.cuda.current_stream(device).cuda_stream)
        # phase 1 = conv + x_proj (tcgen05 kernel P3) followed by the delta kernel; phase 2 = the scan
        for phase, name in ((1, "m1_conv_xproj_tc+m1_delta_kernel" if delta is not None else "m1_conv_xproj"), (2, "m1_scan_kernel")):
            for _ in range(3):
                _cabi.check(lib.dm_mamba1_scan_phase(C.byref(a), phase, st), "phase")
            for i in range(iters):
                flush.zero_()
                ev[i][0].record()
                lib.dm_mamba1_scan_phase(C.byref(a), phase, st)
                ev[i][1].record()
            torch.cuda.synchronize(device)
            res[name] = sorted(e0.elapsed_time(e1) for e0, e1 in ev)[iters // 2] * 1e-3
        upstream = upstream_cuda_scan_us(2 * B * 3, D, L, device, flush) if with_upstream else None
        token_scans = 2 * B * 3 * L
        # algorithmic bytes of the scan kernel per token-scan: read u, z [, delta fp16] (D*2 B each) + x_dbl
        # (64*4 B), write y*silu(z) (D*2 B)
        bytes_per = (4 if delta is not None else 3) * D * 2 + 256
        dom, t = "m1_scan_kernel", res["m1_scan_kernel"]
        # MUFU ops of the SCAN kernel per (token, channel): 14 of the 16 decays (one state pair is evaluated by a
        # polynomial on the FMA pipe, DM_POLY_PAIRS = 1 in csrc/dm_mamba1.cu) + the gate's tanh; with the delta hand-over
        # the softplus (2 more) runs in m1_delta_kernel
        exps = token_scans * D * (15 if delta is not None else 17)
        fma_exps = token_scans * D * 2                # the decay pair evaluated by the FMA-pipe polynomial
    else:
        upstream = None
        Cin = 2 * D + 32 + 16
        zx = [torch.randn(B, L, Cin, generator=g).to(device, torch.bfloat16) for _ in range(2)]
        w = [ops.Mamba2Weights((torch.randn(D + 32, 4, generator=g) * 0.4).to(device), torch.zeros(D + 32, device=device),
                               (torch.randn(16, generator=g) - 3).to(device),
                               -(1 + 15 * torch.rand(16, generator=g)).to(device), torch.ones(16, device=device))
             for _ in range(2)]
        for _ in range(3):
            ops.mamba2_ssd_raw(zx, w, plan, D, 16, 16)
        for i in range(iters):
            flush.zero_()
            ev[i][0].record()
            ops.mamba2_ssd_raw(zx, w, plan, D, 16, 16)       # includes the sumsq memset (tiny)
            ev[i][1].record()
        torch.cuda.synchronize(device)
        res["m2_ssd_kernel"] = sorted(e0.elapsed_time(e1) for e0, e1 in ev)[iters // 2] * 1e-3
        token_scans = 2 * B * 3 * L
        bytes_per = (2 * D + 32 + 16) * 2 + D * 2     # read z, x, B, C, dt ; write v
        dom, t = "m2_ssd_kernel", res["m2_ssd_kernel"]
        exps = token_scans * D * 6
        fma_exps = 0
    achieved = token_scans * bytes_per / t / 1e9
    peak = peaks.get("hbm_gbs", 6650.0)
    sm_mhz = peaks.get("sm_max_mhz", 1965.0)
    mufu_peak = 148 * 16 * sm_mhz * 1e6
    traffic, traffic_src = None, None
    try:        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tj = json.load(f)
        traffic = tj.get(f"{dom}@B{B}_L{L}")
        traffic_src = tj.get("_source")
    except (OSError, ValueError):
        pass
    return {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src,
            "shape": {"mixers": 2, "batch": B, "directions": 3, "tokens": L, "d_inner": D},
            "peak_source": "MEASURED_PEAKS.json (burst copy)" if "hbm_gbs" in peaks else "fallback B200_PROFILING.md",
            "launch_us": {k: round(v * 1e6, 2) for k, v in res.items()},
            "token_scans_per_launch": token_scans, "algorithmic_bytes_per_token_scan": bytes_per,
            "binding_pipe": "mufu" if not mamba2 else "fp32", "upstream_cuda_scan": upstream,
            "mufu": {"achieved_gexp_s": round(exps / t / 1e9, 1), "peak_gexp_s": round(mufu_peak / 1e9, 1),
                     "frac": round(exps / t / mufu_peak, 4),
                     "frac_counting_fma_pipe_exponentials": round((exps + fma_exps) / t / mufu_peak, 4),
                     "note": "148 SM x 16 MUFU/clk x sm_max_mhz; the scan is instruction-bound on this pipe (DESIGN.md); "
                             "`frac` counts the MUFU ops issued, the second figure also the exponentials moved to the FMA pipe"}}


def cpu_model_name():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_oracle_rate(args, budget_s, threads=None):
    """images/s of the oracle (CPU restatement of the reference path, fp32, sequential scan) on a bounded sample."""
    from diffma_b200 import synth
    from oracle import ref_model
    from diffma_b200.model import DiffMa_models, _DEPTH
    if threads:
        torch.set_num_threads(threads)
    size, patch = args.model.split("-")[1].split("/")
    depth, patch = _DEPTH[size], int(patch)
    torch.manual_seed(0)
    net = DiffMa_models[args.model](input_size=args.input_size, dt_rank=16, d_state=16, use_mamba2=args.mamba2)
    synth.fill_trained_like_(net, seed=11)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    L = (args.input_size // patch) ** 2
    bs = args.batch                      # the SAME per-step batch as the GPU arm (same config, like for like)
    b = synth.synthetic_batch(bs, input_size=args.input_size, tokens=L, seed=3)
    grid = args.input_size // patch
    x = torch.randn(bs, L, 512)
    c = torch.randn(bs, 1024)
    c1, x1, w1 = c[:1], x[:1], b["w"][:1]
    orders = ref_model.block_orders("spiral", grid, 0)
    with torch.no_grad():        # probe on one image, then scale: one block at the full batch can exceed the step budget
        t0 = time.perf_counter()
        ref_model.block_ref(sd, "blocks.0.", "spiral", x1, c1, w1, orders, args.mamba2)
        t_block = (time.perf_counter() - t0) * bs
    # how many blocks of the model fit the budget; the rest is extrapolated linearly (blocks are identical in cost)
    nb = max(1, min(depth, int(budget_s / max(t_block, 1e-3))))

    def step():
        with torch.no_grad():
            t0 = time.perf_counter()
            if nb == depth:
                ref_model.diffma_forward_ref(sd, dict(depth=depth, patch_size=patch, block_type="spiral",
                                                      use_mamba2=args.mamba2), b["x"], b["t"], b["y"], b["y2"], b["w"])
                return time.perf_counter() - t0
            h = x
            for i in range(nb):
                h = ref_model.block_ref(sd, f"blocks.{i}.", "spiral", h, c, b["w"],
                                        ref_model.block_orders("spiral", grid, i), args.mamba2)
            return (time.perf_counter() - t0) * depth / nb
    sample = (f"oracle (torch CPU fp32, sequential selective_scan_ref) on batch {bs}: "
              f"{nb} of {depth} blocks per step" + ("" if nb == depth else f", time scaled x{depth}/{nb}")
              + ("" if nb < depth else " + embed/final layers"))
    return step, bs, sample, torch.get_num_threads()


def run_reference(args, world, rank):
    if rank != 0:
        return
    total_budget = args.budget_s
    per_step = total_budget / max(1, args.steps + args.warmup)
    step, bs, sample, threads = cpu_oracle_rate(args, per_step)
    for _ in range(args.warmup):
        step()
    ts = [step() for _ in range(args.steps)]
    t = sum(ts) / len(ts)
    v = bs / t
    cfg = workload(args)
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 4), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(t * 1e3, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": round(v, 4), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "host_cpus": os.cpu_count(), "cpu_model": cpu_model_name()},
            "e2e": {"value": round(v, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference = the repo's CPU oracle port of the reference's selective_scan_ref/mamba_inner_ref path "
                    "(the reference's own CUDA wheels mamba_ssm 2.0.4 / causal_conv1d 1.2.2 are not installable offline "
                    "and ship no sm_100 code; see DESIGN.md)"}
    print(json.dumps(line), flush=True)


def timed_steps(sampler, steps, world, device, min_seconds=0.5, max_repeats=50):
    """Time EXACTLY ``steps`` graph replays, bracketed by barrier + synchronize; a region that short (tens of ms) is
    repeated until >= min_seconds have been timed in total and the MEDIAN repetition is reported (max over ranks)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps, total = [], 0.0
    while True:
        torch.cuda.synchronize(device)
        barrier(world)
        e0.record()
        for _ in range(steps):
            sampler.step()
        e1.record()
        torch.cuda.synchronize(device)
        barrier(world)
        t = max_over_ranks(e0.elapsed_time(e1) * 1e-3, world, device)
        reps.append(t)
        total += t
        if total >= min_seconds or len(reps) >= max_repeats:
            break
    reps.sort()
    return reps[len(reps) // 2], len(reps)


# y2 enters the model through its token mean only: dm_step_head takes that mean inside the step's first kernel (same device-resident
# rate, and the end-to-end path needs no separate reduction after each step's H2D: 9 988 -> 10 065 images/s).  DIFFMA_POOL_Y2=1
# pools once per batch outside the graph instead (GraphedSampler(pool_y2=True)).
POOL_Y2 = os.environ.get("DIFFMA_POOL_Y2", "0") != "0"


def side_config(args, device, world, model, batch, input_size, mamba2, steps=10):
    """Device-resident images/s of another BASELINE config in the same process (same method as the headline value)."""
    import copy
    from diffma_b200 import ops, synth
    from diffma_b200.diffusion import GraphedSampler
    a = copy.copy(args)
    a.model, a.batch, a.input_size, a.mamba2 = model, batch, input_size, mamba2
    net, diffusion = build_model(a, device)
    patch = int(model.split("/")[1])
    L = (input_size // patch) ** 2
    b = synth.synthetic_batch(batch, input_size=input_size, tokens=L, seed=100, device=device)

    def model_fn(x, t, **kwargs):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return net(x, t, **kwargs).float()

    sampler = GraphedSampler(diffusion, model_fn, tuple(b["x"].shape), dict(y=b["y"], y2=b["y2"], w=b["w"]), device,
                             clip_denoised=False, warmup=2, use_graph=not args.no_graph, pool_y2=POOL_Y2)
    sampler.reset(b["x"])
    for _ in range(3):
        sampler.step()
    t, reps = timed_steps(sampler, steps, world, device, min_seconds=0.3)
    out = {"workload": workload(a)["workload"], "value": round(world * batch * steps / t, 2), "unit": UNIT,
           "ms_per_step": round(t / steps * 1e3, 4), "steps": steps, "repeats": reps,
           "gpu_launches_per_step": sampler.kernels_per_step}
    del sampler, net
    torch.cuda.empty_cache()
    return out


def main():
    args = parse()
    world, rank, local = dist_setup(args)
    if args.impl == "reference":
        run_reference(args, world, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    from diffma_b200 import _cabi, ops
    from diffma_b200.diffusion import GraphedSampler
    _cabi.lib()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass

    net, diffusion = build_model(args, device)
    from diffma_b200 import synth
    patch = int(args.model.split("/")[1])
    L = (args.input_size // patch) ** 2
    host = synth.synthetic_batch(args.batch, input_size=args.input_size, tokens=L, seed=100 + rank)
    host = {k: v.pin_memory() for k, v in host.items()}
    dev_in = {k: v.to(device) for k, v in host.items()}
    kw = dict(y=dev_in["y"], y2=dev_in["y2"], w=dev_in["w"])

    def model_fn(x, t, **kwargs):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return net(x, t, **kwargs).float()

    shape = tuple(dev_in["x"].shape)
    ops.LAUNCH_COUNTER["kernels"] = 0
    sampler = GraphedSampler(diffusion, model_fn, shape, kw, device, clip_denoised=False, warmup=2,
                             use_graph=not args.no_graph, pool_y2=POOL_Y2)
    launches_per_step = sampler.kernels_per_step

    # ---- device-resident timing --------------------------------------------------------------------
    sampler.reset(dev_in["x"])
    for _ in range(max(3, args.warmup)):
        sampler.step()
    torch.cuda.synchronize(device)
    with ClockSampler(local) as clocks:
        t_dev, repeats = timed_steps(sampler, args.steps, world, device)
    value = world * args.batch * args.steps / t_dev

    # ---- end to end through the public API with host buffers ------------------------------------
    out_host = torch.empty(shape, dtype=torch.float32).pin_memory()
    t_host = torch.full((args.batch,), diffusion.num_timesteps - 1, dtype=torch.long).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values() if v is not host["t"]) + t_host.numel() * 8
    d2h = out_host.numel() * 4

    kw_host = {"y": host["y"], "y2": host["y2"], "w": host["w"]}
    main = torch.cuda.current_stream(device)

    out_hosts = [out_host, torch.empty_like(out_host).pin_memory()]
    done = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_run(n):
        # every step: H2D of that step's pinned inputs (double buffered: the copy of step i+1 travels on a copy stream
        # while step i computes), the step, D2H of the step's sample into a pinned buffer.  The host waits for EVERY
        # step's result, one step behind the device (results double buffered too): a serving loop consumes sample i
        # while step i+1 is already queued, so launch latency does not idle the GPU.
        sampler.prefetch(0, host["x"], t_host, kw_host)
        for i in range(n):
            s = i & 1
            sampler.load_staged(s)
            if i + 1 < n:
                sampler.prefetch(1 - s, host["x"], t_host, kw_host)
            sampler.step()
            out_hosts[s].copy_(sampler.x, non_blocking=True)
            done[s].record(main)
            if i > 0:
                done[1 - s].synchronize()              # sample i-1 is on the host now
        done[(n - 1) & 1].synchronize()

    e2e_run(3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_reps, e2e_total = [], 0.0
    while e2e_total < 0.5 and len(e2e_reps) < 50:
        barrier(world)
        torch.cuda.synchronize(device)
        e0.record()
        e2e_run(args.steps)
        e1.record()
        torch.cuda.synchronize(device)
        barrier(world)
        e2e_reps.append(max_over_ranks(e0.elapsed_time(e1) * 1e-3, world, device))
        e2e_total += e2e_reps[-1]
    e2e_reps.sort()
    t_e2e = e2e_reps[len(e2e_reps) // 2]
    e2e_value = world * args.batch * args.steps / t_e2e

    line = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": round(t_dev / args.steps * 1e3, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload(args),
            "timing": {"repeats": repeats, "e2e_repeats": len(e2e_reps),
                       "note": "each repetition times exactly `steps` steps (barrier + synchronize both sides, CUDA events, "
                               "max over ranks); repetitions continue until >= 0.5 s are timed, the median is reported"},
            "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(t_e2e / args.steps * 1e3, 4)},
            "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
            "cuda_graph": not args.no_graph, "clocks": clocks.summary()}
    del sampler
    if rank == 0:
        line["roofline"] = kernel_roofline(args, device, peaks)
    # ---- the other BASELINE configs, from the same process (BASELINE.json configs[1..3]) ----
    default_run = (args.model == "DiffMa-B/2" and args.batch == 16 and args.input_size == 28 and not args.mamba2
                   and not args.no_configs)
    if default_run:
        cfgs = {}
        try:
            cfgs["C2_L784"] = side_config(args, device, world, "DiffMa-B/2", 16, 56, False)
            cfgs["C3_mamba2_L2_b32"] = side_config(args, device, world, "DiffMa-L/2", 32, 28, True)
            c5 = side_config(args, device, world, "DiffMa-XXL/2", 8, 28, False)
            # BASELINE configs[4]: 250 respaced steps, batch 64 over 8 GPUs = 8 per GPU; the loop is 250 replays of this step
            c5["p_sample_loop_250_steps_s"] = round(c5["ms_per_step"] * 250 / 1e3, 3)
            c5["loop_images_per_s"] = round(world * 8 / (c5["ms_per_step"] * 250 / 1e3), 3)
            cfgs["C5_sampling_XXL2_b8_per_gpu"] = c5
            if rank == 0:
                # north_star's target: the fused selective scan at DiffMa-L/2, batch 32, L = 784, bf16
                a2 = argparse.Namespace(**vars(args))
                a2.model = "DiffMa-L/2"
                r = kernel_roofline(a2, device, peaks, batch=32, input_size=56, mamba2=False)
                cfgs["north_star_scan_L2_b32_L784"] = {
                    "kernel": r["kernel"], "launch_us": r["launch_us"], "shape": r["shape"],
                    "hbm": {"achieved_gbs": r["achieved"], "peak_gbs": r["peak"], "frac": r["frac"], "traffic": r["traffic"]},
                    "mufu": r["mufu"], "upstream_cuda_scan": r["upstream_cuda_scan"]}
                r2 = kernel_roofline(a2, device, peaks, batch=32, input_size=28, mamba2=True)
                cfgs["C3_ssd_kernel_L2_b32"] = {"kernel": r2["kernel"], "launch_us": r2["launch_us"], "shape": r2["shape"],
                                                "hbm": {"achieved_gbs": r2["achieved"], "peak_gbs": r2["peak"],
                                                        "frac": r2["frac"], "traffic": r2["traffic"]}}
        except Exception as e:      # noqa: BLE001 -- side results must never take the headline line down
            cfgs["error"] = repr(e)[:300]
        line["configs"] = cfgs
    if default_run and not args.no_train_leg:
        try:
            import train_bench
            line.setdefault("configs", {})["C4_training_XL4"] = train_bench.run(
                model="DiffMa-XL/4", batch=32, steps=5, warmup=3, world=world, rank=rank, device=device)
        except Exception as e:      # noqa: BLE001
            line.setdefault("configs", {})["C4_training_XL4"] = {"error": repr(e)[:300]}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            step, bs, sample, threads = cpu_oracle_rate(args, 8.0)
            step()
            t = min(step(), step())
            line["cpu_baseline"] = {"value": round(bs / t, 4), "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": sample + "; best of 2 after 1 warm-up", "host_cpus": os.cpu_count(),
                                    "cpu_model": cpu_model_name()}
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
