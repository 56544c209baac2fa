#!/usr/bin/env python
"""Training-step benchmark (BASELINE config C4): DiffMa-XL/4, synthetic brain.yaml shapes, bf16 autocast,
fwd + bwd + DDP gradient all-reduce (NCCL over NVLink, overlapped with the backward) + AdamW + EMA, one process per GPU.

    python train_bench.py [--steps 10 --warmup 3 --model DiffMa-XL/4 --batch 32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        train_bench.py --gpus 8 [--global-batch 256]

Mirrors the reference loop train.py:225-265: t ~ U{0..999}, training_losses(model, x, t, {y, y2, w}), loss.mean().backward()
(DDP all-reduce inside), AdamW(lr 1e-4, wd 0).step(), update_ema(ema, model) -- ALL of it inside the timed step.
Weak scaling by default (per-GPU batch fixed; reference global batch 256 = 8 x 32); ``--global-batch 256`` fixes the
global batch instead (strong scaling).  Prints ONE JSON line on rank 0; time = CUDA events, max over ranks.
``run()`` is also called by bench.py for its short C4 leg.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def run(model="DiffMa-XL/4", batch=32, steps=10, warmup=3, world=1, rank=0, device=None, mamba2=False, fp32=False,
        use_graph=True, overlap=True, global_batch=0, measure_exposed=True, ema=True, repeats=5):
    """Build the model + flat training state, capture the step, time ``steps`` steps.  Returns the result dict (every
    rank; only rank 0's is printed by callers).  The process group must already exist when world > 1."""
    from diffma_b200 import _cabi, create_model_and_diffusion, ops, synth
    from diffma_b200.ddp import FlatTrainState, autocast_leaf_params
    _cabi.lib()
    if global_batch:
        if global_batch % world:
            raise SystemExit(f"--global-batch {global_batch} is not divisible by {world} ranks")
        batch = global_batch // world
    dist = torch.distributed
    torch.manual_seed(rank)                       # train.py:99 seeds per rank
    net, diffusion = create_model_and_diffusion(model, use_mamba2=mamba2, respacing="")
    synth.fill_trained_like_(net, seed=11)
    net = net.to(device).train()
    # bf16 leaves for the block GEMM weights (fp32 masters in the flat state): no per-weight casts in the step
    lowp = autocast_leaf_params(net) if (not fp32 and os.environ.get("DIFFMA_LOWP", "1") != "0") else None
    state = FlatTrainState(net.parameters(), world, lr=1e-4, weight_decay=0.0, ema_decay=0.9999 if ema else None,
                           overlap=overlap, lowp=lowp)
    patch = int(model.split("/")[1])
    L = (28 // patch) ** 2
    b = synth.synthetic_batch(batch, tokens=L, seed=100 + rank, device=device)
    kw = dict(y=b["y"], y2=b["y2"], w=b["w"])
    g = torch.Generator(device=device).manual_seed(rank)
    t_buf = torch.zeros(batch, dtype=torch.long, device=device)
    noise_buf = torch.zeros_like(b["x"])
    loss_buf = torch.zeros((), device=device)

    def body(sync=True):
        state.begin_step()                               # one memset of the flat gradient buffer, buckets re-armed
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not fp32):
            loss = diffusion.training_losses(net, b["x"], t_buf, kw, noise=noise_buf)["loss"].mean()
        if not sync:                                     # comparison graph: same step without the collective
            state._fired = [True] * len(state.buckets)   # (the hooks then find every bucket already handled)
        loss.backward()                                  # bucket all-reduces fork onto the comm stream from the hooks
        state.finish_backward(reduce=sync)               # join: gradients hold the SUM over ranks
        loss_buf.copy_(loss.detach())
        state.optimizer_step()                           # AdamW + EMA, 1/world folded in: one kernel

    side = torch.cuda.Stream(device=device)
    graphs = {}

    def draw():
        # fresh timesteps / noise every step, drawn outside the graph into static buffers (train.py:243, q_sample)
        t_buf.copy_(torch.randint(0, diffusion.num_timesteps, (batch,), device=device, generator=g))
        noise_buf.normal_(generator=g)

    def step(kind="sync"):
        draw()
        gr = graphs.get(kind)
        if gr is None:
            body(sync=kind == "sync")
        else:
            gr.replay()
            state.replayed()                             # weights changed inside the graph: drop inference weight caches
        return loss_buf

    capture_note = None
    if use_graph:
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        kinds = ["sync"] + (["nosync"] if (world > 1 and measure_exposed) else [])
        for kind in kinds:
            try:
                gr = torch.cuda.CUDAGraph()
                # thread_local: the NCCL watchdog thread polls events while we capture; only this thread's calls are
                # held to the capture rules
                with torch.cuda.graph(gr, stream=side, capture_error_mode="thread_local"):
                    body(sync=kind == "sync")
                state.check_views()
                graphs[kind] = gr
            except RuntimeError as e:       # capture refused: run that variant eagerly
                capture_note = f"CUDA-graph capture of the '{kind}' step failed ({str(e).splitlines()[0][:160]}); eager"
                if rank == 0:
                    print("train_bench: " + capture_note, file=sys.stderr)
                torch.cuda.synchronize(device)
                break

    def timed(kind, n):
        for _ in range(max(3, warmup)):
            step(kind)
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step(kind)
        e1.record()
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n0 = ops.LAUNCH_COUNTER["kernels"]
    # every repetition times exactly `steps` steps; the median repetition is reported (single short runs on a B200 that is
    # still settling its clocks under the step's power draw differ by up to 5 %)
    from bench import ClockSampler                       # nvidia-smi clocks / throttle reasons during the timed region
    secs, clocks = [], []
    for _ in range(max(1, repeats)):
        with ClockSampler(device.index if device.index is not None else 0) as cs:
            secs.append(timed("sync", steps))
        clocks.append(cs.summary())
    sec = sorted(secs)[len(secs) // 2]
    loss = float(loss_buf.item())
    own_launches = (ops.LAUNCH_COUNTER["kernels"] - n0) // max(1, repeats)
    exposed = None
    if world > 1 and measure_exposed and "nosync" in graphs:
        # NOTE: without the collective the ranks' weights drift apart; this variant is timed AFTER the real one and only
        # to quantify how much of the all-reduce is not hidden behind the backward
        sec_ns = sorted(timed("nosync", steps) for _ in range(max(1, repeats)))[max(1, repeats) // 2]
        exposed = round((sec - sec_ns) / steps * 1e3, 3)
    nparam = sum(p.numel() for p in net.parameters() if p.requires_grad)
    return {
        "metric": "training_images_per_s", "value": round(world * batch * steps / sec, 2), "unit": "images/s",
        "n_gpus": world, "steps": steps, "warmup": max(3, warmup), "ms_per_step": round(sec / steps * 1e3, 3),
        "higher_is_better": True, "scaling": "strong" if global_batch else "weak", "dtype": "f32" if fp32 else "bf16",
        "data": "synthetic",
        "config": {"workload": f"{model}{' --use-mamba2' if mamba2 else ''} training step (fwd + bwd + "
                               f"{'DDP all-reduce + ' if world > 1 else ''}AdamW + EMA), L={L}, per-GPU batch {batch}",
                   "global_batch": world * batch, "grad_allreduce_mib": round(nparam * 4 / 2 ** 20, 1),
                   "buckets": len(state.buckets)},
        "clocks": {"sm_mhz_each": [c["sm_mhz"] for c in clocks], "sm_max_mhz": clocks[0]["sm_max_mhz"],
                   "reasons": sorted({r for c in clocks for r in c["reasons"]})},
        "timing": {"repeats": len(secs), "ms_per_step_each": [round(x / steps * 1e3, 3) for x in secs],
                   "note": "each repetition times exactly `steps` steps (CUDA events, max over ranks); median reported"},
        "loss": round(loss, 5), "cuda_graph": "sync" in graphs, "capture_note": capture_note,
        "exposed_allreduce_ms": exposed,
        "grad_sync": "none" if world == 1 else (
            f"FlatTrainState: {len(state.buckets)} buckets all-reduced on a comm stream from post-accumulate hooks, "
            "overlapped with the backward, inside the step's CUDA graph" if overlap else
            "FlatTrainState: one blocking all-reduce after the backward"),
        "optimizer": "dm_adamw_ema_step (AdamW + EMA, one kernel over flat fp32 state)",
        "own_kernel_launches_python_side": own_launches}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--model", default="DiffMa-XL/4")
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (weak scaling: fixed as the GPU count grows)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="strong scaling: fix the GLOBAL batch (reference config: 256) and split it over the ranks")
    ap.add_argument("--fp32", action="store_true")
    ap.add_argument("--mamba2", action="store_true", help="train with the Mamba-2 mixers (reference: train.py --use-mamba2)")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--no-overlap", action="store_true", help="one blocking all-reduce after the backward")
    ap.add_argument("--no-ema", action="store_true")
    ap.add_argument("--repeats", type=int, default=5, help="timed repetitions of `steps` steps each; the median is reported")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", rank=rank, world_size=world)
    res = run(model=args.model, batch=args.batch, steps=args.steps, warmup=args.warmup, world=world, rank=rank,
              device=device, mamba2=args.mamba2, fp32=args.fp32, use_graph=not args.no_graph,
              overlap=not args.no_overlap, global_batch=args.global_batch, ema=not args.no_ema, repeats=args.repeats)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
