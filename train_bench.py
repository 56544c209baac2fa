#!/usr/bin/env python
"""Training-step benchmark (BASELINE config C4): DiffMa-XL/4, synthetic brain.yaml shapes, bf16 autocast,
fwd + bwd + DDP gradient all-reduce (NCCL over NVLink) + AdamW, one process per GPU.

    python train_bench.py [--steps 10 --warmup 3 --model DiffMa-XL/4 --batch 32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        train_bench.py --gpus 8

Mirrors the reference loop train.py:225-265: t ~ U{0..999}, training_losses(model, x, t, {y, y2, w}), loss.mean().backward(),
AdamW(lr 1e-4, wd 0), EMA update.  Weak scaling (per-GPU batch fixed, reference global batch 256 = 8 x 32).
Prints ONE JSON line on rank 0; time = CUDA events, max over ranks.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


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
    ap.add_argument("--torch-ddp", action="store_true",
                    help="multi-GPU: use torch DistributedDataParallel launched eagerly (host-bound) instead of "
                         "diffma_b200.ddp.FlatGradSync between two CUDA graphs")
    args = ap.parse_args()
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and args.torch_ddp:
        # DDP's reducer hooks inside a whole-step capture deadlocked on this stack (torch 2.11 / NCCL 2.28): with torch
        # DDP the multi-GPU step is launched eagerly
        args.no_graph = True
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit(f"--global-batch {args.global_batch} is not divisible by {world} ranks")
        args.batch = args.global_batch // world
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", rank=rank, world_size=world)

    from diffma_b200 import _cabi, create_model_and_diffusion, ops, synth
    _cabi.lib()
    torch.manual_seed(rank)                       # train.py:99 seeds per rank
    net, diffusion = create_model_and_diffusion(args.model, use_mamba2=args.mamba2, respacing="")
    synth.fill_trained_like_(net, seed=11)
    net = net.to(device).train()
    model = net
    side = torch.cuda.Stream(device=device)
    sync = None
    if world > 1 and args.torch_ddp:
        # DDP is built (and warmed up, below) on the side stream: its AccumulateGrad / bucket hooks remember the stream
        # they were created under
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], gradient_as_bucket_view=True)
        torch.cuda.current_stream(device).wait_stream(side)
    elif world > 1:
        from diffma_b200.ddp import FlatGradSync
        sync = FlatGradSync(net.parameters(), world)         # same semantics as DDP; see diffma_b200/ddp.py
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=0, fused=True, capturable=not args.no_graph)
    patch = int(args.model.split("/")[1])
    L = (28 // patch) ** 2
    b = synth.synthetic_batch(args.batch, tokens=L, seed=100 + rank, device=device)
    kw = dict(y=b["y"], y2=b["y2"], w=b["w"])
    g = torch.Generator(device=device).manual_seed(rank)
    t_buf = torch.zeros(args.batch, dtype=torch.long, device=device)
    noise_buf = torch.zeros_like(b["x"])
    loss_buf = torch.zeros((), device=device)

    def fwd_bwd():
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not args.fp32):
            loss = diffusion.training_losses(model, b["x"], t_buf, kw, noise=noise_buf)["loss"].mean()
        if sync is not None:
            sync.zero()                                  # grads are views into one flat buffer: one memset, views stay
        else:
            opt.zero_grad(set_to_none=True)
        loss.backward()
        loss_buf.copy_(loss.detach())

    def body():
        fwd_bwd()
        if sync is not None:
            sync.allreduce()
        opt.step()

    graph = graph_opt = None

    def step():
        # fresh timesteps / noise every step, drawn outside the graph into static buffers (train.py:243, q_sample)
        t_buf.copy_(torch.randint(0, diffusion.num_timesteps, (args.batch,), device=device, generator=g))
        noise_buf.normal_(generator=g)
        if graph is None:
            body()
        elif sync is None:
            graph.replay()                               # whole step: forward, backward, fused AdamW
        else:
            graph.replay()                               # forward + backward into the flat gradient buffer
            sync.allreduce()                             # ONE eager NCCL all-reduce (NVLink / NVSwitch)
            graph_opt.replay()                           # fused AdamW
        return loss_buf

    if not args.no_graph:
        # single GPU: the whole step (forward, backward incl. dm_mamba1_scan_bwd, fused AdamW) is one CUDA graph -- the
        # eager step is host-bound (~5000 launches for XL/4).  Multi-GPU: two graphs around one eager all-reduce.
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        try:
            g1 = torch.cuda.CUDAGraph()
            if sync is None:
                opt.zero_grad(set_to_none=True)
                with torch.cuda.graph(g1, stream=side):
                    fwd_bwd()
                    opt.step()
                graph = g1
            else:
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g1, stream=side):
                    fwd_bwd()
                sync.check_views()
                with torch.cuda.graph(g2, stream=side):
                    opt.step()
                graph, graph_opt = g1, g2
        except RuntimeError as e:       # capture refused: run eagerly
            if rank == 0:
                print(f"train_bench: CUDA-graph capture failed ({str(e).splitlines()[0]}); falling back to eager steps",
                      file=sys.stderr)
            graph = graph_opt = None
            torch.cuda.synchronize(device)

    for _ in range(max(3, args.warmup)):
        loss = step()
    torch.cuda.synchronize(device)
    if world > 1:
        torch.distributed.barrier()
    n0 = ops.LAUNCH_COUNTER["kernels"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize(device)
    if world > 1:
        torch.distributed.barrier()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    sec = float(t.item())
    if rank == 0:
        nparam = sum(p.numel() for p in net.parameters() if p.requires_grad)
        print(json.dumps({
            "metric": "training_images_per_s", "value": round(world * args.batch * args.steps / sec, 2), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(sec / args.steps * 1e3, 3),
            "higher_is_better": True, "scaling": "strong" if args.global_batch else "weak", "dtype": "f32" if args.fp32 else "bf16", "data": "synthetic",
            "config": {"workload": f"{args.model}{' --use-mamba2' if args.mamba2 else ''} training step (fwd+bwd+{'DDP all-reduce+' if world > 1 else ''}AdamW), "
                                   f"L={L}, per-GPU batch {args.batch}", "global_batch": world * args.batch,
                       "grad_allreduce_mib": round(nparam * 4 / 2 ** 20, 1)},
            "loss": round(float(loss.item()), 5), "cuda_graph": graph is not None,
            "grad_sync": "none" if world == 1 else ("torch DDP (eager)" if args.torch_ddp else
                                                      "FlatGradSync: graph(fwd+bwd) -> 1 NCCL all-reduce -> graph(AdamW)")}),
              flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
