"""world_size-2 gloo runs on CPU: the host-side multi-process logic around the path.

* sampling shards the batch with no data-path collective: the per-rank seeds differ and timing is max over ranks;
* training is DDP: gradients of ``SpacedDiffusion.training_losses`` averaged over 2 ranks equal the single-process
  gradient on the concatenated batch (the reference relies on exactly this, train.py:153,259);
* ``bench.py --impl reference`` under torchrun: rank 0 alone prints, the other rank exits 0 silently.
"""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Tiny(torch.nn.Module):
    """Stand-in denoiser with DiffMa's call signature: (x, t, y, y2, w) -> (N, 8, H, W)."""

    def __init__(self):
        super().__init__()
        self.c = torch.nn.Conv2d(4, 8, 3, padding=1)
        self.t = torch.nn.Linear(1, 8)

    def forward(self, x, t, y, y2, w):
        return self.c(x) + self.t(t.float().view(-1, 1) / 1000.0).view(-1, 8, 1, 1) + y.mean(1).view(-1, 1, 1, 1)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bench
    from diffma_b200.diffusion import create_diffusion
    torch.manual_seed(0)
    net = _Tiny()
    ddp = torch.nn.parallel.DistributedDataParallel(net)
    d = create_diffusion("")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 4, 8, 8, generator=g)
    noise = torch.randn(4, 4, 8, 8, generator=g)
    y = torch.randn(4, 16, generator=g)
    t = torch.tensor([3, 250, 600, 999])
    sl = slice(rank * 2, rank * 2 + 2)
    with torch.enable_grad():
        loss = d.training_losses(ddp, x[sl], t[sl], dict(y=y[sl], y2=None, w=None), noise=noise[sl])["loss"].mean()
        loss.backward()
    grads = [p.grad.clone() for p in net.parameters()]
    mx = bench.max_over_ranks(float(rank + 1), world, torch.device("cpu"))
    bench.barrier(world)
    if rank == 0:
        q.put(([g.numpy() for g in grads], mx))
    dist.destroy_process_group()


def test_ddp_gloo_gradients_equal_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    grads, mx = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert mx == 2.0                                      # timing is the max over ranks
    sys.path.insert(0, ROOT)
    from diffma_b200.diffusion import create_diffusion
    torch.manual_seed(0)
    net = _Tiny()
    d = create_diffusion("")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 4, 8, 8, generator=g)
    noise = torch.randn(4, 4, 8, 8, generator=g)
    y = torch.randn(4, 16, generator=g)
    t = torch.tensor([3, 250, 600, 999])
    with torch.enable_grad():     # other test modules switch grad mode off globally
        d.training_losses(net, x, t, dict(y=y, y2=None, w=None), noise=noise)["loss"].mean().backward()
    for a, p in zip(grads, net.parameters()):
        torch.testing.assert_close(torch.from_numpy(a), p.grad, rtol=1e-5, atol=1e-6)


def _flat_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from diffma_b200.ddp import FlatGradSync
    from diffma_b200.diffusion import create_diffusion
    torch.manual_seed(rank)                       # different initial weights per rank: the broadcast must fix that
    net = _Tiny()
    sync = FlatGradSync(net.parameters(), world)
    d = create_diffusion("")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 4, 8, 8, generator=g)
    noise = torch.randn(4, 4, 8, 8, generator=g)
    y = torch.randn(4, 16, generator=g)
    t = torch.tensor([3, 250, 600, 999])
    sl = slice(rank * 2, rank * 2 + 2)
    out = []
    with torch.enable_grad():
        for _ in range(2):                        # second pass: zero() really clears, views survive a step
            sync.zero()
            d.training_losses(net, x[sl], t[sl], dict(y=y[sl], y2=None, w=None), noise=noise[sl])["loss"].mean().backward()
            sync.check_views()
            sync.allreduce()
            out.append([p.grad.clone().numpy() for p in net.parameters()])
    if rank == 0:
        q.put((out, [p.detach().clone().numpy() for p in net.parameters()]))
    dist.destroy_process_group()


def test_flat_grad_sync_equals_single_process_gradient():
    """diffma_b200.ddp.FlatGradSync (graph-friendly DDP: .grad views into one flat buffer, one all-reduce) averages
    gradients exactly like DDP: the 2-rank result equals the single-process gradient on the concatenated batch, and
    rank 0's weights were broadcast."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 200
    procs = [ctx.Process(target=_flat_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, weights = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    from diffma_b200.diffusion import create_diffusion
    torch.manual_seed(0)
    net = _Tiny()                                  # rank 0's initial weights
    for w, p in zip(weights, net.parameters()):
        torch.testing.assert_close(torch.from_numpy(w), p.detach())
    d = create_diffusion("")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 4, 8, 8, generator=g)
    noise = torch.randn(4, 4, 8, 8, generator=g)
    y = torch.randn(4, 16, generator=g)
    t = torch.tensor([3, 250, 600, 999])
    with torch.enable_grad():
        d.training_losses(net, x, t, dict(y=y, y2=None, w=None), noise=noise)["loss"].mean().backward()
    for step_grads in out:
        for a, p in zip(step_grads, net.parameters()):
            torch.testing.assert_close(torch.from_numpy(a), p.grad, rtol=1e-5, atol=1e-6)


def _flat_state_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from diffma_b200.ddp import FlatTrainState
    from diffma_b200.diffusion import create_diffusion
    torch.manual_seed(rank)                       # different initial weights per rank: the broadcast must fix that
    net = _Tiny()
    state = FlatTrainState(net.parameters(), world, bucket_mib=0.0005)      # ~130 elements per bucket: several buckets
    assert len(state.buckets) >= 2 and state.buckets[0][1] == state.total and state.buckets[-1][0] == 0
    d = create_diffusion("")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 4, 8, 8, generator=g)
    noise = torch.randn(4, 4, 8, 8, generator=g)
    y = torch.randn(4, 16, generator=g)
    t = torch.tensor([3, 250, 600, 999])
    sl = slice(rank * 2, rank * 2 + 2)
    out = []
    with torch.enable_grad():
        for _ in range(2):                        # second pass: begin_step() really clears and re-arms the buckets
            state.begin_step()
            d.training_losses(net, x[sl], t[sl], dict(y=y[sl], y2=None, w=None), noise=noise[sl])["loss"].mean().backward()
            fired_by_hooks = sum(state._fired)
            state.finish_backward()
            state.check_views()
            out.append([(p.grad / world).clone().numpy() for p in net.parameters()])
    if rank == 0:
        q.put((out, [p.detach().clone().numpy() for p in net.parameters()], fired_by_hooks, len(state.buckets)))
    dist.destroy_process_group()


def test_flat_train_state_bucketed_sync_equals_single_process_gradient():
    """diffma_b200.ddp.FlatTrainState: parameters and gradients as views into flat buffers, buckets reduced from the
    END of the buffer by post-accumulate hooks (the overlap path; on gloo the collectives simply run in hook order).
    SUM / world == the single-process gradient on the concatenated batch; rank 0's weights were broadcast."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29100 + os.getpid() % 200
    procs = [ctx.Process(target=_flat_state_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out, weights, fired, n_buckets = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert fired == n_buckets                      # every bucket was launched from a hook, none left for finish_backward
    sys.path.insert(0, ROOT)
    from diffma_b200.diffusion import create_diffusion
    torch.manual_seed(0)
    net = _Tiny()
    for w, p in zip(weights, net.parameters()):
        torch.testing.assert_close(torch.from_numpy(w), p.detach())
    d = create_diffusion("")
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 4, 8, 8, generator=g)
    noise = torch.randn(4, 4, 8, 8, generator=g)
    y = torch.randn(4, 16, generator=g)
    t = torch.tensor([3, 250, 600, 999])
    with torch.enable_grad():
        d.training_losses(net, x, t, dict(y=y, y2=None, w=None), noise=noise)["loss"].mean().backward()
    for step_grads in out:
        for a, p in zip(step_grads, net.parameters()):
            torch.testing.assert_close(torch.from_numpy(a), p.grad, rtol=1e-5, atol=1e-6)


def _lowp_state_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from diffma_b200.ddp import FlatTrainState
    torch.manual_seed(rank)
    net = torch.nn.Sequential(torch.nn.Linear(12, 16), torch.nn.SiLU(), torch.nn.Linear(16, 3))
    lowp = [net[0].weight, net[2].weight]
    state = FlatTrainState(net.parameters(), world, bucket_mib=0.0002, lowp=lowp)
    assert net[0].weight.dtype == torch.bfloat16 and net[0].bias.dtype == torch.float32
    x = torch.randn(4, 12, generator=torch.Generator().manual_seed(7 + rank))
    ok = True
    with torch.enable_grad():
        for _ in range(2):
            state.begin_step()
            with torch.autocast("cpu", dtype=torch.bfloat16):
                net(x).float().square().mean().backward()
            local = [p.grad.detach().float().clone() for p in lowp]          # bf16 results of this rank's backward
            state.finish_backward()
            state.check_views()
            for p, mine in zip(lowp, local):
                parts = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(parts, mine)
                i = next(i for i, q_ in enumerate(state.params) if q_ is p)
                ok = ok and torch.allclose(state._g_views[i], sum(parts), rtol=1e-6, atol=1e-7)
    # every rank holds rank 0's weights: fp32 masters and their bf16 shadows
    w0 = [state.flat_p.clone(), state.flat_s.float().clone()]
    for t in w0:
        dist.broadcast(t, src=0)
    ok = ok and torch.equal(w0[0], state.flat_p) and torch.equal(w0[1], state.flat_s.float())
    if rank == 0:
        q.put(bool(ok))
    dist.destroy_process_group()


def test_flat_train_state_bf16_leaves_sum_over_ranks():
    """FlatTrainState(lowp=...) on 2 gloo ranks: the bf16 gradients autograd hands to the leaf weights are copied into the
    flat fp32 buffer bucket by bucket before that bucket's all-reduce; the result is the SUM over ranks, and masters +
    shadows were broadcast from rank 0."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29350 + os.getpid() % 200
    procs = [ctx.Process(target=_lowp_state_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok


def test_reference_arm_under_torchrun_prints_once():
    env = dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(29900 + os.getpid() % 90), os.path.join(ROOT, "bench.py"), "--impl", "reference",
           "--gpus", "2", "--steps", "1", "--warmup", "0", "--model", "DiffMa-S/7", "--budget-s", "4"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["metric"] == "diffusion_step_images_per_s" and d["value"] > 0
