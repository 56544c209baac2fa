"""Host logic of diffma_b200.ddp.FlatGradSync that needs no process group (world size 1); the 2-rank behaviour is in
test_multiproc_gloo.py."""


def test_flat_grad_sync_detects_detached_views():
    """diffma_b200.ddp.FlatGradSync: every .grad aliases one flat buffer; zero() keeps the aliasing, zero_grad(set_to_none)
    breaks it and check_views() says so (host logic of the graph-friendly DDP, no process group needed at world 1)."""
    import pytest
    import torch
    from diffma_b200.ddp import FlatGradSync
    net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    sync = FlatGradSync(net.parameters(), world_size=1)
    with torch.enable_grad():
        net(torch.ones(5, 4)).sum().backward()
    sync.check_views()
    assert sync.flat.abs().sum() > 0 and sync.flat.numel() == sum(p.numel() for p in net.parameters())
    sync.allreduce()                                  # world 1: no collective, gradients untouched
    g0 = net[0].weight.grad.clone()
    sync.zero()
    assert float(sync.flat.abs().sum()) == 0.0 and float(net[0].weight.grad.abs().sum()) == 0.0
    with torch.enable_grad():
        net(torch.ones(5, 4)).sum().backward()
    torch.testing.assert_close(net[0].weight.grad, g0)      # accumulated in place into the zeroed view
    net.zero_grad(set_to_none=True)
    with pytest.raises(RuntimeError):
        sync.check_views()
