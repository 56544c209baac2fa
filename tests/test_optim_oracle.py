"""oracle/ref_optim.py against torch.optim.AdamW and the reference's update_ema loop (train.py:34-43,201,262-264)."""
from collections import OrderedDict

import torch

from oracle import ref_optim


def test_oracle_adamw_ema_matches_torch_adamw_and_reference_ema_loop():
    g = torch.Generator().manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(n, generator=g, dtype=torch.float64)) for n in (7, 64, 129)]
    ema = OrderedDict((str(i), p.detach().clone()) for i, p in enumerate(params))
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=0.01)
    flat_p = torch.cat([p.detach().reshape(-1) for p in params]).clone()
    flat_e = flat_p.clone()
    m, v = torch.zeros_like(flat_p), torch.zeros_like(flat_p)
    for step in range(1, 6):
        grads = [torch.randn(p.shape, generator=g, dtype=torch.float64) for p in params]
        for p, gr in zip(params, grads):
            p.grad = gr.clone()
        opt.step()
        for (name, e), p in zip(ema.items(), params):           # update_ema, reference train.py:34-43
            e.mul_(0.999).add_(p.data, alpha=1 - 0.999)
        flat_p, m, v, flat_e = ref_optim.adamw_ema_ref(flat_p, torch.cat([x.reshape(-1) for x in grads]), m, v, flat_e, step,
                                                       lr=1e-3, weight_decay=0.01, ema_decay=0.999)
        torch.testing.assert_close(flat_p, torch.cat([p.detach().reshape(-1) for p in params]), rtol=1e-12, atol=1e-14)
        torch.testing.assert_close(flat_e, torch.cat([e.reshape(-1) for e in ema.values()]), rtol=1e-12, atol=1e-14)
