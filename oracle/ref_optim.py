"""CPU restatement of the reference's optimizer step + EMA update (TEST INFRASTRUCTURE, never imported by the product).

Follows reference train.py:201 (``torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0)``), :262 ``opt.step()``
and :34-43 / :264 ``update_ema(ema, model.module)``: ``ema.mul_(decay).add_(param, alpha=1 - decay)``.  The AdamW
update is torch's documented single-tensor algorithm (torch/optim/adamw.py, ``_single_tensor_adamw``, amsgrad=False,
maximize=False).  Pinned against ``torch.optim.AdamW`` itself in tests/test_optim_oracle.py.
"""
import torch


def adamw_ema_ref(param, grad, exp_avg, exp_avg_sq, ema, step, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                  ema_decay=0.9999, grad_scale=1.0, dtype=torch.float64):
    """One step on flat tensors; ``step`` = number of steps including this one.  Returns new (param, m, v, ema)."""
    b1, b2 = betas
    p, g = param.to(dtype), grad.to(dtype) * grad_scale
    m, v = exp_avg.to(dtype), exp_avg_sq.to(dtype)
    p = p * (1 - lr * weight_decay)
    m = m + (g - m) * (1 - b1)                                   # exp_avg.lerp_(grad, 1 - beta1)
    v = v * b2 + g * g * (1 - b2)
    denom = v.sqrt() / (1 - b2 ** step) ** 0.5 + eps
    p = p - (lr / (1 - b1 ** step)) * (m / denom)
    e = None if ema is None else ema.to(dtype) * ema_decay + p * (1 - ema_decay)
    return p, m, v, e
