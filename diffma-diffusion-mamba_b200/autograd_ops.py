"""autograd wrappers of the scan ops (training path).

``Mamba1ScanFn`` / ``Mamba2SsdFn`` run the same C-ABI forward as inference and keep the intermediates the
backward kernels need (u, x_dbl for Mamba-1).  The backward entry points are the next build step; until they
exist ``backward`` raises instead of silently falling back to a slow path.
"""
from __future__ import annotations

import torch

from . import ops

_W1 = ("conv_weight", "conv_bias", "x_proj_weight", "dt_proj_weight", "dt_bias", "A", "D")
_W2 = ("conv_weight", "conv_bias", "dt_bias", "A", "D")


def flatten_weights(weights):
    fields = _W1 if isinstance(weights[0], ops.Mamba1Weights) else _W2
    return [getattr(w, f) for w in weights for f in fields]


class Mamba1ScanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, G, *tensors):
        xz = list(tensors[:G])
        flat = tensors[G:]
        n = len(_W1)
        weights = [ops.Mamba1Weights(*flat[g * n:(g + 1) * n]) for g in range(G)]
        out, u, x_dbl = ops.mamba1_scan_raw([t.detach() for t in xz],
                                            [ops.Mamba1Weights(*[None if v is None else v.detach() for v in flat[g * n:(g + 1) * n]])
                                             for g in range(G)], plan)
        ctx.plan, ctx.G = plan, G
        ctx.save_for_backward(*xz, *[t for t in flat if t is not None], u, x_dbl)
        ctx.none_mask = [t is None for t in flat]
        del weights
        return out

    @staticmethod
    def backward(ctx, grad_out):
        raise NotImplementedError("diffma_b200: dm_mamba1_scan_bwd (reverse scan + conv/proj gradients) is not built yet; "
                                  "training through the Mamba-1 mixer is the next milestone (DESIGN.md)")


class Mamba2SsdFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, G, d_inner, d_state, nheads, gate, want_sumsq, *tensors):
        zx = [t.detach() for t in tensors[:G]]
        flat = tensors[G:]
        n = len(_W2)
        weights = [ops.Mamba2Weights(*[None if v is None else v.detach() for v in flat[g * n:(g + 1) * n]])
                   for g in range(G)]
        v, ss = ops.mamba2_ssd_raw(zx, weights, plan, d_inner, d_state, nheads, gate, want_sumsq)
        if ss is None:
            ss = v.new_zeros(())
        ctx.mark_non_differentiable(ss)
        return v, ss

    @staticmethod
    def backward(ctx, gv, gss):
        raise NotImplementedError("diffma_b200: the Mamba-2 backward kernel is not built yet; training with "
                                  "--use-mamba2 is a later milestone (DESIGN.md)")
