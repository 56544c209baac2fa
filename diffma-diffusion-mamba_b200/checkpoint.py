"""DiffMa checkpoints in the reference's on-disk format.

The reference's ``train.py:293-300`` writes ``torch.save({"model": model.module.state_dict(), "ema": ema.state_dict(),
"opt": opt.state_dict(), "args": args}, f"{dir}/{steps:07d}.pt")`` and ``sample.py:19-27`` (``find_model``) loads one
of its entries (``--load-ckpt-type``: "ema" or "model"), or a bare state dict, into ``DiffMa_models[name](...)``.
``load_checkpoint`` / ``save_checkpoint`` read and write exactly that format for this package's modules (same
state-dict keys and shapes, tests/test_checkpoint_compat.py), including from a ``ddp.FlatTrainState`` whose EMA and Adam
moments live in flat buffers: the optimizer entry is rebuilt in ``torch.optim.AdamW.state_dict()`` layout so the
reference's ``opt.load_state_dict`` accepts it (and vice versa for resuming here).
"""
from __future__ import annotations

import argparse
from typing import Optional

import torch

from . import ops


def _strip_module(sd: dict) -> dict:
    """DDP-wrapped models save ``module.``-prefixed keys (the reference saves ``model.module.state_dict()``, so its own
    files have none; checkpoints written from a wrapped module elsewhere do)."""
    if sd and all(k.startswith("module.") for k in sd):
        return {k[len("module."):]: v for k, v in sd.items()}
    return sd


def read(path: str, map_location="cpu") -> dict:
    """torch.load of a reference checkpoint.  ``args`` is an argparse.Namespace (train.py:297): allow-listed so the
    safe ``weights_only`` loader accepts it."""
    torch.serialization.add_safe_globals([argparse.Namespace])
    return torch.load(path, map_location=map_location, weights_only=True)


def select(checkpoint: dict, kind: str = "ema") -> dict:
    """``find_model`` semantics (sample.py:19-27): ``checkpoint[kind]`` if present, else the dict itself is a state dict."""
    sd = checkpoint[kind] if kind in checkpoint else checkpoint
    return _strip_module(sd)


def load_checkpoint(path_or_dict, net: torch.nn.Module, kind: str = "ema", strict: bool = True) -> dict:
    """Load the ``kind`` ("ema" | "model") entry of a reference-format checkpoint (or a bare state dict) into ``net`` IN
    PLACE (so parameters that are views into a FlatTrainState's buffers stay views) and drop the inference weight
    caches.  Returns the whole checkpoint dict (``args``, ``opt`` ... for the caller)."""
    ck = read(path_or_dict) if isinstance(path_or_dict, str) else path_or_dict
    res = net.load_state_dict(select(ck, kind), strict=strict)
    if strict and (res.missing_keys or res.unexpected_keys):
        raise RuntimeError(f"checkpoint does not match the model: missing {res.missing_keys}, unexpected {res.unexpected_keys}")
    ops.invalidate_weight_caches()
    return ck if isinstance(ck, dict) else {}


def adamw_state_dict(net: torch.nn.Module, state, lr: Optional[float] = None) -> dict:
    """``torch.optim.AdamW(net.parameters()).state_dict()`` rebuilt from a ``ddp.FlatTrainState``: one param group over
    ALL of ``net.parameters()`` in order (train.py:201 hands the frozen pos_embed to the optimizer too; it simply never
    gets a state entry), per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq`` for the trainable ones."""
    params = list(net.parameters())
    offs = {id(p): o for p, o in zip(state.params, state.offsets)}
    step = state.step_t.detach().cpu().clone()
    st = {}
    for i, p in enumerate(params):
        if id(p) not in offs:
            continue
        o = offs[id(p)]
        st[i] = {"step": step.clone(), "exp_avg": state.exp_avg[o:o + p.numel()].view_as(p).detach().clone(),
                 "exp_avg_sq": state.exp_avg_sq[o:o + p.numel()].view_as(p).detach().clone()}
    group = {"lr": state.lr if lr is None else lr, "betas": tuple(state.betas), "eps": state.eps,
             "weight_decay": state.weight_decay, "amsgrad": False, "maximize": False, "foreach": None, "capturable": False,
             "differentiable": False, "fused": None, "params": list(range(len(params)))}
    return {"state": st, "param_groups": [group]}


def load_adamw_state(net: torch.nn.Module, state, opt_sd: dict) -> None:
    """Inverse of ``adamw_state_dict``: resume a FlatTrainState from a reference checkpoint's ``opt`` entry."""
    params = list(net.parameters())
    offs = {id(p): o for p, o in zip(state.params, state.offsets)}
    steps = set()
    with torch.no_grad():
        for i, p in enumerate(params):
            e = opt_sd["state"].get(i)
            if e is None or id(p) not in offs:
                continue
            o = offs[id(p)]
            state.exp_avg[o:o + p.numel()].view_as(p).copy_(e["exp_avg"])
            state.exp_avg_sq[o:o + p.numel()].view_as(p).copy_(e["exp_avg_sq"])
            steps.add(float(e["step"]))
        if len(steps) > 1:
            raise RuntimeError(f"per-parameter step counts differ ({sorted(steps)}): FlatTrainState keeps one counter")
        if steps:
            state.step_t.fill_(steps.pop())


def save_checkpoint(path: str, net: torch.nn.Module, state=None, args=None) -> dict:
    """Write ``{"model", "ema", "opt", "args"}`` like train.py:293-300.  ``state``: the ``ddp.FlatTrainState`` that trains
    ``net`` (EMA + Adam moments); without one, ``ema`` is a copy of the model and ``opt`` is empty."""
    model_sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    if state is not None and hasattr(state, "master_state"):
        # bf16 leaf weights (FlatTrainState lowp): the checkpoint holds the fp32 masters, as the reference's would
        for k, v in state.master_state(net.named_parameters()).items():
            if k in model_sd:
                model_sd[k] = v.detach().cpu().clone()
    if state is not None and state.ema is not None:
        ema_named = state.ema_state(net.named_parameters())
        ema_sd = {k: (ema_named[k].detach().cpu().clone() if k in ema_named else v.clone()) for k, v in model_sd.items()}
        opt_sd = adamw_state_dict(net, state)
        opt_sd["state"] = {i: {k: v.cpu() for k, v in e.items()} for i, e in opt_sd["state"].items()}
    else:
        ema_sd, opt_sd = {k: v.clone() for k, v in model_sd.items()}, {}
    ck = {"model": model_sd, "ema": ema_sd, "opt": opt_sd, "args": args if args is not None else argparse.Namespace()}
    torch.save(ck, path)
    return ck
