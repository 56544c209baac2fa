"""Checkpoint compatibility (SURVEY.md section 8f rank 4, section 5 'checkpoint / resume'): state dicts written by the
reference load into this package's modules unchanged.  Needs the reference checkout (build container only)."""
import argparse
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
CKPT = os.path.join(REF, "pretrain_ct_vision_embedder", "brain_patch_size_2.pt")

needs_ref = pytest.mark.skipif(not os.path.exists(CKPT), reason="reference checkout with shipped checkpoints not present")


def _load():
    torch.serialization.add_safe_globals([argparse.Namespace])
    return torch.load(CKPT, map_location="cpu", weights_only=True)


@needs_ref
def test_shipped_ct_embedder_checkpoint_loads_strict_and_matches_reference_forward():
    """The CT soft-mask embedder shipped with the reference ({model, ema, opt, args}; train_embedder.py) loads with
    strict=True into diffma_b200.ct_encoder.CT_Encoder and produces the reference module's outputs."""
    from diffma_b200.ct_encoder import CT_Encoder
    ck = _load()
    assert {"model", "ema"} <= set(ck)
    ours = CT_Encoder(img_size=28, patch_size=2, in_channels=4, embed_dim=512, contain_mask_token=True).eval()
    missing = ours.load_state_dict(ck["ema"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    sys.path.insert(0, REF)
    try:
        from block.CT_encoder import CT_Encoder as RefEnc
    finally:
        sys.path.remove(REF)
    ref = RefEnc(img_size=28, patch_size=2, in_channels=4, embed_dim=512, contain_mask_token=True).eval()
    ref.load_state_dict(ck["ema"], strict=True)
    x = torch.randn(3, 4, 28, 28, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        w0, y0 = ref(x)
        w1, y1 = ours(x)
    np.testing.assert_allclose(w1.numpy(), w0.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(y1.numpy(), y0.numpy(), rtol=1e-4, atol=1e-5)


@needs_ref
def test_reference_model_state_dict_loads_into_mirror():
    """A state dict produced by the reference's own DiffMa (built over the product shims) loads with strict=True into
    diffma_b200.model.DiffMa for both mixer generations (keys AND shapes)."""
    import subprocess
    code = (
        "import sys, torch\n"
        "import model as R\n"
        "from diffma_b200 import model as M\n"
        "for m2 in (False, True):\n"
        "    ref = R.DiffMa_models['DiffMa-S/2'](input_size=28, dt_rank=16, d_state=16, use_mamba2=m2)\n"
        "    ours = M.DiffMa_models['DiffMa-S/2'](input_size=28, dt_rank=16, d_state=16, use_mamba2=m2)\n"
        "    res = ours.load_state_dict(ref.state_dict(), strict=True)\n"
        "    assert not res.missing_keys and not res.unexpected_keys\n"
        "print('ok')\n")
    shims = os.path.join(ROOT, "diffma-diffusion-mamba_b200", "shims")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([shims, ROOT, REF]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-1500:]


def test_diffma_checkpoint_roundtrip_in_reference_format(tmp_path):
    """{model, ema, opt, args} as train.py:293-300 writes it: saved from a FlatTrainState, read back the way sample.py's
    find_model does (--load-ckpt-type ema | model), and the ``opt`` entry loads into a stock torch.optim.AdamW."""
    from diffma_b200 import checkpoint, model as M
    from diffma_b200.ddp import FlatTrainState
    torch.manual_seed(0)
    net = M.DiffMa_models["DiffMa-S/7"](input_size=28, dt_rank=16, d_state=16, use_mamba2=False)
    st = FlatTrainState(net.parameters(), 1, ema_decay=0.5)
    with torch.no_grad():                                   # pretend one optimizer step happened (no GPU here)
        st.flat_p.add_(0.01)
        st.ema.mul_(0.5).add_(st.flat_p, alpha=0.5)
        st.exp_avg.normal_()
        st.exp_avg_sq.uniform_()
        st.step_t.fill_(3.0)
    path = str(tmp_path / "0000003.pt")
    checkpoint.save_checkpoint(path, net, st, argparse.Namespace(model="DiffMa-S/7", image_size=224))
    ck = checkpoint.read(path)
    assert set(ck) == {"model", "ema", "opt", "args"} and ck["args"].model == "DiffMa-S/7"
    assert sorted(ck["model"]) == sorted(net.state_dict()) == sorted(ck["ema"])
    k = "blocks.1.mamba1.in_proj.weight"
    torch.testing.assert_close(ck["ema"][k], ck["model"][k] - 0.005)
    # sample.py:19-27 find_model + model.load_state_dict
    fresh = M.DiffMa_models["DiffMa-S/7"](input_size=28, dt_rank=16, d_state=16, use_mamba2=False)
    checkpoint.load_checkpoint(path, fresh, kind="ema")
    torch.testing.assert_close(fresh.state_dict()[k], ck["ema"][k])
    checkpoint.load_checkpoint({("module." + n): v for n, v in ck["model"].items()}, fresh, kind="model")   # bare, DDP-prefixed
    torch.testing.assert_close(fresh.state_dict()[k], ck["model"][k])
    # the optimizer entry is a torch.optim.AdamW state dict (train.py:201), and resumes a FlatTrainState
    opt = torch.optim.AdamW(fresh.parameters(), lr=1e-4, weight_decay=0)
    opt.load_state_dict(ck["opt"])
    p0 = next(p for p in fresh.parameters() if p.requires_grad)
    assert len(opt.state) == sum(p.requires_grad for p in fresh.parameters())      # frozen pos_embed: in the group, no state
    assert float(opt.state[p0]["step"]) == 3.0
    st2 = FlatTrainState(fresh.parameters(), 1, ema_decay=0.5)
    checkpoint.load_adamw_state(fresh, st2, ck["opt"])
    back = checkpoint.adamw_state_dict(fresh, st2)["state"]            # per-parameter views (the flat buffers also hold padding)
    for i, e in ck["opt"]["state"].items():
        torch.testing.assert_close(back[i]["exp_avg"], e["exp_avg"])
        torch.testing.assert_close(back[i]["exp_avg_sq"], e["exp_avg_sq"])
    assert float(st2.step_t) == 3.0


def test_bf16_leaf_state_saves_fp32_masters(tmp_path):
    """FlatTrainState(lowp=autocast_leaf_params(net)): the modules hold bf16 leaves, the checkpoint still holds the fp32
    masters (what the reference's model.module.state_dict() would contain), and the bf16 gradients autograd hands back
    land in the flat fp32 gradient buffer."""
    from diffma_b200 import checkpoint, model as M
    from diffma_b200.ddp import FlatTrainState, autocast_leaf_params
    torch.manual_seed(0)
    net = M.DiffMa_models["DiffMa-S/7"](input_size=28, dt_rank=16, d_state=16, use_mamba2=False)
    ref = {k: v.clone() for k, v in net.state_dict().items()}
    lowp = autocast_leaf_params(net)
    assert len(lowp) == 12 * len(net.blocks) + 2       # per block 2 x (in, out, x_proj, dt_proj) + adaLN W, b + attention W, b; + final adaLN
    st = FlatTrainState(net.parameters(), 1, ema_decay=0.5, lowp=lowp)
    k = "blocks.1.mamba1.in_proj.weight"
    p = dict(net.named_parameters())[k]
    assert p.dtype == torch.bfloat16 and p.is_leaf and p.requires_grad and p.grad is None
    torch.testing.assert_close(p.detach().float(), ref[k].to(torch.bfloat16).float(), rtol=0, atol=0)
    assert dict(net.named_parameters())["blocks.1.norm1.weight"].dtype == torch.float32
    st.check_views()
    # a bf16 gradient arrives (as autograd would deliver it) and is moved into the flat fp32 buffer
    st.begin_step()
    with torch.enable_grad():                               # other test modules switch grad mode off process-wide
        (p.float().sum() * 2.0).backward()
    assert p.grad is not None and p.grad.dtype == torch.bfloat16
    st.finish_backward()
    i = next(i for i, q in enumerate(st.params) if q is p)
    torch.testing.assert_close(st.flat_g[st.offsets[i]:st.offsets[i] + p.numel()], torch.full((p.numel(),), 2.0))
    path = str(tmp_path / "0000001.pt")
    ck = checkpoint.save_checkpoint(path, net, st, argparse.Namespace(model="DiffMa-S/7"))
    assert ck["model"][k].dtype == torch.float32
    torch.testing.assert_close(ck["model"][k], ref[k], rtol=0, atol=0)
    fresh = M.DiffMa_models["DiffMa-S/7"](input_size=28, dt_rank=16, d_state=16, use_mamba2=False)
    checkpoint.load_checkpoint(path, fresh, kind="model")
    torch.testing.assert_close(fresh.state_dict()[k], ref[k], rtol=0, atol=0)
    # resuming INTO an existing bf16-leaf state goes through the masters (net.load_state_dict would only touch the shadows)
    newer = {n: v + 0.25 for n, v in ck["model"].items() if v.dtype == torch.float32}
    st.load_master_state(net.named_parameters(), newer)
    torch.testing.assert_close(st.master_state(net.named_parameters())[k], ref[k] + 0.25, rtol=0, atol=0)
    torch.testing.assert_close(p.detach().float(), (ref[k] + 0.25).to(torch.bfloat16).float(), rtol=0, atol=0)
    st.check_views()
