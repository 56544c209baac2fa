"""Generate the committed golden fixtures by running the REFERENCE itself in the build container.

    python tests/golden/make_golden.py            # needs /root/reference; writes tests/golden/*

What comes from where:

* ``scan_orders.json``     -- ``/root/reference/tools.py`` imported as is (integer tables + sha256).
* ``diffusion.npz``        -- ``/root/reference/diffusion`` imported as is (tables, timestep map,
                              p_mean_variance / training_losses on a closed-form stand-in model).
* ``ct_encoder.npz``       -- ``/root/reference/block/CT_encoder.py`` imported as is.
* ``model_*.npz``, ``mixer_*.npz`` -- ``/root/reference/model.py`` + ``block/*.py`` imported UNMODIFIED;
  the wheels they import (``mamba_ssm``, ``causal_conv1d``, ``timm``; absent from this image) are
  provided by ``oracle/ref_shims`` (CPU restatement of the upstream reference functions).  These pin
  the reference-owned orchestration; the inner-op arithmetic stays "parity unpinned".

Weights and inputs are NOT stored: both sides regenerate them from
``diffma_b200.synth.fill_trained_like_`` / ``synthetic_batch`` (name-keyed, seeded), so fixtures stay small.
``/root/reference`` does not exist on the GPU box; tests only read the files written here.
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "oracle", "ref_shims"), ROOT, REF]

from diffma_b200 import synth  # noqa: E402


SUB = 7   # token subsampling stride for the larger activations (keeps fixtures small)


def stats(a):
    """Whole-tensor float64 moments, so the untouched tokens are still covered."""
    a = a.astype(np.float64)
    return np.array([a.sum(), np.abs(a).sum(), (a * a).sum()])


def sha(obj):
    return hashlib.sha256(json.dumps(obj, separators=(",", ":")).encode()).hexdigest()


def make_scan_orders():
    import tools
    out = {"full": {}, "sha256": {}}
    for n in (2, 4, 7, 14):
        ml, inv = tools.spiral(n)
        out["full"][f"spiral_{n}"] = [ml, inv]
        out["full"][f"zig_{n}"] = [[list(tools.zig(n, i)[0]) for i in range(8)],
                                   [list(tools.zig(n, i)[1]) for i in range(8)]]
        vm = tools.vmamba_(n)
        out["full"][f"vmamba_{n}"] = [vm[0], vm[1]]
    for n in (4, 7, 14, 28, 56):
        ml, inv = tools.spiral(n)
        out["sha256"][f"spiral_{n}_orders"] = sha(ml)
        out["sha256"][f"spiral_{n}_inverses"] = sha(inv)
        out["sha256"][f"zig_{n}_orders"] = sha([list(tools.zig(n, i)[0]) for i in range(8)])
        out["sha256"][f"vmamba_{n}_orders"] = sha(list(tools.vmamba_(n)[0]))
    with open(os.path.join(HERE, "scan_orders.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))


def fake_model(x, t, **kw):
    """Closed-form stand-in for the network: (eps, var-logits) as smooth functions of (x, t)."""
    return torch.cat([0.3 * x + 0.001 * t.view(-1, 1, 1, 1).float(), torch.tanh(x)], dim=1)


def make_diffusion():
    import diffusion as R
    out = {}
    for tag, resp in (("s250", "250"), ("full", "")):
        d = R.create_diffusion(resp)
        out[f"{tag}_timestep_map"] = np.array(d.timestep_map)
        for k in ("betas", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                  "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                  "posterior_mean_coef1", "posterior_mean_coef2"):
            out[f"{tag}_{k}"] = getattr(d, k)
        g = torch.Generator().manual_seed(7)
        x = torch.randn(4, 4, 28, 28, generator=g)
        n = torch.randn(4, 4, 28, 28, generator=g)
        t = torch.tensor([0, 1, d.num_timesteps // 2, d.num_timesteps - 1])
        pmv = d.p_mean_variance(fake_model, x, t, clip_denoised=False)
        for k in ("mean", "variance", "log_variance", "pred_xstart"):
            out[f"{tag}_pmv_{k}"] = pmv[k].numpy()
        tl = d.training_losses(fake_model, x, t, noise=n)
        for k in ("loss", "mse", "vb"):
            out[f"{tag}_loss_{k}"] = tl[k].numpy()
        out[f"{tag}_q_sample"] = d.q_sample(x, t, noise=n).numpy()
        # p_sample with the noise made explicit: sample = mean + [t!=0] exp(.5 logvar) noise
        nz = (t != 0).float().view(-1, 1, 1, 1)
        out[f"{tag}_p_sample"] = (pmv["mean"] + nz * torch.exp(0.5 * pmv["log_variance"]) * n).numpy()
    np.savez_compressed(os.path.join(HERE, "diffusion.npz"), **out)


def make_ct_encoder():
    from block.CT_encoder import CT_Encoder
    torch.manual_seed(0)
    enc = CT_Encoder(img_size=28, patch_size=2, in_channels=4, embed_dim=512, contain_mask_token=True).eval()
    synth.fill_trained_like_(enc, seed=3)
    x = synth.synthetic_batch(3, seed=5)["x"]
    with torch.no_grad():
        w, y2 = enc(x)
    np.savez_compressed(os.path.join(HERE, "ct_encoder.npz"), weight=w.numpy(), y2_sub=y2.numpy()[:, ::SUB],
                        y2_stats=stats(y2.numpy()), keys=np.array(sorted(enc.state_dict().keys())))


MODEL_CASES = [
    # tag, registry key, use_mamba2, batch
    ("diffma_s2_m1", "DiffMa-S/2", False, 2),
    ("diffma_s2_m2", "DiffMa-S/2", True, 2),
    ("diffma_s4_m1", "DiffMa-S/4", False, 2),
    ("diffma_s7_m1", "DiffMa-S/7", False, 2),
    ("zigma_s4_m1", "ZigMa-S/4", False, 2),
    ("zigma_s4_m2", "ZigMa-S/4", True, 2),
    ("vim_s4_m1", "ViM-S/4", False, 2),
    ("vim_s4_m2", "ViM-S/4", True, 2),
    ("vmamba_s4_m1", "VMamba-S/4", False, 2),
    ("vmamba_s4_m2", "VMamba-S/4", True, 2),
    ("emamba_s2_m1", "EMamba-S/2", False, 1),
]


def make_models():
    import model as M
    for tag, key, m2, batch in MODEL_CASES:
        torch.manual_seed(0)
        net = M.DiffMa_models[key](input_size=28, dt_rank=16, d_state=16, use_mamba2=m2).eval()
        synth.fill_trained_like_(net, seed=11)
        T = net.x_embedder.num_patches
        b = synth.synthetic_batch(batch, tokens=T, seed=21)
        with torch.no_grad():
            out = net(b["x"], b["t"], b["y"], b["y2"], b["w"])
        np.savez_compressed(os.path.join(HERE, f"model_{tag}.npz"), out=out.numpy(),
                            key=np.array(key), use_mamba2=np.array(m2), batch=np.array(batch),
                            state_keys=np.array(list(net.state_dict().keys())),
                            state_shapes=np.array([",".join(map(str, v.shape)) for v in net.state_dict().values()]),
                            pos_embed_sha=np.array(hashlib.sha256(net.pos_embed.numpy().tobytes()).hexdigest()))
        print(tag, tuple(out.shape), float(out.abs().mean()))


def make_mixers():
    """One mixer call in isolation (the unit the fused kernels replace), every scan type."""
    import tools
    from block.mamba import Mamba
    from block.mamba2 import Mamba2
    ml, inv = tools.spiral(14)
    spiral_kw = dict(token_list=ml[2], token_list_reversal=ml[3], origina_list=inv[2], origina_list_reversal=inv[3])
    z = tools.zig(14, 3)
    zig_kw = dict(token_list=z[0], origina_list=z[1])
    vm = tools.vmamba_(14)
    vm_kw = dict(token_list=vm[0], origina_list=vm[1])
    out = {}
    g = torch.Generator().manual_seed(99)
    h = torch.randn(2, 196, 512, generator=g)
    for name, cls in (("m1", Mamba), ("m2", Mamba2)):
        for scan, kw in (("spiral", spiral_kw), ("zigma", zig_kw), ("vim", {}), ("vmamba", vm_kw), ("eff", {})):
            if name == "m2" and scan == "eff":
                continue            # broken in the reference
            torch.manual_seed(0)
            mix = cls(d_model=512, d_state=16, d_conv=4, expand=2, **kw).eval()
            synth.fill_trained_like_(mix, seed=5)
            with torch.no_grad():
                full = mix(h, scan).numpy()
            out[f"{name}_{scan}_sub"] = full[:, ::SUB]          # every SUB-th token, all features
            out[f"{name}_{scan}_stats"] = stats(full)
            print(name, scan, float(np.abs(full).mean()))
    np.savez_compressed(os.path.join(HERE, "mixer.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["orders", "diffusion", "ct", "mixers", "models"]
    if "orders" in which:
        make_scan_orders()
    if "diffusion" in which:
        make_diffusion()
    if "ct" in which:
        make_ct_encoder()
    if "mixers" in which:
        make_mixers()
    if "models" in which:
        make_models()
