"""The drop-in import surface: every name the reference's block/mamba.py:11-23 and block/mamba2.py:9-21 import must
resolve through ``diffma-diffusion-mamba_b200/shims`` (no compute here: that is tests/test_gpu_parity.py)."""
import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIMS = os.path.join(ROOT, "diffma-diffusion-mamba_b200", "shims")

SURFACE = {
    "mamba_ssm.ops.selective_scan_interface": ["selective_scan_fn", "mamba_inner_fn"],
    "causal_conv1d": ["causal_conv1d_fn", "causal_conv1d_update"],
    "mamba_ssm.ops.triton.selective_state_update": ["selective_state_update"],
    "mamba_ssm.ops.triton.layernorm": ["RMSNorm", "layer_norm_fn", "rms_norm_fn"],
    "mamba_ssm.ops.triton.layernorm_gated": ["RMSNorm"],
    "mamba_ssm.distributed.tensor_parallel": ["ColumnParallelLinear", "RowParallelLinear"],
    "mamba_ssm.distributed.distributed_utils": ["all_reduce", "reduce_scatter"],
    "mamba_ssm.ops.triton.ssd_combined": ["mamba_chunk_scan_combined", "mamba_split_conv1d_scan_combined"],
    "timm.models.vision_transformer": ["Attention", "Mlp"],
    "timm.models.layers": ["to_2tuple", "DropPath"],
}


def test_surface_resolves_in_subprocess():
    code = ("import importlib, json, sys\n"
            f"surface = {SURFACE!r}\n"
            "for mod, names in surface.items():\n"
            "    m = importlib.import_module(mod)\n"
            "    assert 'shims' in m.__file__, (mod, m.__file__)\n"
            "    for n in names: assert hasattr(m, n), (mod, n)\n"
            "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([SHIMS, ROOT]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference checkout not present (GPU box)")
def test_reference_model_imports_unchanged_over_shims():
    """The reference's own model.py builds its registry over the product shims (construction only, CPU)."""
    code = ("import model, torch\n"
            "net = model.DiffMa_models['DiffMa-S/2'](input_size=28, dt_rank=16, d_state=16, use_mamba2=False)\n"
            "net2 = model.DiffMa_models['DiffMa-S/2'](input_size=28, dt_rank=16, d_state=16, use_mamba2=True)\n"
            "import block.mamba as bm\n"
            "assert 'diffma_b200' in bm.mamba_inner_fn.__module__\n"
            "print(len(net.state_dict()), len(net2.state_dict()))\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([SHIMS, ROOT, "/root/reference"]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp")
    assert r.returncode == 0, r.stderr[-2000:]


def test_cpu_tensors_fail_loudly():
    """No CPU fallback: the product ops refuse CPU tensors instead of silently computing somewhere else."""
    import torch
    sys.path.insert(0, ROOT)
    from diffma_b200 import ops
    x = torch.zeros(1, 8, 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        ops.mamba_inner_fn(x, torch.zeros(4, 1, 4), None, torch.zeros(36, 4), torch.zeros(4, 4), torch.zeros(2, 4), None,
                           torch.zeros(4, 16))
