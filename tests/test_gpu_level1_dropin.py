"""Level-1 drop-in ON HARDWARE (-m gpu): the reference's own ``model.py`` / ``block/*.py``, byte for byte, executing over
``diffma-diffusion-mamba_b200/shims`` on the GPU, against this package's fast path on the same name-keyed weights.

The reference files are not part of this repository: ``__graft_entry__.build()`` stages them under the git-ignored
``baseline/_ref/`` when the checkout is present (the build container), and they travel to the GPU box with the tree.
Without them the test skips.  INTEGRATION.md section 1 is the claim this test executes: reference sample.py:42-53 /
model.py:264 running unchanged, every mamba_inner_fn / mamba_split_conv1d_scan_combined call landing in the C-ABI.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
SHIMS = os.path.join(ROOT, "diffma-diffusion-mamba_b200", "shims")

SCRIPT = r'''
import sys, torch
import model as ref_model                       # the reference's model.py (baseline/_ref), over the shims
import block.mamba as bm, block.mamba2 as bm2
assert "diffma_b200" in bm.mamba_inner_fn.__module__ and "diffma_b200" in bm2.mamba_split_conv1d_scan_combined.__module__
from diffma_b200 import ops, synth
from diffma_b200.model import DiffMa_models
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
key, use_m2 = sys.argv[1], sys.argv[2] == "1"
torch.manual_seed(0)
ref = ref_model.DiffMa_models[key](input_size=28, dt_rank=16, d_state=16, use_mamba2=use_m2)
ours = DiffMa_models[key](input_size=28, dt_rank=16, d_state=16, use_mamba2=use_m2)
assert sorted(ref.state_dict().keys()) == sorted(ours.state_dict().keys())
synth.fill_trained_like_(ref, seed=11)
synth.fill_trained_like_(ours, seed=11)
ref, ours = ref.to(dev).eval(), ours.to(dev).eval()
b = synth.synthetic_batch(2, tokens=ours.x_embedder.num_patches, seed=21, device=dev)
n0 = ops.LAUNCH_COUNTER["kernels"]
with torch.no_grad():
    o_ref = ref(b["x"], b["t"], b["y"], b["y2"], b["w"])
    n_ref = ops.LAUNCH_COUNTER["kernels"] - n0
    o_ours = ours(b["x"], b["t"], b["y"], b["y2"], b["w"])
torch.cuda.synchronize()
err = (o_ref - o_ours).abs().max().item()
print("RESULT", err, o_ref.abs().max().item(), n_ref)
'''


@pytest.mark.parametrize("key,use_m2", [("DiffMa-S/2", False), ("DiffMa-S/2", True), ("ZigMa-S/4", False)])
def test_reference_model_executes_over_shims_on_gpu(key, use_m2):
    if not os.path.exists(os.path.join(REF, "model.py")):
        pytest.skip("reference sources not staged under baseline/_ref (run __graft_entry__.build() where /root/reference exists)")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([SHIMS, ROOT, REF]))
    r = subprocess.run([sys.executable, "-c", SCRIPT, key, "1" if use_m2 else "0"], capture_output=True, text=True, env=env,
                       cwd="/tmp", timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][-1].split()
    err, scale, launches = float(line[1]), float(line[2]), int(line[3])
    assert launches > 0, "the reference's forward did not reach this package's kernels"
    assert err <= 2e-3 * max(1.0, scale), (err, scale)          # fp32 both sides: two evaluation orders of the same model
