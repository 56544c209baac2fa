"""The JSON lines bench.py printed on the B200 (committed under profiles/) keep the driver's contract: required keys,
roofline / cpu_baseline / e2e objects, internally consistent numbers.  CPU-only (reads the committed files)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    with open(path) as f:
        rows = [l for l in f.read().splitlines() if l.startswith("{")]
    return json.loads(rows[-1])


@pytest.mark.parametrize("name", ["r01_bench_m1.json", "r01_bench_m2.json", "r01_bench_L2_b32_L784.json"])
def test_bench_line_contract(name):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in d, k
    assert d["metric"] == "diffusion_step_images_per_s" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "bf16"
    assert "workload" in d["config"] and d["warmup"] >= 3 and d["gpu_launches"] > 0
    batch = d["config"]["per_gpu_batch"] * d["n_gpus"]
    assert abs(d["value"] - batch / (d["ms_per_step"] * 1e-3)) <= 1e-3 * d["value"]          # value = images / step time
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert 0.5 * d["value"] < e["value"] <= 1.02 * d["value"]                                  # host copies cannot speed it up
    c = d["clocks"]
    assert c["sm_mhz"] <= c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 2e-3
    t = r["launch_us"][r["kernel"].replace("_kernel", "_kernel")] * 1e-6
    assert abs(r["achieved"] - r["token_scans_per_launch"] * r["algorithmic_bytes_per_token_scan"] / t / 1e9) < 0.01 * r["achieved"]
    assert r["traffic"] is None or r["traffic"] > 0
    assert 0 < r["mufu"]["frac"] <= 1.0


def test_headline_line_has_cpu_baseline_and_upstream_comparator():
    d = _line("r01_bench_m1.json")
    b = d["cpu_baseline"]
    assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["value"] > 0 and b["unit"] == d["unit"] and b["sample"]
    assert d["value"] > 1000 * b["value"]                 # the GPU path is not the CPU path in disguise
    u = d["roofline"]["upstream_cuda_scan"]
    assert u and u["launch_us"] > 5 * d["roofline"]["launch_us"]["m1_scan_kernel"]


def test_reference_arm_line_contract():
    d = _line("r01_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == "diffusion_step_images_per_s" and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"
