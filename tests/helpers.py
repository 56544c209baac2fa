"""Shared helpers for the parity tests (builds the reference-shaped state dicts without the reference)."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

REGISTRY_CFG = {
    "DiffMa": "spiral", "ZigMa": "zig", "ViM": "vim", "VMamba": "vmamba", "EMamba": "efficientVMamba",
}
DEPTH = {"S": 4, "B": 8, "L": 16, "XL": 28, "XXL": 56, "BL": 13}


def cfg_of(key, use_mamba2=False):
    fam, rest = key.split("-")
    size, patch = rest.split("/")
    return {"depth": DEPTH[size], "patch_size": int(patch), "block_type": REGISTRY_CFG[fam],
            "use_mamba2": bool(use_mamba2)}


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def stats(a):
    a = np.asarray(a, dtype=np.float64)
    return np.array([a.sum(), np.abs(a).sum(), (a * a).sum()])
