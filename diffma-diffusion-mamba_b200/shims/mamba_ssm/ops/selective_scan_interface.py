"""mamba_ssm.ops.selective_scan_interface as imported at reference block/mamba.py:11."""
from diffma_b200.ops import mamba_inner_fn, selective_scan_fn  # noqa: F401
