"""mamba_ssm.ops.triton.ssd_combined as imported at reference block/mamba2.py:20-21."""
from diffma_b200.ops import mamba_chunk_scan_combined, mamba_split_conv1d_scan_combined  # noqa: F401
