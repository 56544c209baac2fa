"""mamba_ssm.ops.triton.layernorm_gated as imported at reference block/mamba2.py:17."""
from diffma_b200.ops import RMSNormGated as RMSNorm  # noqa: F401
