"""causal_conv1d as imported at reference block/mamba.py:13 and block/mamba2.py:10."""
from diffma_b200.ops import causal_conv1d_fn, causal_conv1d_update  # noqa: F401

__version__ = "1.2.2.post1+diffma_b200"
