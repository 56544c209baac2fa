"""Drop-in for the mamba-ssm wheel (2.0.4 surface used by DiffMa), backed by diffma_b200 CUDA kernels."""
__version__ = "2.0.4+diffma_b200"
