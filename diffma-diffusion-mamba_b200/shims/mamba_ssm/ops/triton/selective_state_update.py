"""mamba_ssm.ops.triton.selective_state_update as imported at reference block/mamba.py:17 (decode: dead code)."""


def selective_state_update(*a, **k):
    raise NotImplementedError("diffma_b200: single-token decode is never used by a diffusion model")
