"""Oracle-backed stand-in for mamba_ssm.ops.selective_scan_interface (imported by block/mamba.py:11)."""
from oracle.ref_ops import mamba_inner_ref, selective_scan_ref


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False):
    return selective_scan_ref(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state)


def mamba_inner_fn(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight,
                   out_proj_bias, A, B=None, C=None, D=None, delta_bias=None, B_proj_bias=None,
                   C_proj_bias=None, delta_softplus=True):
    return mamba_inner_ref(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight,
                           out_proj_weight, out_proj_bias, A, B, C, D, delta_bias, B_proj_bias,
                           C_proj_bias, delta_softplus)
