"""Oracle-backed stand-in for mamba_ssm.ops.triton.ssd_combined (block/mamba2.py:20-21)."""
from oracle.ref_ops import mamba_chunk_scan_combined_ref, mamba_split_conv1d_scan_ref


def mamba_chunk_scan_combined(x, dt, A, B, C, chunk_size, D=None, z=None, dt_bias=None, initial_states=None,
                              seq_idx=None, cu_seqlens=None, dt_softplus=False, dt_limit=(0.0, float("inf")),
                              return_final_states=False, return_varlen_states=False):
    assert initial_states is None and seq_idx is None and not return_final_states
    return mamba_chunk_scan_combined_ref(x, dt, A, B, C, chunk_size, D, z, dt_bias, dt_softplus, dt_limit)


def mamba_split_conv1d_scan_combined(zxbcdt, conv1d_weight, conv1d_bias, dt_bias, A, D, chunk_size,
                                     initial_states=None, seq_idx=None, dt_limit=(0.0, float("inf")),
                                     return_final_states=False, activation="silu", rmsnorm_weight=None,
                                     rmsnorm_eps=1e-6, outproj_weight=None, outproj_bias=None, headdim=None,
                                     ngroups=1, norm_before_gate=True):
    assert initial_states is None and seq_idx is None and not return_final_states
    return mamba_split_conv1d_scan_ref(zxbcdt, conv1d_weight, conv1d_bias, dt_bias, A, D, chunk_size,
                                       dt_limit, activation, rmsnorm_weight, rmsnorm_eps, outproj_weight,
                                       outproj_bias, headdim, ngroups, norm_before_gate)
