"""Oracle-backed stand-in for causal_conv1d (block/mamba.py:13, block/mamba2.py:10)."""
from oracle.ref_ops import causal_conv1d_ref


def causal_conv1d_fn(x, weight, bias=None, seq_idx=None, initial_states=None, return_final_states=False,
                     final_states_out=None, activation=None):
    assert seq_idx is None and initial_states is None and not return_final_states
    return causal_conv1d_ref(x, weight, bias, activation)


def causal_conv1d_update(*a, **k):
    raise NotImplementedError("decode path is dead code in DiffMa")
