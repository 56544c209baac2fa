"""mamba_ssm.distributed.distributed_utils as imported at reference block/mamba2.py:19 (never called)."""


def all_reduce(x, process_group=None):
    raise NotImplementedError("diffma_b200: dead code in DiffMa (process_group is always None)")


reduce_scatter = all_reduce
