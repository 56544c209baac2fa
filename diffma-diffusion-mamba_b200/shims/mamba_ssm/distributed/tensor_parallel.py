"""mamba_ssm.distributed.tensor_parallel as imported at reference block/mamba2.py:18 (process_group is always None)."""
import torch


class ColumnParallelLinear(torch.nn.Linear):
    def __init__(self, *a, process_group=None, sequence_parallel=True, **k):
        raise NotImplementedError("diffma_b200: tensor parallelism is dead code in DiffMa")


RowParallelLinear = ColumnParallelLinear
