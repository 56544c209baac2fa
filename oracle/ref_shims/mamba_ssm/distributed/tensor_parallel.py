import torch


class ColumnParallelLinear(torch.nn.Linear):
    def __init__(self, *a, process_group=None, sequence_parallel=True, **k):
        raise NotImplementedError("tensor parallelism is dead code in DiffMa (process_group is always None)")


RowParallelLinear = ColumnParallelLinear
