def all_reduce(x, process_group=None):
    raise NotImplementedError("dead code in DiffMa (process_group is always None)")


reduce_scatter = all_reduce
