def selective_state_update(*a, **k):
    raise NotImplementedError("decode path is dead code in DiffMa (SURVEY 2.1 row 17)")
