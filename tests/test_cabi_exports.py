"""The C-ABI library loads and exports every symbol ``include/diffma_b200.h`` declares (no compute: CPU-only box)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "diffma_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dm_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from diffma_b200 import _cabi
    lib = _cabi.lib()
    names = _declared()
    assert "dm_mamba1_scan_fwd" in names and "dm_mamba2_ssd_fwd" in names and "dm_spiral_pre" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(_cabi.EXPORTS) == names, (sorted(_cabi.EXPORTS), names)
    assert lib.dm_version() == _cabi.DM_ABI_VERSION
    assert b"sm_100a" in lib.dm_build_info()


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors of the argument structs: field counts / sizes as the C compiler lays them out."""
    from diffma_b200 import _cabi
    assert ctypes.sizeof(_cabi.Mamba1Group) == 16 * 8
    assert ctypes.sizeof(_cabi.Mamba1Args) == 10 * 4 + 8 + 4 * ctypes.sizeof(_cabi.Mamba1Group) + 16   # + sched workspace ptr/size (ABI 3)
    assert ctypes.sizeof(_cabi.Mamba2Group) == 15 * 8
    assert ctypes.sizeof(_cabi.Mamba2Args) == 11 * 4 + 4 + 8 + 4 * ctypes.sizeof(_cabi.Mamba2Group)


def test_invalid_arguments_return_status_not_crash():
    from diffma_b200 import _cabi
    lib = _cabi.lib()
    assert lib.dm_mamba1_scan_fwd(None, None) == _cabi.DM_ERR_INVALID_ARG
    a = _cabi.Mamba1Args()
    assert lib.dm_mamba1_scan_fwd(ctypes.byref(a), None) == _cabi.DM_ERR_INVALID_ARG
    a.batch = a.n_dir = a.seqlen = a.n_groups = 1
    a.out_order, a.act_dtype, a.d_state, a.d_conv, a.dt_rank, a.d_inner = 0, 7, 16, 4, 32, 1024
    assert lib.dm_mamba1_scan_fwd(ctypes.byref(a), None) == _cabi.DM_ERR_UNSUPPORTED       # unknown dtype
    a.act_dtype, a.d_state = _cabi.DM_BF16, 64
    assert lib.dm_mamba1_scan_fwd(ctypes.byref(a), None) == _cabi.DM_ERR_UNSUPPORTED       # d_state this build lacks
    a.d_state = 16
    assert lib.dm_mamba1_scan_fwd(ctypes.byref(a), None) == _cabi.DM_ERR_INVALID_ARG       # null group pointers
    assert lib.dm_status_string(_cabi.DM_ERR_UNSUPPORTED).startswith(b"unsupported")
