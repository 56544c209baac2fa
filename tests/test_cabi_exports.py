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
    assert ctypes.sizeof(_cabi.Mamba1Group) == 18 * 8            # + chunk_states (ABI 4), delta (ABI 5)
    assert ctypes.sizeof(_cabi.Mamba1Args) == 10 * 4 + 8 + 4 * ctypes.sizeof(_cabi.Mamba1Group) + 16 + 8   # + sched workspace ptr/size (ABI 3), z_is_gated (ABI 5)
    assert ctypes.sizeof(_cabi.GemmArgs) == 13 * 8 + 6 * 4                                              # 104 + 24 = 128
    assert ctypes.sizeof(_cabi.Mamba2Group) == 15 * 8
    assert ctypes.sizeof(_cabi.AdamwArgs) == 8 * 8 + 7 * 8                                               # ABI 6
    assert ctypes.sizeof(_cabi.SpiralFoldArgs) == 176                                                   # ABI 6
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


def test_new_entry_points_reject_bad_arguments():
    """ABI 3/4 additions: scheduler workspace sizing / alignment, backward chunking + checkpoints, fused row kernel."""
    from diffma_b200 import _cabi
    assert ctypes.sizeof(_cabi.Mamba1BwdGroup) == 12 * 8
    lib = _cabi.lib()
    assert lib.dm_version() == _cabi.DM_ABI_VERSION
    # workspace: 64 B header + 64 queue slots and 4 KB of state per (sequence, 64-channel) unit
    units = 2 * 16 * 3 * (1024 // 64)
    need = lib.dm_mamba1_sched_workspace_bytes(16, 3, 1024, 2)
    assert need >= units * 4096 + units * 64 * 4 and need % 256 == 0
    assert lib.dm_mamba1_sched_workspace_bytes(0, 3, 1024, 2) == 0
    assert lib.dm_mamba1_bwd_chunk_tokens() in (4, 8)
    a = _cabi.Mamba1Args()
    a.batch = a.n_dir = a.seqlen = a.n_groups = 1
    a.out_order, a.act_dtype, a.d_state, a.d_conv, a.dt_rank, a.d_inner = 0, _cabi.DM_BF16, 16, 4, 32, 1024
    a.sched_workspace, a.sched_workspace_bytes = 8, 1 << 20          # misaligned scratch pointer
    assert lib.dm_mamba1_scan_fwd(ctypes.byref(a), None) == _cabi.DM_ERR_INVALID_ARG
    # fused row kernel: null pointers / unsupported width are statuses, not crashes
    assert lib.dm_spiral_post_mix_pre(None, None, None, None, None, None, None, 0, None, None, None, None, None, 0, None,
                                      None, 1, 1, 512, 1e-5, _cabi.DM_BF16, None) == _cabi.DM_ERR_INVALID_ARG
    p = ctypes.c_void_p(1 << 12)
    assert lib.dm_spiral_post_mix_pre(p, None, p, p, p, p, p, 1536, p, None, p, p, p, 1536, None, p, 1, 1, 384, 1e-5,
                                      _cabi.DM_BF16, None) == _cabi.DM_ERR_UNSUPPORTED


def test_optimizer_entry_point_rejects_bad_arguments():
    """ABI 5: dm_adamw_ema_step -- null / misaligned buffers and out-of-range coefficients are statuses, not crashes."""
    from diffma_b200 import _cabi
    lib = _cabi.lib()
    p = ctypes.c_void_p(1 << 12)
    args = (1e-4, 0.9, 0.999, 1e-8, 0.0, 0.9999, 1.0, None)
    assert lib.dm_adamw_ema_step(None, p, p, p, p, p, 16, *args) == _cabi.DM_ERR_INVALID_ARG
    assert lib.dm_adamw_ema_step(p, p, p, p, p, p, 0, *args) == _cabi.DM_ERR_INVALID_ARG
    assert lib.dm_adamw_ema_step(ctypes.c_void_p((1 << 12) + 4), p, p, p, p, p, 16, *args) == _cabi.DM_ERR_INVALID_ARG
    assert lib.dm_adamw_ema_step(p, p, p, p, p, p, 16, 1e-4, 1.5, 0.999, 1e-8, 0.0, 0.9999, 1.0, None) == _cabi.DM_ERR_INVALID_ARG
