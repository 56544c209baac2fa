"""diffma_b200 -- B200-native (sm_100a) implementation of DiffMa's Mamba selective-scan hot path.

Sub-modules (imported lazily by callers; importing this package does not need a GPU):

* ``scan_orders``  spiral / zig / vmamba token orders (bit-exact with reference tools.py)
* ``_cabi``        ctypes binding of ``libdiffma_b200.so`` (the C-ABI in ``include/diffma_b200.h``)
* ``ops``          torch-facing operators with the upstream ``mamba_ssm`` signatures
* ``mixer``        ``Mamba`` / ``Mamba2`` with the reference's ctor + ``forward(h, scan_type)``
* ``blocks``, ``model``, ``ct_encoder``, ``diffusion``  host-side mirror of the callers
* ``shims``        drop-in ``mamba_ssm`` / ``causal_conv1d`` / ``timm`` import surface
"""
__version__ = "0.1.0"
