"""diffma_b200 -- B200-native (sm_100a) implementation of DiffMa's Mamba selective-scan hot path.

Sub-modules (imported lazily by callers; importing this package does not need a GPU):

* ``scan_orders``  spiral / zig / vmamba token orders (bit-exact with reference tools.py)
* ``_cabi``        ctypes binding of ``lib/libdiffma_b200.so`` (the C-ABI in ``include/diffma_b200.h``)
* ``build``        nvcc recipe for that library (sm_100a only)
* ``ops``          torch-facing operators, incl. the upstream ``mamba_ssm`` signatures
* ``mixer``        ``Mamba`` / ``Mamba2`` with the reference's ctor + ``forward(h, scan_type)``
* ``blocks``, ``model``, ``ct_encoder``, ``diffusion``  host-side mirror of the callers
* ``shims``        drop-in ``mamba_ssm`` / ``causal_conv1d`` / ``timm`` import surface
"""
__version__ = "0.1.0"


def create_model_and_diffusion(name: str = "DiffMa-B/2", input_size: int = 28, use_mamba2: bool = False,
                               respacing: str = "250", d_state: int = 16, dt_rank: int = 16):
    """The two objects the reference's ``train.py:130-155`` / ``sample.py:42-53`` build."""
    from .diffusion import create_diffusion
    from .model import DiffMa_models
    net = DiffMa_models[name](input_size=input_size, dt_rank=dt_rank, d_state=d_state, use_mamba2=use_mamba2)
    diffusion = create_diffusion(respacing)
    net.t_table_rows = max(1000, int(getattr(diffusion, "original_num_steps", 1000)))   # integer-timestep embedding table
    return net, diffusion
