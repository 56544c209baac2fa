"""Oracle-backed stand-in for mamba_ssm.ops.triton.layernorm_gated (block/mamba2.py:17)."""
import torch

from oracle.ref_ops import rmsnorm_gated_ref


class RMSNorm(torch.nn.Module):
    def __init__(self, hidden_size, eps=1e-5, group_size=None, norm_before_gate=True, device=None, dtype=None):
        super().__init__()
        self.eps = eps
        self.weight = torch.nn.Parameter(torch.ones(hidden_size, device=device, dtype=dtype))
        self.register_parameter("bias", None)
        self.group_size = group_size
        self.norm_before_gate = norm_before_gate

    def forward(self, x, z=None):
        return rmsnorm_gated_ref(x, self.weight, z=z, eps=self.eps, group_size=self.group_size,
                                 norm_before_gate=self.norm_before_gate)
