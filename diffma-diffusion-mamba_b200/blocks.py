"""adaLN-modulated Mamba blocks with the reference's interface: ``forward(x (N,T,D), c (N,2D), w (N,T,1)) -> (N,T,D)``.

Row a9 of SURVEY.md section 8 ("DiffMaBlock forward() signature").  Constructor kwargs, sub-module and parameter
names follow reference block/mamba_block.py (Spiral :13-130, Zig :137-205, ViM :208-268, VMamba :271-340,
EfficientVMamba :343-397, DiT :400-418) so checkpoints load.  The Spiral block hands BOTH of its mixers to
``mixer.mix_groups`` so their three directions each run in one launch pair.
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F
from torch import nn

from .mixer import Mamba, Mamba2, mix_groups


# DIFFMA_GEMM=tcgen05 routes the dense projections of the fused (inference) block path through the hand-written tcgen05
# GEMM (dm_gemm_bf16_tn_ex): in-projection with the SiLU(z) gate in its epilogue, out-projection with the CrossMerge
# direction sum as its A producer (K = d_inner instead of 3 * d_inner), attention_network[1] with its bias in the
# epilogue.  Default is the library GEMM: these projections are plain GEMMs with K <= 3072 over 3 136 rows and are bound
# by L2 -> SM operand traffic, where cuBLAS' 256x192 two-CTA tiles move ~22 % fewer bytes per output than our
# persistent 128x256 one-CTA tiles; measured r02 (profiles/r02_notes.md): 19.2 vs 14.2 us (in), 19.8 vs 15.5 us (out,
# even with 3x fewer flops), 9.0 vs 5.6 us (attention Linear) -- and the scan's gain from the hoisted gate (131 -> 125 us)
# does not pay for the difference.
_LN_FOLD = os.environ.get("DIFFMA_LN_FOLD", "1") != "0"     # inference: attention LayerNorm folded around its Linear
_USE_TCGEN05_GEMM = os.environ.get("DIFFMA_GEMM", "cublas") == "tcgen05"
# DIFFMA_FUSED_TRAIN=0: differentiate the block glue op by op with torch autograd (the module path below) instead of the
# fused row kernels + their hand-written adjoints (A/B runs and the gradient parity tests)
_FUSED_TRAIN = os.environ.get("DIFFMA_FUSED_TRAIN", "1") != "0"


def modulate(x, shift, scale):
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def _basic_init(module):
    if isinstance(module, nn.Linear):
        torch.nn.init.xavier_uniform_(module.weight)
        if module.bias is not None:
            nn.init.constant_(module.bias, 0)


def _mixer(use_mamba2, D_dim, d_state, **orders):
    cls = Mamba2 if use_mamba2 else Mamba
    return cls(d_model=D_dim, d_state=d_state, d_conv=4, expand=2, **orders)


class Spiral_MambaBlock(nn.Module):
    def __init__(self, D_dim, E_dim, dt_rank, dim_inner, d_state, token_list, token_list_reversal, origina_list,
                 origina_list_reversal, use_mamba2):
        super().__init__()
        self.D_dim, self.E_dim, self.dt_rank, self.dim_inner, self.d_state = D_dim, E_dim, dt_rank, dim_inner, d_state
        orders = dict(token_list=token_list, token_list_reversal=token_list_reversal, origina_list=origina_list,
                      origina_list_reversal=origina_list_reversal)
        self.norm1 = nn.LayerNorm(D_dim)
        self.mamba1 = _mixer(use_mamba2, D_dim, d_state, **orders)
        self.mamba2 = _mixer(use_mamba2, D_dim, d_state, **orders)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(D_dim * 2, D_dim * 3, bias=True))
        self.attention_network = nn.Sequential(nn.LayerNorm(2 * D_dim), nn.Linear(2 * D_dim, D_dim, bias=True),
                                               nn.SiLU(), nn.Linear(D_dim, 1, bias=True), nn.Sigmoid())
        self.sigmoid = nn.Sigmoid()
        self.initialize_weights()

    def forward(self, x, c, w, skip=None):
        """``skip`` (optional, not in the reference signature): the long-skip tensor model.py:290-292 adds to the
        block input; passing it here lets the fused prologue / epilogue do the add."""
        if (not torch.is_grad_enabled() and x.is_cuda and x.dtype == torch.float32 and x.shape[-1] == 512
                and x.is_contiguous()):
            return self._forward_fused(x, c, w, skip)
        if (_FUSED_TRAIN and torch.is_grad_enabled() and x.is_cuda and x.dtype == torch.float32 and x.shape[-1] == 512
                and w is not None and not w.requires_grad):
            # (the patch embedding hands over a transposed view, and elementwise results inherit its strides: one copy
            # here puts the whole residual stream into the row-major layout the row kernels read)
            return self._forward_train_fused(x.contiguous(), c, w, skip)
        if skip is not None:
            x = x + skip
        shift, scale, gate = self.adaLN_modulation(c).chunk(3, dim=1)
        x_ssm = modulate(self.norm1(x), shift, scale)
        w_ssm = x_ssm * w
        a, b = mix_groups([self.mamba1, self.mamba2], [x_ssm, w_ssm], "spiral")
        alpha = self.attention_network(torch.cat([a, b], dim=-1))
        return x + gate.unsqueeze(1) * (alpha * a + (1 - alpha) * b)

    # ---- training path: the same row kernels as inference with autograd Functions around them (their adjoints live in
    #      csrc/dm_block_bwd.cu); the mixers go through mix_groups (autograd of dm_mamba1_scan_bwd) ---------------------
    def _forward_train_fused(self, x, c, w, skip=None):
        from .autograd_ops import SpiralPostFn, SpiralPreFn
        from .mixer import _act_dtype
        B, L, D = x.shape
        act = _act_dtype(x)
        mod = self.adaLN_modulation(c).float()                                                  # (B, 3D) = [shift | scale | gate]
        wrow = None if w is None else w.detach().reshape(B * L).float().contiguous()
        sk = None if skip is None else skip.contiguous()
        x2 = SpiralPreFn.apply(x, sk, self.norm1.weight, self.norm1.bias, mod, wrow, act)      # (2, B*L, D)
        a, b = mix_groups([self.mamba1, self.mamba2], [x2[0].view(B, L, D), x2[1].view(B, L, D)], "spiral")
        ab = torch.stack([a.reshape(B * L, D), b.reshape(B * L, D)]).to(act)
        an = self.attention_network
        return SpiralPostFn.apply(x, sk, ab, an[0].weight, an[0].bias, an[1].weight, an[1].bias, an[3].weight, an[3].bias, mod)

    # ---- inference path: 8 launches per block (reference: ~100), weights cached in the act dtype ----------
    def _fused_weights(self, act):
        m1, m2 = self.mamba1, self.mamba2
        from . import ops
        params = [m1.in_proj.weight, m2.in_proj.weight, m1.out_proj.weight, m2.out_proj.weight,
                  self.adaLN_modulation[1].weight, self.adaLN_modulation[1].bias, self.attention_network[1].weight,
                  self.attention_network[1].bias, self.attention_network[3].weight, self.attention_network[3].bias,
                  self.norm1.weight, self.norm1.bias, self.attention_network[0].weight, self.attention_network[0].bias]
        key = ops.weights_key(params, act, str(params[0].device))
        cache = getattr(self, "_fcache", None)
        if cache is not None and cache["key"] == key:
            return cache
        K = 3
        cache = {
            "key": key,
            "w_in": torch.stack([m1.in_proj.weight, m2.in_proj.weight]).to(act).transpose(1, 2).contiguous(),
            "w_out": torch.stack([m1.out_proj.weight.repeat(1, K), m2.out_proj.weight.repeat(1, K)]).to(act)
                          .transpose(1, 2).contiguous(),
            "w_in_nk": torch.stack([m1.in_proj.weight, m2.in_proj.weight]).to(act).contiguous(),          # (2, N, K)
            "w_out_nk": torch.stack([m1.out_proj.weight, m2.out_proj.weight]).to(act).contiguous(),       # (2, N, d_inner)
            "att_b32": self.attention_network[1].bias.detach().float().reshape(1, -1).contiguous(),
            "ada_w": self.adaLN_modulation[1].weight.to(act).contiguous(),
            "ada_b": self.adaLN_modulation[1].bias.to(act).contiguous(),
            "att_w": self.attention_network[1].weight.to(act).contiguous(),
            "att_b": self.attention_network[1].bias.to(act).contiguous(),
            "w3": self.attention_network[3].weight.detach().float().reshape(-1).contiguous(),
            "b3": self.attention_network[3].bias.detach().float().reshape(-1).contiguous(),
            "ln1": (self.norm1.weight.detach().float().contiguous(), self.norm1.bias.detach().float().contiguous()),
            "ln2": (self.attention_network[0].weight.detach().float().contiguous(),
                    self.attention_network[0].bias.detach().float().contiguous()),
        }
        # attention LayerNorm folded around attention_network[1] (dm_spiral_post_mix_fold): W' = W * gamma in the act dtype,
        # split into the a- and b-halves of its input; colsum from the ROUNDED W' (what the GEMM multiplies by)
        with torch.no_grad():
            an = self.attention_network
            Dm = an[1].weight.shape[0]
            wf = (an[1].weight.float() * an[0].weight.float()[None, :]).to(act)                    # (D, 2D)
            cache["att_wf"] = torch.stack([wf[:, :Dm].t(), wf[:, Dm:].t()]).contiguous()           # (2, K = D, N = D)
            cache["att_colsum"] = wf.float().sum(1).contiguous()
            cache["att_cvec"] = (an[1].weight.float() @ an[0].bias.float() + an[1].bias.float()).contiguous()
            cache["ln2_eps"] = float(an[0].eps)
        self._fcache = cache
        return cache

    def _m2_out_weights(self, act, nk=False):
        """(2, d_inner, d_model) [nk: (2, d_model, d_inner)]: out_proj weight with the gated-RMSNorm weight folded in."""
        m1, m2 = self.mamba1, self.mamba2
        from . import ops
        ps = [m1.out_proj.weight, m2.out_proj.weight, m1.norm.weight, m2.norm.weight]
        key = ops.weights_key(ps, act, str(ps[0].device))
        c = getattr(self, "_m2cache", None)
        if c is None or c[0] != key:
            w = torch.stack([(m.out_proj.weight.float() * m.norm.weight.float()[None, :]) for m in (m1, m2)])
            c = (key, w.to(act).transpose(1, 2).contiguous(), w.to(act).contiguous())
            self._m2cache = c
        return c[2] if nk else c[1]

    def _forward_fused(self, x, c, w, skip=None, mod=None):
        from . import ops
        from .mixer import _act_dtype
        B, L, D = x.shape
        act = _act_dtype(x)
        W = self._fused_weights(act)
        with torch.autocast("cuda", enabled=False):
            if mod is None:
                mod = F.linear(F.silu(c.float()).to(act), W["ada_w"], W["ada_b"]).float()       # (B, 3D)
            wrow = None if w is None else w.reshape(B * L).float().contiguous()
            x2 = ops.spiral_pre(x, skip, W["ln1"][0], W["ln1"][1], mod, wrow, act)            # (2, B*L, D)
            ab, hidden = self._fused_core(x2, B, L, act)
            return self._post_mix(x, skip, ab, hidden, W, mod)

    def _post_mix(self, x, skip, ab, hidden, W, mod, pre=None):
        """Close the block: sigmoid mix + gated residual [+ the next block's / final layer's LN + modulate, ``pre`` =
        (skip_next, ln_weight, ln_bias, mod_next, w, eps)].  ``hidden`` is what ``_fused_core`` returned: the attention
        Linear's output, or -- LayerNorm folded (``_LN_FOLD``) -- the pair of raw products (2, B*L, D)."""
        from . import ops
        if hidden.dim() == 3:
            return ops.spiral_post_mix_fold(x, skip, ab, hidden, W["att_colsum"], W["att_cvec"], W["ln2_eps"], W["w3"],
                                            W["b3"], mod, pre=pre)
        if pre is None:
            return ops.spiral_post_mix(x, skip, ab, hidden, W["w3"], W["b3"], mod)
        skip_next, ln_w, ln_b, mod_next, w, eps = pre
        return ops.spiral_post_mix_pre(x, skip, ab, hidden, W["w3"], W["b3"], mod, skip_next, ln_w, ln_b, mod_next, w,
                                       eps=eps)

    def _fused_core(self, x2, B, L, act):
        """The part of the block between the two row kernels: in_proj GEMM -> all directions of both mixers (one scan
        launch pair) -> merge + out_proj GEMM -> LN(cat(a, b)) -> attention_network[1].  x2 (2, B*L, D) is
        ``spiral_pre``'s output; returns (ab (2, B*L, D), hidden (B*L, D)).  ``DiffMa.forward`` calls this directly so
        that one row kernel (``spiral_post_mix_pre``) can close block i and open block i+1."""
        from . import ops
        from .mixer import Mamba2
        W = self._fused_weights(act)
        m1, m2 = self.mamba1, self.mamba2
        is_m2 = isinstance(m1, Mamba2)
        D = x2.shape[-1]
        with torch.autocast("cuda", enabled=False):
            tc = act == torch.bfloat16 and _USE_TCGEN05_GEMM       # hand-written tcgen05 GEMM (dm_gemm_bf16_tn_ex)
            gated = tc and not is_m2                               # Mamba-1: the in-projection's epilogue emits silu(z)
            if tc:
                proj = ops.gemm_bf16_tn(x2, W["w_in_nk"], silu_from=m1.d_inner if gated else None)   # (2, B*L, d_in_proj)
            else:
                proj = torch.bmm(x2, W["w_in"])
            plan = m1.plan("spiral", L, x2.device)
            xs = [proj[0].view(B, L, -1), proj[1].view(B, L, -1)]
            if not is_m2:
                y = ops.mamba1_scan(xs, [m1.scan_weights(act), m2.scan_weights(act)], plan, z_gated=gated)   # (2, B, L, K, d_inner)
                if tc:      # merge of the K directions in the GEMM's A producer: contraction over d_inner, not K * d_inner
                    ab = ops.gemm_bf16_tn(y.view(2, B * L, plan.n_dir, -1), W["w_out_nk"])      # (2, B*L, D)
                else:
                    ab = torch.bmm(y.view(2, B * L, -1), W["w_out"])
            elif tc:
                v, ss = ops.mamba2_ssd(xs, [m1.scan_weights(), m2.scan_weights()], plan, m1.d_inner, m1.d_state,
                                       m1.nheads, gate=True, want_sumsq=True)                     # (2,B,L,K,d), (2,B,K,L)
                K = plan.n_dir
                rstd = torch.rsqrt(ss / m1.d_inner + m1.norm.eps).transpose(2, 3).reshape(2, B * L * K).contiguous()
                o = ops.gemm_bf16_tn(v.view(2, B * L * K, -1), self._m2_out_weights(act, nk=True), row_scale=rstd)
                ab = o.view(2, B * L, K, D).sum(2)
            else:
                v, ss = ops.mamba2_ssd(xs, [m1.scan_weights(), m2.scan_weights()], plan, m1.d_inner, m1.d_state,
                                       m1.nheads, gate=True, want_sumsq=True)                     # (2,B,L,K,d), (2,B,K,L)
                # gated RMSNorm + merge + out_proj: rstd is a per-(token, direction) scalar and the norm weight a
                # per-channel one, so  sum_k rstd_k (v_k * w_norm) W^T = sum_k rstd_k * (v_k (W * w_norm)^T):
                # the weight is folded into W_out once (cached) and rstd applied to the 512-wide GEMM output.
                K = plan.n_dir
                o = torch.bmm(v.view(2, B * L * K, -1), self._m2_out_weights(act))               # (2, B*L*K, D)
                rstd = torch.rsqrt(ss / m1.d_inner + m1.norm.eps).transpose(2, 3)                # (2, B, L, K)
                ab = (o.view(2, B * L, K, D) * rstd.reshape(2, B * L, K, 1).to(act)).sum(2)
            if _LN_FOLD and not tc:
                # LayerNorm folded around the Linear: two K = D products on the raw a, b (fp32 out: the row kernel subtracts
                # mean * colsum from them); the row kernel that follows supplies mean / rstd
                return ab, torch.bmm(ab, W["att_wf"], out_dtype=torch.float32) if act != torch.float32 else torch.bmm(ab, W["att_wf"])
            lnab = ops.spiral_post_ln(ab, W["ln2"][0], W["ln2"][1])                              # (B*L, 2D)
            if tc:
                hidden = ops.gemm_bf16_tn(lnab.unsqueeze(0), W["att_w"].unsqueeze(0), bias=W["att_b32"])[0]
            else:
                hidden = F.linear(lnab, W["att_w"], W["att_b"])                                  # (B*L, D)
            return ab, hidden

    def initialize_weights(self):
        self.apply(_basic_init)
        for i in (1, 3):
            nn.init.constant_(self.attention_network[i].weight, 0)
            nn.init.constant_(self.attention_network[i].bias, 0)


class _SingleMixerBlock(nn.Module):
    scan_type = ""

    def __init__(self, D_dim, E_dim, dt_rank, dim_inner, d_state, use_mamba2, **orders):
        super().__init__()
        self.D_dim, self.E_dim, self.dt_rank, self.dim_inner, self.d_state = D_dim, E_dim, dt_rank, dim_inner, d_state
        self.norm1 = nn.LayerNorm(D_dim)
        self.mamba = _mixer(use_mamba2, D_dim, d_state, **orders)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(D_dim * 2, D_dim * 3, bias=True))
        self.apply(_basic_init)

    def forward(self, x, c, w):
        shift, scale, gate = self.adaLN_modulation(c).chunk(3, dim=1)
        x_ssm = modulate(self.norm1(x), shift, scale)
        return x + gate.unsqueeze(1) * self.mamba(x_ssm, self.scan_type)


class Zig_MambaBlock(_SingleMixerBlock):
    scan_type = "zigma"

    def __init__(self, D_dim, E_dim, dt_rank, dim_inner, d_state, token_list, origina_list, use_mamba2):
        super().__init__(D_dim, E_dim, dt_rank, dim_inner, d_state, use_mamba2, token_list=token_list,
                         origina_list=origina_list)


class ViM_MambaBlock(_SingleMixerBlock):
    scan_type = "vim"

    def __init__(self, D_dim, E_dim, dt_rank, dim_inner, d_state, use_mamba2):
        super().__init__(D_dim, E_dim, dt_rank, dim_inner, d_state, use_mamba2)


class VMamba_MambaBlock(_SingleMixerBlock):
    scan_type = "vmamba"

    def __init__(self, D_dim, E_dim, dt_rank, dim_inner, d_state, token_list, origina_list, use_mamba2):
        super().__init__(D_dim, E_dim, dt_rank, dim_inner, d_state, use_mamba2, token_list=token_list,
                         origina_list=origina_list)


class EfficientVMamba_MambaBlock(_SingleMixerBlock):
    scan_type = "eff"

    def __init__(self, D_dim, E_dim, dt_rank, dim_inner, d_state, use_mamba2):
        super().__init__(D_dim, E_dim, dt_rank, dim_inner, d_state, use_mamba2)


class _Attention(nn.Module):
    """timm-style multi-head self-attention (names qkv / proj as in timm.models.vision_transformer.Attention)."""

    def __init__(self, dim, num_heads=8, qkv_bias=True):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        q, k, v = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        return self.proj(F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, C))


class _Mlp(nn.Module):
    def __init__(self, in_features, hidden_features):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = nn.GELU(approximate="tanh")
        self.fc2 = nn.Linear(hidden_features, in_features)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class DiTBlock(nn.Module):
    """The paper's transformer baseline (reference block/mamba_block.py:400-418); not on the Mamba hot path."""

    def __init__(self, hidden_size, num_heads, mlp_ratio=4.0, **block_kwargs):
        super().__init__()
        self.norm1 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.attn = _Attention(hidden_size, num_heads=num_heads, qkv_bias=True)
        self.norm2 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.mlp = _Mlp(hidden_size, int(hidden_size * mlp_ratio))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size * 2, 6 * hidden_size, bias=True))

    def forward(self, x, c, w):
        s1, sc1, g1, s2, sc2, g2 = self.adaLN_modulation(c).chunk(6, dim=1)
        x = x + g1.unsqueeze(1) * self.attn(modulate(self.norm1(x), s1, sc1))
        return x + g2.unsqueeze(1) * self.mlp(modulate(self.norm2(x), s2, sc2))
