"""Host-side algebra of the inference fast path, checked on CPU against plain torch modules (no kernels involved).

* the attention LayerNorm folded around its Linear (blocks._fused_weights: att_wf / att_colsum / att_cvec, consumed by
  dm_spiral_post_mix_fold): rstd * (a W'_a^T + b W'_b^T - mean * colsum) + cvec == Linear(LayerNorm(cat(a, b)))
  (reference block/mamba_block.py:110);
* the patch-embedding / timestep tables dm_step_head consumes (model._patch_tables, model._t_table) against the modules they
  replace (reference model.py:264-276)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _net():
    from diffma_b200 import model as M, synth
    torch.manual_seed(0)
    net = M.DiffMa_models["DiffMa-S/4"](input_size=28, dt_rank=16, d_state=16, use_mamba2=False).eval()
    synth.fill_trained_like_(net, seed=11)
    return net


def test_folded_attention_layernorm_weights_reproduce_ln_plus_linear():
    net = _net()
    blk = net.blocks[1]
    W = blk._fused_weights(torch.float32)
    an = blk.attention_network
    D = an[1].weight.shape[0]
    g = torch.Generator().manual_seed(1)
    ab = torch.randn(2, 37, D, generator=g) + 2.0 * torch.randn(1, 37, 1, generator=g)      # rows with a common offset
    with torch.no_grad():
        want = an[1](an[0](torch.cat([ab[0], ab[1]], 1)))
        g2 = torch.bmm(ab, W["att_wf"])                                                      # (2, rows, D): a W'_a^T, b W'_b^T
        x = torch.cat([ab[0], ab[1]], 1)
        mean = x.mean(1, keepdim=True)
        rstd = torch.rsqrt(x.var(1, unbiased=False, keepdim=True) + W["ln2_eps"])
        got = rstd * (g2[0] + g2[1] - mean * W["att_colsum"][None]) + W["att_cvec"][None]
    assert W["att_wf"].shape == (2, D, D) and W["att_colsum"].shape == (D,) and W["att_cvec"].shape == (D,)
    torch.testing.assert_close(got, want, rtol=2e-5, atol=2e-5)


def test_step_head_tables_reproduce_patch_embed_and_timestep_embedder():
    net = _net()
    p = net.patch_size
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 4, 28, 28, generator=g)
    with torch.no_grad():
        want_h = net.x_embedder(x) + net.pos_embed                                           # reference model.py:272
        wp, posb = net._patch_tables()
        gsz = 28 // p
        patches = x.view(3, 4, gsz, p, gsz, p).permute(0, 2, 4, 1, 3, 5).reshape(3, gsz * gsz, 4 * p * p)   # dm_step_head's gather order
        got_h = patches @ wp + posb[None]
        t = torch.tensor([0, 17, 999])
        want_t = net.t_embedder(t)
        got_t = net._t_table()[t]
    assert wp.shape == (4 * p * p, 512) and posb.shape == (gsz * gsz, 512) and net._t_table().shape == (1000, 512)
    torch.testing.assert_close(got_h, want_h, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(got_t, want_t, rtol=1e-5, atol=1e-6)


def test_autocast_leaf_parameter_names_exist_in_every_block():
    from diffma_b200.ddp import autocast_leaf_params
    net = _net()
    names = {n for n, _ in net.named_parameters()}
    for i in range(len(net.blocks)):
        for m in ("mamba1", "mamba2"):
            for leaf in ("in_proj.weight", "out_proj.weight", "x_proj.weight", "dt_proj.weight"):
                assert f"blocks.{i}.{m}.{leaf}" in names
        assert f"blocks.{i}.adaLN_modulation.1.weight" in names and f"blocks.{i}.attention_network.1.bias" in names
    leaves = autocast_leaf_params(net)
    assert len(leaves) == 12 * len(net.blocks) + 2
    # nothing the kernels read in fp32 is among them
    by_id = {id(p): n for n, p in net.named_parameters()}
    for p in leaves:
        n = by_id[id(p)]
        assert not any(k in n for k in ("A_log", ".D", "conv1d", "norm", "dt_proj.bias")), n
