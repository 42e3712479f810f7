"""Live check of the oracle against the reference modules imported from /root/reference (skipped where that tree
is absent, e.g. on the GPU box).  CPU only."""
import pytest
import torch

from oracle import cases as Cs, ref_loader as R, restated as O, weights as W

pytestmark = pytest.mark.skipif(not R.available(), reason="/root/reference not mounted")


@torch.no_grad()
def test_dit_realistic_heads():
    # head dims of the shipped configs: 24 (M), 72 (XL), 48 (XXL) at reduced width/depth
    for D, H in ((96, 4), (144, 2), (96, 2)):
        cfg = dict(in_channels=20, context_dim=48, hidden_size=D, num_heads=H, depth=2)
        sd = W.dit_state_dict(**cfg, seed=7)
        m = R.build_dit(sd, **cfg, max_len=64)
        x, ctx = Cs.dit_inputs(cfg, N=3, T=40, L=7)
        t = torch.tensor([0, 500, 999])
        assert O.max_rel_err(O.dit_forward(sd, x, t, ctx, heads=H, max_len=64), m(x, t, context=ctx)) < 2e-5


@torch.no_grad()
def test_ntk_rope_override():
    # scripts/video2audio_flow_inpaint.py:223-235 overwrites freqs_cis on the live module
    cfg = Cs.DIT_TINY
    sd = W.dit_state_dict(**cfg, seed=3)
    m = R.build_dit(sd, **cfg, max_len=100)
    m.freqs_cis = m.precompute_freqs_cis(16, 100, ntk_factor=3.0)
    x, ctx = Cs.dit_inputs(cfg)
    t = torch.tensor([41, 958])
    out = O.dit_forward(sd, x, t, ctx, heads=4, rope=O.rope_table(16, 100, ntk_factor=3.0))
    assert O.max_rel_err(out, m(x, t, context=ctx)) < 2e-5


@torch.no_grad()
def test_bigvgan_base_layout_and_amp2():
    for h in (dict(W.BIGVGAN_BASE_256X, upsample_initial_channel=64),
              dict(W.BIGVGAN_BASE_256X, upsample_initial_channel=64, resblock="2",
                   resblock_dilation_sizes=[[1, 3], [1, 3], [1, 3]])):
        sd = W.bigvgan_state_dict(h)
        g = R.build_bigvgan(sd, h)
        mel = Cs.mel_inputs(B=1, T=6)
        assert O.max_rel_err(O.bigvgan_forward(sd, mel, h), g(mel)) < 2e-5
