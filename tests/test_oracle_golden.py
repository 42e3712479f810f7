"""Pin the oracle (oracle/restated.py) to golden vectors produced by the reference's own modules
(tools/make_golden.py).  CPU only."""
import torch

from oracle import cases as Cs, restated as O, weights as W

TOL = 2e-5  # fp32 vs fp32, different op order


def test_dit_text(golden):
    for name, cfg in (("dit_tiny", Cs.DIT_TINY), ("dit_small", Cs.DIT_SMALL)):
        sd = W.dit_state_dict(**cfg, seed=3)
        x, ctx = Cs.dit_inputs(cfg)
        out = O.dit_forward(sd, x, torch.tensor([41, 958]), ctx, heads=cfg["num_heads"], max_len=100)
        assert O.max_rel_err(out, golden[name]) < TOL


def test_dit_video_moe(golden):
    sd = W.dit_state_dict(**Cs.DIT_TINY, video=True, num_experts=4, seed=4)
    x, ctx = Cs.dit_inputs(Cs.DIT_TINY)
    out = O.dit_forward(sd, x, torch.tensor([260, 958]), ctx, heads=4, video=True, num_experts=4, max_len=100)
    assert O.max_rel_err(out, golden["dit_moe_tiny"]) < TOL


def test_sampler(golden):
    sd = W.dit_state_dict(**Cs.DIT_TINY, seed=3)
    x0, c, uc = Cs.cfm_inputs(Cs.DIT_TINY)
    vel = lambda x, t, ctx: O.dit_forward(sd, x, t, ctx, heads=4, max_len=100)
    xf, traj, _ = O.sample_cfg(vel, x0, c, uc, 3.0, n_points=6)
    assert O.max_rel_err(xf, golden["cfm_cfg_final"]) < 1e-4
    assert O.max_rel_err(traj, golden["cfm_cfg_traj"]) < 1e-4
    xp, _ = O.sample_plain(vel, x0, c, n_points=6)
    assert O.max_rel_err(xp, golden["cfm_plain_final"]) < 1e-4
    xs, _, _ = O.sample_cfg(vel, x0, c, uc, 3.0, n_points=6, t_start=2)
    assert O.max_rel_err(xs, golden["cfm_cfg_tstart2_final"]) < 1e-4


def test_timestep_ints(golden):
    ints, dts = O.timestep_ints(25)
    assert ints == golden["t_ints_25"].tolist()
    assert ints[:4] == [0, 41, 83, 125] and ints[-1] == 958 and len(dts) == 24


def test_vae_decode(golden):
    sd = W.vae_decoder_state_dict(Cs.VAE_TINY, 20)
    out = O.vae_decode(sd, Cs.latent_inputs(), Cs.VAE_TINY, scale_factor=float(golden["scale_factor"]))
    assert out.shape == golden["vae_tiny"].shape
    assert O.max_rel_err(out, golden["vae_tiny"]) < TOL


def test_bigvgan(golden):
    for name, h, T in (("bigvgan_tiny", Cs.BIGVGAN_TINY, 12), ("bigvgan_small", Cs.BIGVGAN_SMALL, 40)):
        out = O.bigvgan_forward(W.bigvgan_state_dict(h), Cs.mel_inputs(T=T), h)
        assert O.max_rel_err(out, golden[name]) < TOL


def test_activation1d(golden):
    x, al, be = Cs.act_inputs()
    sd = {"a.act.alpha": al, "a.act.beta": be}
    h = dict(activation="snakebeta", snake_logscale=True)
    assert O.max_rel_err(O.activation1d(x, sd, "a", h), golden["act1d"]) < TOL
    assert O.max_rel_err(O.activation1d(x[..., :3], sd, "a", h), golden["act1d_short"]) < TOL
    assert torch.allclose(O.kaiser_sinc_filter(), golden["filter"], atol=1e-7)


def test_fold_weight_norm():
    g = Cs.gen(5)
    v = torch.randn(6, 4, 3, generator=g)
    gg = torch.rand(6, 1, 1, generator=g) + 0.5
    sd = O.fold_weight_norm({"c.weight_g": gg, "c.weight_v": v, "c.bias": torch.zeros(6)})
    ref = torch._weight_norm(v, gg, 0)
    assert torch.allclose(sd["c.weight"], ref, atol=1e-6) and "c.weight_v" not in sd


# ------------------------------------------------------------------------------------------------ round 2 additions
import os as _os

import pytest as _pytest


@_pytest.fixture(scope="module")
def golden2():
    root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    return torch.load(_os.path.join(root, "tests", "golden", "ref_golden_r02.pt"))


def test_vae_encode(golden2):
    sd = W.vae_encoder_state_dict(Cs.VAE_TINY, 20)
    mom = O.vae_encode(sd, Cs.mel_inputs(B=2, T=48), Cs.VAE_TINY)
    assert mom.shape == golden2["vae_enc_moments"].shape
    assert O.max_rel_err(mom, golden2["vae_enc_moments"]) < TOL
    mean, _ = O.posterior_mode_sample(mom)
    assert O.max_rel_err(mean, golden2["vae_enc_mode"]) < TOL


def test_dit_qk_norm(golden2):
    cfg = Cs.DIT_SMALL
    sd = W.dit_state_dict(**cfg, seed=9, qk_norm=True)
    x, ctx = Cs.dit_inputs(cfg)
    out = O.dit_forward(sd, x, torch.tensor([41, 958]), ctx, heads=cfg["num_heads"], max_len=100)
    assert O.max_rel_err(out, golden2["dit_small_qknorm"]) < TOL


def test_melnet(golden2):
    out = O.melnet(Cs.wave_inputs(), Cs.MEL_HP)
    assert out.shape == golden2["melnet"].shape
    assert O.max_rel_err(out, golden2["melnet"]) < 1e-4


def test_slaney_filterbank_matches_torchaudio():
    """The filterbank is restated from librosa's published algorithm (librosa is not installed); torchaudio ships the
    same construction (norm='slaney', mel_scale='slaney')."""
    ta = _pytest.importorskip("torchaudio")
    hp = Cs.MEL_HP
    ref = ta.functional.melscale_fbanks(hp["fft_size"] // 2 + 1, hp["fmin"], hp["fmax"], hp["audio_num_mel_bins"],
                                        hp["audio_sample_rate"], norm="slaney", mel_scale="slaney").t()
    fb = O.slaney_mel_filterbank(hp["audio_sample_rate"], hp["fft_size"], hp["audio_num_mel_bins"], hp["fmin"], hp["fmax"])
    assert fb.shape == ref.shape == (80, 513)
    assert float((fb - ref).abs().max()) < 1e-6
