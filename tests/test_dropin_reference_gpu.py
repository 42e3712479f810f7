"""Row a4 / (b) of SURVEY.md section 8 on hardware: the reference's OWN `CFM` object, built by its own
`instantiate_from_config` (ldm/util.py:110-125) with only the two `target:` strings of INTEGRATION.md section 1 changed,
runs its own `sample_cfg` loop (cfm1_audio.py:89-111 -> Wrapper_cfg :145-161 -> apply_model ddpm_audio.py:437-460 ->
DiffusionWrapper.forward ddpm.py:1406-1420) into the B200 drop-in DiT and `decode_first_stage` into the drop-in VAE --
compared with the same object built from the unmodified reference classes (fp32 on the same GPU) on the same weights.

The reference sources come from oracle/_ref (git-ignored copy written by `__graft_entry__.build()`, which travels to
the GPU box); without it the test is skipped with that reason."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ref_loader as R, restated as O, weights as W  # noqa: E402


@pytest.fixture(scope="module")
def stacks():
    if not R.available():
        pytest.skip("no copy of the reference on this box (oracle/_ref is written by __graft_entry__.build())")
    from ma3_b200.pipeline import MODEL_CONFIGS, VAE_DDCONFIG
    O.strict_fp32()
    cfg = dict(MODEL_CONFIGS["M"])
    wcfg = {k: v for k, v in cfg.items() if k != "max_len"}
    dsd = W.dit_state_dict(**wcfg, seed=5)
    vsd = W.vae_decoder_state_dict(VAE_DDCONFIG, 20)
    ref = R.build_cfm(cfg, VAE_DDCONFIG, 20, dsd, vsd, mel_length=312).cuda()
    ours = R.build_cfm(cfg, VAE_DDCONFIG, 20, dsd, vsd, mel_length=312,
                       unet_target="ma3_b200.dit.TxtFlagLargeImprovedDiTV2",
                       first_stage_target="ma3_b200.vae.AutoencoderKL").cuda()
    return ref, ours


def test_reference_cfm_drives_dropin_modules(stacks):
    from ma3_b200 import dit as D, vae as V
    ref, ours = stacks
    assert type(ours).__module__ == "ldm.models.diffusion.cfm1_audio"          # the wrapper is the reference's
    assert isinstance(ours.model.diffusion_model, D.TxtFlagLargeImprovedDiTV2)
    assert isinstance(ours.first_stage_model, V.AutoencoderKL)
    c, uc, x0 = W.synthetic_inputs(prompts=2, latent_ch=20, T=312, L=154, Cd=1024)
    c, uc, x0 = c.cuda(), uc.cuda(), x0.cuda()
    zr, traj_r = ref.sample_cfg(c, 3.0, uc, 2, timesteps=25, x_latent=x0)
    zo, traj_o = ours.sample_cfg(c, 3.0, uc, 2, timesteps=25, x_latent=x0)     # the reference's loop, our DiT
    assert traj_o.shape == traj_r.shape == (25, 2, 20, 312)
    assert O.cosine(zo.cpu(), zr.cpu()) >= 0.999
    mel_r, mel_o = ref.decode_first_stage(zr), ours.decode_first_stage(zr)
    assert mel_o.shape == mel_r.shape == (2, 80, 624)
    assert O.cosine(mel_o.cpu(), mel_r.cpu()) >= 0.999
    # list conditioning through the reference's apply_model (conditioning_key 'crossattn', ddpm_audio.py:445-449)
    t = torch.tensor([333, 333], device="cuda")
    halves = [c[:, :77].contiguous(), c[:, 77:].contiguous()]     # list conditioning: concatenated on dim 1 (ddpm.py:1413-1416)
    vd_r, vd_o = ref.apply_model(x0, t, halves), ours.apply_model(x0, t, halves)
    assert O.max_rel_err(vd_o.cpu(), vd_r.cpu()) <= 1e-2
    assert torch.equal(vd_o, ours.apply_model(x0, t, c))
    # no-CFG branch (Wrapper, cfm1_audio.py:134-142) and default latent shape (cfm1_audio.py:61-66)
    zp_r, _ = ref.sample(c, 2, timesteps=5, x_latent=x0)
    zp_o, _ = ours.sample(c, 2, timesteps=5, x_latent=x0)
    assert O.cosine(zp_o.cpu(), zp_r.cpu()) >= 0.999
    torch.manual_seed(7)
    zs, _ = ours.sample_cfg(c, 3.0, uc, 2, timesteps=3)
    assert zs.shape == (2, 20, 312)


def test_cfm_sampler_on_reference_wrapper(stacks):
    """The engine (`CFMSampler`, CUDA-graph step loop) finds the drop-in DiT inside the reference's wrapper as
    model.model.diffusion_model (cfm1_audio_sampler.py:26-33) and reproduces the reference loop's result."""
    from ma3_b200.sampler import CFMSampler
    ref, ours = stacks
    c, uc, x0 = W.synthetic_inputs(prompts=2, latent_ch=20, T=312, L=154, Cd=1024, rank=1)
    c, uc, x0 = c.cuda(), uc.cuda(), x0.cuda()
    zr, _ = ref.sample_cfg(c, 3.0, uc, 2, timesteps=25, x_latent=x0)
    s = CFMSampler(ours)
    for _ in range(2):
        z, traj = s.sample_cfg(c, 3.0, uc, 2, timesteps=25, x_latent=x0)
    assert traj.shape == (25, 2, 20, 312)
    assert O.cosine(z.cpu(), zr.cpu()) >= 0.999
    zdef, _ = s.sample_cfg(c, 3.0, uc, 2, timesteps=3)        # default shape from the wrapper: (B, mel_dim, mel_length)
    assert zdef.shape == (2, 20, 312)
