"""Generate tests/golden/ref_golden.pt by running the REFERENCE's own modules (imported from /root/reference through
oracle/_shims) on the seeded cases of oracle/cases.py.  Run in the build container only:  python tools/make_golden.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import cases as Cs, ref_loader as R, weights as W

out = {}
with torch.no_grad():
    # 1. text DiT forward (flag_large_dit.py:177-210)
    for name, cfg in (("dit_tiny", Cs.DIT_TINY), ("dit_small", Cs.DIT_SMALL)):
        sd = W.dit_state_dict(**cfg, seed=3)
        m = R.build_dit(sd, **cfg, max_len=100)
        x, ctx = Cs.dit_inputs(cfg)
        out[name] = m(x, torch.tensor([41, 958]), context=ctx).clone()
    # 2. video / MoE DiT forward (flag_large_dit_moe.py:664-698)
    sd = W.dit_state_dict(**Cs.DIT_TINY, video=True, num_experts=4, seed=4)
    m = R.build_dit(sd, **Cs.DIT_TINY, max_len=100, video=True, num_experts=4)
    x, ctx = Cs.dit_inputs(Cs.DIT_TINY)
    out["dit_moe_tiny"] = m(x, torch.tensor([260, 958]), context=ctx).clone()
    # 3. the reference CFM stack: sample_cfg / sample / decode_first_stage (cfm1_audio.py:60-111, ddpm_audio.py:358-371)
    dsd = W.dit_state_dict(**Cs.DIT_TINY, seed=3)
    vsd = W.vae_decoder_state_dict(Cs.VAE_TINY, 20)
    cfm = R.build_cfm(dict(Cs.DIT_TINY, max_len=100), Cs.VAE_TINY, 20, dsd, vsd)
    x0, c, uc = Cs.cfm_inputs(Cs.DIT_TINY)
    xf, traj = cfm.sample_cfg(c, 3.0, uc, 2, timesteps=6, x_latent=x0)
    out["cfm_cfg_final"], out["cfm_cfg_traj"] = xf.clone(), traj.clone()
    xf2, traj2 = cfm.sample(c, 2, timesteps=6, x_latent=x0)
    out["cfm_plain_final"] = xf2.clone()
    xf3, _ = cfm.sample_cfg(c, 3.0, uc, 2, timesteps=6, x_latent=x0, t_start=2)
    out["cfm_cfg_tstart2_final"] = xf3.clone()
    out["scale_factor"] = torch.as_tensor(float(cfm.scale_factor))
    out["vae_tiny"] = cfm.decode_first_stage(Cs.latent_inputs()).clone()
    # 4. BigVGAN (vocoder/bigvgan/models.py:183-205)
    for name, h, T in (("bigvgan_tiny", Cs.BIGVGAN_TINY, 12), ("bigvgan_small", Cs.BIGVGAN_SMALL, 40)):
        g = R.build_bigvgan(W.bigvgan_state_dict(h), h)
        out[name] = g(Cs.mel_inputs(T=T)).clone()
    # 5. Activation1d alone (alias_free_torch/act.py:23-28 with SnakeBeta, activations.py:107-119)
    R.setup()
    from vocoder.bigvgan.activations import SnakeBeta
    from vocoder.bigvgan.alias_free_torch.act import Activation1d
    xa, al, be = Cs.act_inputs()
    act = Activation1d(activation=SnakeBeta(24, alpha_logscale=True))
    act.act.alpha.data.copy_(al)
    act.act.beta.data.copy_(be)
    out["act1d"] = act(xa).clone()
    out["act1d_short"] = act(xa[..., :3]).clone()   # shorter than the filter: replicate padding dominates
    out["filter"] = act.upsample.filter.flatten().clone()
    # 6. integer timesteps seen by the DiT for the default 25 points (cfm1_audio.py:103,156)
    seen = []

    class Probe(torch.nn.Module):
        def forward(self, x, t, **kw):
            seen.append(int(t[0]))
            return torch.zeros_like(x)
    cfm.model.diffusion_model = Probe()
    cfm.sample_cfg(c, 3.0, uc, 2, timesteps=25, x_latent=x0)
    out["t_ints_25"] = torch.tensor(seen)

path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_golden.pt")
torch.save({k: v.contiguous() for k, v in out.items()}, path)
print("wrote", path, {k: tuple(v.shape) for k, v in out.items()}, os.path.getsize(path), "bytes")
