"""TEST INFRASTRUCTURE: import the *reference's own modules* through the import shims in oracle/_shims (SURVEY.md
section 8(c)).  The tree is taken from $MA3_REFERENCE_ROOT, else /root/reference (mounted in the build container
only), else oracle/_ref (the git-ignored copy made by oracle/build_ref.py at build() time, which travels to the GPU
box).  Callers on the `-m gpu` / bench paths must check `available()` first and say so when it is False."""
import copy
import os
import sys

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    for cand in (os.environ.get("MA3_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "ldm")):
            return cand
    return "/root/reference"


REF_ROOT = _find_root()
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_shims")
_ready = False


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "ldm"))


def setup():
    global _ready
    if _ready:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")  # @torch.compile at flag_large_dit_moe.py:484 fails on CPU here
    for p in (_SHIMS, REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not torch.cuda.is_available():
        # precompute_freqs_cis hard-codes .cuda() (flag_large_dit.py:245)
        torch.Tensor.cuda = lambda self, *a, **k: self
    _ready = True


def dit_class(video=False):
    setup()
    if video:
        from ldm.modules.diffusionmodules.flag_large_dit_moe import VideoFlagLargeDiT
        return VideoFlagLargeDiT
    from ldm.modules.diffusionmodules.flag_large_dit import TxtFlagLargeImprovedDiTV2
    return TxtFlagLargeImprovedDiTV2


def build_dit(sd, *, in_channels, context_dim, hidden_size, num_heads, depth, max_len=1000, video=False,
              num_experts=0):
    import contextlib
    import io
    cls = dit_class(video)
    kw = dict(in_channels=in_channels, context_dim=context_dim, hidden_size=hidden_size, num_heads=num_heads,
              depth=depth, max_len=max_len)
    if video:
        kw["num_experts"] = num_experts
    with contextlib.redirect_stdout(io.StringIO()):
        m = cls(**kw)
    missing, unexpected = m.load_state_dict(sd, strict=True), None
    return m.eval()


def build_vae(sd, ddconfig, embed_dim):
    setup()
    import contextlib
    import io
    from ldm.models.autoencoder1d import AutoencoderKL
    with contextlib.redirect_stdout(io.StringIO()):
        m = AutoencoderKL(embed_dim=embed_dim, ddconfig=dict(ddconfig), lossconfig={"target": "torch.nn.Identity"})
    r = m.load_state_dict(sd, strict=False)
    assert not r.unexpected_keys, r.unexpected_keys
    assert all(k.startswith(("encoder.", "quant_conv.", "loss.")) for k in r.missing_keys), r.missing_keys
    return m.eval()


class _AttrDict(dict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.__dict__ = self


def build_bigvgan(sd, h):
    setup()
    import contextlib
    import io
    from vocoder.bigvgan.models import BigVGAN
    with contextlib.redirect_stdout(io.StringIO()):
        m = BigVGAN(_AttrDict(copy.deepcopy(h)))
        m.remove_weight_norm()
    r = m.load_state_dict(sd, strict=False)
    assert not r.unexpected_keys, r.unexpected_keys
    assert all(k.endswith("filter") for k in r.missing_keys), r.missing_keys
    return m.eval()


def build_cfm(unet_params, ddconfig, embed_dim, dit_sd, vae_sd, *, video=False, mel_dim=20, mel_length=256,
              unet_target=None, first_stage_target=None):
    """The reference's own CFM -> LatentDiffusion_audio -> DDPM -> DiffusionWrapper stack
    (ldm/models/diffusion/cfm1_audio.py:30), with an Identity conditioner (synthetic embeddings)."""
    setup()
    import contextlib
    import io
    from ldm.util import instantiate_from_config
    target = unet_target or ("ldm.modules.diffusionmodules.flag_large_dit_moe.VideoFlagLargeDiT" if video else
                             "ldm.modules.diffusionmodules.flag_large_dit.TxtFlagLargeImprovedDiTV2")
    cfg = {
        "target": "ldm.models.diffusion.cfm1_audio.CFM",
        "params": {
            "linear_start": 0.00085, "linear_end": 0.012, "num_timesteps_cond": 1, "log_every_t": 200,
            "timesteps": 1000, "first_stage_key": "image", "cond_stage_key": "caption", "mel_dim": mel_dim,
            "mel_length": mel_length, "channels": 0, "cond_stage_trainable": True, "conditioning_key": "crossattn",
            "monitor": "val/loss_simple_ema", "scale_by_std": True, "use_ema": False,
            "unet_config": {"target": target, "params": dict(unet_params)},
            "first_stage_config": {"target": first_stage_target or "ldm.models.autoencoder1d.AutoencoderKL",
                                   "params": {"embed_dim": embed_dim, "monitor": "val/rec_loss",
                                              "ddconfig": dict(ddconfig),
                                              "lossconfig": {"target": "torch.nn.Identity"}}},
            "cond_stage_config": {"target": "torch.nn.Identity"},
        },
    }
    with contextlib.redirect_stdout(io.StringIO()):
        model = instantiate_from_config(cfg)
    model.model.diffusion_model.load_state_dict(dit_sd, strict=True)
    r = model.first_stage_model.load_state_dict(vae_sd, strict=False)
    assert not r.unexpected_keys
    return model.eval()
