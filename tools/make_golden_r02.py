"""Round-2 golden vectors (tests/golden/ref_golden_r02.pt), produced by the REFERENCE's own modules:
VAE encoder moments (AutoencoderKL.encode), a qk_norm=True Next-DiT forward, MelNet.forward.  Run in the build
container only:  python tools/make_golden_r02.py

MelNet imports librosa.filters.mel, which is not installed here: the filterbank is supplied by a shim that calls the
oracle's restatement (oracle.restated.slaney_mel_filterbank, cross-checked against torchaudio in the tests), so the
golden pins the reference's clamp / reflect pad / STFT / magnitude / log10 chain, not the filterbank."""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import cases as Cs, ref_loader as R, restated as O, weights as W

out = {}
with torch.no_grad():
    # 1. AutoencoderKL.encode (autoencoder1d.py:49-53): moments of the posterior
    R.setup()
    import contextlib, io
    from ldm.models.autoencoder1d import AutoencoderKL
    with contextlib.redirect_stdout(io.StringIO()):
        vae = AutoencoderKL(embed_dim=20, ddconfig=dict(Cs.VAE_TINY), lossconfig={"target": "torch.nn.Identity"})
    sd = dict(W.vae_encoder_state_dict(Cs.VAE_TINY, 20), **W.vae_decoder_state_dict(Cs.VAE_TINY, 20))
    r = vae.load_state_dict(sd, strict=False)
    assert not r.unexpected_keys and all(k.startswith("loss.") for k in r.missing_keys), r
    mel = Cs.mel_inputs(B=2, T=48)
    post = vae.encode(mel)
    out["vae_enc_moments"] = post.parameters.clone()
    out["vae_enc_mode"] = post.mode().clone()
    # 2. TxtFlagLargeDiT with qk_norm=True (flag_large_dit.py:128-210, flag_large_dit_moe.py:199-207,345-346)
    from ldm.modules.diffusionmodules.flag_large_dit import TxtFlagLargeDiT
    cfg = Cs.DIT_SMALL
    dsd = W.dit_state_dict(**cfg, seed=9, qk_norm=True)
    with contextlib.redirect_stdout(io.StringIO()):
        m = TxtFlagLargeDiT(cfg["in_channels"], cfg["context_dim"], hidden_size=cfg["hidden_size"], depth=cfg["depth"],
                            num_heads=cfg["num_heads"], max_len=100, qk_norm=True)
    m.load_state_dict(dsd, strict=True)
    x, ctx = Cs.dit_inputs(cfg)
    out["dit_small_qknorm"] = m.eval()(x, torch.tensor([41, 958]), context=ctx).clone()
    # 3. MelNet.forward (preprocess/NAT_mel.py:42-85) with the filterbank shim
    lib = types.ModuleType("librosa"); lib.filters = types.ModuleType("librosa.filters")
    lib.filters.mel = lambda sr, n_fft, n_mels, fmin, fmax: O.slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax).numpy()
    sys.modules["librosa"], sys.modules["librosa.filters"] = lib, lib.filters
    sys.path.insert(0, os.path.join(R.REF_ROOT))
    from preprocess.NAT_mel import MelNet
    net = MelNet(Cs.MEL_HP)
    out["melnet"] = net(Cs.wave_inputs()).clone()

path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_golden_r02.pt")
torch.save({k: v.contiguous() for k, v in out.items()}, path)
print("wrote", path, {k: tuple(v.shape) for k, v in out.items()}, os.path.getsize(path), "bytes")
