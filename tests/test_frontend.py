"""Inference front-end: WAV writer and the batched GenSamples / manifest loop on a stub
model (CPU); the GPU variant drives the real pipeline."""
import os
import types

import numpy as np
import pytest
import torch


def _frontend():
    from ma3_b200 import frontend
    return frontend


def test_write_wav_pcm16_round_trip(tmp_path):
    F = _frontend()
    x = np.concatenate([np.linspace(-1.2, 1.2, 1000), [0.0, 1.0, -1.0, 0.5 / 32768, 0.49 / 32768]]).astype(np.float32)
    p = F.write_wav(str(tmp_path / "a.wav"), x, 16000)
    y, sr = F.read_wav(p)
    assert sr == 16000 and y.shape == x.shape
    assert os.path.getsize(p) == 44 + 2 * x.size
    # float -> PCM_16 as libsndfile: scale by 2^15, round, clip to [-32768, 32767]
    ref = np.clip(np.rint(x.astype(np.float64) * 32768), -32768, 32767) / 32768
    assert np.array_equal(y, ref.astype(np.float32))
    import wave
    with wave.open(p) as w:                                   # a standard reader accepts the header
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (1, 2, 16000, x.size)


class _StubModel:
    """Stands in for CFM + conditioner: deterministic 'sampling' so the plumbing can be checked on CPU."""
    channels = 0

    def get_learned_conditioning(self, prompt):
        if isinstance(prompt, dict):
            prompt = prompt["ori_caption"]
        if isinstance(prompt, str):
            prompt = [prompt]
        return torch.stack([torch.full((4, 8), float(len(p))) for p in prompt])

    def sample_cfg(self, c, scale, uc, B, timesteps=None, x_latent=None, **kw):
        assert c.shape[0] == B == uc.shape[0] and x_latent.shape == (B, 20, 6)
        return c.mean(dim=(1, 2)).view(B, 1, 1).expand(B, 20, 6).contiguous(), None

    def sample(self, c, B, timesteps=None, x_latent=None, **kw):
        return self.sample_cfg(c, 1.0, c, B, timesteps, x_latent)

    def decode_first_stage(self, z):
        return z.mean(1, keepdim=True).expand(-1, 80, -1).repeat(1, 1, 2).contiguous()


class _StubVocoder:
    def vocode_tensor(self, mel):
        return torch.tanh(mel.mean(1) * 0.01).repeat_interleave(4, dim=1)


def test_gensamples_stub(tmp_path):
    F = _frontend()
    opt = types.SimpleNamespace(H=20, W=6, scale=3.0, ddim_steps=5, n_iter=2, sample_rate=16000)
    g = F.GenSamples(opt, _StubModel(), str(tmp_path), vocoder=_StubVocoder(), save_mel=True, save_wav=True)
    recs = g.gen_test_sample({"ori_caption": ["a dog barks"], "struct_caption": ["<dog, bark>"]}, wav_name="dog")
    g.flush()
    assert len(recs) == 2 and all(os.path.exists(r["audio_path"]) and os.path.exists(r["mel_path"]) for r in recs)
    assert recs[0]["caption"] == "a dog barks"
    items = [{"caption": "rain", "f_name": "vid1_0"}, {"caption": "thunder", "f_name": "vid1_1"},
             {"caption": "birds sing", "f_name": "clip_b_7"}]
    rows = g.generate_manifest(items, batch_size=2)
    assert [os.path.basename(r["audio_path"]) for r in rows] == ["vid1_sample_0_0.wav", "vid1_sample_1_0.wav", "clip_b_sample_7_0.wav"]
    y, sr = F.read_wav(rows[2]["audio_path"])
    assert sr == 16000 and y.shape == (48,) and abs(float(y[0]) - np.tanh(0.01 * len("birds sing"))) < 1e-3
    lines = open(os.path.join(str(tmp_path), "result.csv")).read().strip().split("\n")
    assert lines[0].split("\t") == ["caption", "mel_path", "audio_path"] and len(lines) == 4


@pytest.mark.gpu
def test_gensamples_real_pipeline(tmp_path):
    """GenSamples over the B200 pipeline: 3 prompts of precomputed embeddings in one batch; every written WAV equals
    the pipeline's own waveform for that prompt (PCM 16 quantisation apart)."""
    from ma3_b200 import dit as D
    from ma3_b200.pipeline import Txt2AudioPipeline
    from ma3_b200.vae import AutoencoderKL
    from ma3_b200.vocoder import VocoderBigVGAN
    from oracle import cases as Cs, weights as W
    F = _frontend()
    cfg = Cs.DIT_SMALL
    dit = D.TxtFlagLargeImprovedDiTV2(**dict(cfg, max_len=100))
    dit.load_state_dict(W.dit_state_dict(**cfg, seed=3), strict=True)
    vae = AutoencoderKL(embed_dim=20, ddconfig=dict(Cs.VAE_TINY))
    vae.load_state_dict(W.vae_decoder_state_dict(Cs.VAE_TINY, 20), strict=False)
    h = Cs.BIGVGAN_SMALL
    pipe = Txt2AudioPipeline(dit.cuda(), vae.cuda(), VocoderBigVGAN(h=h, state_dict=W.bigvgan_state_dict(h)), use_graph=False)
    opt = types.SimpleNamespace(H=20, W=24, scale=3.0, ddim_steps=5, n_iter=1, sample_rate=16000)
    x0, c, uc = Cs.cfm_inputs(cfg, B=3, T=24, L=10)
    g = F.GenSamples(opt, pipe, str(tmp_path), vocoder=pipe.vocoder, save_mel=True, save_wav=True)
    recs = g.gen_batch(c.cuda(), uc.cuda(), names=["a", "b", "c"], captions=["x", "y", "z"], x_latent=x0.cuda())
    g.flush()
    wav = pipe.generate(c.cuda(), uc.cuda(), x0.cuda(), scale=3.0, timesteps=5).cpu().numpy()
    for i, r in enumerate(recs):
        y, sr = F.read_wav(r["audio_path"])
        assert sr == 16000 and y.shape == wav[i].shape
        assert np.abs(y - np.clip(wav[i], -1, 32767 / 32768)).max() <= 1.0 / 32768 + 1e-6
        mel = np.load(r["mel_path"])
        assert mel.shape == (80, 48)
