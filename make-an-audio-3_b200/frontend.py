"""Inference front-end: the caller of the sampling path.

`GenSamples` mirrors scripts/txt2audio_for_2cap_flow.py:133-217 -- same constructor, same
`gen_test_sample(prompt, mel_name, wav_name, gt, video)` call and the same record dicts / files on disk -- with the
three serialisations of the reference removed:

  * prompts are sampled as a BATCH (the reference runs batch 1 inside an `n_iter` loop): `gen_batch` takes a list of
    prompts, conditions them with the model's own `get_learned_conditioning`, and runs one `sample_cfg` for all of them;
  * the mel never leaves the GPU between `decode_first_stage` and the vocoder (the reference goes through numpy per
    clip, txt2audio_for_2cap_flow.py:181-188; `VocoderBigVGAN.vocode_tensor` takes the device tensor);
  * waveforms come back with ONE asynchronous device-to-host copy into pinned memory per batch, and files are written by
    a small thread pool while the next batch samples.

`write_wav` replaces `soundfile.write(path, wav, sr)` for float input: RIFF/WAVE, PCM 16 bit, the sample scaled by 2^15
and clipped exactly as libsndfile does for float -> PCM_16.  `generate_manifest` is the `testset` loop of the script's
main() (:240-262) including result.csv (tab separated, same columns).

`model.get_learned_conditioning` is whatever the wrapped model provides: the reference's FrozenCLAP/T5 stack, this
package's GPU conditioner (`conditioners.FrozenCLAPFLANEmbedder` attached as `pipeline.cond_stage_model`, section 8(f)
rank 1), or the identity for precomputed embeddings.
"""
import csv
import os
import struct
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch


def write_wav(path, samples, sample_rate):
    """float32 [n] (or [n, channels]) in [-1, 1] -> 16-bit PCM WAV (what soundfile.write does for float input)."""
    x = np.asarray(samples, dtype=np.float32)
    ch = 1 if x.ndim == 1 else x.shape[1]
    pcm = np.clip(np.rint(x.astype(np.float64) * 32768.0), -32768, 32767).astype("<i2")
    data = pcm.tobytes()
    hdr = b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVE" + b"fmt " + struct.pack(
        "<IHHIIHH", 16, 1, ch, int(sample_rate), int(sample_rate) * ch * 2, ch * 2, 16) + b"data" + struct.pack("<I", len(data))
    with open(path, "wb") as f:
        f.write(hdr)
        f.write(data)
    return path


def read_wav(path):
    """Inverse of write_wav (16-bit PCM only) -> (float32 samples, sample_rate); used by the tests."""
    with open(path, "rb") as f:
        b = f.read()
    assert b[:4] == b"RIFF" and b[8:12] == b"WAVE" and b[12:16] == b"fmt "
    _, fmt, ch, sr, _, _, bits = struct.unpack("<IHHIIHH", b[16:36])
    assert fmt == 1 and bits == 16 and b[36:40] == b"data"
    n = struct.unpack("<I", b[40:44])[0]
    x = np.frombuffer(b[44:44 + n], dtype="<i2").astype(np.float32) / 32768.0
    return (x if ch == 1 else x.reshape(-1, ch)), sr


def _caption_of(prompt):
    if isinstance(prompt, dict):
        c = prompt.get("ori_caption", "")
        return c[0] if isinstance(c, (list, tuple)) and c else c
    return prompt


class GenSamples:
    """scripts/txt2audio_for_2cap_flow.py:133-217 with batched sampling and a device-resident mel -> wav hand-over."""

    def __init__(self, opt, model, outpath, config=None, vocoder=None, save_mel=True, save_wav=True, io_threads=4):
        self.opt, self.model, self.outpath, self.config = opt, model, outpath, config
        if save_wav:
            assert vocoder is not None
        self.vocoder = vocoder
        self.save_mel, self.save_wav = save_mel, save_wav
        self.channel_dim = getattr(model, "channels", 0)
        self._pool = ThreadPoolExecutor(max_workers=io_threads)
        self._pending = []
        os.makedirs(outpath, exist_ok=True)

    # ------------------------------------------------------------------ conditioning (the model's own conditioner)
    def _condition(self, prompts):
        m = self.model
        try:
            return m.get_learned_conditioning(prompts)
        except Exception:
            return m.get_learned_conditioning([_caption_of(p) for p in prompts] if isinstance(prompts, list) else _caption_of(prompts))

    def _uncond(self, n, like):
        m = self.model
        try:
            uc = m.get_learned_conditioning({"ori_caption": [""] * n, "struct_caption": [""] * n})
        except Exception:
            uc = m.get_learned_conditioning([""] * n)
        if torch.is_tensor(uc) and uc.shape[0] == 1 and n > 1:
            uc = uc.expand(n, *uc.shape[1:]).contiguous()
        return uc

    # ------------------------------------------------------------------ batched generation
    @torch.no_grad()
    def gen_batch(self, cond, uncond=None, names=None, captions=None, x_latent=None):
        """cond [B, L, Cd] (already conditioned) -> list of record dicts; files are written asynchronously (call
        `flush()` before reading them).  B = prompts x n_iter when called through gen_test_sample."""
        opt, m = self.opt, self.model
        B = cond.shape[0]
        dev = cond.device
        shape = (B, opt.H, opt.W)
        x0 = torch.randn(shape, device=dev) if x_latent is None else x_latent
        if opt.scale == 1:
            z, _ = m.sample(cond, B, timesteps=opt.ddim_steps, x_latent=x0)
        else:
            z, _ = m.sample_cfg(cond, opt.scale, uncond, B, timesteps=opt.ddim_steps, x_latent=x0)
        mel = m.decode_first_stage(z)                       # [B, 80, 2W] on the device
        names = names or [f"sample_{i}" for i in range(B)]
        captions = captions or [""] * B
        recs = [{"caption": captions[i]} for i in range(B)]
        mel_h = wav_h = None
        on_gpu = mel.is_cuda      # (the CPU branch only serves the plumbing tests on a stub model)

        def to_host(t):
            if not on_gpu:
                return t.float()
            hbuf = torch.empty(t.shape, dtype=torch.float32, pin_memory=True)
            return hbuf.copy_(t, non_blocking=True)

        if self.save_mel or not hasattr(self.vocoder, "vocode_tensor"):
            mel_h = to_host(mel)
        if self.save_wav and hasattr(self.vocoder, "vocode_tensor"):
            wav_h = to_host(self.vocoder.vocode_tensor(mel))
        ev = torch.cuda.Event() if on_gpu else None
        if ev is not None:
            ev.record()
        sr = getattr(opt, "sample_rate", 16000)
        for i in range(B):
            if self.save_mel:
                recs[i]["mel_path"] = os.path.join(self.outpath, names[i] + "_0.npy")
            if self.save_wav:
                recs[i]["audio_path"] = os.path.join(self.outpath, names[i] + "_0.wav")

        def write(i):
            if ev is not None:
                ev.synchronize()
            if self.save_mel:
                np.save(recs[i]["mel_path"], mel_h[i].numpy())
            if self.save_wav:
                w = wav_h[i].numpy() if wav_h is not None else self.vocoder.vocode(mel_h[i].numpy())
                write_wav(recs[i]["audio_path"], w, sr)

        self._pending += [self._pool.submit(write, i) for i in range(B)]
        return recs

    def flush(self):
        for f in self._pending:
            f.result()
        self._pending = []

    # ------------------------------------------------------------------ reference surface
    @torch.no_grad()
    def gen_test_sample(self, prompt, mel_name=None, wav_name=None, gt=None, video=None):
        """One prompt, `opt.n_iter` samples of it -- generated as one batch of n_iter."""
        n = max(1, int(getattr(self.opt, "n_iter", 1)))
        c = self._condition(prompt)
        if torch.is_tensor(c) and c.shape[0] == 1 and n > 1:
            c = c.expand(n, *c.shape[1:]).contiguous()
        uc = self._uncond(c.shape[0], c) if self.opt.scale != 1.0 else None
        base = wav_name or mel_name or "sample"
        names = [base if n == 1 else f"{base}_iter{i}" for i in range(c.shape[0])]
        recs = self.gen_batch(c, uc, names=names, captions=[_caption_of(prompt)] * c.shape[0])
        if gt is not None and self.save_wav:
            wav_gt = self.vocoder.vocode(gt)
            self.flush()
            write_wav(os.path.join(self.outpath, (wav_name or "ground_truth_audio") + "_gt.wav"), wav_gt, 16000)
        return recs

    @torch.no_grad()
    def generate_manifest(self, items, batch_size=8, csv_name="result.csv"):
        """The `testset` loop of the reference's main() (:240-262), `batch_size` prompts at a time.  items: iterable of
        dicts with 'caption' (str or the reference's two-caption dict) and 'f_name' ("<video>_<num>")."""
        items = list(items)
        rows = []
        for b0 in range(0, len(items), batch_size):
            chunk = items[b0:b0 + batch_size]
            prompts = [it["caption"] for it in chunk]
            names = []
            for it in chunk:
                f = it["f_name"]
                k = f.rfind("_")
                names.append(f"{f[:k]}_sample_{f[k + 1:]}" if k > 0 else f"{f}_sample")
            try:
                c = torch.cat([self._condition(p) for p in prompts])
            except Exception:
                c = self._condition(prompts)
            uc = self._uncond(c.shape[0], c) if self.opt.scale != 1.0 else None
            rows += self.gen_batch(c, uc, names=names, captions=[_caption_of(p) for p in prompts])
        self.flush()
        if rows:
            with open(os.path.join(self.outpath, csv_name), "w", newline="") as f:
                w = csv.DictWriter(f, fieldnames=list(rows[0].keys()), delimiter="\t")
                w.writeheader()
                w.writerows(rows)
        return rows
