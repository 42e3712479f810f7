"""Mel front-end: drop-in for preprocess/NAT_mel.py:42-85 (`MelNet`), the wav -> log-mel transform either side of the
sampling path (ground-truth mels for inpainting, CLAP-score evaluation).

Same constructor (`hparams` dict with fft_size / hop_size / win_size / audio_num_mel_bins / audio_sample_rate / fmin /
fmax) and the same `forward(y, center=False, complex=False)` -> log10-mel [B, n_mels, frames].  On B200 the STFT is a
4-tap GEMM over hop-sized rows of the reflect-padded signal against the Hann-windowed DFT basis (the frames are never
materialised) and the filterbank a second GEMM, both on tcgen05 with (hi, lo) bf16 operand splits and fp32 accumulation
(csrc/melnet.cu).  librosa's `filters.mel` (slaney scale, slaney norm) is restated here -- librosa is not a dependency.
"""
import math

import torch
import torch.nn as nn

from . import lib as L
from . import ops


def mel_filterbank(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) (htk=False, norm='slaney') -> float32 [n_mels, n_fft/2 + 1]."""
    f_sp, min_log_hz, logstep = 200.0 / 3, 1000.0, math.log(6.4) / 27.0
    min_log_mel = min_log_hz / f_sp

    def hz_to_mel(f):
        return min_log_mel + math.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp

    def mel_to_hz(m):
        return torch.where(m >= min_log_mel, min_log_hz * torch.exp(logstep * (m - min_log_mel)), f_sp * m)

    fft_f = torch.linspace(0, sr / 2, n_fft // 2 + 1, dtype=torch.float64)
    mel_f = mel_to_hz(torch.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2, dtype=torch.float64))
    fdiff = mel_f[1:] - mel_f[:-1]
    ramps = mel_f[:, None] - fft_f[None, :]
    w = torch.clamp(torch.minimum(-ramps[:-2] / fdiff[:-1, None], ramps[2:] / fdiff[1:, None]), min=0)
    return (w * (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]).float()


class MelNet(nn.Module):
    def __init__(self, hparams, device="cuda"):
        super().__init__()
        self.n_fft, self.num_mels = hparams["fft_size"], hparams["audio_num_mel_bins"]
        self.sampling_rate, self.hop_size, self.win_size = hparams["audio_sample_rate"], hparams["hop_size"], hparams["win_size"]
        self.fmin, self.fmax = hparams["fmin"], hparams["fmax"]
        if self.n_fft % self.hop_size or self.n_fft // self.hop_size > 5 or self.win_size > self.n_fft or self.hop_size % 16:
            raise ValueError("MelNet on B200 needs fft_size = k * hop_size with k <= 5, win_size <= fft_size, hop_size % 16 == 0")
        self.device = torch.device(device)
        self.mel_basis = mel_filterbank(self.sampling_rate, self.n_fft, self.num_mels, self.fmin, self.fmax)
        self.hann_window = torch.hann_window(self.win_size)
        self._packed = None

    def to(self, device, **kw):
        self.device = torch.device(device)
        self._packed = None
        return self

    def _pack(self):
        L.require_device()
        n, hop, bins = self.n_fft, self.hop_size, self.n_fft // 2 + 1
        win = torch.zeros(n, dtype=torch.float64)
        off = (n - self.win_size) // 2                       # torch.stft centres a shorter window inside the frame
        win[off:off + self.win_size] = torch.hann_window(self.win_size, dtype=torch.float64)
        ang = 2 * math.pi * torch.arange(bins, dtype=torch.float64)[:, None] * torch.arange(n, dtype=torch.float64)[None] / n
        basis = torch.stack([torch.cos(ang) * win, -torch.sin(ang) * win], 1).reshape(2 * bins, n)   # (re, im) per bin
        self.nout = (2 * bins + 15) // 16 * 16
        k = n // hop
        taps_w = torch.zeros(k, self.nout, hop, dtype=torch.float64)
        for j in range(k):
            taps_w[j, :2 * bins] = basis[:, j * hop:(j + 1) * hop]
        wb = taps_w.reshape(k * self.nout, hop).float()
        self.bins, self.bins_pad, self.k = bins, (bins + 15) // 16 * 16, k
        mb = torch.zeros(self.num_mels, self.bins_pad)
        mb[:, :bins] = self.mel_basis
        self._packed = {"basis": ops.split_weight(wb.to(self.device)), "mel": ops.split_weight(mb.to(self.device))}

    @torch.no_grad()
    def forward(self, y, center=False, complex=False):
        """y: np.ndarray / Tensor [n] or [B, n] -> log10-mel fp32 [B, n_mels, frames] on the device."""
        if center or complex:
            raise NotImplementedError("MelNet on B200 implements the path the reference uses: center=False, complex=False")
        if not torch.is_tensor(y):
            y = torch.as_tensor(y, dtype=torch.float32)
        if y.dim() == 1:
            y = y.unsqueeze(0)
        if self._packed is None:
            self._pack()
        y = y.to(self.device, torch.float32).contiguous()
        B, n = y.shape
        hop, k, nout = self.hop_size, self.k, self.nout
        pad = int((self.n_fft - hop) / 2)
        if n <= pad:
            raise ValueError(f"signal of {n} samples is shorter than the reflect pad ({pad})")
        Lp = n + 2 * pad
        F = (Lp - self.n_fft) // hop + 1
        nh = F + k - 1
        bf = torch.bfloat16
        hops = torch.empty(B, 2, nh, hop, device=self.device, dtype=bf)
        ops._call("ma3_melnet_prep", L.ptr(y), L.ptr(hops), B, n, pad, nh, hop)
        # STFT[r] = sum_j hops[r + j] . Basis_j^T, each product as hi.hi + lo.hi + hi.lo
        nb = k * nout
        taps = [(j, j * nout) for j in range(k)] + [(nh + j, j * nout) for j in range(k)] + [(j, nb + j * nout) for j in range(k)]
        S = torch.empty(B, F, nout, device=self.device, dtype=torch.float32)
        ops.gemm(hops, self._packed["basis"], M=F, N=nout, K=hop, batch=B, a_rows=2 * nh, a_batch_stride=2 * nh * hop,
                 b_rows=2 * nb, taps=taps, out=S, out_batch_stride=F * nout)
        mag = torch.empty(B, 2, F, self.bins_pad, device=self.device, dtype=bf)
        ops._call("ma3_melnet_mag", L.ptr(S), nout, L.ptr(mag), B, F, self.bins, self.bins_pad)
        mel = torch.empty(B, F, self.num_mels, device=self.device, dtype=torch.float32)
        nm = self.num_mels
        ops.gemm(mag, self._packed["mel"], M=F, N=nm, K=self.bins_pad, batch=B, a_rows=2 * F,
                 a_batch_stride=2 * F * self.bins_pad, b_rows=2 * nm, taps=((0, 0), (F, 0), (0, nm)), out=mel,
                 out_batch_stride=F * nm)
        out = torch.empty(B, nm, F, device=self.device, dtype=torch.float32)
        ops._call("ma3_melnet_log", L.ptr(mel), L.ptr(out), B, F, nm)
        return out
