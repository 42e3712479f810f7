"""Drop-in BigVGAN vocoder: `VocoderBigVGAN(ckpt_vocoder_dir, device).vocode(mel)` and `BigVGAN(h).forward(mel)` of
vocoder/bigvgan/models.py:135-215,394-595.

The nn.Module tree holds the (weight-norm-folded) parameters under the reference's state_dict names; checkpoints
still carrying `weight_g` / `weight_v` are folded on load (the reference does it with remove_weight_norm(),
models.py:551).  The forward pass runs channels-last in fp16: every dense / dilated Conv1d and each phase of the
ConvTranspose1d upsamplers is a tap-GEMM on tcgen05, every Activation1d is the fused anti-aliased Snake kernel, the
residual add and the 1/num_kernels averaging of the AMP blocks live in the GEMM epilogues, and tanh is the epilogue of
conv_post.
"""
import json
import os

import numpy as np
import torch
import torch.nn as nn

from . import lib as L
from . import ops
from .convs import PackedConv, PackedConvTranspose, cpad


class AttrDict(dict):
    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.__dict__ = self


class _P(nn.Module):
    def __init__(self, wshape, nbias):
        super().__init__()
        fan = wshape[1] * wshape[2] if len(wshape) == 3 else wshape[-1]
        a = (1.0 / max(fan, 1)) ** 0.5
        self.weight = nn.Parameter(torch.empty(*wshape).uniform_(-a, a))
        self.bias = nn.Parameter(torch.empty(nbias).uniform_(-a, a))


class _Act(nn.Module):
    """Activation1d holder: `.act.alpha` / `.act.beta` (Snake has no beta)."""

    def __init__(self, c, beta=True):
        super().__init__()
        self.act = nn.Module()
        self.act.alpha = nn.Parameter(torch.zeros(c))
        if beta:
            self.act.beta = nn.Parameter(torch.zeros(c))


class _AMP(nn.Module):
    def __init__(self, kind, c, k, dils, beta):
        super().__init__()
        mk = lambda: nn.ModuleList([_P((c, c, k), c) for _ in dils])
        if kind == "1":
            self.convs1, self.convs2 = mk(), mk()
            n = 2 * len(dils)
        else:
            self.convs = mk()
            n = len(dils)
        self.activations = nn.ModuleList([_Act(c, beta) for _ in range(n)])


def fold_weight_norm(sd):
    """W = g * v / ||v|| with the norm over every dim but 0 (torch weight_norm default, dim=0)."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            base = k[: -len(".weight_g")]
            vv = sd[base + ".weight_v"].float()
            nrm = vv.reshape(vv.shape[0], -1).norm(dim=1).view(-1, *([1] * (vv.dim() - 1)))
            out[base + ".weight"] = v.float() * vv / nrm
        elif not k.endswith(".weight_v"):
            out[k] = v
    return out


class BigVGAN(nn.Module):
    """Generator (vocoder/bigvgan/models.py:135-215), parameterised on the hyper-parameter object `h`."""

    def __init__(self, h):
        super().__init__()
        h = h if isinstance(h, dict) else dict(vars(h))
        self.h = AttrDict(h)
        h = self.h
        if h.activation not in ("snake", "snakebeta"):
            raise NotImplementedError("activation incorrectly specified. check the config file and look for 'activation'.")
        beta = h.activation == "snakebeta"
        self.num_kernels, self.num_upsamples = len(h.resblock_kernel_sizes), len(h.upsample_rates)
        C0 = h.upsample_initial_channel
        self.conv_pre = _P((C0, h.num_mels, 7), C0)
        self.ups = nn.ModuleList()
        self.resblocks = nn.ModuleList()
        ch = C0
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            cin, ch = C0 // (2 ** i), C0 // (2 ** (i + 1))
            self.ups.append(nn.ModuleList([_P((cin, ch, k), ch)]))
            for rk, d in zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes):
                self.resblocks.append(_AMP(str(h.resblock), ch, rk, d, beta))
        self.activation_post = _Act(ch, beta)
        self.conv_post = _P((1, ch, 7), 1)
        self._packed = None
        self._bufs = {}

    def remove_weight_norm(self):
        """Kept for interface parity: weights are folded when the state_dict is loaded."""

    def load_state_dict(self, state_dict, strict=True, **kw):
        sd = fold_weight_norm(dict(state_dict))
        sd = {k: v for k, v in sd.items() if not k.endswith("filter")}  # registered FIR buffers: recomputed here
        self._packed = None
        return super().load_state_dict(sd, strict=strict, **kw)

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    # ---------------------------------------------------------------- packing
    def _pack(self):
        dev = self.conv_pre.weight.device
        if dev.type != "cuda":
            raise L.Ma3Error("ma3_b200 modules run on CUDA only (no CPU fallback): call .cuda() first")
        L.require_device()
        h, dt = self.h, torch.float16
        pc = lambda m, **kw: PackedConv(m.weight, m.bias, dtype=dt, device=dev, **kw)

        def act(m, c):
            cp = cpad(c)
            al = torch.zeros(cp, device=dev)
            al[:c] = m.act.alpha.detach().float()
            be = None
            if hasattr(m.act, "beta"):
                be = torch.zeros(cp, device=dev)
                be[:c] = m.act.beta.detach().float()
            return al, be

        p = {"pre": pc(self.conv_pre), "stages": []}
        C0 = h.upsample_initial_channel
        for i, (u, k) in enumerate(zip(h.upsample_rates, h.upsample_kernel_sizes)):
            ch = C0 // (2 ** (i + 1))
            up = self.ups[i][0]
            st = {"up": PackedConvTranspose(up.weight, up.bias, stride=u, padding=(k - u) // 2, dtype=dt, device=dev),
                  "ch": ch, "blocks": []}
            for j, (rk, dils) in enumerate(zip(h.resblock_kernel_sizes, h.resblock_dilation_sizes)):
                rb = self.resblocks[i * self.num_kernels + j]
                layers = []
                for l, d in enumerate(dils):
                    if str(h.resblock) == "1":
                        layers.append((act(rb.activations[2 * l], ch), pc(rb.convs1[l], dilation=d),
                                       act(rb.activations[2 * l + 1], ch), pc(rb.convs2[l])))
                    else:
                        layers.append((act(rb.activations[l], ch), pc(rb.convs[l], dilation=d), None, None))
                st["blocks"].append(layers)
            p["stages"].append(st)
        p["act_post"] = act(self.activation_post, ch)
        p["post"] = pc(self.conv_post)
        self._packed = p

    def _buf(self, name, shape, zero=True, dtype=torch.float16):
        key = (name, tuple(shape), dtype)
        b = self._bufs.get(key)
        if b is None:
            # zero-filled once: padded channels are never written by the GEMM epilogues and stay exactly zero
            b = torch.zeros(*shape, device=self.conv_pre.weight.device, dtype=dtype)
            self._bufs[key] = b
        return b

    # ---------------------------------------------------------------- forward
    @torch.no_grad()
    def forward(self, x):
        """mel fp32 [B, num_mels, T] -> waveform fp32 [B, 1, T * prod(upsample_rates)] (models.py:183-205)."""
        if self._packed is None:
            self._pack()
        p, h = self._packed, self.h
        dev = self.conv_pre.weight.device
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        B, nm, T = x.shape
        logscale = bool(h.snake_logscale)
        mel = ops.nct_to_ntc(x, self._buf("mel", (B, T, p["pre"].cin_pad)))
        cur = p["pre"](mel, self._buf("pre", (B, T, cpad(p["pre"].cout))))
        inv_nk = 1.0 / self.num_kernels
        for si, st in enumerate(p["stages"]):
            T = st["up"].out_len(T)
            cp = cpad(st["ch"])
            xs = st["up"](cur, self._buf(f"x{si}", (B, T, cp)))
            a1 = self._buf(f"a{si}", (B, T, cp))
            mid = self._buf(f"m{si}", (B, T, cp))
            xa = self._buf(f"xa{si}", (B, T, cp))
            xb = self._buf(f"xb{si}", (B, T, cp))
            acc = self._buf(f"s{si}", (B, T, cp))
            for j, layers in enumerate(st["blocks"]):
                xcur = xs
                for l, (act_a, conv_a, act_b, conv_b) in enumerate(layers):
                    last = l == len(layers) - 1
                    # AMP blocks are averaged: the last conv of block j accumulates (x_j)/num_kernels into `acc`
                    dst = acc if last else (xb if xcur.data_ptr() == xa.data_ptr() else xa)
                    ops.act1d(xcur, a1, act_a[0], act_a[1], logscale)
                    if conv_b is None:  # AMPBlock2: x = conv(act(x)) + x
                        conv_a(a1, dst, res=xcur, alpha=inv_nk if last else 1.0, accumulate=last and j > 0)
                    else:               # AMPBlock1: x = conv2(act(conv1(act(x)))) + x
                        conv_a(a1, mid)
                        ops.act1d(mid, a1, act_b[0], act_b[1], logscale)
                        conv_b(a1, dst, res=xcur, alpha=inv_nk if last else 1.0, accumulate=last and j > 0)
                    xcur = dst
            cur = acc
        cp = cur.shape[-1]
        a1 = ops.act1d(cur, self._buf("apost", (B, T, cp)), p["act_post"][0], p["act_post"][1], logscale)
        wav = torch.empty(B, 1, T, device=dev, dtype=torch.float32)
        p["post"](a1, wav.view(B, T, 1), act=3)
        return wav


class VocoderBigVGAN(nn.Module):
    """vocoder/bigvgan/models.py:394-595.  `ckpt_vocoder_dir` holds config.json (or args.yml) and the generator
    weights; the YAML configs pass the directory as `ckpt_vocoder` (configs/txt2audio-cfm-cfg.yaml:101), the scripts
    positionally -- both are accepted.  For benchmarks / tests without a checkpoint pass `h` (+ optional state_dict)."""

    WEIGHT_NAMES = ["generator.pth.tar", "generator.pt", "generator.pth", "g_02500000", "g_02500000.pth",
                    "bigvgan_generator.pt", "best_netG.pt"]

    def __init__(self, ckpt_vocoder_dir=None, device="cuda", ckpt_vocoder=None, h=None, state_dict=None):
        super().__init__()
        self.device = torch.device(device)
        ckpt_dir = ckpt_vocoder_dir if ckpt_vocoder_dir is not None else ckpt_vocoder
        if h is None:
            if ckpt_dir is None:
                raise FileNotFoundError("VocoderBigVGAN needs a checkpoint directory or explicit hyper-parameters `h`")
            cj, cy = os.path.join(ckpt_dir, "config.json"), os.path.join(ckpt_dir, "args.yml")
            if os.path.exists(cj):
                with open(cj) as f:
                    h = json.load(f)
            elif os.path.exists(cy):
                import yaml
                with open(cy) as f:
                    h = yaml.safe_load(f)
            else:
                raise FileNotFoundError(f"neither 'config.json' nor 'args.yml' found in the vocoder directory: {ckpt_dir}")
            if h is None:
                raise ValueError("could not load the vocoder configuration")
        self.generator = BigVGAN(h)
        if state_dict is None and ckpt_dir is not None:
            path = next((os.path.join(ckpt_dir, n) for n in self.WEIGHT_NAMES if os.path.exists(os.path.join(ckpt_dir, n))),
                        None)
            if path is None:
                raise FileNotFoundError(f"no recognised generator weights in {ckpt_dir}; looked for {self.WEIGHT_NAMES}")
            ckpt = torch.load(path, map_location="cpu", weights_only=False)
            if isinstance(ckpt, dict) and "generator" in ckpt:
                state_dict = ckpt["generator"]
            elif isinstance(ckpt, dict) and "model" in ckpt:
                state_dict = ckpt["model"]
            elif isinstance(ckpt, dict) and "state_dict" in ckpt:
                state_dict = {k.replace("generator.", ""): v for k, v in ckpt["state_dict"].items()
                              if k.startswith("generator.")} or ckpt["state_dict"]
            else:
                state_dict = ckpt
        if state_dict is not None:
            self.generator.load_state_dict(state_dict, strict=False)
        self.generator.to(self.device).eval()

    @torch.no_grad()
    def vocode_tensor(self, mel):
        """Device-resident variant: mel [B, num_mels, T] tensor -> waveform tensor [B, T*hop] on the device (no
        host round trip; used by the pipeline / benchmark)."""
        return self.generator(mel).squeeze(1)

    def vocode(self, mel_spectrogram_input):
        with torch.no_grad():
            if isinstance(mel_spectrogram_input, np.ndarray):
                mel = torch.from_numpy(mel_spectrogram_input).unsqueeze(0)
            elif isinstance(mel_spectrogram_input, torch.Tensor):
                mel = mel_spectrogram_input
                if mel.dim() == 2:
                    mel = mel.unsqueeze(0)
            else:
                raise TypeError(f"vocoder input must be a NumPy array or a PyTorch tensor, got {type(mel_spectrogram_input)}")
            mel = mel.to(dtype=torch.float32, device=self.device)
            n_mels = getattr(self.generator.h, "num_mels", None)
            if n_mels is not None and mel.shape[1] != n_mels:
                raise ValueError(f"shape mismatch for n_mels: input has {mel.shape[1]}, vocoder expects {n_mels}")
            wav = self.generator(mel)
            return wav.squeeze(1).squeeze(0).cpu().numpy()

    def __call__(self, mel_spectrogram_input):
        return self.vocode(mel_spectrogram_input)
