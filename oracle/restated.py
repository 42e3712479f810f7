"""TEST INFRASTRUCTURE -- the parity oracle.  NOT product code: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module.

A CPU (torch fp32) restatement of the arithmetic on Make-An-Audio-3's sampling path, written functionally over
state_dicts.  Every function cites the reference lines it restates (paths relative to the reference repo).
Pinning: tests/test_oracle_golden.py checks it against golden vectors produced by the *reference modules
themselves* (tools/make_golden.py, run where /root/reference is mounted); tests/test_oracle_vs_reference.py
checks it live against the reference when that tree is present.  The only unpinned boundary is the fixed-step
Euler loop of torchdyn (un-vendored, un-pinned dependency of the reference; see oracle/_shims/torchdyn).
"""
import math

import torch
import torch.nn.functional as Fn


# ------------------------------------------------------------------------------------------------ DiT pieces
def rope_table(head_dim, end, theta=10000.0, rope_scaling_factor=1.0, ntk_factor=1.0, device=None):
    """flag_large_dit.py:212-251 (precompute_freqs_cis) -> (cos, sin), each [end, head_dim/2].  The table is always
    computed on the CPU (as the GPU tests' reference of record) and moved to `device` afterwards."""
    theta = theta * ntk_factor
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    pos = torch.arange(end, dtype=torch.float32) / rope_scaling_factor
    ang = torch.outer(pos, inv).float()
    return torch.cos(ang).to(device), torch.sin(ang).to(device)


def apply_rope(x, cos, sin):
    """flag_large_dit_moe.py:240-271: interleaved pairs (x[2i], x[2i+1]) rotated by the angle of (pos, i).
    x: [N, T, H, hd]."""
    T = x.shape[1]
    c = cos[:T][None, :, None, :]
    s = sin[:T][None, :, None, :]
    xe, xo = x[..., 0::2], x[..., 1::2]
    out = torch.empty_like(x)
    out[..., 0::2] = xe * c - xo * s
    out[..., 1::2] = xe * s + xo * c
    return out


def rmsnorm(x, w, eps=1e-5):
    """flag_large_dit_moe.py:63,76-77."""
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


def timestep_embedding(t, dim=256, max_period=10000):
    """flag_large_dit_moe.py:110-127."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _softmax_attention(q, k, v):
    """softmax(q k^T / sqrt(hd)) v with q,k,v [N, H, S, hd] (the fp32 SDPA branch, flag_large_dit_moe.py:382-402;
    the masks are all-ones on this path)."""
    s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(q.shape[-1])
    return torch.matmul(torch.softmax(s, dim=-1), v)


def attention(sd, p, x, y, cos, sin, heads):
    """flag_large_dit_moe.py:325-408 (qk_norm is Identity in every shipped config)."""
    N, T, D = x.shape
    hd = D // heads
    q, k = x @ sd[p + "wq.weight"].t(), x @ sd[p + "wk.weight"].t()
    if p + "q_norm.weight" in sd:   # qk_norm=True: nn.LayerNorm over the full model dim (flag_large_dit_moe.py:199-207,345-346)
        q = Fn.layer_norm(q, (D,), sd[p + "q_norm.weight"], sd[p + "q_norm.bias"], 1e-5)
        k = Fn.layer_norm(k, (D,), sd[p + "k_norm.weight"], sd[p + "k_norm.bias"], 1e-5)
    q, k = q.view(N, T, heads, hd), k.view(N, T, heads, hd)
    v = (x @ sd[p + "wv.weight"].t()).view(N, T, heads, hd)
    q, k = apply_rope(q, cos, sin), apply_rope(k, cos, sin)
    qh = q.permute(0, 2, 1, 3)
    o = _softmax_attention(qh, k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3))
    L = y.shape[1]
    yk = y @ sd[p + "wk_y.weight"].t()
    if p + "ky_norm.weight" in sd:
        yk = Fn.layer_norm(yk, (D,), sd[p + "ky_norm.weight"], sd[p + "ky_norm.bias"], 1e-5)
    yk = yk.view(N, L, heads, hd).permute(0, 2, 1, 3)
    yv = (y @ sd[p + "wv_y.weight"].t()).view(N, L, heads, hd).permute(0, 2, 1, 3)
    oy = _softmax_attention(qh, yk, yv) * torch.tanh(sd[p + "gate"]).view(1, heads, 1, 1)
    o = (o + oy).permute(0, 2, 1, 3).reshape(N, T, D)
    return o @ sd[p + "wo.weight"].t()


def swiglu_ffn(sd, p, x):
    """flag_large_dit_moe.py:484-489."""
    return (Fn.silu(x @ sd[p + "w1.weight"].t()) * (x @ sd[p + "w3.weight"].t())) @ sd[p + "w2.weight"].t()


def moe_ffn(sd, p, x, t_int, num_experts):
    """flag_large_dit_moe.py:516-538: time expert e = t // 250 per sample, then per-band frequency experts."""
    N, T, D = x.shape
    y = torch.zeros_like(x)
    for n in range(N):
        e = int(t_int[n]) // 250
        if 0 <= e < num_experts:
            y[n] = swiglu_ffn(sd, p + f"time_experts.{e}.", x[n])
    band = D // num_experts
    z = torch.zeros_like(y)
    for e in range(num_experts):
        masked = torch.zeros_like(y)
        masked[..., band * e: band * (e + 1)] = y[..., band * e: band * (e + 1)]
        z[..., band * e: band * (e + 1)] = swiglu_ffn(sd, p + f"freq_experts.{e}.", masked)[..., band * e: band * (e + 1)]
    return z


def dit_forward(sd, x, t, context, *, heads, video=False, num_experts=0, rope=None, max_len=1000):
    """TxtFlagLargeDiT.forward (flag_large_dit.py:177-210) / VideoFlagLargeDiT.forward
    (flag_large_dit_moe.py:664-698).  x [N,C,T] fp32, t [N] int64, context [N,L,Cd] -> [N,C,T]."""
    D = sd["proj_in.weight"].shape[0]
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    hd = D // heads
    cos, sin = rope if rope is not None else rope_table(hd, max_len, device=x.device)
    h = x.transpose(1, 2) @ sd["proj_in.weight"].t() + sd["proj_in.bias"]
    te = timestep_embedding(t)
    te = Fn.silu(te @ sd["t_embedder.mlp.0.weight"].t() + sd["t_embedder.mlp.0.bias"])
    te = te @ sd["t_embedder.mlp.2.weight"].t() + sd["t_embedder.mlp.2.bias"]
    if video:
        c = context @ sd["c_embedder.mlp.0.weight"].t() + sd["c_embedder.mlp.0.bias"]
        c = Fn.gelu(c) @ sd["c_embedder.mlp.2.weight"].t() + sd["c_embedder.mlp.2.bias"]
        y = Fn.layer_norm(c, (D,), sd["c_embedder.mlp.3.weight"], sd["c_embedder.mlp.3.bias"], 1e-5)
    else:
        y = context
    pool = y.mean(dim=1)
    cap = Fn.layer_norm(pool, (pool.shape[-1],), sd["cap_embedder.0.weight"], sd["cap_embedder.0.bias"], 1e-5)
    cap = cap @ sd["cap_embedder.1.weight"].t() + sd["cap_embedder.1.bias"]
    a = te + cap
    sa = Fn.silu(a)
    for i in range(depth):
        p = f"blocks.{i}."
        mod = sa @ sd[p + "adaLN_modulation.1.weight"].t() + sd[p + "adaLN_modulation.1.bias"]
        sh1, sc1, g1, sh2, sc2, g2 = [m.unsqueeze(1) for m in mod.chunk(6, dim=1)]
        u = rmsnorm(h, sd[p + "attention_norm.weight"]) * (1 + sc1) + sh1
        yn = rmsnorm(y, sd[p + "attention_y_norm.weight"])
        h = h + g1 * attention(sd, p + "attention.", u, yn, cos, sin, heads)
        zz = rmsnorm(h, sd[p + "ffn_norm.weight"]) * (1 + sc2) + sh2
        if num_experts:
            f = moe_ffn(sd, p + "feed_forward.", zz, t, num_experts)
        else:
            f = swiglu_ffn(sd, p + "feed_forward.", zz)
        h = h + g2 * f
    mod = sa @ sd["final_layer.adaLN_modulation.1.weight"].t() + sd["final_layer.adaLN_modulation.1.bias"]
    shift, scale = [m.unsqueeze(1) for m in mod.chunk(2, dim=1)]
    h = Fn.layer_norm(h, (D,), None, None, 1e-6) * (1 + scale) + shift
    out = h @ sd["final_layer.linear.weight"].t() + sd["final_layer.linear.bias"]
    return out.transpose(1, 2)


# ------------------------------------------------------------------------------------------------ sampler
def timestep_ints(n_points, t_start=None):
    """Integer timesteps the DiT sees: long(t*1000) with t advanced as t+dt in fp32
    (cfm1_audio.py:103-105,156; Euler restated in oracle/_shims/torchdyn)."""
    ts = torch.linspace(0, 1, n_points)
    if t_start is not None:
        ts = ts[t_start:]
    t = ts[0]
    dt = ts[1] - t
    ints, dts = [], []
    for k in range(1, len(ts)):
        ints.append(int((t * 1000).long()))
        dts.append(float(dt))
        t = t + dt
        if k < len(ts) - 1:
            dt = ts[k + 1] - t
    return ints, dts


def sample_cfg(velocity, x0, cond, uncond, scale, n_points=25, t_start=None):
    """CFM.sample_cfg + Wrapper_cfg.forward (cfm1_audio.py:89-111,145-161).  velocity(x, t_int64[N], ctx) is the
    DiT; batch order is [uncond, cond].  Returns (x_final, traj [n_points, B, C, T], per-step guided velocities)."""
    ints, dts = timestep_ints(n_points, t_start)
    x = x0
    traj, vels = [x], []
    B = x.shape[0]
    for ti, dt in zip(ints, dts):
        t = torch.full((2 * B,), ti, dtype=torch.long, device=x.device)
        v = velocity(torch.cat([x, x]), t, torch.cat([uncond, cond]))
        vu, vc = v[:B], v[B:]
        vg = vu + scale * (vc - vu)
        vels.append(vg)
        x = x + dt * vg
        traj.append(x)
    return x, torch.stack(traj), vels


def sample_plain(velocity, x0, cond, n_points=25, t_start=None):
    """CFM.sample + Wrapper.forward (cfm1_audio.py:60-82,134-142)."""
    ints, dts = timestep_ints(n_points, t_start)
    x = x0
    traj = [x]
    for ti, dt in zip(ints, dts):
        t = torch.full((x.shape[0],), ti, dtype=torch.long, device=x.device)
        x = x + dt * velocity(x, t, cond)
        traj.append(x)
    return x, torch.stack(traj)


# ------------------------------------------------------------------------------------------------ VAE decoder
def _gn(x, sd, name, groups=32, eps=1e-6):
    return Fn.group_norm(x, groups, sd[name + ".weight"], sd[name + ".bias"], eps)


def _swish(x):
    return x * torch.sigmoid(x)


def _c1d(x, sd, name, padding=0, dilation=1):
    return Fn.conv1d(x, sd[name + ".weight"], sd[name + ".bias"], padding=padding, dilation=dilation)


def _resblock(x, sd, name):
    """autoencoder1d.py:215-235 (temb is None; dropout 0); kernel 3 in the decoder, `kernel_size` in the encoder."""
    pad = sd[name + ".conv1.weight"].shape[-1] // 2
    h = _c1d(_swish(_gn(x, sd, name + ".norm1")), sd, name + ".conv1", padding=pad)
    h = _c1d(_swish(_gn(h, sd, name + ".norm2")), sd, name + ".conv2", padding=pad)
    if name + ".nin_shortcut.weight" in sd:
        x = _c1d(x, sd, name + ".nin_shortcut")
    return x + h


def _attnblock(x, sd, name):
    """autoencoder1d.py:257-278; the scale is C**-0.5 (the reference names the channel count `t`)."""
    h = _gn(x, sd, name + ".norm")
    q, k, v = _c1d(h, sd, name + ".q"), _c1d(h, sd, name + ".k"), _c1d(h, sd, name + ".v")
    C = q.shape[1]
    w = torch.softmax(torch.bmm(q.transpose(1, 2), k) * (C ** -0.5), dim=2)  # [B, Tq, Tk]
    h = torch.bmm(v, w.transpose(1, 2))
    return x + _c1d(h, sd, name + ".proj_out")


def vae_decode(sd, z, ddconfig, scale_factor=1.0):
    """decode_first_stage (ddpm_audio.py:358-371) -> AutoencoderKL.decode (autoencoder1d.py:59-62) ->
    Decoder1D.forward (autoencoder1d.py:484-517)."""
    ch_mult = list(ddconfig["ch_mult"])
    nrb = ddconfig["num_res_blocks"]
    ks = ddconfig.get("kernel_size", 3)
    down_layers = [i + 1 for i in ddconfig.get("down_layers", [])]
    attn_layers = ddconfig.get("attn_layers", [])
    z = (1.0 / scale_factor) * z
    h = _c1d(z, sd, "post_quant_conv")
    h = _c1d(h, sd, "decoder.conv_in", padding=ks // 2)
    h = _resblock(h, sd, "decoder.mid.block_1")
    h = _attnblock(h, sd, "decoder.mid.attn_1")
    h = _resblock(h, sd, "decoder.mid.block_2")
    for lvl in reversed(range(len(ch_mult))):
        for ib in range(nrb + 1):
            h = _resblock(h, sd, f"decoder.up.{lvl}.block.{ib}")
            if lvl in attn_layers:
                h = _attnblock(h, sd, f"decoder.up.{lvl}.attn.{ib}")
        if lvl in down_layers:
            h = h.repeat_interleave(2, dim=2)  # F.interpolate(scale_factor=2, mode='nearest'), :291-295
            h = _c1d(h, sd, f"decoder.up.{lvl}.upsample.conv", padding=1)
    h = _swish(_gn(h, sd, "decoder.norm_out"))
    return _c1d(h, sd, "decoder.conv_out", padding=ks // 2)


def vae_encode(sd, x, ddconfig):
    """AutoencoderKL.encode (autoencoder1d.py:49-53) -> Encoder1D.forward (autoencoder1d.py:319-413) -> quant_conv.
    x [B, in_channels, T] -> moments [B, 2 * embed_dim, T / 2^len(down_layers)] (mean | logvar, before the clamp of
    DiagonalGaussianDistribution, ldm/modules/distributions/distributions.py:24-35)."""
    ch_mult = list(ddconfig["ch_mult"])
    nrb = ddconfig["num_res_blocks"]
    ks = ddconfig.get("kernel_size", 3)
    down_layers = list(ddconfig.get("down_layers", []))
    attn_layers = ddconfig.get("attn_layers", [])
    h = _c1d(x, sd, "encoder.conv_in", padding=ks // 2)
    for lvl in range(len(ch_mult)):
        for ib in range(nrb):
            h = _resblock(h, sd, f"encoder.down.{lvl}.block.{ib}")
            if lvl in attn_layers:
                h = _attnblock(h, sd, f"encoder.down.{lvl}.attn.{ib}")
        if lvl in down_layers:   # Downsample1D: zero-pad one sample on the right, conv k3 stride 2 (:296-317)
            h = Fn.conv1d(Fn.pad(h, (0, 1)), sd[f"encoder.down.{lvl}.downsample.conv.weight"],
                          sd[f"encoder.down.{lvl}.downsample.conv.bias"], stride=2)
    h = _resblock(h, sd, "encoder.mid.block_1")
    h = _attnblock(h, sd, "encoder.mid.attn_1")
    h = _resblock(h, sd, "encoder.mid.block_2")
    h = _c1d(_swish(_gn(h, sd, "encoder.norm_out")), sd, "encoder.conv_out", padding=ks // 2)
    return _c1d(h, sd, "quant_conv")


def posterior_mode_sample(moments, noise=None):
    """DiagonalGaussianDistribution (distributions.py:24-44): mean, logvar clamped to [-30, 20], std = exp(logvar / 2)."""
    mean, logvar = moments.chunk(2, dim=1)
    std = torch.exp(0.5 * logvar.clamp(-30.0, 20.0))
    return mean, (mean + std * noise if noise is not None else None)


# ------------------------------------------------------------------------------------------------ mel front-end
def slaney_mel_filterbank(sr, n_fft, n_mels, fmin, fmax):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) with its defaults (htk=False, norm='slaney') -- restated from
    the published algorithm because librosa is not installed here (preprocess/NAT_mel.py:54 calls it): linear below
    1 kHz (200/3 Hz per mel), logarithmic above (step log(6.4)/27), triangular filters on the FFT bin centres,
    each scaled by 2 / (f_hi - f_lo).  -> [n_mels, n_fft/2 + 1]"""
    def hz_to_mel(f):
        f = torch.as_tensor(f, dtype=torch.float64)
        lin = f / (200.0 / 3)
        log = 15.0 + torch.log(f.clamp_min(1e-10) / 1000.0) / (math.log(6.4) / 27.0)
        return torch.where(f >= 1000.0, log, lin)

    def mel_to_hz(m):
        lin = m * (200.0 / 3)
        log = 1000.0 * torch.exp((math.log(6.4) / 27.0) * (m - 15.0))
        return torch.where(m >= 15.0, log, lin)

    fft_f = torch.linspace(0, sr / 2, n_fft // 2 + 1, dtype=torch.float64)
    mel_f = mel_to_hz(torch.linspace(float(hz_to_mel(fmin)), float(hz_to_mel(fmax)), n_mels + 2, dtype=torch.float64))
    fdiff = mel_f[1:] - mel_f[:-1]
    ramps = mel_f[:, None] - fft_f[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = torch.clamp(torch.minimum(lower, upper), min=0)
    w = w * (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.float()


def melnet(y, hp, mel_basis=None):
    """MelNet.forward (preprocess/NAT_mel.py:65-85, center=False, complex=False): clamp to [-1, 1], reflect-pad
    (n_fft - hop) / 2 on both sides, STFT (Hann window of win_size), magnitude sqrt(re^2 + im^2 + 1e-9), mel filterbank,
    log10(clamp(., 1e-5)).  y [B, samples] -> [B, n_mels, frames]."""
    n_fft, hop, win = hp["fft_size"], hp["hop_size"], hp["win_size"]
    if mel_basis is None:
        mel_basis = slaney_mel_filterbank(hp["audio_sample_rate"], n_fft, hp["audio_num_mel_bins"], hp["fmin"], hp["fmax"])
    y = y.clamp(-1.0, 1.0)
    pad = int((n_fft - hop) / 2)
    y = Fn.pad(y.unsqueeze(1), [pad, pad], mode="reflect").squeeze(1)
    spec = torch.stft(y, n_fft, hop_length=hop, win_length=win, window=torch.hann_window(win).to(y.device), center=False,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    mag = torch.sqrt(spec.real.pow(2) + spec.imag.pow(2) + 1e-9)
    return torch.log10(torch.clamp(torch.matmul(mel_basis.to(y.device), mag), min=1e-5))


# ------------------------------------------------------------------------------------------------ BigVGAN
def kaiser_sinc_filter(cutoff=0.25, half_width=0.3, kernel_size=12):
    """vocoder/bigvgan/alias_free_torch/filter.py:28-57 -> [kernel_size] taps (sum 1)."""
    half = kernel_size // 2
    delta_f = 4 * half_width
    A = 2.285 * (half - 1) * math.pi * delta_f + 7.95
    if A > 50.0:
        beta = 0.1102 * (A - 8.7)
    elif A >= 21.0:
        beta = 0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0)
    else:
        beta = 0.0
    window = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    if kernel_size % 2 == 0:
        time = torch.arange(-half, half) + 0.5
    else:
        time = torch.arange(kernel_size) - half
    f = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    return f / f.sum()


def up2(x, f):
    """UpSample1d.forward, ratio 2, 12 taps (alias_free_torch/resample.py:25-33)."""
    C = x.shape[1]
    K = f.numel()
    pad = K // 2 - 1
    pad_left = pad * 2 + (K - 2) // 2
    pad_right = pad * 2 + (K - 2 + 1) // 2
    x = Fn.pad(x, (pad, pad), mode="replicate")
    x = 2 * Fn.conv_transpose1d(x, f.view(1, 1, K).expand(C, 1, K), stride=2, groups=C)
    return x[..., pad_left:-pad_right]


def down2(x, f):
    """DownSample1d / LowPassFilter1d.forward, stride 2 (alias_free_torch/filter.py:86-95)."""
    C = x.shape[1]
    K = f.numel()
    x = Fn.pad(x, (K // 2 - 1, K // 2), mode="replicate")
    return Fn.conv1d(x, f.view(1, 1, K).expand(C, 1, K), stride=2, groups=C)


def snakebeta(x, alpha, beta, logscale=True):
    """vocoder/bigvgan/activations.py:107-119."""
    a = alpha.view(1, -1, 1)
    b = beta.view(1, -1, 1)
    if logscale:
        a, b = torch.exp(a), torch.exp(b)
    return x + (1.0 / (b + 1e-9)) * torch.sin(x * a) ** 2


def snake(x, alpha, logscale=True):
    """vocoder/bigvgan/activations.py:48-59."""
    a = alpha.view(1, -1, 1)
    if logscale:
        a = torch.exp(a)
    return x + (1.0 / (a + 1e-9)) * torch.sin(x * a) ** 2


def activation1d(x, sd, name, h, f=None):
    """Activation1d.forward (alias_free_torch/act.py:23-28): up x2 -> snake(beta) -> down x2."""
    f = (kaiser_sinc_filter() if f is None else f).to(x.device)
    u = up2(x, f)
    if h["activation"] == "snakebeta":
        u = snakebeta(u, sd[name + ".act.alpha"], sd[name + ".act.beta"], h["snake_logscale"])
    else:
        u = snake(u, sd[name + ".act.alpha"], h["snake_logscale"])
    return down2(u, f)


def _amp1(x, sd, p, h, k, dils, f):
    """AMPBlock1.forward (vocoder/bigvgan/models.py:74-83)."""
    for l, d in enumerate(dils):
        xt = activation1d(x, sd, p + f"activations.{2 * l}", h, f)
        xt = _c1d(xt, sd, p + f"convs1.{l}", padding=(k * d - d) // 2, dilation=d)
        xt = activation1d(xt, sd, p + f"activations.{2 * l + 1}", h, f)
        xt = _c1d(xt, sd, p + f"convs2.{l}", padding=(k - 1) // 2)
        x = xt + x
    return x


def _amp2(x, sd, p, h, k, dils, f):
    """AMPBlock2.forward (vocoder/bigvgan/models.py:124-129)."""
    for l, d in enumerate(dils):
        xt = activation1d(x, sd, p + f"activations.{l}", h, f)
        xt = _c1d(xt, sd, p + f"convs.{l}", padding=(k * d - d) // 2, dilation=d)
        x = xt + x
    return x


def bigvgan_forward(sd, mel, h):
    """BigVGAN.forward (vocoder/bigvgan/models.py:183-205) on weight-norm-folded weights.  mel [B,80,T] ->
    [B,1,T*hop]."""
    f = kaiser_sinc_filter().to(mel.device)
    nk = len(h["resblock_kernel_sizes"])
    amp = _amp1 if h["resblock"] == "1" else _amp2
    x = _c1d(mel, sd, "conv_pre", padding=3)
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = Fn.conv_transpose1d(x, sd[f"ups.{i}.0.weight"], sd[f"ups.{i}.0.bias"], stride=u, padding=(k - u) // 2)
        xs = None
        for j, (rk, dils) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            o = amp(x, sd, f"resblocks.{i * nk + j}.", h, rk, dils, f)
            xs = o if xs is None else xs + o
        x = xs / nk
    x = activation1d(x, sd, "activation_post", h, f)
    x = _c1d(x, sd, "conv_post", padding=3)
    return torch.tanh(x)


def fold_weight_norm(sd):
    """remove_weight_norm (vocoder/bigvgan/models.py:207-215): W = g * v / ||v||, norm over all dims but 0."""
    out = {}
    for k, v in sd.items():
        if k.endswith(".weight_g"):
            base = k[: -len(".weight_g")]
            vv = sd[base + ".weight_v"]
            nrm = vv.reshape(vv.shape[0], -1).norm(dim=1).view(-1, *([1] * (vv.dim() - 1)))
            out[base + ".weight"] = v * vv / nrm
        elif k.endswith(".weight_v"):
            continue
        else:
            out[k] = v
    return out


# ------------------------------------------------------------------------------------------------ device
def strict_fp32():
    """The oracle is fp32 arithmetic.  On a CUDA device torch would otherwise run convolutions (cuDNN) in TF32;
    callers that put the oracle on `cuda` (the full-size parity tests, bench.py's library baseline in fp32 mode)
    call this first."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")


def to_device(sd, device):
    return {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in sd.items()}


# ------------------------------------------------------------------------------------------------ metrics
def max_rel_err(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30))


def snr_db(ref, out):
    ref, out = ref.double().flatten(), out.double().flatten()
    return float(10 * torch.log10(ref.pow(2).sum() / (ref - out).pow(2).sum().clamp_min(1e-30)))
