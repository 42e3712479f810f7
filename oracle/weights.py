"""TEST INFRASTRUCTURE (oracle): deterministic random-init weights for the sampling path, keyed exactly like the
reference's state_dicts so the same tensors load into the reference modules and into the B200 drop-ins.

The reference's own constructors zero-initialise adaLN / final layer / attention gate
(ldm/modules/diffusionmodules/flag_large_dit.py:288-297, flag_large_dit_moe.py:192), which makes a fresh model
output exactly 0; parity on such weights proves nothing, so every tensor here is drawn non-zero from a seeded
generator.  Nothing here depends on /root/reference (it must run on the GPU box too).
"""
import math

import torch


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _lin(g, out_f, in_f, gain=1.0):
    a = gain * math.sqrt(3.0 / in_f)
    return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * a


def _vec(g, n, std=0.02, mean=0.0):
    return torch.randn(n, generator=g) * std + mean


def ffn_hidden(dim, multiple_of=256):
    h = int(2 * (4 * dim) / 3)
    return multiple_of * ((h + multiple_of - 1) // multiple_of)


def dit_state_dict(*, in_channels, context_dim, hidden_size, num_heads, depth, video=False, num_experts=0, seed=0,
                   qk_norm=False):
    """Keys of TxtFlagLargeImprovedDiTV2 (flag_large_dit.py:256-299) or VideoFlagLargeDiT
    (flag_large_dit_moe.py:613-662)."""
    g = _gen(seed)
    D, Cd = hidden_size, context_dim
    F = ffn_hidden(D)
    y_dim = D if video else Cd
    sd = {}
    sd["t_embedder.mlp.0.weight"] = _lin(g, D, 256)
    sd["t_embedder.mlp.0.bias"] = _vec(g, D)
    sd["t_embedder.mlp.2.weight"] = _lin(g, D, D)
    sd["t_embedder.mlp.2.bias"] = _vec(g, D)
    sd["proj_in.weight"] = _lin(g, D, in_channels)
    sd["proj_in.bias"] = _vec(g, D)
    if video:
        sd["c_embedder.mlp.0.weight"] = _lin(g, D, Cd)
        sd["c_embedder.mlp.0.bias"] = _vec(g, D)
        sd["c_embedder.mlp.2.weight"] = _lin(g, D, D)
        sd["c_embedder.mlp.2.bias"] = _vec(g, D)
        sd["c_embedder.mlp.3.weight"] = _vec(g, D, 0.1, 1.0)
        sd["c_embedder.mlp.3.bias"] = _vec(g, D, 0.05)
    cap_in = D if video else Cd
    sd["cap_embedder.0.weight"] = _vec(g, cap_in, 0.1, 1.0)
    sd["cap_embedder.0.bias"] = _vec(g, cap_in, 0.05)
    sd["cap_embedder.1.weight"] = _lin(g, D, cap_in)
    sd["cap_embedder.1.bias"] = _vec(g, D)
    for i in range(depth):
        p = f"blocks.{i}."
        for n in ("wq", "wk", "wv", "wo"):
            sd[p + f"attention.{n}.weight"] = _lin(g, D, D)
        sd[p + "attention.wk_y.weight"] = _lin(g, D, y_dim)
        sd[p + "attention.wv_y.weight"] = _lin(g, D, y_dim)
        sd[p + "attention.gate"] = _vec(g, num_heads, 0.5)
        if qk_norm:   # nn.LayerNorm(D) on q, k and the cross k (flag_large_dit_moe.py:199-207)
            for n in ("q_norm", "k_norm", "ky_norm"):
                sd[p + f"attention.{n}.weight"] = _vec(g, D, 0.1, 1.0)
                sd[p + f"attention.{n}.bias"] = _vec(g, D, 0.05)
        if num_experts:
            for kind in ("time_experts", "freq_experts"):
                for e in range(num_experts):
                    q = p + f"feed_forward.{kind}.{e}."
                    sd[q + "w1.weight"] = _lin(g, F, D)
                    sd[q + "w2.weight"] = _lin(g, D, F)
                    sd[q + "w3.weight"] = _lin(g, F, D)
        else:
            sd[p + "feed_forward.w1.weight"] = _lin(g, F, D)
            sd[p + "feed_forward.w2.weight"] = _lin(g, D, F)
            sd[p + "feed_forward.w3.weight"] = _lin(g, F, D)
        sd[p + "attention_norm.weight"] = _vec(g, D, 0.1, 1.0)
        sd[p + "ffn_norm.weight"] = _vec(g, D, 0.1, 1.0)
        sd[p + "attention_y_norm.weight"] = _vec(g, y_dim, 0.1, 1.0)
        # small but non-zero modulation so gates / shifts / scales all matter
        sd[p + "adaLN_modulation.1.weight"] = _lin(g, 6 * D, D, gain=0.5)
        sd[p + "adaLN_modulation.1.bias"] = _vec(g, 6 * D, 0.1)
    sd["final_layer.linear.weight"] = _lin(g, in_channels, D)
    sd["final_layer.linear.bias"] = _vec(g, in_channels)
    sd["final_layer.adaLN_modulation.1.weight"] = _lin(g, 2 * D, D, gain=0.5)
    sd["final_layer.adaLN_modulation.1.bias"] = _vec(g, 2 * D, 0.1)
    return sd


def _conv(g, sd, name, cout, cin, k, gain=1.0):
    a = gain * math.sqrt(1.0 / (cin * k))
    sd[name + ".weight"] = (torch.rand(cout, cin, k, generator=g) * 2 - 1) * a
    sd[name + ".bias"] = (torch.rand(cout, generator=g) * 2 - 1) * a


def _norm(g, sd, name, c):
    sd[name + ".weight"] = _vec(g, c, 0.1, 1.0)
    sd[name + ".bias"] = _vec(g, c, 0.05)


def vae_decoder_state_dict(ddconfig, embed_dim, seed=1):
    """Keys of AutoencoderKL.{post_quant_conv, decoder} (ldm/models/autoencoder1d.py:18-62, 415-482)."""
    g = _gen(seed)
    ch, ch_mult = ddconfig["ch"], list(ddconfig["ch_mult"])
    nrb, zc, ks = ddconfig["num_res_blocks"], ddconfig["z_channels"], ddconfig.get("kernel_size", 3)
    down_layers = [i + 1 for i in ddconfig.get("down_layers", [])]
    attn_layers = ddconfig.get("attn_layers", [])
    nl = len(ch_mult)
    sd = {}
    _conv(g, sd, "post_quant_conv", zc, embed_dim, 1)
    block_in = ch * ch_mult[nl - 1]
    _conv(g, sd, "decoder.conv_in", block_in, zc, ks)

    def res(name, cin, cout):
        _norm(g, sd, name + ".norm1", cin)
        _conv(g, sd, name + ".conv1", cout, cin, 3)
        _norm(g, sd, name + ".norm2", cout)
        _conv(g, sd, name + ".conv2", cout, cout, 3)
        if cin != cout:
            _conv(g, sd, name + ".nin_shortcut", cout, cin, 1)

    def attn(name, c):
        _norm(g, sd, name + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            _conv(g, sd, name + "." + n, c, c, 1)

    res("decoder.mid.block_1", block_in, block_in)
    attn("decoder.mid.attn_1", block_in)
    res("decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(nl)):
        block_out = ch * ch_mult[lvl]
        for ib in range(nrb + 1):
            res(f"decoder.up.{lvl}.block.{ib}", block_in, block_out)
            block_in = block_out
            if lvl in attn_layers:
                attn(f"decoder.up.{lvl}.attn.{ib}", block_in)
        if lvl in down_layers:
            _conv(g, sd, f"decoder.up.{lvl}.upsample.conv", block_in, block_in, 3)
    _norm(g, sd, "decoder.norm_out", block_in)
    _conv(g, sd, "decoder.conv_out", ddconfig["out_ch"], block_in, ks)
    return sd


def vae_encoder_state_dict(ddconfig, embed_dim, seed=7):
    """Keys of AutoencoderKL.{encoder, quant_conv} (ldm/models/autoencoder1d.py:18-53, 319-413)."""
    g = _gen(seed)
    ch, ch_mult = ddconfig["ch"], list(ddconfig["ch_mult"])
    nrb, zc, ks = ddconfig["num_res_blocks"], ddconfig["z_channels"], ddconfig.get("kernel_size", 3)
    down_layers = list(ddconfig.get("down_layers", []))
    attn_layers = ddconfig.get("attn_layers", [])
    sd = {}
    _conv(g, sd, "encoder.conv_in", ch, ddconfig["in_channels"], ks)

    def res(name, cin, cout):
        _norm(g, sd, name + ".norm1", cin)
        _conv(g, sd, name + ".conv1", cout, cin, ks)     # the encoder's ResNet blocks take kernel_size (:345-349)
        _norm(g, sd, name + ".norm2", cout)
        _conv(g, sd, name + ".conv2", cout, cout, ks)
        if cin != cout:
            _conv(g, sd, name + ".nin_shortcut", cout, cin, 1)

    def attn(name, c):
        _norm(g, sd, name + ".norm", c)
        for n in ("q", "k", "v", "proj_out"):
            _conv(g, sd, name + "." + n, c, c, 1)

    block_in = ch
    for lvl in range(len(ch_mult)):
        block_out = ch * ch_mult[lvl]
        for ib in range(nrb):
            res(f"encoder.down.{lvl}.block.{ib}", block_in, block_out)
            block_in = block_out
            if lvl in attn_layers:
                attn(f"encoder.down.{lvl}.attn.{ib}", block_in)
        if lvl in down_layers:
            _conv(g, sd, f"encoder.down.{lvl}.downsample.conv", block_in, block_in, 3)
    res("encoder.mid.block_1", block_in, block_in)
    attn("encoder.mid.attn_1", block_in)
    res("encoder.mid.block_2", block_in, block_in)
    _norm(g, sd, "encoder.norm_out", block_in)
    _conv(g, sd, "encoder.conv_out", 2 * zc if ddconfig.get("double_z", True) else zc, block_in, ks)
    _conv(g, sd, "quant_conv", 2 * embed_dim, 2 * zc, 1)
    return sd


# Benchmark default for the BigVGAN hyper-parameters (not in the reference repo).
BIGVGAN_LARGE_256X = dict(
    resblock="1", num_mels=80, upsample_rates=[4, 4, 2, 2, 2, 2], upsample_kernel_sizes=[8, 8, 4, 4, 4, 4],
    upsample_initial_channel=1536, resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], activation="snakebeta", snake_logscale=True,
    sampling_rate=16000, hop_size=256)

BIGVGAN_BASE_256X = dict(
    resblock="1", num_mels=80, upsample_rates=[8, 8, 2, 2], upsample_kernel_sizes=[16, 16, 4, 4],
    upsample_initial_channel=512, resblock_kernel_sizes=[3, 7, 11],
    resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], activation="snakebeta", snake_logscale=True,
    sampling_rate=16000, hop_size=256)


def bigvgan_state_dict(h, seed=2, snake_std=0.3):
    """Keys of BigVGAN after remove_weight_norm() (vocoder/bigvgan/models.py:135-215): plain .weight/.bias."""
    g = _gen(seed)
    sd = {}
    C0 = h["upsample_initial_channel"]
    _conv(g, sd, "conv_pre", C0, h["num_mels"], 7)
    nk = len(h["resblock_kernel_sizes"])
    ch = C0
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        cin, cout = C0 // (2 ** i), C0 // (2 ** (i + 1))
        a = math.sqrt(1.0 / (cin * k / u))
        sd[f"ups.{i}.0.weight"] = (torch.rand(cin, cout, k, generator=g) * 2 - 1) * a
        sd[f"ups.{i}.0.bias"] = (torch.rand(cout, generator=g) * 2 - 1) * a
        ch = cout
        for j, (rk, dils) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            p = f"resblocks.{i * nk + j}."
            if h["resblock"] == "1":
                for l in range(len(dils)):
                    _conv(g, sd, p + f"convs1.{l}", ch, ch, rk)
                    _conv(g, sd, p + f"convs2.{l}", ch, ch, rk)
                nact = 2 * len(dils)
            else:
                for l in range(len(dils)):
                    _conv(g, sd, p + f"convs.{l}", ch, ch, rk)
                nact = len(dils)
            for m in range(nact):
                sd[p + f"activations.{m}.act.alpha"] = _vec(g, ch, snake_std)
                sd[p + f"activations.{m}.act.beta"] = _vec(g, ch, snake_std)
    sd["activation_post.act.alpha"] = _vec(g, ch, snake_std)
    sd["activation_post.act.beta"] = _vec(g, ch, snake_std)
    _conv(g, sd, "conv_post", 1, ch, 7)
    return sd


def synthetic_inputs(*, prompts, latent_ch, T, L, Cd, rank=0):
    """context ~ N(0,1) seed 1234+rank, uncond seed 4321, x0 seed 2024+rank."""
    c = torch.randn(prompts, L, Cd, generator=_gen(1234 + rank))
    uc = torch.randn(1, L, Cd, generator=_gen(4321)).expand(prompts, L, Cd).contiguous()
    x0 = torch.randn(prompts, latent_ch, T, generator=_gen(2024 + rank))
    return c, uc, x0
