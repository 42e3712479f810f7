"""TEST INFRASTRUCTURE: the small seeded cases shared by tools/make_golden.py (reference -> fixtures),
tests/test_oracle_golden.py (oracle vs fixtures) and the GPU parity tests (CUDA path vs oracle)."""
import torch

from . import weights as W


def gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


DIT_TINY = dict(in_channels=20, context_dim=32, hidden_size=64, num_heads=4, depth=2)
DIT_SMALL = dict(in_channels=20, context_dim=64, hidden_size=192, num_heads=8, depth=3)  # head_dim 24 like config 1
VAE_TINY = dict(double_z=True, in_channels=80, out_ch=80, z_channels=20, kernel_size=5, ch=64, ch_mult=[1, 2, 4],
                num_res_blocks=2, attn_layers=[3], down_layers=[0], dropout=0.0)
BIGVGAN_TINY = dict(W.BIGVGAN_LARGE_256X, upsample_initial_channel=96)
BIGVGAN_SMALL = dict(W.BIGVGAN_LARGE_256X, upsample_initial_channel=384, upsample_rates=[4, 2, 2],
                     upsample_kernel_sizes=[8, 4, 4], hop_size=16)


def dit_inputs(cfg, N=2, T=24, L=10, seed=11):
    g = gen(seed)
    x = torch.randn(N, cfg["in_channels"], T, generator=g)
    ctx = torch.randn(N, L, cfg["context_dim"], generator=g)
    return x, ctx


def cfm_inputs(cfg, B=2, T=24, L=10, seed=12):
    g = gen(seed)
    x0 = torch.randn(B, cfg["in_channels"], T, generator=g)
    c = torch.randn(B, L, cfg["context_dim"], generator=g)
    uc = torch.randn(1, L, cfg["context_dim"], generator=g).expand(B, L, -1).contiguous()
    return x0, c, uc


def latent_inputs(B=2, T=24, seed=13):
    return torch.randn(B, 20, T, generator=gen(seed))


def mel_inputs(B=2, T=12, seed=14):
    return torch.randn(B, 80, T, generator=gen(seed)) * 1.5 - 2.0


MEL_HP = dict(fft_size=1024, audio_num_mel_bins=80, audio_sample_rate=16000, hop_size=256, win_size=1024, fmin=0, fmax=8000)


def wave_inputs(B=2, n=4096, seed=16):
    g = gen(seed)
    t = torch.arange(n) / 16000.0
    tone = 0.4 * torch.sin(2 * 3.14159265 * 440.0 * t)[None] + 0.2 * torch.sin(2 * 3.14159265 * 3100.0 * t)[None]
    return (tone + 0.3 * torch.randn(B, n, generator=g)) * torch.tensor([[1.0], [2.5]])[:B]   # second row clips at +-1


def act_inputs(B=2, C=24, T=50, seed=15):
    g = gen(seed)
    return torch.randn(B, C, T, generator=g), torch.randn(C, generator=g) * 0.3, torch.randn(C, generator=g) * 0.3
