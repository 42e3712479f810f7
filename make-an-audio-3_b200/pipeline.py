"""Text/video-to-audio sampling pipeline: CFM sampling -> decode_first_stage -> vocode, the hot path of
scripts/txt2audio_for_2cap_flow.py:151-190 (GenSamples.gen_test_sample) with the conditioner replaced by caller-
supplied embeddings.  Everything stays on the device between the three stages (the reference round-trips the mel
through numpy, txt2audio_for_2cap_flow.py:181-188).

Multi-GPU: prompts are independent, so `shard_prompts` gives rank r the prompts r::world and `gather_waveforms` is the
only collective (one NCCL all-gather of the finished waveforms).
"""
import contextlib
import os

import torch

from . import lib as L
from . import sampler as _sampler
from .dit import TxtFlagLargeImprovedDiTV2, VideoFlagLargeDiT
from .sampler import CFMSampler
from .vae import AutoencoderKL
from .vocoder import VocoderBigVGAN

# the reference's shipped model configs (configs/*.yaml: unet_config.params), keyed by the names BASELINE.json uses
MODEL_CONFIGS = {
    "M": dict(in_channels=20, context_dim=1024, hidden_size=768, num_heads=32, depth=16, max_len=1000),
    "XL": dict(in_channels=20, context_dim=1024, hidden_size=1152, num_heads=16, depth=28, max_len=1000),
    "XXL": dict(in_channels=20, context_dim=1024, hidden_size=1536, num_heads=32, depth=32, max_len=1000),
    "MOE": dict(in_channels=20, context_dim=768, hidden_size=768, num_heads=32, depth=16, max_len=1000, num_experts=4),
}
# first_stage_config.params.ddconfig of every shipped config (configs/txt2audio-cfm-cfg.yaml:50-66)
VAE_DDCONFIG = dict(double_z=True, in_channels=80, out_ch=80, z_channels=20, kernel_size=5, ch=384, ch_mult=[1, 2, 4],
                    num_res_blocks=2, attn_layers=[3], down_layers=[0], dropout=0.0)


_NVTX = os.environ.get("MA3_NVTX", "0") == "1"


@contextlib.contextmanager
def stage_range(name):
    """NVTX range around one stage of the path (the reference has no tracing; MA3_NVTX=1 marks
    sample_cfg / decode_first_stage / vocode so that a timeline shows the stages by name).  Off by default: a range
    push / pop is a host call per stage, harmless, but the graph replays need none."""
    if not _NVTX:
        yield
        return
    torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


class Txt2AudioPipeline:
    """model.sample_cfg(...) -> model.decode_first_stage(...) -> vocoder.vocode(...) on one GPU."""

    def __init__(self, dit, vae, vocoder, scale_factor=1.0, mel_dim=20, mel_length=256, use_graph=True):
        self.dit, self.first_stage_model, self.vocoder = dit, vae, vocoder
        self.scale_factor = float(scale_factor)
        self.mel_dim, self.mel_length, self.channels = mel_dim, mel_length, 0
        self.sampler = CFMSampler(self, use_graph=use_graph)
        self.use_graph = use_graph
        self.cond_stage_model = None      # optional text conditioner (conditioners.py), ddpm_audio.py:343-356
        self._tail = {}

    # CFMSampler looks the DiT up as model.model.diffusion_model (the reference's nesting, ddpm.py:1402)
    @property
    def model(self):
        return self

    @property
    def diffusion_model(self):
        return self.dit

    def sample(self, cond, batch_size=16, timesteps=None, shape=None, x_latent=None, t_start=None, **kw):
        return self.sampler.sample(cond, batch_size, timesteps, shape, x_latent, t_start)

    def sample_cfg(self, cond, unconditional_guidance_scale, unconditional_conditioning, batch_size=16, timesteps=None,
                   shape=None, x_latent=None, t_start=None, **kw):
        return self.sampler.sample_cfg(cond, unconditional_guidance_scale, unconditional_conditioning, batch_size,
                                       timesteps, shape, x_latent, t_start)

    @property
    def device(self):
        return self.dit.proj_in.weight.device

    def get_learned_conditioning(self, c):
        """ddpm_audio.py:343-356: the conditioner's `encode` when one is attached (`cond_stage_model`, e.g.
        `conditioners.FrozenCLAPFLANEmbedder`: the reference's `cond_stage_config`), otherwise the identity for
        precomputed embeddings [B, L, context_dim]."""
        m = getattr(self, "cond_stage_model", None)
        if torch.is_tensor(c):
            return c.to(self.device, torch.float32)
        if m is None:
            raise TypeError("Txt2AudioPipeline has no cond_stage_model: pass precomputed embeddings [B, L, context_dim] or "
                            "attach a conditioner (ma3_b200.conditioners.FrozenCLAPFLANEmbedder)")
        c = m.encode(c) if hasattr(m, "encode") and callable(m.encode) else m(c)
        return c.to(self.device, torch.float32)

    @torch.no_grad()
    def decode_first_stage(self, z):
        """ddpm_audio.py:358-371: z / scale_factor -> first_stage_model.decode."""
        return self.first_stage_model.decode((1.0 / self.scale_factor) * z)

    @torch.no_grad()
    def encode_first_stage(self, x):
        """ddpm_audio.py:372-374: mel [B, 80, 2T] -> posterior over the latent (VAE Encoder1D on the GPU)."""
        return self.first_stage_model.encode(x)

    def get_first_stage_encoding(self, encoder_posterior, sample=True):
        """ddpm.py get_first_stage_encoding: scale_factor * (posterior sample | tensor)."""
        z = encoder_posterior
        if hasattr(encoder_posterior, "sample"):
            z = encoder_posterior.sample() if sample else encoder_posterior.mode()
        return self.scale_factor * z

    @torch.no_grad()
    def generate_windows(self, conds, unconds, scale=3.0, timesteps=25, length=None, x0=None):
        """Long-form generation over independent windows (scripts/video2audio_flow.py:483-523: 40-frame video windows,
        one sample_cfg + decode_first_stage per window, mels concatenated along time, then ONE vocoder pass).  The
        reference runs the windows one after the other at batch 1; they are independent, so here they are one batch.
        conds / unconds: [W, L, Cd] (window-major); length: latent frames per window (default mel_length; callers that
        stretch a window overwrite dit.freqs_cis for NTK scaling first, as the reference script does)."""
        Wn = conds.shape[0]
        T = length or self.mel_length
        if x0 is None:
            x0 = torch.randn(Wn, self.mel_dim, T, device=self.device)
        z, _ = self.sample_cfg(conds, scale, unconds, Wn, timesteps=timesteps, x_latent=x0)
        mel = self.decode_first_stage(z)                                  # [W, 80, 2T]
        mel = mel.permute(1, 0, 2).reshape(1, mel.shape[1], -1)            # np.concatenate(mel_list, 1) of the script
        return self.vocoder.vocode_tensor(mel.contiguous()), mel

    @torch.no_grad()
    def generate(self, cond, uncond, x0, scale=3.0, timesteps=25):
        """cond/uncond [B, L, Cd], x0 [B, 20, T] (device tensors) -> waveforms [B, 2T*hop] on the device."""
        B = x0.shape[0]
        with stage_range("ma3.sample_cfg"):
            z, _ = self.sample_cfg(cond, scale, uncond, B, timesteps=timesteps, x_latent=x0)
        with stage_range("ma3.decode_first_stage+vocode"):
            return self.decode_and_vocode(z)

    @torch.no_grad()
    def decode_and_vocode(self, z):
        """decode_first_stage + vocode of latents z [B, 20, T] -> waveforms [B, 2T*hop].  With use_graph the ~450
        launches of the two stages are captured once per latent shape and replayed (as the sampler does for its step
        loop), so no host launch latency sits between the short kernels of the VAE and of the early vocoder stages."""
        L.auto_pdl(2 * z.shape[0] * z.shape[-1])   # same rule as the sampler (whose batch is CFG-doubled)
        if os.environ.get("MA3_PDL_TAIL") in ("0", "1"):   # experiment: pin it for the decode / vocode tail only
            L.load().ma3_set_pdl(int(os.environ["MA3_PDL_TAIL"]))
        if not self.use_graph:
            with stage_range("ma3.decode_first_stage"):
                mel = self.decode_first_stage(z)
            with stage_range("ma3.vocode"):
                return self.vocoder.vocode_tensor(mel)
        key = (tuple(z.shape), id(self.first_stage_model), id(self.vocoder))
        st = self._tail.get(key)
        if st is None:
            zin = z.clone()
            self.vocoder.vocode_tensor(self.decode_first_stage(zin))   # eager pass: packs weights, allocates buffers
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = L.launch_count()
            with torch.cuda.graph(g):
                wav = self.vocoder.vocode_tensor(self.decode_first_stage(zin))
            st = {"zin": zin, "wav": wav, "graph": g, "launches": L.launch_count() - n0}
            self._tail = {key: st}             # keep one plan: the vocoder buffers are large
        st["zin"].copy_(z)
        st["graph"].replay()
        _sampler.GRAPH_REPLAY_LAUNCHES += st["launches"]
        return st["wav"].clone()


def build_random_pipeline(model="M", vocoder_h=None, seed=0, device="cuda", use_graph=True, state_dicts=None):
    """Random-init pipeline of a shipped config (no checkpoints exist offline).  `state_dicts` = (dit, vae, vocoder)
    state_dicts keyed like the reference's; when omitted the constructors' own random init is used."""
    L.require_device()
    cfg = dict(MODEL_CONFIGS[model])
    torch.manual_seed(seed)
    if "num_experts" in cfg:
        dit = VideoFlagLargeDiT(**cfg)
    else:
        dit = TxtFlagLargeImprovedDiTV2(**cfg)
    vae = AutoencoderKL(embed_dim=20, ddconfig=dict(VAE_DDCONFIG), lossconfig=None)
    voc = VocoderBigVGAN(device=device, h=vocoder_h, state_dict=state_dicts[2] if state_dicts else None)
    for blk in dit.blocks:   # the constructor leaves the cross-attention gate at zero; make it count
        blk.attention.gate.data.normal_(0.0, 0.5)
    if state_dicts:
        dit.load_state_dict(state_dicts[0], strict=True)
        vae.load_state_dict(state_dicts[1], strict=True)
    return Txt2AudioPipeline(dit.to(device), vae.to(device), voc, use_graph=use_graph)


def shard_prompts(n_prompts, rank, world):
    """Indices of the prompts rank `rank` generates (round-robin)."""
    return list(range(rank, n_prompts, world))


def gather_waveforms(wav, group=None):
    """The path's single collective: all-gather [B_local, samples] -> [world * B_local, samples] (rank-major)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return wav
    out = [torch.empty_like(wav) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, wav.contiguous(), group=group)
    return torch.cat(out)
