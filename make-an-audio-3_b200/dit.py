"""Drop-in Next-DiT backbones for the reference's `unet_config.target`.

`TxtFlagLargeImprovedDiTV2` / `TxtFlagLargeDiT` mirror ldm/modules/diffusionmodules/flag_large_dit.py:128-299 and
`VideoFlagLargeDiT` mirrors ldm/modules/diffusionmodules/flag_large_dit_moe.py:613-740: same constructor arguments,
`forward(x, t, context)` signature, `freqs_cis` attribute / `precompute_freqs_cis` static method and `state_dict`
key names.  The nn.Module tree below only *holds* the fp32 parameters (so reference checkpoints load with
load_state_dict); all arithmetic runs in the sm_100a kernels of libma3b200.so on packed bf16 copies of the weights.

Step-invariant work is hoisted: the normalised context and every layer's cross-attention
K/V are computed once per call to `prepare_context`, the adaLN modulation of every layer for a whole list of
timesteps in one GEMM by `prepare_timesteps`; `run_blocks` then costs 7 launches per block.
"""
import math
import os

import torch
import torch.nn as nn

from . import lib as L
from . import ops


def _ffn_hidden(dim, multiple_of=256, ffn_dim_multiplier=None):
    hidden = int(2 * (4 * dim) / 3)
    if ffn_dim_multiplier is not None:
        hidden = int(ffn_dim_multiplier * hidden)
    return multiple_of * ((hidden + multiple_of - 1) // multiple_of)


class _Weight(nn.Module):
    """Parameter holder with a `.weight` (RMSNorm scale)."""

    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))


class _Lin(nn.Module):
    """Parameter holder shaped like nn.Linear (never called)."""

    def __init__(self, fin, fout, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(fout, fin).uniform_(-1, 1) * math.sqrt(3.0 / fin))
        if bias:
            self.bias = nn.Parameter(torch.zeros(fout))
        else:
            self.register_parameter("bias", None)


class _LN(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))


class _Seq(nn.Module):
    """Holder reproducing nn.Sequential's integer child names at chosen indices."""

    def __init__(self, **children):
        super().__init__()
        for k, v in children.items():
            self.add_module(k.lstrip("_"), v)


class _Attention(nn.Module):
    def __init__(self, dim, n_heads, y_dim, qk_norm=False):
        super().__init__()
        if qk_norm:   # nn.LayerNorm over the full model dim (flag_large_dit_moe.py:199-207)
            self.q_norm, self.k_norm, self.ky_norm = _LN(dim), _LN(dim), _LN(dim)
        self.wq, self.wk, self.wv = _Lin(dim, dim, False), _Lin(dim, dim, False), _Lin(dim, dim, False)
        self.wk_y, self.wv_y = _Lin(y_dim, dim, False), _Lin(y_dim, dim, False)
        self.gate = nn.Parameter(torch.zeros(n_heads))
        self.wo = _Lin(dim, dim, False)


class _FFN(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.w1, self.w2, self.w3 = _Lin(dim, hidden, False), _Lin(hidden, dim, False), _Lin(dim, hidden, False)


class _MoE(nn.Module):
    def __init__(self, dim, hidden, num_experts):
        super().__init__()
        self.time_experts = nn.ModuleDict({str(i): _FFN(dim, hidden) for i in range(num_experts)})
        self.freq_experts = nn.ModuleDict({str(i): _FFN(dim, hidden) for i in range(num_experts)})


class _Block(nn.Module):
    def __init__(self, dim, n_heads, y_dim, hidden, num_experts, qk_norm=False):
        super().__init__()
        self.attention = _Attention(dim, n_heads, y_dim, qk_norm)
        self.feed_forward = _MoE(dim, hidden, num_experts) if num_experts else _FFN(dim, hidden)
        self.attention_norm, self.ffn_norm = _Weight(dim), _Weight(dim)
        self.adaLN_modulation = _Seq(_1=_Lin(dim, 6 * dim))
        self.attention_y_norm = _Weight(y_dim)


class _Final(nn.Module):
    def __init__(self, dim, out_channels):
        super().__init__()
        self.linear = _Lin(dim, out_channels)
        self.adaLN_modulation = _Seq(_1=_Lin(dim, 2 * dim))


_ROWNORM_MIN_CTAS = 100
_EPI_NORM_MAX_ROWS = int(os.environ.get("MA3_EPI_NORM_ROWS", "1024"))


class _Work:
    """Device buffers for one (N, T, L) problem size."""
    pass


class TxtFlagLargeDiT(nn.Module):
    """B200 drop-in for flag_large_dit.py:128-251 (same ctor / forward / freqs_cis surface)."""

    _video = False

    def __init__(self, in_channels, context_dim, hidden_size=1152, depth=28, num_heads=16, max_len=1000,
                 n_kv_heads=None, multiple_of: int = 256, ffn_dim_multiplier=None, norm_eps=1e-5, qk_norm=None,
                 rope_scaling_factor: float = 1.0, ntk_factor: float = 1.0, num_experts=0):
        super().__init__()
        self.qk_norm = bool(qk_norm)
        if n_kv_heads not in (None, num_heads):
            raise NotImplementedError("grouped KV heads are not used by any shipped config")
        if hidden_size % num_heads or (hidden_size // num_heads) % 8:
            raise ValueError("head_dim must be a multiple of 8")
        self.in_channels = self.out_channels = in_channels
        self.context_dim, self.hidden_size, self.depth, self.num_heads = context_dim, hidden_size, depth, num_heads
        self.head_dim = hidden_size // num_heads
        self.norm_eps = norm_eps
        self.num_experts = num_experts if self._video else 0
        self.ffn_hidden = _ffn_hidden(hidden_size, multiple_of, ffn_dim_multiplier)
        D = hidden_size
        y_dim = D if self._video else context_dim
        self.y_dim = y_dim
        self.t_embedder = nn.Module()
        self.t_embedder.mlp = _Seq(_0=_Lin(256, D), _2=_Lin(D, D))
        if self._video:
            self.c_embedder = nn.Module()
            self.c_embedder.mlp = _Seq(_0=_Lin(context_dim, D), _2=_Lin(D, D), _3=_LN(D))
        self.proj_in = _Lin(in_channels, D)
        self.cap_embedder = _Seq(_0=_LN(y_dim), _1=_Lin(y_dim, D))
        self.blocks = nn.ModuleList([_Block(D, num_heads, y_dim, self.ffn_hidden, self.num_experts, self.qk_norm)
                                     for _ in range(depth)])
        self.final_layer = _Final(D, in_channels)
        self.rope_scaling_factor, self.ntk_factor = rope_scaling_factor, ntk_factor
        self.freqs_cis = self.precompute_freqs_cis(self.head_dim, max_len, rope_scaling_factor=rope_scaling_factor,
                                                   ntk_factor=ntk_factor)
        self._packed = None
        self._rope_src = None
        self._work = {}
        self._ctx = None
        # bumped whenever a device buffer a captured CUDA graph may point at is re-allocated (packed weights, RoPE
        # table, context buffers, workspaces): CFMSampler keys its plans on it
        self.generation = 0
        self.register_load_state_dict_post_hook(lambda m, k: m.invalidate())

    # ---------------------------------------------------------------- reference surface
    @staticmethod
    def precompute_freqs_cis(dim: int, end: int, theta: float = 10000.0, rope_scaling_factor: float = 1.0,
                             ntk_factor: float = 1.0):
        """complex64 [end, dim/2] = exp(i * pos/rope_scaling * (theta*ntk)^(-2j/dim)) (flag_large_dit.py:212-251)."""
        theta = theta * ntk_factor
        inv = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        pos = torch.arange(end, dtype=torch.float32) / rope_scaling_factor
        ang = torch.outer(pos, inv).float()
        return torch.polar(torch.ones_like(ang), ang)

    def invalidate(self):
        self._packed = None
        self._ctx = None
        self._work = {}
        self._rope_src = None
        self.generation += 1

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    @torch.no_grad()
    def forward(self, x, t, context):
        """x [N, C, T] fp32, t [N] int64, context [N, L, Cd] fp32 -> [N, C, T] fp32."""
        x = x.contiguous().float()
        N, _, T = x.shape
        self.prepare_context(context)
        cond = self.prepare_timesteps(t.reshape(-1).to(torch.int64), per_sample=True)
        w = self.run_blocks(x, cond, 0, t_ints=[int(v) for v in t.reshape(-1).tolist()] if self.num_experts else None)
        p = self._packed
        v = torch.empty(N, self.out_channels, T, device=x.device, dtype=torch.float32)
        ops.final_layer(w.h, cond["mod"][0], p["final_off"], p["final_off"] + self.hidden_size, p["final_w"],
                        p["final_b"], N, T, v)
        return v

    # ---------------------------------------------------------------- packing
    def _pack(self):
        dev = self.proj_in.weight.device
        if dev.type != "cuda":
            raise L.Ma3Error("ma3_b200 modules run on CUDA only (no CPU fallback): call .cuda() first")
        L.require_device()
        D, H, F = self.hidden_size, self.num_heads, self.ffn_hidden
        bf = torch.bfloat16
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        b16 = lambda t: t.detach().to(device=dev, dtype=bf).contiguous()
        # the step-invariant conditioning path runs as hi/lo split bf16 GEMMs (~fp32 accuracy, ops.gemm_split)
        sp = lambda t: ops.split_weight(t.to(dev))
        il = lambda w1, w3: torch.stack([w1.detach(), w3.detach()], 1).reshape(2 * w1.shape[0], w1.shape[1])
        p = {}
        p["proj_w"], p["proj_b"] = f32(self.proj_in.weight.detach().t()), f32(self.proj_in.bias)  # [C, D]
        m = self.t_embedder.mlp
        p["t_w1"], p["t_b1"] = sp(getattr(m, "0").weight), f32(getattr(m, "0").bias)
        p["t_w2"], p["t_b2"] = sp(getattr(m, "2").weight), f32(getattr(m, "2").bias)
        ce = self.cap_embedder
        p["cap_ln_w"], p["cap_ln_b"] = f32(getattr(ce, "0").weight), f32(getattr(ce, "0").bias)
        p["cap_w"], p["cap_b"] = sp(getattr(ce, "1").weight), f32(getattr(ce, "1").bias)
        if self._video:
            cm = self.c_embedder.mlp
            p["c_w1"], p["c_b1"] = b16(getattr(cm, "0").weight), f32(getattr(cm, "0").bias)
            p["c_w2"], p["c_b2"] = b16(getattr(cm, "2").weight), f32(getattr(cm, "2").bias)
            p["c_ln_w"], p["c_ln_b"] = f32(getattr(cm, "3").weight), f32(getattr(cm, "3").bias)
        ada_w, ada_b, blocks = [], [], []
        # Dense blocks whose width is a multiple of 384 run wo / w2 as row-owning cluster GEMMs that also emit the
        # normalised, modulated operand of the next projection (ma3_gemm_rownorm): 5 launches per block, no stand-alone
        # RMSNorm pass.  MA3_ROWNORM=0 restores the 7-launch form.
        fused = not self.num_experts and D % 384 == 0 and D // 384 <= 4 and os.environ.get("MA3_ROWNORM", "1") != "0"
        # Batches too small to fill the row-owning GEMM's single wave fold the norm into the epilogues of the tiled GEMMs
        # instead (ma3_gemm_t.norm_out / row_ss: the wo / w2 epilogue emits h_new, g = 16-bit(h_new * wn_s) and partial
        # sums of squares; the QKV / w1|w3 GEMM scales its accumulator rows by rstd and adds shift_s W^T): also 5 launches
        # per block, which is what counts when every launch is latency-bound.  MA3_EPI_NORM=0 switches it off.
        epi = fused and not self.qk_norm and D % 32 == 0 and os.environ.get("MA3_EPI_NORM", "1") != "0"
        if fused:
            p["norm_w"] = torch.stack([torch.stack([f32(b_.attention_norm.weight), f32(b_.ffn_norm.weight)])
                                       for b_ in self.blocks]).contiguous()          # [depth, 2, D]
        if epi:
            # one contiguous tensor per weight kind, so that the per-block shift_s W^T tables are ONE batched GEMM
            p["wqkv_all"] = torch.empty(self.depth, 3 * D, D, device=dev, dtype=bf)
            p["w13_all"] = torch.empty(self.depth, 2 * F, D, device=dev, dtype=bf)
        for bi, blk in enumerate(self.blocks):
            a = blk.attention
            q = {}
            if epi:
                p["wqkv_all"][bi].copy_(torch.cat([a.wq.weight, a.wk.weight, a.wv.weight]).detach())
                q["wqkv"] = p["wqkv_all"][bi]
            else:
                q["wqkv"] = b16(torch.cat([a.wq.weight, a.wk.weight, a.wv.weight]))
            yw = blk.attention_y_norm.weight.detach()[None, :]  # RMSNorm_y scale folded into the projections
            q["wkv_y"] = b16(torch.cat([a.wk_y.weight.detach() * yw, a.wv_y.weight.detach() * yw]))
            q["wo"] = b16(a.wo.weight)
            q["gate"] = f32(a.gate)
            if self.qk_norm:
                q["qn"] = (f32(a.q_norm.weight), f32(a.q_norm.bias))
                q["kn"] = (f32(a.k_norm.weight), f32(a.k_norm.bias))
                q["kyn"] = (f32(a.ky_norm.weight), f32(a.ky_norm.bias))
            q["attn_norm"], q["ffn_norm"] = f32(blk.attention_norm.weight), f32(blk.ffn_norm.weight)
            ff = blk.feed_forward
            if self.num_experts:
                band = D // self.num_experts
                q["t_w13"] = [b16(il(e.w1.weight, e.w3.weight)) for e in ff.time_experts.values()]
                q["t_w2"] = [b16(e.w2.weight) for e in ff.time_experts.values()]
                # frequency expert j only sees / produces band j: slice the weights once (exact), stacked
                # [E][...] so that the E band GEMMs of a block are ONE batched launch per projection
                q["f_w13"] = torch.stack([b16(il(e.w1.weight[:, j * band:(j + 1) * band], e.w3.weight[:, j * band:(j + 1) * band]))
                                          for j, e in enumerate(ff.freq_experts.values())]).contiguous()
                q["f_w2"] = torch.stack([b16(e.w2.weight[j * band:(j + 1) * band, :])
                                         for j, e in enumerate(ff.freq_experts.values())]).contiguous()
            else:
                if epi:
                    p["w13_all"][bi].copy_(il(ff.w1.weight, ff.w3.weight))
                    q["w13"] = p["w13_all"][bi]
                else:
                    q["w13"] = b16(il(ff.w1.weight, ff.w3.weight))
                q["w2"] = b16(ff.w2.weight)
            blocks.append(q)
            ada_w.append(getattr(blk.adaLN_modulation, "1").weight.detach())
            ada_b.append(getattr(blk.adaLN_modulation, "1").bias.detach())
        fl = self.final_layer
        ada_w.append(getattr(fl.adaLN_modulation, "1").weight.detach())
        ada_b.append(getattr(fl.adaLN_modulation, "1").bias.detach())
        p["ada_w"], p["ada_b"] = sp(torch.cat(ada_w)), f32(torch.cat(ada_b))
        p["final_off"] = 6 * D * self.depth
        p["mod_cols"] = 6 * D * self.depth + 2 * D
        # row pitch of the modulation buffer: dense models append wn_s = w * (1 + scale_s) of both norms of every block
        p["fused"] = fused
        p["epi"] = epi
        p["wn_off"] = p["mod_cols"]
        p["mod_ld"] = p["mod_cols"] + (2 * D * self.depth if fused else 0)
        p["final_w"], p["final_b"] = f32(fl.linear.weight), f32(fl.linear.bias)
        p["blocks"] = blocks
        self._packed = p
        self._rope_src = None
        self._ctx = None

    def _ensure(self):
        if self._packed is None:
            self._pack()
        fc = self.freqs_cis
        if self._rope_src is not fc:  # callers overwrite the table on the live module (NTK scaling)
            dev = self.proj_in.weight.device
            tab = torch.view_as_real(fc.to("cpu")).float().contiguous().to(dev)
            old = self._packed.get("rope")
            if old is not None and old.shape == tab.shape:
                old.copy_(tab)             # same storage: captured graphs keep pointing at a live, updated table
            else:
                self._packed["rope"] = tab
                self.generation += 1
            self._rope_src = fc
        return self._packed

    def _workspace(self, N, T):
        key = (N, T)
        w = self._work.get(key)
        if w is None:
            dev = self.proj_in.weight.device
            D, H, F = self.hidden_size, self.num_heads, self.ffn_hidden
            hd = self.head_dim
            hdp = 64 if hd <= 64 else 128
            if hd > 128:
                raise L.Ma3Error("head_dim > 128 not supported")
            Tp = (T + 7) // 8 * 8
            bf = torch.bfloat16
            w = _Work()
            w.hdp, w.Tp = hdp, Tp
            w.h = torch.empty(N * T, D, device=dev, dtype=torch.float32)
            w.u = torch.empty(N * T, D, device=dev, dtype=bf)
            w.q = torch.zeros(N, H, T, hdp, device=dev, dtype=bf)
            w.k = torch.zeros(N, H, T, hdp, device=dev, dtype=bf)
            w.vt = ops.alloc_vt(N, H, hd=hd, hdp=hdp, tokens_pad=Tp, device=dev, dtype=bf)
            w.att = torch.empty(N * T, D, device=dev, dtype=bf)
            w.mid = torch.empty(N * T, F, device=dev, dtype=bf)
            if self._ensure()["epi"]:
                w.g = torch.empty(N * T, D, device=dev, dtype=bf)                       # h * wn_s (epilogue-fused RMSNorm operand)
                # per-chunk sums of squares of h (row pitch padded to 4 floats; the pad columns stay zero)
                w.ss = torch.zeros(N * T, (D // 32 + 3) // 4 * 4, device=dev, dtype=torch.float32)
            if self.qk_norm:
                w.qkv_raw = torch.empty(N * T, 3 * D, device=dev, dtype=torch.float32)
            if self.num_experts:
                w.y1 = torch.empty(N * T, D, device=dev, dtype=bf)
                w.mid_e = torch.empty(self.num_experts, N * T, F, device=dev, dtype=bf)   # one SwiGLU buffer per band expert
            self._work[key] = w
            self.generation += 1
        return w

    # ---------------------------------------------------------------- step-invariant work
    @torch.no_grad()
    def prepare_context(self, context):
        """RMSNorm_y(y), every layer's cross K/V and the caption embedding -- computed once per conditioning
        (the reference recomputes them in every block of every step, flag_large_dit_moe.py:390-392)."""
        p = self._ensure()
        dev = self.proj_in.weight.device
        context = context.to(device=dev, dtype=torch.float32).contiguous()
        N, Lc, Cd = context.shape
        D, H, hd = self.hidden_size, self.num_heads, self.head_dim
        hdp = 64 if hd <= 64 else 128
        Lp = (Lc + 7) // 8 * 8
        bf = torch.bfloat16
        c = self._ctx
        if c is None or c["shape"] != (N, Lc, Cd):
            c = {"shape": (N, Lc, Cd)}
            c["yn"] = torch.empty(N * Lc, self.y_dim, device=dev, dtype=bf)
            c["ky"] = torch.zeros(self.depth, N, H, Lc, hdp, device=dev, dtype=bf)
            c["vyt"] = ops.alloc_vt(self.depth, N, H, hd=hd, hdp=hdp, tokens_pad=Lp, device=dev, dtype=bf)
            c["pool"] = torch.empty(N, self.y_dim, device=dev, dtype=torch.float32)
            c["cap"] = torch.empty(N, D, device=dev, dtype=torch.float32)
            if self._video:
                c["c1"] = torch.empty(N * Lc, D, device=dev, dtype=bf)
                c["c2"] = torch.empty(N * Lc, D, device=dev, dtype=torch.float32)
                c["y"] = torch.empty(N * Lc, D, device=dev, dtype=torch.float32)
                c["ctx16"] = torch.empty(N * Lc, Cd, device=dev, dtype=bf)
            self._ctx = c
            self.generation += 1
        if self._video:
            # c = LayerNorm(W2 gelu(W1 ctx + b1) + b2)   (flag_large_dit_moe.py:151-162, 680)
            ops.cast(context.view(N * Lc, Cd), c["ctx16"])
            ops.gemm(c["ctx16"], p["c_w1"], M=N * Lc, N=D, K=Cd, out=c["c1"], bias=p["c_b1"], act=2)
            ops.gemm(c["c1"], p["c_w2"], M=N * Lc, N=D, K=D, out=c["c2"], bias=p["c_b2"])
            ops.layernorm_rows(c["c2"], p["c_ln_w"], p["c_ln_b"], c["y"])
            y = c["y"].view(N, Lc, D)
        else:
            y = context
        ops.pool_layernorm(y, p["cap_ln_w"], p["cap_ln_b"], c["pool"])
        ops.gemm_split(c["pool"], p["cap_w"], M=N, N=D, K=self.y_dim, out=c["cap"], bias=p["cap_b"])
        ops.rmsnorm_modulate(y.view(N * Lc, self.y_dim), None, c["yn"], eps=self.norm_eps)
        for i, q in enumerate(p["blocks"]):
            if self.qk_norm:   # ky_norm: LayerNorm over the full dim of the cross K before the head split
                raw = torch.empty(N * Lc, 2 * D, device=dev, dtype=torch.float32)
                ops.gemm(c["yn"], q["wkv_y"], M=N * Lc, N=2 * D, K=self.y_dim, out=raw)
                ops.qknorm_rope(raw, first_section=1, qn=None, kn=q["kyn"], rope=None, q_out=None, k_out=c["ky"][i],
                                vt_out=c["vyt"][i], tokens=Lc, tokens_pad=Lp, D=D, hd=hd, hdp=hdp)
                continue
            ops.gemm(c["yn"], q["wkv_y"], M=N * Lc, N=2 * D, K=self.y_dim, epi=L.EPI_QKV_ROPE, q_out=c["ky"][i],
                     k_out=c["ky"][i], vt_out=c["vyt"][i], rope=None, model_dim=D, head_dim=hd, head_dim_pad=hdp,
                     tokens=Lc, tokens_pad=Lp, first_section=1)
        return c

    def _norm_mode(self, N, T):
        """How the RMSNorm + modulate between two projections is realised for a batch of N x T token rows: 'row' = the
        row-owning cluster GEMM (its single wave of ceil(M / 128) * D / 384 CTAs must fill most of the GPU: XL x 8 prompts =
        117 of 148 SMs), 'epi' = folded into the epilogues of the tiled GEMMs (small batches), 'plain' = stand-alone pass."""
        p = self._ensure()
        if not p["fused"]:
            return "plain"
        n_cta = (N * T + 127) // 128 * (self.hidden_size // 384)
        if _ROWNORM_MIN_CTAS <= n_cta <= ops.sm_count() or os.environ.get("MA3_ROWNORM") == "force":
            return "row"
        # Epilogue-fused norms pay while every launch is latency-bound (M x 1 prompt, 624 rows: 28.2 -> 27.1 ms per clip); at
        # XXL / T = 936 (1872 rows) the heavier epilogues cost more than the two saved launches (137.1 -> 140.5 ms).
        # A warp's 32 accumulator rows must lie in at most two samples (T >= 32).
        return "epi" if p["epi"] and T >= 32 and N * T <= _EPI_NORM_MAX_ROWS else "plain"

    def cond_buffers(self, S, N, T=None):
        """Device buffers of one (steps, batch) conditioning: `mod` fp32 [S, N, mod_ld] = the adaLN modulation of every
        block and of the final layer; fused models append wn_s = w * (1 + scale_s) of both norms of every block.  When the
        batch (N x T rows, T given) will run the epilogue-fused norm, also its per-block bias tables
        b2q [depth, S*N, 3D] = shift_1 Wqkv^T and b2f [depth, S*N, 2F] = shift_2 W13^T."""
        p = self._ensure()
        dev = self.proj_in.weight.device
        D, F = self.hidden_size, self.ffn_hidden
        c = {"S": S, "N": N, "mod": torch.empty(S, N, p["mod_ld"], device=dev, dtype=torch.float32)}
        if T is not None and self._norm_mode(N, T) == "epi":
            c["b2q"] = torch.empty(self.depth, S * N, 3 * D, device=dev, dtype=torch.float32)
            c["b2f"] = torch.empty(self.depth, S * N, 2 * F, device=dev, dtype=torch.float32)
            c["sh2"] = torch.empty(2 * self.depth, 2 * S * N, D, device=dev, dtype=torch.bfloat16)
        return c

    @torch.no_grad()
    def prepare_timesteps(self, t, per_sample=False, out=None):
        """adaLN modulation of every block (+ final layer) for all requested timesteps in ONE GEMM (hi/lo split bf16:
        ~fp32 accuracy), plus -- fused models -- wn_s = w (1 + scale_s) of every norm behind the modulation columns.
        t: int64 [S] (per_sample=False: every sample shares t[s]) or [N] (per_sample=True: one step, t per sample).
        Returns the cond_buffers dict (filled in place when `out` is given)."""
        p = self._ensure()
        c = self._ctx
        dev = self.proj_in.weight.device
        N = c["shape"][0]
        D, F = self.hidden_size, self.ffn_hidden
        t = t.to(dev)
        Mt = t.numel()
        S = 1 if per_sample else Mt
        if per_sample and Mt != N:
            raise ValueError(f"t has {Mt} entries for a batch of {N}")
        cb = out if out is not None else self.cond_buffers(S, N)
        assert cb["S"] == S and cb["N"] == N
        f32 = torch.float32
        e0 = torch.empty(Mt, 256, device=dev, dtype=f32)
        e1 = torch.empty(Mt, D, device=dev, dtype=f32)
        temb = torch.empty(Mt, D, device=dev, dtype=f32)
        ops.timestep_embed(t, e0)
        ops.gemm_split(e0, p["t_w1"], M=Mt, N=D, K=256, out=e1, bias=p["t_b1"], act=1)
        ops.gemm_split(e1, p["t_w2"], M=Mt, N=D, K=D, out=temb, bias=p["t_b2"])
        a = torch.empty(S * N, D, device=dev, dtype=f32)
        ops.adaln_input(temb, c["cap"], a, S, N, 0 if per_sample else 1, 1 if per_sample else 0)
        mod = cb["mod"]
        R = S * N
        ops.gemm_split(a, p["ada_w"], M=R, N=p["mod_cols"], K=D, out=mod, out_ld=p["mod_ld"], bias=p["ada_b"])
        if p["fused"]:
            m2 = mod.view(R, p["mod_ld"])
            ops.norm_weights(m2, p["norm_w"], self.depth, D, p["wn_off"])
            if "b2q" in cb:
                # shift_s W^T with W the bf16 weights the main GEMMs use: A = (hi, lo) halves of the fp32 shift vectors
                sh = cb["sh2"]
                ops.split_bf16(m2, sh[:self.depth], col0=0, col_step=6 * D, nb=self.depth, cols=D)            # shift_1
                ops.split_bf16(m2, sh[self.depth:], col0=3 * D, col_step=6 * D, nb=self.depth, cols=D)        # shift_2
                taps = ((0, 0), (R, 0))
                ops.gemm(sh[:self.depth], p["wqkv_all"], M=R, N=3 * D, K=D, batch=self.depth, a_rows=2 * R,
                         a_batch_stride=2 * R * D, b_rows=3 * D, b_batch_stride=3 * D * D, taps=taps, out=cb["b2q"],
                         out_batch_stride=R * 3 * D)
                ops.gemm(sh[self.depth:], p["w13_all"], M=R, N=2 * F, K=D, batch=self.depth, a_rows=2 * R,
                         a_batch_stride=2 * R * D, b_rows=2 * F, b_batch_stride=2 * F * D, taps=taps, out=cb["b2f"],
                         out_batch_stride=R * 2 * F)
        return cb

    # ---------------------------------------------------------------- per-step work
    @torch.no_grad()
    def run_blocks(self, x, cond, k=0, t_ints=None):
        """x [xB, C, T] fp32 (xB divides N: the CFG halves share x); cond = prepare_timesteps(...) and k the step index
        into it -> workspace with w.h filled.

        Fused models run 5 launches per block: QKV GEMM (+RoPE) -> attention -> wo -> w1|w3 GEMM (+SwiGLU) -> w2, where
        wo and w2 are row-owning cluster GEMMs (ma3_gemm_rownorm) that update the fp32 residual stream in place and
        write the RMS-normalised, adaLN-modulated 16-bit operand of the next projection.  Only the first norm of a step
        (on the proj_in output) is a stand-alone launch.  MoE models and other widths keep the 7(+8)-launch form."""
        p = self._ensure()
        c = self._ctx
        N = c["shape"][0]
        Lc = c["shape"][1]
        T = x.shape[-1]
        D, H, hd, F = self.hidden_size, self.num_heads, self.head_dim, self.ffn_hidden
        w = self._workspace(N, T)
        M = N * T
        mod = cond["mod"][k]
        mode = self._norm_mode(N, T)
        if mode == "epi" and "b2q" not in cond:
            mode = "plain"                       # conditioning prepared without the bias tables (cond_buffers(S, N) without T)
        fused = mode == "row"
        qs = math.log2(math.e) / math.sqrt(hd)
        eps = self.norm_eps
        if T > p["rope"].shape[0]:
            raise ValueError(f"sequence length {T} exceeds the RoPE table ({p['rope'].shape[0]} positions)")
        ops.proj_in(x, p["proj_w"], p["proj_b"], w.h, N)
        for i, q in enumerate(p["blocks"]):
            o = 6 * D * i
            gate1, gate2 = mod[:, o + 2 * D:o + 3 * D], mod[:, o + 5 * D:o + 6 * D]
            if mode == "epi":
                # 5 launches per block with the norms folded into the tiled GEMMs' epilogues (small batches)
                qkv_kw = dict(M=M, N=3 * D, K=D, epi=L.EPI_QKV_ROPE, q_out=w.q, k_out=w.k, vt_out=w.vt, rope=p["rope"],
                              model_dim=D, head_dim=hd, head_dim_pad=w.hdp, tokens=T, tokens_pad=w.Tp, q_scale=qs)
                if i > 0:   # attention_norm of this block was folded into the previous block's w2 epilogue (w.g, w.ss)
                    ops.gemm(w.g, q["wqkv"], rows_per_sample=T, row_ss=w.ss, ss_dim=D, ss_eps=eps,
                             col_bias2=cond["b2q"][i, k * N:(k + 1) * N], **qkv_kw)
                else:
                    ops.rmsnorm_modulate(w.h, q["attn_norm"], w.u, mod=mod, shift_off=o, scale_off=o + D,
                                         rows_per_sample=T, eps=eps)
                    ops.gemm(w.u, q["wqkv"], **qkv_kw)
                ops.attention(w.q, w.k, w.vt, c["ky"][i] if Lc else None, c["vyt"][i] if Lc else None, q["gate"], w.att,
                              hd=hd)
                wn = p["wn_off"] + 2 * D * i
                ops.gemm(w.att, q["wo"], M=M, N=D, K=D, epi=L.EPI_GATE_RES, out=w.h, gate=gate1, rows_per_sample=T,
                         norm_out=w.g, norm_w=mod[:, wn + D:wn + 2 * D], ss_out=w.ss)            # -> ffn_norm
                ops.gemm(w.g, q["w13"], M=M, N=2 * F, K=D, epi=L.EPI_SWIGLU, out=w.mid, out_ld=F, rows_per_sample=T,
                         row_ss=w.ss, ss_dim=D, ss_eps=eps, col_bias2=cond["b2f"][i, k * N:(k + 1) * N])
                if i + 1 < self.depth:
                    ops.gemm(w.mid, q["w2"], M=M, N=D, K=F, epi=L.EPI_GATE_RES, out=w.h, gate=gate2, rows_per_sample=T,
                             norm_out=w.g, norm_w=mod[:, wn + 2 * D:wn + 3 * D], ss_out=w.ss)    # -> next attention_norm
                else:
                    ops.gemm(w.mid, q["w2"], M=M, N=D, K=F, epi=L.EPI_GATE_RES, out=w.h, gate=gate2, rows_per_sample=T)
                continue
            if not fused or i == 0:
                ops.rmsnorm_modulate(w.h, q["attn_norm"], w.u, mod=mod, shift_off=o, scale_off=o + D, rows_per_sample=T,
                                     eps=eps)
            if self.qk_norm:
                # LayerNorm(q), LayerNorm(k) need whole rows: raw fp32 projections, then one fused norm+RoPE+scatter pass
                ops.gemm(w.u, q["wqkv"], M=M, N=3 * D, K=D, out=w.qkv_raw)
                ops.qknorm_rope(w.qkv_raw, first_section=0, qn=q["qn"], kn=q["kn"], rope=p["rope"], q_out=w.q, k_out=w.k,
                                vt_out=w.vt, tokens=T, tokens_pad=w.Tp, D=D, hd=hd, hdp=w.hdp, q_scale=qs)
            else:
                ops.gemm(w.u, q["wqkv"], M=M, N=3 * D, K=D, epi=L.EPI_QKV_ROPE, q_out=w.q, k_out=w.k, vt_out=w.vt,
                         rope=p["rope"], model_dim=D, head_dim=hd, head_dim_pad=w.hdp, tokens=T, tokens_pad=w.Tp,
                         q_scale=qs)
            ops.attention(w.q, w.k, w.vt, c["ky"][i] if Lc else None, c["vyt"][i] if Lc else None, q["gate"], w.att,
                          hd=hd)
            if fused:
                wn = p["wn_off"] + 2 * D * i
                ops.gemm_rownorm(w.att, q["wo"], w.h, gate1, rows_per_sample=T, eps=eps, u_out=w.u,
                                 wn=mod[:, wn + D:wn + 2 * D], shift=mod[:, o + 3 * D:o + 4 * D])            # -> ffn_norm
                ops.gemm(w.u, q["w13"], M=M, N=2 * F, K=D, epi=L.EPI_SWIGLU, out=w.mid, out_ld=F)
                if i + 1 < self.depth:   # -> attention_norm of the next block
                    ops.gemm_rownorm(w.mid, q["w2"], w.h, gate2, rows_per_sample=T, eps=eps, u_out=w.u,
                                     wn=mod[:, wn + 2 * D:wn + 3 * D], shift=mod[:, o + 6 * D:o + 7 * D])
                else:
                    ops.gemm_rownorm(w.mid, q["w2"], w.h, gate2, rows_per_sample=T)
                continue
            ops.gemm(w.att, q["wo"], M=M, N=D, K=D, epi=L.EPI_GATE_RES, out=w.h, gate=gate1, rows_per_sample=T)
            ops.rmsnorm_modulate(w.h, q["ffn_norm"], w.u, mod=mod, shift_off=o + 3 * D, scale_off=o + 4 * D,
                                 rows_per_sample=T, eps=eps)
            if self.num_experts:
                self._moe_ffn(q, w, gate2, t_ints, N, T)
            else:
                ops.gemm(w.u, q["w13"], M=M, N=2 * F, K=D, epi=L.EPI_SWIGLU, out=w.mid, out_ld=F)
                ops.gemm(w.mid, q["w2"], M=M, N=D, K=F, epi=L.EPI_GATE_RES, out=w.h, gate=gate2, rows_per_sample=T)
        return w

    def _moe_ffn(self, q, w, gate2, t_ints, N, T):
        """flag_large_dit_moe.py:516-538: time expert e = t // 250 (per sample; a pointer switch, no routing kernel),
        then per-band frequency experts on sliced weights (exact, because masked columns contribute exact zeros)."""
        D, F, E = self.hidden_size, self.ffn_hidden, self.num_experts
        band = D // E
        n = 0
        while n < N:  # runs of consecutive samples sharing an expert
            e = int(t_ints[n]) // 250
            m = n + 1
            while m < N and int(t_ints[m]) // 250 == e:
                m += 1
            rows = slice(n * T, m * T)
            if 0 <= e < E:
                ops.gemm(w.u[rows], q["t_w13"][e], M=(m - n) * T, N=2 * F, K=D, epi=L.EPI_SWIGLU, out=w.mid[rows], out_ld=F)
                ops.gemm(w.mid[rows], q["t_w2"][e], M=(m - n) * T, N=D, K=F, out=w.y1[rows])
            else:
                w.y1[rows].zero_()
            n = m
        # The E band experts as two batched launches (z = band): operand / output column slices of y1, h and the gate are
        # batch strides of `band` columns.  (One launch per expert and projection was 8 of the block's 14 launches, and
        # at batch 2 every one of them is latency-bound.)
        M = N * T
        ops.gemm(w.y1, q["f_w13"], M=M, N=2 * F, K=band, batch=E, a_ld=D, a_batch_stride=band, b_rows=2 * F,
                 b_batch_stride=2 * F * band, epi=L.EPI_SWIGLU, out=w.mid_e, out_ld=F, out_batch_stride=M * F)
        ops.gemm(w.mid_e, q["f_w2"], M=M, N=band, K=F, batch=E, a_batch_stride=M * F, b_rows=band, b_batch_stride=band * F,
                 epi=L.EPI_GATE_RES, out=w.h, out_ld=D, out_batch_stride=band, gate=gate2, gate_batch_stride=band,
                 rows_per_sample=T)


class TxtFlagLargeImprovedDiTV2(TxtFlagLargeDiT):
    """flag_large_dit.py:256-299 (the class every txt2audio / txt2music config targets)."""

    def __init__(self, in_channels, context_dim, hidden_size=1152, depth=28, num_heads=16, max_len=1000):
        super().__init__(in_channels, context_dim, hidden_size, depth, num_heads, max_len)


class VideoFlagLargeDiT(TxtFlagLargeDiT):
    """flag_large_dit_moe.py:613-740: ConditionEmbedder on the context, MoE feed-forward (time + frequency experts)."""

    _video = True

    def __init__(self, in_channels, context_dim, hidden_size=1152, depth=28, num_heads=16, max_len=1000,
                 n_kv_heads=None, multiple_of: int = 256, ffn_dim_multiplier=None, norm_eps=1e-5, qk_norm=None,
                 rope_scaling_factor: float = 1.0, ntk_factor: float = 1, num_experts=8):
        super().__init__(in_channels, context_dim, hidden_size, depth, num_heads, max_len, n_kv_heads, multiple_of,
                         ffn_dim_multiplier, norm_eps, qk_norm, rope_scaling_factor, ntk_factor, num_experts)
