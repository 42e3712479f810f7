"""Drop-in first-stage model: ldm/models/autoencoder1d.py:18-62 (AutoencoderKL) with the 1-D Decoder1D
(autoencoder1d.py:415-517) on the sampling path and the Encoder1D (autoencoder1d.py:319-413) for the inpainting /
ground-truth paths.

Same constructor arguments and state_dict keys (`encoder.*`, `quant_conv.*`, `post_quant_conv.*`, `decoder.*`; loss
keys of a full checkpoint are ignored; a decoder-only state_dict is accepted and leaves encode() disabled).  Both
halves run channels-last in bf16 on the tap-GEMM (implicit-GEMM conv1d on tcgen05) with GroupNorm+swish as the only
separate elementwise pass; the residual stream is kept in fp32.
"""
import torch
import torch.nn as nn

from . import lib as L
from . import ops
from .convs import PackedConv, cpad


class _Conv(nn.Module):
    def __init__(self, cin, cout, k):
        super().__init__()
        a = (1.0 / (cin * k)) ** 0.5
        self.weight = nn.Parameter(torch.empty(cout, cin, k).uniform_(-a, a))
        self.bias = nn.Parameter(torch.empty(cout).uniform_(-a, a))


class _GN(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))


class _Res(nn.Module):
    def __init__(self, cin, cout, k=3):
        super().__init__()
        self.norm1, self.conv1 = _GN(cin), _Conv(cin, cout, k)
        self.norm2, self.conv2 = _GN(cout), _Conv(cout, cout, k)
        if cin != cout:
            self.nin_shortcut = _Conv(cin, cout, 1)


class _Attn(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.norm = _GN(c)
        self.q, self.k, self.v, self.proj_out = _Conv(c, c, 1), _Conv(c, c, 1), _Conv(c, c, 1), _Conv(c, c, 1)


class Decoder1D(nn.Module):
    """Parameter tree of autoencoder1d.py:415-482 (ResNet blocks use k=3 regardless of `kernel_size`; only
    conv_in / conv_out take it, as in the reference)."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_layers=[], down_layers=[],
                 dropout=0.0, kernel_size=3, resamp_with_conv=True, in_channels=None, z_channels, give_pre_end=False,
                 tanh_out=False, **ignorekwargs):
        super().__init__()
        if not resamp_with_conv or give_pre_end:
            raise NotImplementedError("resamp_with_conv=False / give_pre_end=True are not used by any shipped config")
        self.ch, self.ch_mult, self.num_res_blocks = ch, list(ch_mult), num_res_blocks
        self.kernel_size, self.tanh_out, self.out_ch, self.z_channels = kernel_size, tanh_out, out_ch, z_channels
        self.down_layers = [i + 1 for i in down_layers]
        self.attn_layers = list(attn_layers)
        nl = len(self.ch_mult)
        block_in = ch * self.ch_mult[nl - 1]
        self.conv_in = _Conv(z_channels, block_in, kernel_size)
        self.mid = nn.Module()
        self.mid.block_1, self.mid.attn_1, self.mid.block_2 = _Res(block_in, block_in), _Attn(block_in), _Res(block_in, block_in)
        self.up = nn.ModuleList()
        for lvl in reversed(range(nl)):
            block_out = ch * self.ch_mult[lvl]
            up = nn.Module()
            up.block, up.attn = nn.ModuleList(), nn.ModuleList()
            for _ in range(num_res_blocks + 1):
                up.block.append(_Res(block_in, block_out))
                block_in = block_out
                if lvl in self.attn_layers:
                    up.attn.append(_Attn(block_in))
            if lvl in self.down_layers:
                up.upsample = nn.Module()
                up.upsample.conv = _Conv(block_in, block_in, 3)
            self.up.insert(0, up)
        self.norm_out = _GN(block_in)
        self.conv_out = _Conv(block_in, out_ch, kernel_size)


class Encoder1D(nn.Module):
    """Parameter tree of autoencoder1d.py:319-381 (here the ResNet blocks DO take `kernel_size`, unlike the decoder)."""

    def __init__(self, *, ch, out_ch=None, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_layers=[], down_layers=[],
                 dropout=0.0, resamp_with_conv=True, in_channels, z_channels, double_z=True, kernel_size=3,
                 **ignore_kwargs):
        super().__init__()
        if not resamp_with_conv:
            raise NotImplementedError("resamp_with_conv=False is not used by any shipped config")
        self.ch, self.ch_mult, self.num_res_blocks = ch, list(ch_mult), num_res_blocks
        self.kernel_size, self.in_channels = kernel_size, in_channels
        self.down_layers, self.attn_layers = list(down_layers), list(attn_layers)
        self.conv_in = _Conv(in_channels, ch, kernel_size)
        self.down = nn.ModuleList()
        block_in = ch
        for lvl in range(len(self.ch_mult)):
            block_out = ch * self.ch_mult[lvl]
            down = nn.Module()
            down.block, down.attn = nn.ModuleList(), nn.ModuleList()
            for _ in range(num_res_blocks):
                down.block.append(_Res(block_in, block_out, kernel_size))
                block_in = block_out
                if lvl in self.attn_layers:
                    down.attn.append(_Attn(block_in))
            if lvl in self.down_layers:
                down.downsample = nn.Module()
                down.downsample.conv = _Conv(block_in, block_in, 3)
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = _Res(block_in, block_in, kernel_size)
        self.mid.attn_1 = _Attn(block_in)
        self.mid.block_2 = _Res(block_in, block_in, kernel_size)
        self.norm_out = _GN(block_in)
        self.conv_out = _Conv(block_in, 2 * z_channels if double_z else z_channels, kernel_size)


class DiagonalGaussianDistribution:
    """ldm/modules/distributions/distributions.py:24-77: the posterior returned by encode() (tiny tensors: plain torch)."""

    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self):
        return self.mean + self.std * torch.randn(self.mean.shape, device=self.parameters.device)

    def mode(self):
        return self.mean

    def kl(self, other=None):
        if self.deterministic:
            return torch.Tensor([0.0])
        dims = list(range(1, self.mean.dim()))
        if other is None:
            return 0.5 * torch.sum(self.mean.pow(2) + self.var - 1.0 - self.logvar, dim=dims)
        return 0.5 * torch.sum((self.mean - other.mean).pow(2) / other.var + self.var / other.var - 1.0 - self.logvar
                               + other.logvar, dim=dims)

    def nll(self, sample, dims=[1, 2, 3]):
        if self.deterministic:
            return torch.Tensor([0.0])
        import math
        return 0.5 * torch.sum(math.log(2.0 * math.pi) + self.logvar + (sample - self.mean).pow(2) / self.var, dim=dims)


class AutoencoderKL(nn.Module):
    """B200 drop-in for ldm.models.autoencoder1d.AutoencoderKL."""

    def __init__(self, embed_dim, ddconfig, lossconfig=None, ckpt_path=None, ignore_keys=[], image_key="image",
                 monitor=None):
        super().__init__()
        assert ddconfig["double_z"]
        self.image_key, self.embed_dim = image_key, embed_dim
        self.encoder = Encoder1D(**ddconfig)
        self.decoder = Decoder1D(**ddconfig)
        self.quant_conv = _Conv(2 * ddconfig["z_channels"], 2 * embed_dim, 1)
        self.post_quant_conv = _Conv(embed_dim, ddconfig["z_channels"], 1)
        self._has_encoder = True
        if monitor is not None:
            self.monitor = monitor
        self._packed = None
        self._bufs = {}
        self.register_load_state_dict_post_hook(lambda m, k: m.invalidate())
        if ckpt_path is not None:
            self.init_from_ckpt(ckpt_path, ignore_keys=ignore_keys)

    def invalidate(self):
        self._packed = None

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def init_from_ckpt(self, path, ignore_keys=list()):
        sd = torch.load(path, map_location="cpu", weights_only=False)["state_dict"]
        for k in list(sd.keys()):
            if any(k.startswith(ik) for ik in ignore_keys):
                del sd[k]
        self.load_state_dict(sd, strict=False)

    def load_state_dict(self, state_dict, strict=True, **kw):
        """As nn.Module.load_state_dict; a state_dict without `encoder.*` / `quant_conv.*` (the sampling path only needs
        the decoder) is accepted under strict=True and disables encode(); `loss.*` keys of a full checkpoint are ignored."""
        sd = {k: v for k, v in state_dict.items() if not k.startswith("loss.")}
        r = super().load_state_dict(sd, strict=False, **kw)
        enc_missing = [k for k in r.missing_keys if k.startswith(("encoder.", "quant_conv."))]
        other_missing = [k for k in r.missing_keys if k not in enc_missing]
        self._has_encoder = not enc_missing
        if enc_missing:
            r.missing_keys[:] = other_missing
        if strict and (other_missing or r.unexpected_keys):
            raise RuntimeError(f"AutoencoderKL.load_state_dict: missing {other_missing}, unexpected {r.unexpected_keys}")
        return r

    # ---------------------------------------------------------------- packing
    def _pack(self):
        dev = self.post_quant_conv.weight.device
        if dev.type != "cuda":
            raise L.Ma3Error("ma3_b200 modules run on CUDA only (no CPU fallback): call .cuda() first")
        L.require_device()
        bf = torch.bfloat16
        pc = lambda m, **kw: PackedConv(m.weight, m.bias, dtype=bf, device=dev, **kw)
        f32 = lambda t: t.detach().float().to(dev).contiguous()
        gn = lambda m: (f32(m.weight), f32(m.bias))
        d = self.decoder
        ks = d.kernel_size

        def res(m):
            r = {"n1": gn(m.norm1), "c1": pc(m.conv1), "n2": gn(m.norm2), "c2": pc(m.conv2)}
            if hasattr(m, "nin_shortcut"):
                r["sc"] = pc(m.nin_shortcut)
            return r

        def attn(m):
            C = m.q.weight.shape[0]
            return {"n": gn(m.norm), "q": pc(m.q), "k": pc(m.k), "p": pc(m.proj_out), "C": C,
                    "wv": m.v.weight.detach().float()[:, :, 0].to(device=dev, dtype=bf).contiguous(), "bv": f32(m.v.bias)}

        p = {"pq": pc(self.post_quant_conv), "cin": pc(d.conv_in, padding=ks // 2),
             "mid1": res(d.mid.block_1), "mida": attn(d.mid.attn_1), "mid2": res(d.mid.block_2), "up": []}
        for lvl in range(len(d.ch_mult)):
            u = d.up[lvl]
            e = {"blocks": [res(b) for b in u.block], "attn": [attn(a) for a in u.attn]}
            if hasattr(u, "upsample"):
                e["upconv"] = pc(u.upsample.conv)
            p["up"].append(e)
        p["nout"] = gn(d.norm_out)
        p["cout"] = pc(d.conv_out, padding=ks // 2)
        if self._has_encoder:
            e = self.encoder
            q = {"cin": pc(e.conv_in), "down": [], "mid1": res(e.mid.block_1), "mida": attn(e.mid.attn_1),
                 "mid2": res(e.mid.block_2), "nout": gn(e.norm_out), "cout": pc(e.conv_out), "qc": pc(self.quant_conv)}
            for lvl in range(len(e.ch_mult)):
                dn = e.down[lvl]
                ent = {"blocks": [res(b) for b in dn.block], "attn": [attn(a) for a in dn.attn]}
                if hasattr(dn, "downsample"):
                    # Downsample1D (autoencoder1d.py:296-317): zero-pad one frame on the right, conv k3 stride 2.  On the
                    # frame-PAIR view [B, T/2, 2C] of the input it is a two-tap GEMM:
                    #   out[t] = [W0 | W1] . pair[t] + [W2 | 0] . pair[t+1]     (pair[T/2] = 0 by TMA zero fill)
                    wt = dn.downsample.conv.weight.detach().float()           # [C, C, 3]
                    C = wt.shape[0]
                    wp = torch.zeros(2 * C, 2 * C)
                    wp[:C, :C], wp[:C, C:], wp[C:, :C] = wt[:, :, 0], wt[:, :, 1], wt[:, :, 2]
                    ent["down_w"] = wp.to(device=dev, dtype=bf).contiguous()
                    ent["down_b"] = f32(dn.downsample.conv.bias)
                q["down"].append(ent)
            p["enc"] = q
        self._packed = p

    def _buf(self, name, shape, dtype, zero=False):
        key = (name, tuple(shape), dtype)
        b = self._bufs.get(key)
        if b is None:
            dev = self.post_quant_conv.weight.device
            b = (torch.zeros if zero else torch.empty)(*shape, device=dev, dtype=dtype)
            self._bufs[key] = b
        return b

    # ---------------------------------------------------------------- decode
    def _res(self, r, h, B, T):
        """h fp32 [B,T,Cin] -> fp32 [B,T,Cout]: GN+swish -> conv k3 -> GN+swish -> conv k3 (+ skip) (:215-235)."""
        bf = torch.bfloat16
        cin, cout = r["c1"].cin, r["c1"].cout
        t1 = ops.groupnorm_swish(h, *r["n1"], self._buf("t1", (B, T, cin), bf))
        c1 = r["c1"](t1, self._buf("c1", (B, T, cout), bf))
        t2 = ops.groupnorm_swish(c1, *r["n2"], self._buf("t2", (B, T, cout), bf))
        if "sc" in r:
            h16 = ops.cast(h, self._buf("h16", (B, T, cin), bf))
            skip = r["sc"](h16, self._buf("skip", (B, T, cout), torch.float32))
        else:
            skip = h
        return r["c2"](t2, self._next_h(B, T, cout, h), res=skip)

    def _next_h(self, B, T, C, cur):
        """Ping-pong fp32 residual-stream buffers (never alias the current one)."""
        a = self._buf("hA", (B, T, C), torch.float32)
        b = self._buf("hB", (B, T, C), torch.float32)
        return b if cur.data_ptr() == a.data_ptr() else a

    def _attn(self, a, h, B, T):
        """Single-head attention over T with scale C^-0.5 (autoencoder1d.py:257-278) as two batched GEMMs."""
        bf = torch.bfloat16
        C = a["C"]
        Tp = (T + 63) // 64 * 64
        hn = ops.groupnorm_swish(h, *a["n"], self._buf("t1", (B, T, C), bf), swish=False)
        q = a["q"](hn, self._buf("aq", (B, T, C), bf))
        k = a["k"](hn, self._buf("ak", (B, T, C), bf))
        vt = self._buf("avt", (B, C, Tp), bf, zero=True)
        ops.gemm(a["wv"], hn, M=C, N=T, K=C, batch=B, a_rows=C, b_rows=T, b_batch_stride=T * C, out=vt, out_ld=Tp,
                 out_batch_stride=C * Tp, bias=a["bv"], bias_per_row=True)
        S = self._buf("aS", (B, T, Tp), torch.float32)
        ops.gemm(q, k, M=T, N=T, K=C, batch=B, a_rows=T, a_batch_stride=T * C, b_rows=T, b_batch_stride=T * C, out=S,
                 out_ld=Tp, out_batch_stride=T * Tp)
        P = ops.softmax_rows(S, self._buf("aP", (B, T, Tp), bf), T, float(C) ** -0.5)
        o = self._buf("ao", (B, T, C), bf)
        ops.gemm(P, vt, M=T, N=C, K=Tp, batch=B, a_rows=T, a_batch_stride=T * Tp, b_rows=C, b_batch_stride=C * Tp,
                 out=o, out_ld=C, out_batch_stride=T * C)
        return a["p"](o, self._next_h(B, T, C, h), res=h)

    @torch.no_grad()
    def decode(self, z):
        """z fp32 [B, z_channels, T] -> mel fp32 [B, out_ch, 2^len(down_layers) * T] (autoencoder1d.py:59-62,484-517)."""
        if self._packed is None:
            self._pack()
        p, d = self._packed, self.decoder
        dev = self.post_quant_conv.weight.device
        z = z.to(device=dev, dtype=torch.float32).contiguous()
        B, zc, T = z.shape
        bf = torch.bfloat16
        zin = ops.nct_to_ntc(z, self._buf("zin", (B, T, p["pq"].cin_pad), bf))
        zq = p["pq"](zin, self._buf("zq", (B, T, p["cin"].cin_pad), bf, zero=True))
        h = p["cin"](zq, self._buf("hA", (B, T, p["cin"].cout), torch.float32))
        h = self._res(p["mid1"], h, B, T)
        h = self._attn(p["mida"], h, B, T)
        h = self._res(p["mid2"], h, B, T)
        for lvl in reversed(range(len(d.ch_mult))):
            e = p["up"][lvl]
            for ib, r in enumerate(e["blocks"]):
                h = self._res(r, h, B, T)
                if e["attn"]:
                    h = self._attn(e["attn"][ib], h, B, T)
            if "upconv" in e:
                C = h.shape[-1]
                h16 = ops.cast(h, self._buf("h16", (B, T, C), bf))
                up = ops.upsample_nearest2(h16, self._buf("up", (B, 2 * T, C), bf))
                T = 2 * T
                h = e["upconv"](up, self._buf("hA", (B, T, C), torch.float32))
        C = h.shape[-1]
        t1 = ops.groupnorm_swish(h, *p["nout"], self._buf("t1", (B, T, C), bf))
        y = p["cout"](t1, self._buf("y", (B, T, d.out_ch), torch.float32), act=3 if d.tanh_out else 0)
        out = torch.empty(B, d.out_ch, T, device=dev, dtype=torch.float32)
        return ops.ntc_to_nct(y, out)

    @torch.no_grad()
    def encode(self, x):
        """x fp32 [B, in_channels, T] -> DiagonalGaussianDistribution over [B, embed_dim, T / 2^len(down_layers)]
        (autoencoder1d.py:49-53, 383-413)."""
        if not self._has_encoder:
            raise L.Ma3Error("this AutoencoderKL was loaded from a decoder-only state_dict: encode() needs encoder.* / quant_conv.*")
        if self._packed is None:
            self._pack()
        q, e = self._packed["enc"], self.encoder
        dev = self.post_quant_conv.weight.device
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        B, _, T = x.shape
        bf = torch.bfloat16
        xin = ops.nct_to_ntc(x, self._buf("xin", (B, T, q["cin"].cin_pad), bf, zero=True))
        h = q["cin"](xin, self._buf("hA", (B, T, q["cin"].cout), torch.float32))
        for lvl, ent in enumerate(q["down"]):
            for ib, r in enumerate(ent["blocks"]):
                h = self._res(r, h, B, T)
                if ent["attn"]:
                    h = self._attn(ent["attn"][ib], h, B, T)
            if "down_w" in ent:
                C = h.shape[-1]
                Te = T + (T & 1)                                  # an odd length gets its zero frame explicitly
                h16 = self._buf("d16", (B, Te, C), bf, zero=True)
                ops.cast(h, h16[:, :T]) if T == Te else [ops.cast(h[b], h16[b, :T]) for b in range(B)]
                To = (T + 1 - 3) // 2 + 1
                out = self._next_h(B, To, C, h)
                ops.gemm(h16, ent["down_w"], M=To, N=C, K=2 * C, batch=B, a_rows=Te // 2, a_batch_stride=Te * C,
                         b_rows=2 * C, taps=((0, 0), (1, C)), out=out, out_ld=C, out_batch_stride=To * C,
                         bias=ent["down_b"])
                h, T = out, To
        h = self._res(q["mid1"], h, B, T)
        h = self._attn(q["mida"], h, B, T)
        h = self._res(q["mid2"], h, B, T)
        C = h.shape[-1]
        t1 = ops.groupnorm_swish(h, *q["nout"], self._buf("t1", (B, T, C), bf))
        mo = q["cout"](t1, self._buf("emo", (B, T, q["qc"].cin_pad), bf, zero=True))
        mq = q["qc"](mo, self._buf("emq", (B, T, q["qc"].cout), torch.float32))
        moments = torch.empty(B, q["qc"].cout, T, device=dev, dtype=torch.float32)
        return DiagonalGaussianDistribution(ops.ntc_to_nct(mq, moments))

    @torch.no_grad()
    def forward(self, input, sample_posterior=True):
        """autoencoder1d.py:64-71: (reconstruction, posterior)."""
        posterior = self.encode(input)
        z = posterior.sample() if sample_posterior else posterior.mode()
        return self.decode(z), posterior
