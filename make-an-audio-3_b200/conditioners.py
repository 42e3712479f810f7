"""Text conditioners on the GPU kernels of this package: the step immediately before the
sampling path.  `FrozenCLAPFLANEmbedder` (ldm/modules/encoders/modules.py:133-191) runs a prompt pair through

  * the CLAP caption encoder = BERT-base (`transformers` AutoModel, encoders/CLAP/clap.py:47-56) followed by the CLAP
    `Projection` (clap.py:17-30) on every token, and
  * the T5 v1.1-large encoder (gated tanh-GELU feed-forward, relative position bias, RMS "T5LayerNorm"),

and concatenates the two [B, 77, 1024] results along the token axis: the [B, 154, 1024] context the DiT is conditioned
on.  Both encoders are called WITHOUT an attention mask by the reference (modules.py:187-190), i.e. full attention over
the 77 padded tokens -- restated here as such.

Everything after the tokenizer runs in the library's kernels: embedding lookup (`ma3_embed_rows`), every Linear as a
tap-GEMM (bias / exact-GELU / residual epilogues; the T5 feed-forward through the gated epilogue with tanh-GELU), the
attention core as two batched GEMMs around `ma3_softmax_rows` (77 tokens: one key tile; T5's additive position bias is
the residual operand of the QK^T GEMM), LayerNorm / RMSNorm rows.  fp32 residual stream, bf16 GEMM operands.

Weights are addressed by the reference module's own state_dict keys (`caption_encoder.base.*`,
`caption_encoder.projection.*`, `t5_transformer.*`).  There are no tokenizer files or checkpoints offline:
`encode(text)` needs the two tokenizers handed in (any callable with the `transformers` tokenizer call signature);
`encode_tokens(ori_tokens, struct_tokens)` is the GPU part and what the parity tests drive.
"""
import math

import torch

from . import lib as L
from . import ops

BERT_BASE = dict(hidden=768, layers=12, heads=12, ffn=3072, vocab=30522, max_pos=512, eps=1e-12)
T5_V11_LARGE = dict(d_model=1024, layers=24, heads=16, d_kv=64, d_ff=2816, vocab=32128, buckets=32, max_distance=128,
                    eps=1e-6)


def t5_relative_buckets(T, num_buckets=32, max_distance=128):
    """Bidirectional T5 bucket of (key j - query i) -> [T, T] int64 (the published T5 bucketing the encoder's
    `relative_attention_bias` embedding is indexed with)."""
    ctx = torch.arange(T)[:, None]
    mem = torch.arange(T)[None, :]
    rel = mem - ctx
    nb = num_buckets // 2
    ret = (rel > 0).long() * nb
    n = rel.abs()
    max_exact = nb // 2
    is_small = n < max_exact
    large = max_exact + (torch.log(n.float().clamp_min(1) / max_exact) / math.log(max_distance / max_exact) *
                         (nb - max_exact)).long()
    large = torch.minimum(large, torch.full_like(large, nb - 1))
    return ret + torch.where(is_small, n, large)


class _Buffers:
    """Named scratch buffers keyed by (name, shape, dtype); re-used across calls of the same batch size."""

    def __init__(self, device):
        self.device, self._b = device, {}

    def get(self, name, shape, dtype, zero=False):
        key = (name, tuple(shape), dtype)
        t = self._b.get(key)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(*shape, device=self.device, dtype=dtype)
            self._b[key] = t
        return t


def _lin(x16, w16, out, *, bias=None, act=0, res=None, accumulate=False):
    """out = act(x16 @ w16^T + bias) (+ res | + out).  x16 [M, K] bf16, w16 [N, K] bf16, out fp32 or bf16 [M, N]."""
    M, K = x16.shape
    N = w16.shape[0]
    return ops.gemm(x16, w16, M=M, N=N, K=K, out=out, out_ld=N, bias=bias, act=act, res=res,
                    res_ld=N if res is not None else None, accumulate=accumulate)


def _attention_core(qk, vt, att, S, P, *, B, T, H, hd, scale, pos_bias=None):
    """att[b*T + i, h*hd + d] = softmax_j(scale * q.k (+ pos_bias[h, i, j])) v for every sample b and head h.
    qk [B*T, 2*H*hd] bf16 (q | k columns), vt [B, H*hd, Tp] bf16 (V^T, keys contiguous), S / P [H, T, Tp] scratch.
    One QK^T GEMM, one softmax, one PV GEMM per sample, batched over the heads (z = head)."""
    D = H * hd
    Tp = vt.shape[-1]
    for b in range(B):
        q = qk[b * T:(b + 1) * T]
        ops.gemm(q, q[:, D:], M=T, N=T, K=hd, batch=H, a_rows=T, a_ld=2 * D, a_batch_stride=hd, b_rows=T, b_ld=2 * D,
                 b_batch_stride=hd, out=S, out_ld=Tp, out_batch_stride=T * Tp, res=pos_bias,
                 res_ld=Tp if pos_bias is not None else None, res_batch_stride=T * Tp if pos_bias is not None else 0)
        ops.softmax_rows(S, P, T, scale)
        ops.gemm(P, vt[b], M=T, N=hd, K=Tp, batch=H, a_rows=T, a_batch_stride=T * Tp, b_rows=hd, b_batch_stride=hd * Tp,
                 out=att[b * T:(b + 1) * T], out_ld=D, out_batch_stride=hd)
    return att


def _vt(w_v, b_v, u, vt, *, B, T, D):
    """V^T of every sample in one launch: vt[b, d, j] = sum_k w_v[d, k] u[b*T + j, k] (+ b_v[d])."""
    Tp = vt.shape[-1]
    ops.gemm(w_v, u, M=D, N=T, K=u.shape[1], batch=B, a_rows=D, b_rows=T, b_batch_stride=T * u.shape[1], out=vt, out_ld=Tp,
             out_batch_stride=D * Tp, bias=b_v, bias_per_row=b_v is not None)
    return vt


class BertTextEncoder:
    """BERT-base encoder stack + CLAP Projection (`TextEncoder.base` / `.projection`, clap.py:47-56) applied to every
    token, as `FrozenCLAPFLANEmbedder.encode` does (modules.py:187-188).  Post-LN blocks, exact GELU, absolute
    positions, token type 0."""

    def __init__(self, state_dict, prefix="caption_encoder.", cfg=None, device="cuda"):
        self.cfg = dict(BERT_BASE if cfg is None else cfg)
        self.device = torch.device(device)
        g = lambda k: state_dict[prefix + k].detach().to(self.device, torch.float32).contiguous()
        b16 = lambda t: t.to(torch.bfloat16).contiguous()
        e = "base.embeddings."
        self.word, self.pos, self.type0 = g(e + "word_embeddings.weight"), g(e + "position_embeddings.weight"), \
            g(e + "token_type_embeddings.weight")[0].contiguous()
        self.emb_ln = (g(e + "LayerNorm.weight"), g(e + "LayerNorm.bias"))
        self.layers = []
        for i in range(self.cfg["layers"]):
            p = f"base.encoder.layer.{i}."
            a = p + "attention.self."
            self.layers.append(dict(
                wqk=b16(torch.cat([g(a + "query.weight"), g(a + "key.weight")])),
                bqk=torch.cat([g(a + "query.bias"), g(a + "key.bias")]).contiguous(),
                wv=b16(g(a + "value.weight")), bv=g(a + "value.bias"),
                wo=b16(g(p + "attention.output.dense.weight")), bo=g(p + "attention.output.dense.bias"),
                ln1=(g(p + "attention.output.LayerNorm.weight"), g(p + "attention.output.LayerNorm.bias")),
                wi=b16(g(p + "intermediate.dense.weight")), bi=g(p + "intermediate.dense.bias"),
                wo2=b16(g(p + "output.dense.weight")), bo2=g(p + "output.dense.bias"),
                ln2=(g(p + "output.LayerNorm.weight"), g(p + "output.LayerNorm.bias"))))
        pj = "projection."
        self.p1, self.p2 = b16(g(pj + "linear1.weight")), b16(g(pj + "linear2.weight"))
        self.pln = (g(pj + "layer_norm.weight"), g(pj + "layer_norm.bias"))
        self.d_out = self.p1.shape[0]
        self._buf = _Buffers(self.device)

    @torch.no_grad()
    def __call__(self, tokens):
        """tokens int64 [B, T] -> fp32 [B, T, d_proj]."""
        L.require_device()
        c, bf, f32 = self.cfg, torch.bfloat16, torch.float32
        B, T = tokens.shape
        D, H, F = c["hidden"], c["heads"], c["ffn"]
        hd, M, Tp = D // H, B * T, (T + 15) // 16 * 16
        if T > c["max_pos"]:
            raise L.Ma3Error(f"BERT: {T} tokens exceed the {c['max_pos']} learned positions")
        if not torch.cuda.is_current_stream_capturing() and (int(tokens.min()) < 0 or int(tokens.max()) >= c["vocab"]):
            raise L.Ma3Error(f"BERT token id outside [0, {c['vocab']})")
        buf = self._buf.get
        ids = tokens.to(self.device).reshape(-1).contiguous()
        tmp = ops.embed_rows(self.word, ids, buf("tmp", (M, D), f32), T=T, pos=self.pos, type0=self.type0)
        h = ops.layernorm_rows(tmp, *self.emb_ln, buf("h", (M, D), f32), eps=c["eps"])
        u, qk, att = buf("u", (M, D), bf), buf("qk", (M, 2 * D), bf), buf("att", (M, D), bf)
        vt, mid = buf("vt", (B, D, Tp), bf, zero=True), buf("mid", (M, F), bf)
        S, P = buf("S", (H, T, Tp), f32), buf("P", (H, T, Tp), bf)
        for q in self.layers:
            ops.cast(h, u)
            _lin(u, q["wqk"], qk, bias=q["bqk"])
            _vt(q["wv"], q["bv"], u, vt, B=B, T=T, D=D)
            _attention_core(qk, vt, att, S, P, B=B, T=T, H=H, hd=hd, scale=hd ** -0.5)
            _lin(att, q["wo"], tmp, bias=q["bo"], res=h)                    # h + attention output
            ops.layernorm_rows(tmp, *q["ln1"], h, eps=c["eps"])
            ops.cast(h, u)
            _lin(u, q["wi"], mid, bias=q["bi"], act=2)                      # exact (erf) GELU
            _lin(mid, q["wo2"], tmp, bias=q["bo2"], res=h)
            ops.layernorm_rows(tmp, *q["ln2"], h, eps=c["eps"])
        # CLAP Projection (clap.py:25-30; dropout is the identity at inference): LN(e1 + linear2(gelu(e1)))
        ops.cast(h, u)
        Dp = self.d_out
        e1, ge = buf("e1", (M, Dp), f32), buf("ge", (M, Dp), bf)
        _lin(u, self.p1, e1)
        _lin(u, self.p1, ge, act=2)
        e2 = _lin(ge, self.p2, buf("e2", (M, Dp), f32), res=e1)
        out = torch.empty(M, Dp, device=self.device, dtype=f32)
        ops.layernorm_rows(e2, *self.pln, out, eps=1e-5)
        return out.view(B, T, Dp)


class T5TextEncoder:
    """T5 v1.1 encoder stack (`T5EncoderModel`, modules.py:189): pre-RMSNorm blocks, unscaled attention logits plus the
    bucketed relative position bias of block 0 (shared by all blocks), gated tanh-GELU feed-forward, final RMSNorm."""

    def __init__(self, state_dict, prefix="t5_transformer.", cfg=None, device="cuda"):
        self.cfg = dict(T5_V11_LARGE if cfg is None else cfg)
        self.device = torch.device(device)
        g = lambda k: state_dict[prefix + k].detach().to(self.device, torch.float32).contiguous()
        b16 = lambda t: t.to(torch.bfloat16).contiguous()
        self.embed = g("shared.weight") if prefix + "shared.weight" in state_dict else g("encoder.embed_tokens.weight")
        self.rel = g("encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight")   # [buckets, heads]
        self.layers = []
        for i in range(self.cfg["layers"]):
            a = f"encoder.block.{i}.layer.0."
            f = f"encoder.block.{i}.layer.1."
            wi0, wi1 = g(f + "DenseReluDense.wi_0.weight"), g(f + "DenseReluDense.wi_1.weight")
            self.layers.append(dict(
                ln1=g(a + "layer_norm.weight"),
                wqk=b16(torch.cat([g(a + "SelfAttention.q.weight"), g(a + "SelfAttention.k.weight")])),
                wv=b16(g(a + "SelfAttention.v.weight")), wo=b16(g(a + "SelfAttention.o.weight")),
                ln2=g(f + "layer_norm.weight"),
                # gated epilogue: rows interleaved (gate row 2i = wi_0[i] -> GELU, value row 2i + 1 = wi_1[i])
                wi=b16(torch.stack([wi0, wi1], 1).reshape(2 * wi0.shape[0], wi0.shape[1])),
                wo2=b16(g(f + "DenseReluDense.wo.weight"))))
        self.final_ln = g("encoder.final_layer_norm.weight")
        self._buf = _Buffers(self.device)
        self._bias = {}

    def _pos_bias(self, T, Tp):
        key = (T, Tp)
        if key not in self._bias:
            c = self.cfg
            bk = t5_relative_buckets(T, c["buckets"], c["max_distance"]).to(self.device)
            bias = torch.zeros(c["heads"], T, Tp, device=self.device, dtype=torch.float32)
            bias[:, :, :T] = self.rel[bk].permute(2, 0, 1)
            self._bias[key] = bias.contiguous()
        return self._bias[key]

    @torch.no_grad()
    def __call__(self, tokens):
        """tokens int64 [B, T] -> last_hidden_state fp32 [B, T, d_model]."""
        L.require_device()
        c, bf, f32 = self.cfg, torch.bfloat16, torch.float32
        B, T = tokens.shape
        D, H, hd, F = c["d_model"], c["heads"], c["d_kv"], c["d_ff"]
        inner, M, Tp = H * hd, B * T, (T + 15) // 16 * 16
        if not torch.cuda.is_current_stream_capturing() and (int(tokens.min()) < 0 or int(tokens.max()) >= self.embed.shape[0]):
            raise L.Ma3Error(f"T5 token id outside [0, {self.embed.shape[0]})")
        buf = self._buf.get
        ids = tokens.to(self.device).reshape(-1).contiguous()
        h = ops.embed_rows(self.embed, ids, buf("h", (M, D), f32), T=T)
        u, qk, att = buf("u", (M, D), bf), buf("qk", (M, 2 * inner), bf), buf("att", (M, inner), bf)
        vt, mid = buf("vt", (B, inner, Tp), bf, zero=True), buf("mid", (M, F), bf)
        S, P = buf("S", (H, T, Tp), f32), buf("P", (H, T, Tp), bf)
        bias = self._pos_bias(T, Tp)
        for q in self.layers:
            ops.rmsnorm_modulate(h, q["ln1"], u, eps=c["eps"])
            _lin(u, q["wqk"], qk)
            _vt(q["wv"], None, u, vt, B=B, T=T, D=inner)
            _attention_core(qk, vt, att, S, P, B=B, T=T, H=H, hd=hd, scale=1.0, pos_bias=bias)
            _lin(att, q["wo"], h, accumulate=True)                          # h += attention output
            ops.rmsnorm_modulate(h, q["ln2"], u, eps=c["eps"])
            ops.gemm(u, q["wi"], M=M, N=2 * F, K=D, epi=L.EPI_SWIGLU, act=4, out=mid, out_ld=F)
            _lin(mid, q["wo2"], h, accumulate=True)
        out = torch.empty(M, D, device=self.device, dtype=f32)
        ops.rmsnorm_modulate(h, self.final_ln, out, eps=c["eps"])
        return out.view(B, T, D)


class FrozenCLAPFLANEmbedder:
    """Drop-in for ldm.modules.encoders.modules.FrozenCLAPFLANEmbedder (modules.py:133-191): `encode({'ori_caption': [...],
    'struct_caption': [...]})` -> [B, 2 * max_length, 1024].

    The reference constructor downloads the CLAP checkpoint, the BERT / T5 tokenizers and the T5 weights; offline the
    caller passes what it has: `state_dict` keyed like the reference module (`caption_encoder.*`, `t5_transformer.*`),
    and the two tokenizers (objects with the `transformers` tokenizer call signature).  `weights_path` is loaded with
    the reference's own recipe when given (modules.py:140-145)."""

    def __init__(self, weights_path=None, t5version="google/t5-v1_1-large", freeze=True, device="cuda", max_length=77,
                 state_dict=None, clap_tokenizer=None, t5_tokenizer=None, bert_cfg=None, t5_cfg=None, use_graph=True):
        self.use_graph, self._plans = use_graph, {}
        if state_dict is None:
            if weights_path is None:
                raise L.Ma3Error("FrozenCLAPFLANEmbedder: pass `state_dict` (reference-keyed) or `weights_path`; there is no "
                                 "network to download google/t5-v1_1-large or the CLAP checkpoint from")
            ck = torch.load(weights_path, map_location="cpu")
            ck = ck.get("model", ck)
            state_dict = {k: v for k, v in ck.items() if k.startswith(("caption_encoder.", "t5_transformer."))}
        self.max_length, self.device = max_length, torch.device(device)
        self.clap_tokenizer, self.t5_tokenizer = clap_tokenizer, t5_tokenizer
        self.caption_encoder = BertTextEncoder(state_dict, "caption_encoder.", bert_cfg, device)
        self.t5_transformer = T5TextEncoder(state_dict, "t5_transformer.", t5_cfg, device)

    def to(self, device):
        if torch.device(device) != self.device:
            raise L.Ma3Error("FrozenCLAPFLANEmbedder holds packed device weights; construct it on the target device")
        return self

    def freeze(self):
        return self

    def eval(self):
        return self

    @torch.no_grad()
    def encode_tokens(self, ori_tokens, struct_tokens):
        """int64 [B, T] token ids of the two captions -> fp32 [B, 2T, 1024] (CLAP part first, modules.py:187-191).
        The ~1100 short launches of the two encoders are captured into one CUDA graph per (B, T) and replayed (the
        kernels are microseconds long; issued eagerly the host launch latency is the whole cost)."""
        if ori_tokens.shape != struct_tokens.shape or ori_tokens.dim() != 2:
            raise L.Ma3Error("encode_tokens: ori_tokens and struct_tokens must both be [B, T]")
        for name, t, vocab in (("CLAP", ori_tokens, self.caption_encoder.cfg["vocab"]),
                               ("T5", struct_tokens, self.t5_transformer.embed.shape[0])):
            if int(t.min()) < 0 or int(t.max()) >= vocab:
                raise L.Ma3Error(f"encode_tokens: {name} token id outside [0, {vocab})")
        if not self.use_graph:
            return torch.cat([self.caption_encoder(ori_tokens), self.t5_transformer(struct_tokens)], dim=1)
        key = tuple(ori_tokens.shape)
        st = self._plans.get(key)
        if st is None:
            ids_o = ori_tokens.to(self.device).clone()
            ids_s = struct_tokens.to(self.device).clone()
            run = lambda: torch.cat([self.caption_encoder(ids_o), self.t5_transformer(ids_s)], dim=1)
            run()                                    # eager pass: allocates the scratch buffers, packs the position bias
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = run()
            st = {"ids_o": ids_o, "ids_s": ids_s, "graph": g, "out": out}
            self._plans = {key: st}                  # one plan: the scratch buffers are shared between shapes
        st["ids_o"].copy_(ori_tokens)
        st["ids_s"].copy_(struct_tokens)
        st["graph"].replay()
        return st["out"].clone()

    def _tokenize(self, tok, text, which):
        if tok is None:
            raise L.Ma3Error(f"FrozenCLAPFLANEmbedder.encode needs the {which} tokenizer (no tokenizer files offline); pass it "
                             "to the constructor or call encode_tokens with token ids")
        enc = tok(text, truncation=True, max_length=self.max_length, return_length=True, return_overflowing_tokens=False,
                  padding="max_length", return_tensors="pt")
        return enc["input_ids"].to(self.device)

    def encode(self, text):
        return self.encode_tokens(self._tokenize(self.clap_tokenizer, text["ori_caption"], "CLAP (bert-base-uncased)"),
                                  self._tokenize(self.t5_tokenizer, text["struct_caption"], "T5"))

    __call__ = encode


class Video_Feat_Encoder_NoPosembed:
    """Drop-in for ldm.modules.encoders.modules.Video_Feat_Encoder_NoPosembed (modules.py:16-27; the
    cond_stage_config of configs/video2audio-cfm-cfg-moe.yaml:70-75): x [B, L, origin_dim] -> Linear -> [B, L, embed_dim].
    state_dict keys `embedder.0.weight / .bias`.  The fp32 input is carried through the bf16 tensor cores as a hi / lo
    pair against hi / lo weights (three products, ~16 mantissa bits), like the DiT's own conditioning path."""

    def __init__(self, origin_dim, embed_dim, seq_len=40, state_dict=None, device="cuda"):
        self.origin_dim, self.embed_dim, self.seq_len, self.device = origin_dim, embed_dim, seq_len, torch.device(device)
        if state_dict is None:
            lin = torch.nn.Linear(origin_dim, embed_dim)
            state_dict = {"embedder.0.weight": lin.weight.detach(), "embedder.0.bias": lin.bias.detach()}
        self.load_state_dict(state_dict)

    def load_state_dict(self, sd, strict=True):
        want = {"embedder.0.weight", "embedder.0.bias"}
        if strict and set(sd) != want:
            raise RuntimeError(f"Video_Feat_Encoder_NoPosembed: state_dict keys {sorted(sd)} != {sorted(want)}")
        self.weight = sd["embedder.0.weight"].detach().to(self.device, torch.float32).contiguous()
        self.bias = sd["embedder.0.bias"].detach().to(self.device, torch.float32).contiguous()
        self._w2 = None

    def state_dict(self):
        return {"embedder.0.weight": self.weight, "embedder.0.bias": self.bias}

    @torch.no_grad()
    def forward(self, x):
        L.require_device()
        lead, K = x.shape[:-1], x.shape[-1]
        x2 = x.to(self.device, torch.float32).reshape(-1, K).contiguous()
        if self._w2 is None:
            self._w2 = ops.split_weight(self.weight)
        out = torch.empty(x2.shape[0], self.embed_dim, device=self.device, dtype=torch.float32)
        ops.gemm_split(x2, self._w2, M=x2.shape[0], N=self.embed_dim, K=K, out=out, bias=self.bias)
        return out.view(*lead, self.embed_dim)

    __call__ = encode = forward


class Video_Feat_Encoder_NoPosembed_inpaint(Video_Feat_Encoder_NoPosembed):
    """modules.py:31-39: {'mix_video_feat', 'mix_spec'} -> (Linear(video), spec)."""

    @torch.no_grad()
    def forward(self, x):
        return super().forward(x["mix_video_feat"]), x["mix_spec"]

    __call__ = encode = forward


class FrozenFLANEmbedder:
    """Drop-in for ldm.modules.encoders.modules.FrozenFLANEmbedder (modules.py:54-87): T5 encoder only,
    `encode(list_of_str)` -> last_hidden_state [B, max_length, d_model]."""

    def __init__(self, version="google/flan-t5-large", device="cuda", max_length=77, freeze=True, state_dict=None,
                 tokenizer=None, t5_cfg=None):
        if state_dict is None:
            raise L.Ma3Error("FrozenFLANEmbedder: pass `state_dict` (keys `transformer.*`); there is no network to download "
                             f"{version} from")
        self.max_length, self.device, self.tokenizer = max_length, torch.device(device), tokenizer
        self.transformer = T5TextEncoder(state_dict, "transformer.", t5_cfg, device)

    @torch.no_grad()
    def encode_tokens(self, tokens):
        return self.transformer(tokens)

    def encode(self, text):
        if self.tokenizer is None:
            raise L.Ma3Error("FrozenFLANEmbedder.encode needs the T5 tokenizer (no tokenizer files offline)")
        enc = self.tokenizer(text, truncation=True, max_length=self.max_length, return_length=True,
                             return_overflowing_tokens=False, padding="max_length", return_tensors="pt")
        return self.encode_tokens(enc["input_ids"].to(self.device))

    __call__ = forward = encode
