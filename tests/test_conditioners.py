"""the text conditioners (ldm/modules/encoders/modules.py:133-191) on the GPU kernels.

The reference's conditioner IS `transformers.AutoModel('bert-base-uncased')` + its own `Projection` (clap.py:17-30) and
`transformers.T5EncoderModel('google/t5-v1_1-large')`; no checkpoints exist offline, so the oracle here is those same
`transformers` classes built from their published configs with seeded random weights, run in fp32 -- on the same
weights the drop-in is loaded with (reference-keyed state_dict)."""
import pytest
import torch

transformers = pytest.importorskip("transformers")


def _hf_models(bert_layers, t5_layers, seed=0):
    from transformers import BertConfig, BertModel, T5Config, T5EncoderModel
    torch.manual_seed(seed)
    bert = BertModel(BertConfig(num_hidden_layers=bert_layers), add_pooling_layer=False).eval()
    t5 = T5EncoderModel(T5Config(vocab_size=32128, d_model=1024, d_kv=64, d_ff=2816, num_layers=t5_layers, num_heads=16,
                                 feed_forward_proj="gated-gelu", tie_word_embeddings=False)).eval()
    proj = {"linear1.weight": torch.randn(1024, 768) * 768 ** -0.5, "linear2.weight": torch.randn(1024, 1024) * 1024 ** -0.5,
            "layer_norm.weight": 1 + 0.1 * torch.randn(1024), "layer_norm.bias": 0.1 * torch.randn(1024)}
    return bert, t5, proj


def _state_dict(bert, t5, proj):
    sd = {"caption_encoder.base." + k: v for k, v in bert.state_dict().items()}
    sd.update({"caption_encoder.projection." + k: v for k, v in proj.items()})
    sd.update({"t5_transformer." + k: v for k, v in t5.state_dict().items()})
    return sd


def _projection(x, p):
    """clap.py:25-30 at inference (dropout = identity)."""
    e1 = x @ p["linear1.weight"].T
    e2 = torch.nn.functional.gelu(e1) @ p["linear2.weight"].T
    return torch.nn.functional.layer_norm(e1 + e2, (e1.shape[-1],), p["layer_norm.weight"], p["layer_norm.bias"])


def test_t5_relative_buckets_match_transformers():
    from transformers.models.t5.modeling_t5 import T5Attention
    from ma3_b200.conditioners import t5_relative_buckets
    for T in (1, 7, 77, 200):
        rel = torch.arange(T)[None, :] - torch.arange(T)[:, None]
        want = T5Attention._relative_position_bucket(rel, bidirectional=True, num_buckets=32, max_distance=128)
        assert torch.equal(t5_relative_buckets(T), want)


def test_embedder_needs_weights_and_tokenizers():
    from ma3_b200 import conditioners as Cn, lib as L
    with pytest.raises(L.Ma3Error):
        Cn.FrozenCLAPFLANEmbedder()          # nothing to download offline: weights must be handed in
    with pytest.raises(L.Ma3Error):
        Cn.FrozenFLANEmbedder()


@pytest.mark.gpu
@pytest.mark.parametrize("bert_layers,t5_layers,B", [(2, 2, 3), (12, 24, 2)])
def test_clap_flan_embedder_matches_transformers(bert_layers, t5_layers, B):
    """(2, 2): a shallow stack that isolates the per-layer arithmetic; (12, 24): the published depths of
    bert-base-uncased and t5-v1_1-large.  Tolerance: bf16 GEMM operands with fp32 accumulation and an fp32 residual
    stream against the fp32 oracle -- max error relative to the largest activation <= 2e-2, cosine >= 0.999."""
    from ma3_b200 import conditioners as Cn, lib as L
    bert, t5, proj = _hf_models(bert_layers, t5_layers, seed=3)
    cfgb = dict(Cn.BERT_BASE, layers=bert_layers)
    cfgt = dict(Cn.T5_V11_LARGE, layers=t5_layers)
    emb = Cn.FrozenCLAPFLANEmbedder(state_dict=_state_dict(bert, t5, proj), bert_cfg=cfgb, t5_cfg=cfgt)
    g = torch.Generator().manual_seed(11)
    T = 77
    ori = torch.randint(0, 30522, (B, T), generator=g)
    struct = torch.randint(0, 32128, (B, T), generator=g)
    ori[:, 40:] = 0            # padded tails as the tokenizers produce them ([PAD] = 0 in both vocabularies)
    struct[:, 55:] = 0
    out = emb.encode_tokens(ori, struct)
    assert out.shape == (B, 2 * T, 1024) and out.dtype == torch.float32
    with torch.no_grad():
        bert, t5 = bert.cuda(), t5.cuda()
        pj = {k: v.cuda() for k, v in proj.items()}
        z = _projection(bert(input_ids=ori.cuda()).last_hidden_state, pj)
        z2 = t5(input_ids=struct.cuda()).last_hidden_state
        ref = torch.cat([z, z2], 1)
    for name, a, b in (("clap", out[:, :T], ref[:, :T]), ("t5", out[:, T:], ref[:, T:])):
        err = float((a - b).abs().max() / b.abs().max())
        cos = float(torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0))
        assert err <= 2e-2 and cos >= 0.999, (name, err, cos)
    # the tokenizer-facing entry point refuses politely when no tokenizer was supplied
    with pytest.raises(L.Ma3Error):
        emb.encode({"ori_caption": ["a dog barks"], "struct_caption": ["<dog& barks>"]})
    # ... and runs the same path through any tokenizer-like callable
    class Tok:
        def __init__(self, ids): self.ids = ids
        def __call__(self, text, **kw): return {"input_ids": self.ids[:len(text)]}
    emb.clap_tokenizer, emb.t5_tokenizer = Tok(ori), Tok(struct)
    out2 = emb.encode({"ori_caption": ["x"] * B, "struct_caption": ["y"] * B})
    assert torch.equal(out2, out)


@pytest.mark.gpu
def test_video_feat_encoder():
    """modules.py:16-39 (cond_stage_config of the video/MoE config): Linear 512 -> 768 on 40 video tokens."""
    from ma3_b200 import conditioners as Cn
    torch.manual_seed(5)
    lin = torch.nn.Linear(512, 768)
    sd = {"embedder.0.weight": lin.weight.detach(), "embedder.0.bias": lin.bias.detach()}
    x = torch.randn(3, 40, 512)
    ref = lin(x).detach()
    enc = Cn.Video_Feat_Encoder_NoPosembed(512, 768, seq_len=40, state_dict=sd)
    out = enc(x.cuda())
    assert out.shape == (3, 40, 768)
    assert float((out.cpu() - ref).abs().max() / ref.abs().max()) <= 1e-4      # split-bf16 products: ~fp32
    assert set(enc.state_dict()) == set(sd)
    v, spec = Cn.Video_Feat_Encoder_NoPosembed_inpaint(512, 768, state_dict=sd)({"mix_video_feat": x.cuda(), "mix_spec": x})
    assert torch.equal(v, out) and spec is x
    with pytest.raises(RuntimeError):
        enc.load_state_dict({"embedder.0.weight": lin.weight})


@pytest.mark.gpu
def test_token_ids_to_waveform_end_to_end():
    """The widened path in one piece: token ids -> FrozenCLAPFLANEmbedder (attached as the pipeline's cond_stage_model,
    ddpm_audio.py:343-356) -> sample_cfg -> decode_first_stage -> vocode.  The conditioner's output is the [B, 154, 1024]
    context the txt2audio DiT expects; the waveform equals the one generated from the same embeddings handed in
    precomputed."""
    from ma3_b200 import conditioners as Cn
    from ma3_b200.pipeline import build_random_pipeline
    bert, t5, proj = _hf_models(2, 2, seed=9)
    emb = Cn.FrozenCLAPFLANEmbedder(state_dict=_state_dict(bert, t5, proj), bert_cfg=dict(Cn.BERT_BASE, layers=2),
                                    t5_cfg=dict(Cn.T5_V11_LARGE, layers=2))
    B, T = 2, 77
    g = torch.Generator().manual_seed(2)
    ori, struct = torch.randint(0, 30522, (B, T), generator=g), torch.randint(0, 32128, (B, T), generator=g)

    class Tok:   # stands in for the two tokenizers (no vocabulary files offline): caption i -> row i of a fixed id table
        def __init__(self, ids): self.ids = ids
        def __call__(self, text, **kw): return {"input_ids": torch.zeros_like(self.ids[:len(text)]) if text[0] == "" else self.ids[:len(text)]}
    emb.clap_tokenizer, emb.t5_tokenizer = Tok(ori), Tok(struct)
    from oracle.cases import BIGVGAN_TINY      # the vocoder layout with a narrow first stage: a small, fast generator
    pipe = build_random_pipeline("M", vocoder_h=dict(BIGVGAN_TINY), seed=4)
    pipe.cond_stage_model = emb
    c = pipe.get_learned_conditioning({"ori_caption": ["a dog barks"] * B, "struct_caption": ["<dog& barks& all>"] * B})
    uc = pipe.get_learned_conditioning({"ori_caption": [""] * B, "struct_caption": [""] * B})
    assert c.shape == uc.shape == (B, 2 * T, 1024) and not torch.equal(c, uc)
    x0 = torch.randn(B, 20, 312, generator=torch.Generator().manual_seed(3)).cuda()
    wav = pipe.generate(c, uc, x0, scale=3.0, timesteps=5)
    assert wav.shape == (B, 2 * 312 * 256) and bool(torch.isfinite(wav).all())
    wav2 = pipe.generate(emb.encode_tokens(ori, struct), emb.encode_tokens(torch.zeros_like(ori), torch.zeros_like(struct)),
                         x0, scale=3.0, timesteps=5)
    assert torch.equal(wav, wav2)
    with pytest.raises(TypeError):
        build_random_pipeline("M", vocoder_h=dict(BIGVGAN_TINY), seed=4).get_learned_conditioning(["no conditioner attached"])
