"""Per-kernel parity on a B200: each C-ABI entry point against the oracle (oracle/restated.py) or a plain torch fp32
expression of the same op, on seeded inputs.  Tolerances are stated per test (bf16/fp16 storage => ~2^-8 relative)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import cases as Cs, restated as O  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    from ma3_b200 import lib, ops as _ops
    lib.require_device()
    return _ops


def rel(a, b):
    return O.max_rel_err(a.float().cpu(), b.float().cpu())


def g(seed):
    return Cs.gen(seed)


# ------------------------------------------------------------------------------------------------ tap-GEMM
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (1000, 1152, 1152), (300, 80, 96), (300, 48, 48), (77, 16, 32),
                                   (4992, 3456, 1152)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_gemm_plain(ops, M, N, K, dt):
    a = torch.randn(M, K, generator=g(1)).to(dt).cuda()
    b = (torch.randn(N, K, generator=g(2)) / K ** 0.5).to(dt).cuda()
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.gemm(a, b, M=M, N=N, K=K, out=out)
    assert rel(out, a.float() @ b.float().t()) < 1e-4  # fp32 accumulate of identical 16-bit operands


@pytest.mark.parametrize("M,N,K,act", [(24, 1152, 256, 1), (384, 4000, 1152, 0), (16, 768, 1024, 0), (130, 200, 64, 0)])
def test_gemm_split_near_fp32(ops, M, N, K, act):
    """hi/lo split GEMM (three taps over stacked bf16 halves): ~16 mantissa bits, vs 8 for a plain bf16 GEMM."""
    x = torch.randn(M, K, generator=g(3)).cuda()
    w = (torch.randn(N, K, generator=g(4)) / K ** 0.5)
    bias = torch.randn(N, generator=g(5)).cuda()
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.gemm_split(x, ops.split_weight(w.cuda()), M=M, N=N, K=K, out=out, bias=bias, act=act)
    ref = x.double() @ w.cuda().double().t() + bias.double()
    if act:
        ref = ref * torch.sigmoid(ref)
    assert rel(out, ref) < 5e-5
    plain = torch.empty_like(out)
    ops.gemm(x.bfloat16(), w.bfloat16().cuda(), M=M, N=N, K=K, out=plain, bias=bias, act=act)
    assert rel(plain, ref) > 10 * rel(out, ref)


@pytest.mark.parametrize("M,N,K,tn", [(256, 128, 64, 128), (1000, 1152, 1152, 0), (4992, 3456, 1152, 0), (300, 256, 128, 256),
                                      (2496, 768, 768, 0)])
def test_gemm_cta_pair(ops, M, N, K, tn):
    """cta_group::2 (256-row tiles over CTA pairs) against the same fp32 reference, incl. a ragged last pair."""
    a = torch.randn(M, K, generator=g(1)).bfloat16().cuda()
    b = (torch.randn(N, K, generator=g(2)) / K ** 0.5).bfloat16().cuda()
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    ops.gemm(a, b, M=M, N=N, K=K, out=out, tile_n=tn, cta_group=2)
    assert rel(out, a.float() @ b.float().t()) < 1e-4
    o1 = torch.empty_like(out)
    ops.gemm(a, b, M=M, N=N, K=K, out=o1, tile_n=tn, cta_group=1)
    assert torch.equal(out, o1)   # same accumulation order per output element


def test_gemm_epilogue_store_variants(ops):
    from ma3_b200 import lib as L
    M, N, K = 500, 384, 384
    a = torch.randn(M, K, generator=g(3)).bfloat16().cuda()
    b = (torch.randn(N, K, generator=g(4)) / K ** .5).bfloat16().cuda()
    bias = torch.randn(N, generator=g(5)).cuda()
    res = torch.randn(M, N, generator=g(6)).bfloat16().cuda()
    out = torch.randn(M, N, generator=g(7)).bfloat16().cuda()
    old = out.clone()
    ops.gemm(a, b, M=M, N=N, K=K, out=out, bias=bias, res=res, alpha=0.5, accumulate=True)
    ref = (a.float() @ b.float().t() + bias + res.float()) * 0.5 + old.float()
    assert rel(out, ref) < 8e-3
    # per-row bias, ragged N (scalar tail path), fp16 out
    N2 = 20
    b2 = (torch.randn(N2, K, generator=g(8)) / K ** .5).bfloat16().cuda()
    brow = torch.randn(M, generator=g(9)).cuda()
    o2 = torch.zeros(M, 64, device="cuda", dtype=torch.float16)
    ops.gemm(a, b2, M=M, N=N2, K=K, out=o2, out_ld=64, bias=brow, bias_per_row=True)
    assert rel(o2[:, :N2], a.float() @ b2.float().t() + brow[:, None]) < 2e-3
    assert bool((o2[:, N2:] == 0).all())
    # activation epilogues
    for act, fn in ((1, torch.nn.functional.silu), (2, torch.nn.functional.gelu), (3, torch.tanh)):
        o3 = torch.empty(M, N, device="cuda", dtype=torch.float32)
        ops.gemm(a, b, M=M, N=N, K=K, out=o3, bias=bias, act=act)
        assert rel(o3, fn(a.float() @ b.float().t() + bias)) < 1e-3


@pytest.mark.parametrize("k,dil", [(3, 1), (3, 5), (7, 3), (11, 1), (5, 1), (1, 1)])
def test_gemm_as_conv1d(ops, k, dil):
    B, T, Ci, Co = 3, 300, 128, 192
    pad = (k * dil - dil) // 2
    x = torch.randn(B, T, Ci, generator=g(10)).bfloat16().cuda()
    w = (torch.randn(Co, Ci, k, generator=g(11)) / (Ci * k) ** .5).bfloat16().cuda()
    wp = w.permute(2, 0, 1).contiguous().view(k * Co, Ci)
    bias = torch.randn(Co, generator=g(12)).cuda()
    y = torch.empty(B, T, Co, device="cuda", dtype=torch.bfloat16)
    ops.gemm(x, wp, M=T, N=Co, K=Ci, batch=B, a_rows=T, a_batch_stride=T * Ci, b_rows=k * Co,
             taps=[(j * dil - pad, j * Co) for j in range(k)], out=y, out_batch_stride=T * Co, bias=bias)
    ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.float(), bias, padding=pad, dilation=dil)
    assert rel(y, ref.transpose(1, 2)) < 8e-3


@pytest.mark.parametrize("C,k,dil,T", [(48, 7, 5, 1000), (32, 11, 3, 300), (96, 7, 1, 513), (16, 3, 1, 129), (64, 7, 3, 2048),
                                       (96, 11, 3, 700), (96, 11, 5, 300), (128, 3, 1, 257), (96, 3, 1, 40000), (112, 7, 1, 400)])
def test_gemm_narrow_conv_epilogue(ops, C, k, dil, T):
    """Narrow conv layers of the vocoder (C <= 128): ragged 64-wide k-chunks (TMA zero fill beyond K), the row-direct STORE
    epilogue with prefetched residual, scaling and accumulation (fp16), and for 65..128 channels the staged-once kernel
    with two k-chunks and the output columns split over two CTAs (112 channels: not splittable, generic tap-GEMM)."""
    B = 2
    pad = (k * dil - dil) // 2
    x = (torch.randn(B, T, C, generator=g(60))).half().cuda()
    w = (torch.randn(C, C, k, generator=g(61)) / (C * k) ** .5).half().cuda()
    wp = w.permute(2, 0, 1).contiguous().view(k * C, C)
    bias = torch.randn(C, generator=g(62)).cuda()
    res = torch.randn(B, T, C, generator=g(63)).half().cuda()
    old = torch.randn(B, T, C, generator=g(64)).half().cuda()
    taps = [(j * dil - pad, j * C) for j in range(k)]
    ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.float(), bias, padding=pad, dilation=dil).transpose(1, 2)
    y = torch.empty(B, T, C, device="cuda", dtype=torch.float16)
    ops.gemm(x, wp, M=T, N=C, K=C, batch=B, a_rows=T, a_batch_stride=T * C, b_rows=k * C, taps=taps, out=y,
             out_batch_stride=T * C, bias=bias)
    assert rel(y, ref) < 2e-3
    y2 = old.clone()
    ops.gemm(x, wp, M=T, N=C, K=C, batch=B, a_rows=T, a_batch_stride=T * C, b_rows=k * C, taps=taps, out=y2,
             out_batch_stride=T * C, bias=bias, res=res, res_batch_stride=T * C, alpha=1 / 3, accumulate=True)
    assert rel(y2, (ref + res.float()) / 3 + old.float()) < 2e-3
    o32 = torch.empty(B, T, 1, device="cuda", dtype=torch.float32)   # conv_post: one output channel, tanh, fp32
    w1 = wp.view(k, C, C)[:, :1].contiguous().view(k, C)
    ops.gemm(x, w1, M=T, N=1, K=C, batch=B, a_rows=T, a_batch_stride=T * C, b_rows=k, taps=[(j * dil - pad, j) for j in range(k)],
             out=o32, out_ld=1, out_batch_stride=T, bias=bias[:1].contiguous(), act=3)
    assert rel(o32, torch.tanh(ref[..., :1])) < 1e-3


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_gemm_narrow_conv_lean_variants(ops, dt):
    """Lean epilogue of the narrow-conv kernel: no bias, bf16 as well as fp16, output rows written with a row stride
    (ConvTranspose phases: out_row_mul / out_row_off), ragged last tile; the rows in between must stay untouched."""
    B, T, C, k = 2, 333, 48, 3
    x = torch.randn(B, T, C, generator=g(70)).to(dt).cuda()
    w = (torch.randn(C, C, k, generator=g(71)) / (C * k) ** .5).to(dt).cuda()
    wp = w.permute(2, 0, 1).contiguous().view(k * C, C)
    taps = [(j - 1, j * C) for j in range(k)]
    ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.float(), None, padding=1).transpose(1, 2)
    tol = 2e-3 if dt == torch.float16 else 1e-2
    y = torch.empty(B, T, C, device="cuda", dtype=dt)
    ops.gemm(x, wp, M=T, N=C, K=C, batch=B, a_rows=T, a_batch_stride=T * C, b_rows=k * C, taps=taps, out=y,
             out_batch_stride=T * C)
    assert rel(y, ref) < tol
    y2 = torch.full((B, 2 * T, C), 7.0, device="cuda", dtype=dt)
    ops.gemm(x, wp, M=T, N=C, K=C, batch=B, a_rows=T, a_batch_stride=T * C, b_rows=k * C, taps=taps, out=y2,
             out_batch_stride=2 * T * C, out_row_mul=2, out_row_off=1, alpha=0.5)
    assert rel(y2[:, 1::2], 0.5 * ref) < tol
    assert bool((y2[:, 0::2] == 7.0).all())


def test_gemm_gate_residual_and_swiglu(ops):
    from ma3_b200 import lib as L
    M, N, K, T = 624, 768, 768, 312
    a = torch.randn(M, K, generator=g(13)).bfloat16().cuda()
    b = (torch.randn(N, K, generator=g(14)) / K ** .5).bfloat16().cuda()
    h = torch.randn(M, N, generator=g(15)).cuda()
    h0 = h.clone()
    gate = torch.randn(2, N, generator=g(16)).cuda()
    ops.gemm(a, b, M=M, N=N, K=K, epi=L.EPI_GATE_RES, out=h, gate=gate, rows_per_sample=T)
    assert rel(h, h0 + gate.repeat_interleave(T, 0) * (a.float() @ b.float().t())) < 1e-4
    F = 2048
    w1 = (torch.randn(F, K, generator=g(17)) / K ** .5).bfloat16().cuda()
    w3 = (torch.randn(F, K, generator=g(18)) / K ** .5).bfloat16().cuda()
    w13 = torch.stack([w1, w3], 1).reshape(2 * F, K).contiguous()
    o = torch.empty(M, F, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w13, M=M, N=2 * F, K=K, epi=L.EPI_SWIGLU, out=o, out_ld=F)
    ref = torch.nn.functional.silu(a.float() @ w1.float().t()) * (a.float() @ w3.float().t())
    assert rel(o, ref) < 8e-3


@pytest.mark.parametrize("M,T,E,band,F", [(512, 256, 4, 192, 2048), (300, 100, 3, 64, 256)])
def test_gemm_batched_band_experts(ops, M, T, E, band, F):
    """The frequency experts of the MoE feed-forward as two batched launches (flag_large_dit_moe.py:516-538): z selects a
    band of `band` columns of the input, of the fp32 residual stream and of the gate (a/out/gate_batch_stride), its own
    stacked weights and its own SwiGLU buffer.  Against the per-band fp32 expression, and the gated-GELU variant (act=4)
    of the same epilogue against torch's tanh-GELU."""
    from ma3_b200 import lib as L
    D = E * band
    y = torch.randn(M, D, generator=g(50)).bfloat16().cuda()
    w13 = (torch.randn(E, 2 * F, band, generator=g(51)) / band ** .5).bfloat16().cuda()   # rows interleaved: w1_0, w3_0, w1_1, ...
    w2 = (torch.randn(E, band, F, generator=g(52)) / F ** .5).bfloat16().cuda()
    gate = torch.randn(M // T, 3 * D, generator=g(53)).cuda()[:, D:2 * D]                 # a strided slice, like the adaLN table
    h0 = torch.randn(M, D, generator=g(54)).cuda()
    for act, fn in ((0, torch.nn.functional.silu), (4, lambda t: torch.nn.functional.gelu(t, approximate="tanh"))):
        mid = torch.empty(E, M, F, device="cuda", dtype=torch.bfloat16)
        ops.gemm(y, w13, M=M, N=2 * F, K=band, batch=E, a_ld=D, a_batch_stride=band, b_rows=2 * F, b_batch_stride=2 * F * band,
                 epi=L.EPI_SWIGLU, act=act, out=mid, out_ld=F, out_batch_stride=M * F)
        h = h0.clone()
        ops.gemm(mid, w2, M=M, N=band, K=F, batch=E, a_batch_stride=M * F, b_rows=band, b_batch_stride=band * F,
                 epi=L.EPI_GATE_RES, out=h, out_ld=D, out_batch_stride=band, gate=gate, gate_batch_stride=band, rows_per_sample=T)
        ref = h0.clone()
        for j in range(E):
            x = y[:, j * band:(j + 1) * band].float()
            w = w13[j].float()
            m_ref = (fn(x @ w[0::2].t()) * (x @ w[1::2].t()))
            assert rel(mid[j], m_ref) < 8e-3
            ref[:, j * band:(j + 1) * band] += gate[:, j * band:(j + 1) * band].repeat_interleave(T, 0) * \
                (mid[j].float() @ w2[j].float().t())
        assert rel(h, ref) < 1e-4


@pytest.mark.parametrize("M,N,K,cg", [(4992, 1152, 1152, 1), (4992, 1152, 3072, 2), (624, 768, 2048, 1), (1000, 192, 512, 2)])
def test_gemm_gate_residual_stream_k(ops, M, N, K, cg):
    """Stream-K work split (equal k-iteration ranges per SM, partial products added by separate reductions) against
    the same fp32 expression and against the whole-tile schedule (differences only from fp32 summation order)."""
    from ma3_b200 import lib as L
    T = 8
    a = torch.randn(M, K, generator=g(40)).bfloat16().cuda()
    b = (torch.randn(N, K, generator=g(41)) / K ** .5).bfloat16().cuda()
    h0 = torch.randn(M, N, generator=g(42)).cuda()
    gate = torch.randn(M // T, N, generator=g(43)).cuda()
    ref = h0 + gate.repeat_interleave(T, 0) * (a.float() @ b.float().t())
    outs = []
    for sk in (1, -1):
        h = h0.clone()
        ops.gemm(a, b, M=M, N=N, K=K, epi=L.EPI_GATE_RES, out=h, gate=gate, rows_per_sample=T, cta_group=cg, stream_k=sk)
        assert rel(h, ref) < 1e-4
        outs.append(h)
    assert rel(outs[0], outs[1]) < 1e-5


@pytest.mark.parametrize("M,N,K,T,tn,cg,ld", [(200, 176, 64, 40, 48, 1, 176), (333, 100, 96, 37, 0, 1, 104),
                                              (520, 304, 128, 65, 80, 1, 320), (777, 224, 256, 111, 96, 2, 224)])
def test_gemm_gate_residual_ragged(ops, M, N, K, T, tn, cg, ld):
    """The tensor reduce-add epilogue on ragged shapes: row and column tails (clipped by the tensor map), 16-wide last
    chunks (tile_n % 32 == 16), a padded output pitch, several samples per 32-row chunk; untouched columns stay put."""
    from ma3_b200 import lib as L
    ns = (M + T - 1) // T
    Mp = ns * T
    a = torch.randn(Mp, K, generator=g(70)).bfloat16().cuda()
    b = (torch.randn(N, K, generator=g(71)) / K ** .5).bfloat16().cuda()
    h0 = torch.randn(Mp, ld, generator=g(72)).cuda()
    gate = torch.randn(ns, N, generator=g(73)).cuda()
    h = h0.clone()
    ops.gemm(a, b, M=M, N=N, K=K, epi=L.EPI_GATE_RES, out=h, out_ld=ld, gate=gate, rows_per_sample=T, tile_n=tn, cta_group=cg)
    ref = h0.clone()
    ref[:M, :N] += (gate.repeat_interleave(T, 0) * (a.float() @ b.float().t()))[:M]
    assert rel(h[:M, :N], ref[:M, :N]) < 1e-4
    assert torch.equal(h[M:], h0[M:]) and torch.equal(h[:, N:], h0[:, N:])
    h2 = h0.clone()
    ops.gemm(a, b, M=M, N=N, K=K, epi=L.EPI_GATE_RES, out=h2, out_ld=ld, gate=gate, rows_per_sample=T, tile_n=tn, cta_group=cg)
    assert torch.equal(h, h2)


# ------------------------------------------------------------------------------------------------ QKV+RoPE, attention
def _qkv_attention_case(ops, D, H, T, L, Ns, seed):
    """QKV GEMM with RoPE scatter followed by the fused self+cross attention, against oracle.attention pieces."""
    from ma3_b200 import lib as L_
    hd = D // H
    hdp = 64 if hd <= 64 else 128
    Tp, Lp = (T + 7) // 8 * 8, (L + 7) // 8 * 8
    gg = g(seed)
    x = torch.randn(Ns, T, D, generator=gg)
    y = torch.randn(Ns, L, D, generator=gg)
    sd = {"wq.weight": torch.randn(D, D, generator=gg) / D ** .5, "wk.weight": torch.randn(D, D, generator=gg) / D ** .5,
          "wv.weight": torch.randn(D, D, generator=gg) / D ** .5, "wk_y.weight": torch.randn(D, D, generator=gg) / D ** .5,
          "wv_y.weight": torch.randn(D, D, generator=gg) / D ** .5, "gate": torch.randn(H, generator=gg),
          "wo.weight": torch.eye(D)}
    # oracle on bf16-rounded inputs/weights (the kernels see bf16 operands)
    r = lambda t: t.bfloat16().float()
    sdr = {k: (r(v) if k != "gate" else v) for k, v in sd.items()}
    cos, sin = O.rope_table(hd, T)
    ref = O.attention(sdr, "", r(x), r(y), cos, sin, H)  # [Ns, T, D] (wo = identity)

    dev = "cuda"
    xb = x.bfloat16().to(dev).view(Ns * T, D)
    wqkv = torch.cat([sd["wq.weight"], sd["wk.weight"], sd["wv.weight"]]).bfloat16().to(dev)
    rope = torch.stack([cos, sin], -1).contiguous().to(dev)
    q = torch.zeros(Ns, H, T, hdp, device=dev, dtype=torch.bfloat16)
    k = torch.zeros_like(q)
    vt = ops.alloc_vt(Ns, H, hd=hd, hdp=hdp, tokens_pad=Tp, device=dev)
    ops.gemm(xb, wqkv, M=Ns * T, N=3 * D, K=D, epi=L_.EPI_QKV_ROPE, q_out=q, k_out=k, vt_out=vt, rope=rope,
             model_dim=D, head_dim=hd, head_dim_pad=hdp, tokens=T, tokens_pad=Tp,
             q_scale=math.log2(math.e) / math.sqrt(hd))
    yb = y.bfloat16().to(dev).view(Ns * L, D)
    wkv = torch.cat([sd["wk_y.weight"], sd["wv_y.weight"]]).bfloat16().to(dev)
    ky = torch.zeros(Ns, H, L, hdp, device=dev, dtype=torch.bfloat16)
    vyt = ops.alloc_vt(Ns, H, hd=hd, hdp=hdp, tokens_pad=Lp, device=dev)
    ops.gemm(yb, wkv, M=Ns * L, N=2 * D, K=D, epi=L_.EPI_QKV_ROPE, q_out=ky, k_out=ky, vt_out=vyt, rope=None,
             model_dim=D, head_dim=hd, head_dim_pad=hdp, tokens=L, tokens_pad=Lp, first_section=1)
    # layout contract of the epilogue: rows [0, hd) of V^T hold v, the ones row and the zero pads are untouched
    full = (xb.float() @ wqkv.float().t()).view(Ns, T, 3, H, hd)
    assert rel(vt[:, :, :hd, :T], full[:, :, 2].permute(0, 2, 3, 1)) < 1e-2
    if hd < hdp:
        assert bool((vt[:, :, hd] == 1).all()) and bool((vt[:, :, hd + 1:] == 0).all())
    assert bool((q[..., hd:] == 0).all()) and bool((k[..., hd:] == 0).all())
    out = torch.empty(Ns, T, D, device=dev, dtype=torch.bfloat16)
    ops.attention(q, k, vt, ky, vyt, sd["gate"].to(dev), out, hd=hd)
    torch.cuda.synchronize()
    return rel(out, ref)


@pytest.mark.parametrize("D,H,T,L,Ns", [(768, 32, 312, 154, 2), (1152, 16, 312, 154, 2), (768, 16, 936, 154, 1),
                                        (768, 32, 256, 40, 2), (256, 16, 100, 7, 3), (128, 2, 130, 77, 1)])
def test_qkv_rope_attention(ops, D, H, T, L, Ns):
    # bf16 q/k/v/p storage: 2^-8 relative per rounding, a handful of roundings
    assert _qkv_attention_case(ops, D, H, T, L, Ns, seed=20) < 2e-2


@pytest.mark.parametrize("D,H,T,L,Ns", [(1152, 16, 312, 154, 2), (768, 16, 936, 154, 1), (768, 32, 256, 40, 2),
                                        (128, 2, 130, 77, 1), (1536, 16, 312, 154, 1)])
def test_qkv_rope_attention_v3(ops, D, H, T, L, Ns):
    """The opt-in third-generation attention kernel (query tiles of a head as contexts of one CTA, one softmax thread per
    row, per-context MMA warps) against the same oracle; forced through ma3_debug_set_attn_version."""
    from ma3_b200 import lib as L_
    lib = L_.require_device()
    lib.ma3_debug_set_attn_version(3)
    try:
        assert _qkv_attention_case(ops, D, H, T, L, Ns, seed=22) < 2e-2
    finally:
        lib.ma3_debug_set_attn_version(0)


@pytest.mark.parametrize("Ns,H,hd,hdp,reps", [(2, 4, 72, 128, 1), (2, 32, 24, 64, 25)])
def test_attention_lazy_rescale_large_logits(ops, Ns, H, hd, hdp, reps):
    """Logits with a large, growing spread across KV tiles force the in-TMEM rescale of O (running maximum raised by
    more than 2^8) on most tiles; compared with an fp32 softmax of the same bf16 operands.  The second case fills the
    GPU with two CTAs per SM and repeats the launch back to back: every run must be bit-identical (this is the shape
    that exposed an under-declared named-barrier count and out-of-order mbarrier phase waits as intermittent wrong rows
    in the second CTA of an SM)."""
    T, L = 312, 154
    dev = "cuda"
    gg = g(21)
    q = torch.zeros(Ns, H, T, hdp); k = torch.zeros(Ns, H, T, hdp); ky = torch.zeros(Ns, H, L, hdp)
    q[..., :hd] = torch.randn(Ns, H, T, hd, generator=gg) * 0.8
    k[..., :hd] = torch.randn(Ns, H, T, hd, generator=gg) * torch.linspace(0.5, 4.0, T)[None, None, :, None]
    ky[..., :hd] = torch.randn(Ns, H, L, hd, generator=gg) * torch.linspace(0.5, 3.0, L)[None, None, :, None]
    v = torch.randn(Ns, H, T, hd, generator=gg); vy = torch.randn(Ns, H, L, hd, generator=gg)
    gate = torch.randn(H, generator=gg)
    b = lambda t: t.bfloat16()
    Tp, Lp = T, (L + 7) // 8 * 8
    vt = ops.alloc_vt(Ns, H, hd=hd, hdp=hdp, tokens_pad=Tp, device=dev); vt[:, :, :hd, :T] = b(v).transpose(2, 3).to(dev)
    vyt = ops.alloc_vt(Ns, H, hd=hd, hdp=hdp, tokens_pad=Lp, device=dev); vyt[:, :, :hd, :L] = b(vy).transpose(2, 3).to(dev)
    qd, kd, kyd, gd = b(q).to(dev), b(k).to(dev), b(ky).to(dev), gate.to(dev)
    outs = [torch.empty(Ns, T, H * hd, device=dev, dtype=torch.bfloat16) for _ in range(reps)]
    for o in outs:
        ops.attention(qd, kd, vt, kyd, vyt, gd, o, hd=hd)
    torch.cuda.synchronize()
    out = outs[0]
    assert all(torch.equal(o, out) for o in outs[1:])
    ln2 = math.log(2.0)
    qf, kf, kyf = b(q).float()[..., :hd], b(k).float()[..., :hd], b(ky).float()[..., :hd]
    ps = torch.softmax(qf @ kf.transpose(2, 3) * ln2, -1) @ b(v).float()
    pc = torch.softmax(qf @ kyf.transpose(2, 3) * ln2, -1) @ b(vy).float()
    ref = (ps + torch.tanh(gate)[None, :, None, None] * pc).permute(0, 2, 1, 3).reshape(Ns, T, H * hd)
    assert rel(out, ref) < 2e-2


# ------------------------------------------------------------------------------------------------ elementwise
def test_rmsnorm_modulate(ops):
    M, D, T = 624, 1152, 312
    x = torch.randn(M, D, generator=g(30)) * 3
    w = torch.randn(D, generator=g(31))
    mod = torch.randn(2, 6 * D, generator=g(32)) * 0.3
    out = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    ops.rmsnorm_modulate(x.cuda(), w.cuda(), out, mod=mod.cuda(), shift_off=3 * D, scale_off=4 * D, rows_per_sample=T)
    sc = mod[:, 4 * D:5 * D].repeat_interleave(T, 0)
    sh = mod[:, 3 * D:4 * D].repeat_interleave(T, 0)
    assert rel(out, O.rmsnorm(x, w) * (1 + sc) + sh) < 5e-3
    out32 = torch.empty(M, D, device="cuda", dtype=torch.float32)
    ops.rmsnorm_modulate(x.cuda(), None, out32)
    assert rel(out32, O.rmsnorm(x, torch.ones(D))) < 1e-5


@pytest.mark.parametrize("D,T", [(768, 50), (1152, 51), (1536, 8), (256, 33)])   # fast row-pair kernels + the generic one (odd N*T)
def test_final_layer_and_cfg_euler(ops, D, T):
    B, Cc = 3, 20
    N = 2 * B
    h = torch.randn(N * T, D, generator=g(33)) * 2 + 0.3
    mod = torch.randn(N, 2 * D, generator=g(34)) * 0.3
    W = torch.randn(Cc, D, generator=g(35)) / D ** .5
    bias = torch.randn(Cc, generator=g(36))
    x = torch.randn(B, Cc, T, generator=g(37))
    hn = torch.nn.functional.layer_norm(h.view(N, T, D), (D,), None, None, 1e-6)
    v = (hn * (1 + mod[:, None, D:]) + mod[:, None, :D]) @ W.t() + bias
    v = v.transpose(1, 2)  # [N, C, T]
    vout = torch.empty(N, Cc, T, device="cuda")
    ops.final_layer(h.cuda(), mod.cuda(), 0, D, W.cuda(), bias.cuda(), N, T, vout)
    assert rel(vout, v) < 1e-4
    vg = v[:B] + 3.0 * (v[B:] - v[:B])
    xo = torch.empty(B, Cc, T, device="cuda")
    vgo = torch.empty(B, Cc, T, device="cuda")
    ops.final_layer_cfg_euler(h.cuda(), mod.cuda(), 0, D, W.cuda(), bias.cuda(), N, T, 3.0, 1 / 24, x.cuda(), xo, vgo)
    assert rel(vgo, vg) < 1e-4 and rel(xo, x + vg / 24) < 1e-4
    x2 = torch.empty(B, Cc, T, device="cuda")
    ops.cfg_euler_update(vout, x.cuda(), x2, 1 / 24, 3.0, cfg=True)
    assert rel(x2, x + vg / 24) < 1e-4


def test_proj_in_timestep_pool(ops):
    N, Cc, T, D = 4, 20, 37, 192
    x = torch.randn(2, Cc, T, generator=g(38))
    W = torch.randn(D, Cc, generator=g(39))
    b = torch.randn(D, generator=g(40))
    h = torch.empty(N * T, D, device="cuda")
    ops.proj_in(x.cuda(), W.t().contiguous().cuda(), b.cuda(), h, N)
    ref = (x.transpose(1, 2) @ W.t() + b).repeat(2, 1, 1).view(N * T, D)
    assert rel(h, ref) < 1e-5
    t = torch.tensor([0, 41, 500, 958, 999])
    e = torch.empty(5, 256, device="cuda", dtype=torch.float32)
    ops.timestep_embed(t.cuda(), e)
    assert (e.cpu() - O.timestep_embedding(t)).abs().max() < 2e-4
    ctx = torch.randn(3, 11, 96, generator=g(41))
    w = torch.randn(96, generator=g(42))
    bb = torch.randn(96, generator=g(43))
    o = torch.empty(3, 96, device="cuda", dtype=torch.bfloat16)
    ops.pool_layernorm(ctx.cuda(), w.cuda(), bb.cuda(), o)
    assert rel(o, torch.nn.functional.layer_norm(ctx.mean(1), (96,), w, bb, 1e-5)) < 8e-3


def test_groupnorm_softmax_layout(ops):
    B, T, Cc = 2, 60, 384
    x = torch.randn(B, Cc, T, generator=g(44)) * 2 + 0.5
    w = torch.randn(Cc, generator=g(45))
    b = torch.randn(Cc, generator=g(46))
    ref = torch.nn.functional.group_norm(x, 32, w, b, 1e-6)
    ref = ref * torch.sigmoid(ref)
    xin = x.transpose(1, 2).contiguous().cuda()
    o = torch.empty(B, T, Cc, device="cuda", dtype=torch.bfloat16)
    ops.groupnorm_swish(xin, w.cuda(), b.cuda(), o)
    assert rel(o.transpose(1, 2), ref) < 8e-3
    S = torch.randn(2, 50, 50, generator=g(47)) * 4
    P = torch.full((2, 50, 64), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.softmax_rows(S.cuda(), P, 50, 0.25)
    assert rel(P[..., :50], torch.softmax(S * 0.25, -1)) < 8e-3 and bool((P[..., 50:] == 0).all())
    z = torch.randn(2, 20, 24, generator=g(48))
    zz = torch.empty(2, 24, 64, device="cuda", dtype=torch.bfloat16)
    ops.nct_to_ntc(z.cuda(), zz, 0.5)
    assert rel(zz[..., :20], 0.5 * z.transpose(1, 2)) < 5e-3 and bool((zz[..., 20:] == 0).all())
    back = torch.empty(2, 20, 24, device="cuda")
    ops.ntc_to_nct(zz, back)
    assert rel(back, 0.5 * z) < 5e-3
    up = torch.empty(2, 48, 64, device="cuda", dtype=torch.bfloat16)
    ops.upsample_nearest2(zz, up)
    assert torch.equal(up, zz.repeat_interleave(2, dim=1))


# ------------------------------------------------------------------------------------------------ Activation1d
@pytest.mark.parametrize("B,Cc,T", [(2, 32, 50), (1, 64, 3), (2, 96, 513), (1, 48, 1000), (1, 1536, 130), (1, 32, 16),
                                    (1, 16, 17), (2, 192, 2496)])
@pytest.mark.parametrize("dt_in", [torch.float32, torch.float16])
def test_act1d(ops, B, Cc, T, dt_in):
    gg = g(50)
    x = torch.randn(B, Cc, T, generator=gg) * 1.5
    al = torch.randn(Cc, generator=gg) * 0.3
    be = torch.randn(Cc, generator=gg) * 0.3
    xin = x.to(dt_in)
    sd = {"a.act.alpha": al, "a.act.beta": be}
    ref = O.activation1d(xin.float(), sd, "a", dict(activation="snakebeta", snake_logscale=True))
    out = torch.empty(B, T, Cc, device="cuda", dtype=torch.float16)
    ops.act1d(xin.transpose(1, 2).contiguous().cuda(), out, al.cuda(), be.cuda(), logscale=True)
    assert rel(out.transpose(1, 2), ref) < 2e-3  # fp16 output rounding 2^-11, fast sin
    # plain Snake, linear-scale alpha
    ref2 = O.activation1d(xin.float(), {"a.act.alpha": al.abs() + 0.5}, "a", dict(activation="snake", snake_logscale=False))
    ops.act1d(xin.transpose(1, 2).contiguous().cuda(), out, (al.abs() + 0.5).cuda(), None, logscale=False)
    assert rel(out.transpose(1, 2), ref2) < 2e-3


@pytest.mark.parametrize("B,Cc,T", [(2, 64, 9984), (1, 384, 1256), (2, 96, 520), (1, 48, 2048), (2, 16, 1544), (1, 32, 8),
                                    (3, 64, 128), (1, 128, 136), (2, 32, 256), (1, 48, 512), (1, 768, 2496)])
def test_act1d_tma_staged(ops, B, Cc, T):
    """fp16 -> fp16, T % 8 == 0: the TMA-staged tensor-core kernel (channel tiles of 64 / 32 / 16, sequence ends inside,
    at and beyond segment and tile boundaries) against the oracle and against the first-generation kernel."""
    from ma3_b200 import lib as L_
    lib = L_.require_device()
    gg = g(51)
    x = (torch.randn(B, Cc, T, generator=gg) * 1.5).half()
    al = torch.randn(Cc, generator=gg) * 0.3
    be = torch.randn(Cc, generator=gg) * 0.3
    ref = O.activation1d(x.float(), {"a.act.alpha": al, "a.act.beta": be}, "a",
                         dict(activation="snakebeta", snake_logscale=True))
    xin = x.transpose(1, 2).contiguous().cuda()
    out = torch.full((B, T, Cc), float("nan"), device="cuda", dtype=torch.float16)
    ops.act1d(xin, out, al.cuda(), be.cuda(), logscale=True)
    assert rel(out.transpose(1, 2), ref) < 2e-3
    again = torch.full_like(out, float("nan"))
    ops.act1d(xin, again, al.cuda(), be.cuda(), logscale=True)
    assert torch.equal(out, again)
    lib.ma3_debug_set_act_version(1)
    try:
        old = torch.empty_like(out)
        ops.act1d(xin, old, al.cuda(), be.cuda(), logscale=True)
    finally:
        lib.ma3_debug_set_act_version(0)
    assert rel(out.float(), old.float()) < 1e-3


def test_act1d_golden(ops, golden):
    x, al, be = Cs.act_inputs()
    xp = torch.zeros(2, 50, 32)
    xp[..., :24] = x.transpose(1, 2)
    alp, bep = torch.zeros(32), torch.zeros(32)
    alp[:24], bep[:24] = al, be
    out = torch.empty(2, 50, 32, device="cuda", dtype=torch.float32)
    ops.act1d(xp.cuda(), out, alp.cuda(), bep.cuda())
    assert rel(out[..., :24].transpose(1, 2), golden["act1d"]) < 1e-4
    assert bool((out[..., 24:] == 0).all())  # padded channels stay exactly zero


def test_errors(ops):
    from ma3_b200 import lib as L
    a = torch.zeros(16, 24, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(L.Ma3Error):
        ops.gemm(a, a, M=16, N=16, K=24, out=torch.empty(16, 16, device="cuda"))  # K not a multiple of 16
    with pytest.raises(L.Ma3Error):
        ops.act1d(torch.zeros(1, 8, 24, device="cuda"), torch.zeros(1, 8, 24, device="cuda", dtype=torch.float16),
                  torch.zeros(24, device="cuda"), None)  # C not a multiple of 16


# ------------------------------------------------------------------------------------------------ fused RMSNorm
@pytest.mark.parametrize("samples,T,D,K", [(3, 104, 1152, 256), (2, 312, 768, 2048), (2, 40, 64, 64), (16, 312, 1152, 1152)])
def test_fused_rmsnorm_producer(ops, samples, T, D, K):
    """GATE_RES with norm_out: h_new = h + gate*acc (plain load-add-store), g = bf16(h_new * wn_s), per-chunk sums of
    squares of h_new -- against the fp32 expression on the same 16-bit operands."""
    from ma3_b200 import lib as L
    M = samples * T
    a = torch.randn(M, K, generator=g(1)).bfloat16().cuda()
    b = (torch.randn(D, K, generator=g(2)) / K ** 0.5).bfloat16().cuda()
    h = torch.randn(M, D, generator=g(3)).cuda()
    mod = torch.randn(samples, 2 * D + 8, generator=g(4)).cuda()           # [gate | wn | pad]
    gate, wn = mod[:, :D], mod[:, D:2 * D]
    sc = (D // 32 + 3) // 4 * 4
    ss = torch.zeros(M, sc, device="cuda")
    gout = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    acc = a.float() @ b.float().t()
    rows = torch.arange(M, device="cuda") // T
    h_ref = h + gate[rows] * acc
    hh = h.clone()
    ops.gemm(a, b, M=M, N=D, K=K, epi=L.EPI_GATE_RES, out=hh, gate=gate, rows_per_sample=T, norm_out=gout, norm_w=wn,
             ss_out=ss)
    assert rel(hh, h_ref) < 1e-5
    assert rel(gout, (h_ref * wn[rows])) < 5e-3                           # bf16 rounding of the operand
    assert rel(ss.sum(1), h_ref.pow(2).sum(1)) < 1e-5
    assert torch.equal(ss[:, D // 32:], torch.zeros_like(ss[:, D // 32:]))
    # bit-reproducible (no atomics) and identical to the reduction path's h
    h2, h3 = h.clone(), h.clone()
    ops.gemm(a, b, M=M, N=D, K=K, epi=L.EPI_GATE_RES, out=h2, gate=gate, rows_per_sample=T, norm_out=gout, norm_w=wn,
             ss_out=ss)
    assert torch.equal(h2, hh)
    ops.gemm(a, b, M=M, N=D, K=K, epi=L.EPI_GATE_RES, out=h3, gate=gate, rows_per_sample=T)
    assert rel(h3, hh) < 1e-6


@pytest.mark.parametrize("samples,T,D,F", [(3, 104, 1152, 512), (2, 312, 768, 2048), (2, 40, 64, 256)])
def test_fused_rmsnorm_consumer_swiglu(ops, samples, T, D, F):
    """SWIGLU with row_ss / col_bias2: acc <- acc * rstd[m] + (shift_s W^T)[n] before the gate."""
    from ma3_b200 import lib as L
    M = samples * T
    gop = torch.randn(M, D, generator=g(5)).bfloat16().cuda()            # h * wn (un-normalised operand)
    w13 = (torch.randn(2 * F, D, generator=g(6)) / D ** 0.5).bfloat16().cuda()
    sc = (D // 32 + 3) // 4 * 4
    ss = torch.zeros(M, sc, device="cuda")
    ss[:, :D // 32] = torch.rand(M, D // 32, generator=g(7)).cuda() * 40
    b2 = torch.randn(samples, 2 * F, generator=g(8)).cuda()
    out = torch.empty(M, F, device="cuda", dtype=torch.bfloat16)
    ops.gemm(gop, w13, M=M, N=2 * F, K=D, epi=L.EPI_SWIGLU, out=out, out_ld=F, rows_per_sample=T, row_ss=ss, ss_dim=D,
             ss_eps=1e-5, col_bias2=b2)
    rstd = torch.rsqrt(ss.sum(1) / D + 1e-5)
    v = (gop.float() @ w13.float().t()) * rstd[:, None] + b2[torch.arange(M, device="cuda") // T]
    ref = torch.nn.functional.silu(v[:, 0::2]) * v[:, 1::2]
    assert rel(out, ref) < 6e-3


@pytest.mark.parametrize("N,T,D,H", [(2, 312, 1152, 16), (4, 40, 192, 8)])
def test_fused_rmsnorm_qkv_matches_unfused(ops, N, T, D, H):
    """The fused chain (producer epilogue -> QKV GEMM with row scale + bias table) against rmsnorm_modulate + QKV GEMM."""
    from ma3_b200 import lib as L
    hd = D // H
    hdp = 64 if hd <= 64 else 128
    M = N * T
    h = torch.randn(M, D, generator=g(11)).cuda() * 3
    w = torch.randn(D, generator=g(12)).cuda() * 0.1 + 1
    mod = torch.randn(N, 3 * D, generator=g(13)).cuda() * 0.3           # [shift | scale | wn]
    mod[:, 2 * D:] = w * (1 + mod[:, D:2 * D])
    wqkv = (torch.randn(3 * D, D, generator=g(14)) / D ** 0.5).bfloat16().cuda()
    rope = torch.view_as_real(torch.polar(torch.ones(T, hd // 2), torch.outer(torch.arange(T).float(),
                                                                             1.0 / 10000 ** (torch.arange(0, hd, 2).float() / hd)))).contiguous().cuda()
    kw = dict(M=M, N=3 * D, K=D, epi=L.EPI_QKV_ROPE, rope=rope, model_dim=D, head_dim=hd, head_dim_pad=hdp, tokens=T,
              tokens_pad=(T + 7) // 8 * 8, q_scale=0.17)
    def outs():
        return (torch.zeros(N, H, T, hdp, device="cuda", dtype=torch.bfloat16),
                torch.zeros(N, H, T, hdp, device="cuda", dtype=torch.bfloat16),
                torch.zeros(N, H, hdp, (T + 7) // 8 * 8, device="cuda", dtype=torch.bfloat16))
    u = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    ops.rmsnorm_modulate(h, w, u, mod=mod, shift_off=0, scale_off=D, rows_per_sample=T)
    q0, k0, v0 = outs()
    ops.gemm(u, wqkv, q_out=q0, k_out=k0, vt_out=v0, **kw)
    # fused: operand g = bf16(h * wn), sums of squares per 32-column chunk, bias table shift W^T (fp32 here)
    rows = torch.arange(M, device="cuda") // T
    gop = (h * mod[rows, 2 * D:]).bfloat16()
    sc = (D // 32 + 3) // 4 * 4
    ss = torch.zeros(M, sc, device="cuda")
    ss[:, :D // 32] = h.pow(2).view(M, D // 32, 32).sum(-1)
    b2 = mod[:, :D] @ wqkv.float().t()
    q1, k1, v1 = outs()
    ops.gemm(gop, wqkv, q_out=q1, k_out=k1, vt_out=v1, rows_per_sample=T, row_ss=ss, ss_dim=D, ss_eps=1e-5, col_bias2=b2,
             **kw)
    for x0, x1 in ((q0, q1), (k0, k1), (v0, v1)):
        assert rel(x1, x0) < 1.5e-2 and O.cosine(x1.float().cpu(), x0.float().cpu()) > 0.9999


def test_split_bf16_strided_and_norm_weights(ops):
    depth, D, R = 3, 64, 10
    ld = 6 * D * depth + 2 * D + 2 * D * depth
    mod = torch.randn(R, ld, generator=g(21)).cuda()
    nw = torch.randn(depth, 2, D, generator=g(22)).cuda()
    ref = mod.clone()
    tail = 6 * D * depth + 2 * D
    for i in range(depth):
        for j in range(2):
            ref[:, tail + (2 * i + j) * D: tail + (2 * i + j + 1) * D] = nw[i, j] * (1 + mod[:, 6 * D * i + (4 if j else 1) * D: 6 * D * i + (5 if j else 2) * D])
    ops.norm_weights(mod, nw, depth, D, tail)
    assert torch.allclose(mod, ref, rtol=1e-6, atol=1e-6)
    out = torch.empty(depth, 2 * R, D, device="cuda", dtype=torch.bfloat16)
    ops.split_bf16(mod, out, col0=3 * D, col_step=6 * D, nb=depth, cols=D)
    for i in range(depth):
        x = mod[:, 6 * D * i + 3 * D: 6 * D * i + 4 * D]
        assert torch.equal(out[i, :R], x.bfloat16())
        assert rel(out[i, :R].float() + out[i, R:].float(), x) < 2e-5


@pytest.mark.parametrize("samples,T,D,K", [(16, 312, 1152, 1152), (16, 312, 1152, 3072), (2, 312, 768, 2048),
                                           (2, 936, 1536, 1536), (3, 40, 384, 64), (1, 100, 1152, 256), (2, 5, 768, 80)])
def test_gemm_rownorm(ops, samples, T, D, K):
    """Row-owning cluster GEMM: h += gate_s * (a w^T) in place, u = bf16(rms(h_new) * wn_s + shift_s); the row sums of
    squares cross the cluster (D / 384 CTAs) through distributed shared memory.  Ragged M, sample boundaries inside a
    row block, K not a multiple of 64."""
    M = samples * T
    a = torch.randn(M, K, generator=g(31)).bfloat16().cuda()
    w = (torch.randn(D, K, generator=g(32)) / K ** 0.5).bfloat16().cuda()
    h = (torch.randn(M, D, generator=g(33)) * 2).cuda()
    mod = torch.randn(samples, 3 * D + 4, generator=g(34)).cuda()
    gate, wn, shift = mod[:, :D], mod[:, D:2 * D], mod[:, 2 * D:3 * D]
    rows = torch.arange(M, device="cuda") // T
    h_ref = h.double() + gate[rows].double() * (a.double() @ w.double().t())
    u_ref = h_ref * torch.rsqrt(h_ref.pow(2).mean(1, keepdim=True) + 1e-5) * wn[rows].double() + shift[rows].double()
    hh = h.clone()
    u = torch.full((M, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm_rownorm(a, w, hh, gate, rows_per_sample=T, wn=wn, shift=shift, u_out=u)
    assert rel(hh, h_ref) < 1e-5      # fp32 accumulation over K vs float64
    assert rel(u, u_ref) < 5e-3 and bool(torch.isfinite(u.float()).all())
    h2, u2 = h.clone(), torch.empty_like(u)
    ops.gemm_rownorm(a, w, h2, gate, rows_per_sample=T, wn=wn, shift=shift, u_out=u2)
    assert torch.equal(h2, hh) and torch.equal(u2, u)            # no atomics: bit-reproducible
    h3 = h.clone()
    ops.gemm_rownorm(a, w, h3, gate, rows_per_sample=T)          # residual update only (last block)
    assert torch.equal(h3, hh)
