"""Thin Python wrappers over the C ABI (one function per entry point).  Tensors must be CUDA tensors; every
wrapper enqueues on the current torch stream and returns immediately."""
import ctypes as C

import torch

from . import lib as L


# When set to a list, every wrapped launch is bracketed by CUDA events on the launching stream and recorded as
# (kernel tag, start event, end event, algorithmic work: FLOPs for GEMM/attention, bytes for HBM-bound kernels).
# bench.py uses this for the live roofline measurement; None (default) adds no overhead.
PROFILE = None
# With PROFILE_CAPTURED_ONLY only launches issued while the stream is being captured are bracketed: the events become
# event-record nodes of the CUDA graph (`external=True`), so a later replay times every kernel on the device with no
# host launch latency between the two records (eager bracketing over-reports short kernels when the host falls behind).
PROFILE_CAPTURED_ONLY = False


class _Span:
    def __init__(self, tag, work):
        self.tag, self.work = tag, work
        self.on = False

    def __enter__(self):
        self.on = PROFILE is not None and (not PROFILE_CAPTURED_ONLY or torch.cuda.is_current_stream_capturing())
        if self.on:
            self.e0 = torch.cuda.Event(enable_timing=True, external=True)
            self.e1 = torch.cuda.Event(enable_timing=True, external=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.on and exc[0] is None:
            self.e1.record()
            PROFILE.append((self.tag, self.e0, self.e1, self.work))
        return False


def declare(lib):
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    protos = {
        "ma3_rmsnorm_modulate": [vp, vp, vp, i64, i32, i32, i32, vp, i32, i32, i32, f32, vp],
        "ma3_final_layer": [vp, vp, i64, i32, i32, vp, vp, i32, i32, i32, i32, f32, vp, vp],
        "ma3_final_layer_cfg_euler": [vp, vp, i64, i32, i32, vp, vp, i32, i32, i32, i32, f32, f32, f32, vp, vp, vp, vp],
        "ma3_cfg_euler_update": [vp, vp, vp, i64, f32, f32, i32, vp],
        "ma3_proj_in": [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp],
        "ma3_timestep_embed": [vp, vp, i32, i32, i32, vp],
        "ma3_pool_layernorm": [vp, i32, vp, vp, vp, i32, i32, i32, i32, f32, vp],
        "ma3_layernorm_rows": [vp, i32, vp, vp, vp, i32, i32, i32, f32, vp],
        "ma3_groupnorm_swish": [vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, f32, i32, vp],
        "ma3_softmax_rows": [vp, vp, i32, i32, i32, i64, i64, f32, vp],
        "ma3_nct_to_ntc": [vp, vp, i32, i32, i32, i32, i32, f32, vp],
        "ma3_ntc_to_nct": [vp, i32, vp, i32, i32, i32, i64, vp],
        "ma3_upsample_nearest2": [vp, vp, i64, i32, vp],
        "ma3_cast": [vp, i32, vp, i32, i64, vp],
        "ma3_adaln_input": [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp],
        "ma3_act1d_set_filter": [vp, vp],
        "ma3_act1d": [vp, i32, vp, i32, vp, vp, i32, i32, i32, i32, vp],
        "ma3_attention": [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp],
        "ma3_l2_persist": [vp, C.c_size_t, vp],
        "ma3_split_bf16": [vp, i64, i32, i32, i32, i32, i32, vp, vp],
        "ma3_norm_weights": [vp, i64, vp, i32, i32, i32, i32, vp],
        "ma3_qknorm_rope": [vp, i64, i32, vp, vp, vp, vp, f32, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, f32, vp],
        "ma3_melnet_prep": [vp, vp, i32, i32, i32, i32, i32, vp],
        "ma3_melnet_mag": [vp, i64, vp, i32, i32, i32, i32, vp],
        "ma3_melnet_log": [vp, vp, i32, i32, i32, vp],
        "ma3_gemm_rownorm": [vp, i64, vp, i64, i32, i32, i32, i32, vp, vp, vp, vp, i64, i32, vp, f32, vp],
        "ma3_embed_rows": [vp, i64, vp, vp, vp, vp, i32, i32, i32, vp],
    }
    for name, args in protos.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int


# ---------------------------------------------------------------------------------------------------- tap-GEMM
def gemm(a, b, *, M, N, K, batch=1, a_rows=None, a_ld=None, a_batch_stride=0, b_rows=None, b_ld=None,
         b_batch_stride=0, taps=((0, 0),), epi=L.EPI_STORE, out=None, out_ld=None, out_batch_stride=0,
         out_row_mul=1, out_row_off=0, bias=None, bias_per_row=False, res=None, res_ld=None, res_batch_stride=0,
         alpha=1.0, accumulate=False, gate=None, gate_batch_stride=0, rows_per_sample=0, q_out=None, k_out=None, vt_out=None, rope=None,
         model_dim=0, head_dim=0, head_dim_pad=0, tokens=0, tokens_pad=0, q_scale=1.0, first_section=0, act=0,
         tile_n=0, cta_group=0, stream_k=0, norm_out=None, norm_w=None, ss_out=None, row_ss=None, ss_dim=0,
         ss_eps=1e-5, col_bias2=None):
    """acc[z,m,n] = sum_taps sum_k A[z, m + a_shift, k] * B[z, n + b_row, k]; see include/ma3_b200.h."""
    lib = L.require_device()
    assert a.dtype == b.dtype and a.dtype in (torch.bfloat16, torch.float16)
    d = L.GemmDesc()
    d.a = a.data_ptr()
    d.a_rows = a_rows if a_rows is not None else M
    d.a_ld = a_ld if a_ld is not None else K
    d.a_batch_stride = a_batch_stride
    d.b = b.data_ptr()
    d.b_rows = b_rows if b_rows is not None else N * len(taps)
    d.b_ld = b_ld if b_ld is not None else K
    d.b_batch_stride = b_batch_stride
    d.dtype = L.dt(a)
    d.batch, d.M, d.N, d.K = batch, M, N, K
    d.taps = len(taps)
    for i, (s, r) in enumerate(taps):
        d.a_shift[i] = s
        d.b_row[i] = r
    d.epi = epi
    if out is not None:
        d.out = out.data_ptr()
        d.out_dtype = L.dt(out)
    d.out_ld = out_ld if out_ld is not None else N
    d.out_batch_stride = out_batch_stride
    d.out_row_mul, d.out_row_off = out_row_mul, out_row_off
    if bias is not None:
        assert bias.dtype == torch.float32
        d.bias = bias.data_ptr()
    d.bias_per_row = int(bias_per_row)
    if res is not None:
        d.res = res.data_ptr()
        d.res_dtype = L.dt(res)
        d.res_ld = res_ld if res_ld is not None else d.out_ld
        d.res_batch_stride = res_batch_stride
    d.alpha = alpha
    d.accumulate = int(accumulate)
    if gate is not None:
        assert gate.dtype == torch.float32
        d.gate = gate.data_ptr()
        d.gate_ld = gate.stride(0)
        d.gate_batch_stride = gate_batch_stride
    d.rows_per_sample = rows_per_sample
    if q_out is not None:
        d.q_out, d.k_out, d.vt_out = q_out.data_ptr(), k_out.data_ptr(), vt_out.data_ptr()
        d.rope = rope.data_ptr() if rope is not None else None
    d.model_dim, d.head_dim, d.head_dim_pad = model_dim, head_dim, head_dim_pad
    d.tokens, d.tokens_pad = tokens, tokens_pad
    d.q_scale = q_scale
    d.first_section = first_section
    d.act = act
    d.tile_n = tile_n
    d.cta_group = cta_group
    d.stream_k = stream_k
    if norm_out is not None:   # fused-RMSNorm producer (GATE_RES): see include/ma3_b200.h
        assert norm_w.dtype == torch.float32 and norm_w.stride(0) == gate.stride(0) and ss_out.dtype == torch.float32
        assert norm_out.dtype == a.dtype and N % 32 == 0 and ss_out.shape[-1] >= N // 32 and ss_out.shape[-1] % 4 == 0
        d.norm_out, d.norm_w, d.ss_out = norm_out.data_ptr(), norm_w.data_ptr(), ss_out.data_ptr()
        d.ss_cols = ss_out.shape[-1]
        d.stream_k = -1
    if row_ss is not None:     # fused-RMSNorm consumer
        assert row_ss.dtype == torch.float32 and col_bias2.dtype == torch.float32 and rows_per_sample > 0
        d.row_ss, d.ss_cols, d.ss_dim, d.ss_eps = row_ss.data_ptr(), row_ss.shape[-1], ss_dim, ss_eps
        d.col_bias2, d.col_bias2_ld = col_bias2.data_ptr(), col_bias2.stride(0)
    with _Span(f"tap_gemm/{_EPI_NAMES[epi]}/M{M} N{N} K{K} taps{len(taps)} batch{batch}", 2.0 * M * N * K * len(taps) * batch):
        L.check(lib.ma3_gemm(C.byref(d), L.stream_ptr()), "ma3_gemm")
    return out


_SM_COUNT = {}


def sm_count():
    dev = torch.cuda.current_device()
    if dev not in _SM_COUNT:
        _SM_COUNT[dev] = torch.cuda.get_device_properties(dev).multi_processor_count
    return _SM_COUNT[dev]


def gemm_rownorm(a, w, h, gate, *, rows_per_sample, wn=None, shift=None, u_out=None, eps=1e-5):
    """h += gate_s * (a w^T) in place (fp32 [M, D]); with u_out: u_out = 16-bit(rms(h_new) * wn_s + shift_s).
    gate / wn / shift: fp32 [samples, D] views sharing one row pitch.  See include/ma3_b200.h (ma3_gemm_rownorm)."""
    M, K = a.shape
    D = w.shape[0]
    assert a.dtype == w.dtype and h.dtype == torch.float32 and h.is_contiguous() and h.shape == (M, D)
    assert gate.dtype == torch.float32 and a.stride(1) == 1 and w.stride(1) == 1
    ld = gate.stride(0)
    if u_out is not None:
        assert wn.stride(0) == ld and shift.stride(0) == ld and u_out.is_contiguous() and u_out.dtype == a.dtype
    lib = L.require_device()
    with _Span(f"tap_gemm/rownorm/M{M} N{D} K{K} taps1 batch1", 2.0 * M * D * K):
        L.check(lib.ma3_gemm_rownorm(L.ptr(a), a.stride(0), L.ptr(w), w.stride(0), L.dt(a), M, K, D, L.ptr(h), L.ptr(gate),
                                     L.ptr(wn), L.ptr(shift), ld, rows_per_sample, L.ptr(u_out), eps, L.stream_ptr()),
                "ma3_gemm_rownorm")
    return h


_EPI_NAMES = {L.EPI_STORE: "store", L.EPI_GATE_RES: "gate_res", L.EPI_SWIGLU: "swiglu", L.EPI_QKV_ROPE: "qkv_rope"}


# ---------------------------------------------------------------------------------------------------- other ops
def _call(name, *args, work=0.0):
    lib = L.require_device()
    with _Span(name[4:], work):
        L.check(getattr(lib, name)(*args, L.stream_ptr()), name)


def rmsnorm_modulate(x, w, out, *, mod=None, shift_off=0, scale_off=0, rows_per_sample=1, eps=1e-5):
    """out = rms(x) * w * (1 + scale) + shift; x fp32 [M, D]; mod fp32 [samples, ld] or None (plain RMSNorm)."""
    M, D = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous()
    _call("ma3_rmsnorm_modulate", L.ptr(x), L.ptr(w), L.ptr(mod), mod.stride(0) if mod is not None else 0,
          shift_off, scale_off, rows_per_sample, L.ptr(out), L.dt(out), M, D, eps,
          work=float(M) * D * (4 + out.element_size()))
    return out


def final_layer(h, mod, shift_off, scale_off, W, bias, N, T, v_out, eps=1e-6):
    D, Cout = h.shape[-1], W.shape[0]
    _call("ma3_final_layer", L.ptr(h), L.ptr(mod), mod.stride(0), shift_off, scale_off, L.ptr(W), L.ptr(bias), N, T, D,
          Cout, eps, L.ptr(v_out))
    return v_out


def final_layer_cfg_euler(h, mod, shift_off, scale_off, W, bias, N, T, guidance, dt, x_in, x_out, v_out=None,
                          eps=1e-6):
    D, Cout = h.shape[-1], W.shape[0]
    _call("ma3_final_layer_cfg_euler", L.ptr(h), L.ptr(mod), mod.stride(0), shift_off, scale_off, L.ptr(W),
          L.ptr(bias), N, T, D, Cout, eps, guidance, dt, L.ptr(x_in), L.ptr(x_out), L.ptr(v_out))
    return x_out


def cfg_euler_update(v, x, out, dt, guidance, cfg=True):
    _call("ma3_cfg_euler_update", L.ptr(v), L.ptr(x), L.ptr(out), x.numel(), dt, guidance, int(cfg))
    return out


def proj_in(x, Wt, b, h, N):
    """h = x^T Wt + b with Wt the transposed nn.Linear weight [C, D]."""
    xB, Cc, T = x.shape
    _call("ma3_proj_in", L.ptr(x), L.ptr(Wt), L.ptr(b), L.ptr(h), N, xB, Cc, T, Wt.shape[1])
    return h


def timestep_embed(t, out):
    assert t.dtype == torch.int64
    _call("ma3_timestep_embed", L.ptr(t), L.ptr(out), L.dt(out), t.numel(), out.shape[-1])
    return out


def pool_layernorm(ctx, w, b, out, eps=1e-5):
    N, Lc, Cd = ctx.shape
    _call("ma3_pool_layernorm", L.ptr(ctx), L.dt(ctx), L.ptr(w), L.ptr(b), L.ptr(out), L.dt(out), N, Lc, Cd, eps)
    return out


def layernorm_rows(x, w, b, out, eps=1e-5):
    M, D = x.shape
    _call("ma3_layernorm_rows", L.ptr(x), L.dt(x), L.ptr(w), L.ptr(b), L.ptr(out), L.dt(out), M, D, eps)
    return out


def groupnorm_swish(x, w, b, out, *, groups=32, eps=1e-6, swish=True):
    B, T, Cc = x.shape
    _call("ma3_groupnorm_swish", L.ptr(x), L.dt(x), L.ptr(w), L.ptr(b), L.ptr(out), L.dt(out), B, T, Cc, groups, eps,
          int(swish), work=float(B) * T * Cc * (x.element_size() + out.element_size()))
    return out


def softmax_rows(S, P, n, scale):
    rows = S.numel() // S.shape[-1]
    _call("ma3_softmax_rows", L.ptr(S), L.ptr(P), L.dt(P), rows, n, S.shape[-1], P.shape[-1], scale)
    return P


def nct_to_ntc(x, out, scale=1.0):
    B, Cc, T = x.shape
    _call("ma3_nct_to_ntc", L.ptr(x), L.ptr(out), L.dt(out), B, Cc, T, out.shape[-1], scale)
    return out


def ntc_to_nct(x, out):
    B, Cc, T = out.shape
    _call("ma3_ntc_to_nct", L.ptr(x), L.dt(x), L.ptr(out), B, Cc, T, x.shape[-1])
    return out


def upsample_nearest2(x, out):
    B, T, Cc = x.shape
    _call("ma3_upsample_nearest2", L.ptr(x), L.ptr(out), B * T, Cc)
    return out


def split_bf16(x, out, *, col0=0, col_step=0, nb=1, cols=None):
    """x fp32 [M, ld] -> out bf16 [nb, 2M, cols]: slice b = columns [col0 + b*col_step, +cols) of x;
    rows [0, M) = bf16(slice), rows [M, 2M) = bf16(slice - hi)."""
    M = x.shape[0]
    cols = x.shape[1] if cols is None else cols
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1 and out.dtype == torch.bfloat16
    assert out.is_contiguous() and out.numel() == 2 * nb * M * cols
    _call("ma3_split_bf16", L.ptr(x), x.stride(0), col0, col_step, nb, M, cols, L.ptr(out))
    return out


def norm_weights(mod2d, norm_w, depth, D, tail_off):
    """mod2d fp32 [rows, ld] (in place): tail columns <- norm_w[i][j] * (1 + scale slot), see include/ma3_b200.h."""
    assert mod2d.dtype == torch.float32 and mod2d.stride(1) == 1 and norm_w.is_contiguous()
    _call("ma3_norm_weights", L.ptr(mod2d), mod2d.stride(0), L.ptr(norm_w), mod2d.shape[0], depth, D, tail_off)
    return mod2d


def gemm_split(x32, w2, *, M, N, K, out, bias=None, act=0, out_ld=None):
    """out = act(x W^T + bias) to ~16 mantissa bits on the bf16 tensor cores: x32 fp32 [M, K]; w2 bf16 [2N, K] = the
    stacked (hi, lo) halves of W made by `split_weight`.  One tap-GEMM with three taps (hi.hi + lo.hi + hi.lo)."""
    a2 = torch.empty(2 * M, K, device=x32.device, dtype=torch.bfloat16)
    split_bf16(x32, a2)
    return gemm(a2, w2, M=M, N=N, K=K, a_rows=2 * M, b_rows=2 * N, taps=((0, 0), (M, 0), (0, N)), out=out, bias=bias,
                act=act, out_ld=out_ld)


def split_weight(w):
    """fp32 [N, K] -> bf16 [2N, K] (hi rows, then lo rows); done once at pack time."""
    w = w.detach().float()
    hi = w.to(torch.bfloat16)
    return torch.cat([hi, (w - hi.float()).to(torch.bfloat16)]).contiguous()


def qknorm_rope(x, *, first_section, qn, kn, rope, q_out, k_out, vt_out, tokens, tokens_pad, D, hd, hdp, q_scale=1.0,
                eps=1e-5):
    """x fp32 [M, sections*D] raw projections -> LayerNorm(q), LayerNorm(k), RoPE, scale, scatter (qk_norm=True path).
    qn / kn: (weight, bias) fp32 or None."""
    M = x.shape[0]
    assert x.dtype == torch.float32 and x.stride(1) == 1
    qw, qb = qn if qn is not None else (None, None)
    kw, kb = kn if kn is not None else (None, None)
    _call("ma3_qknorm_rope", L.ptr(x), x.stride(0), first_section, L.ptr(qw), L.ptr(qb), L.ptr(kw), L.ptr(kb), eps,
          L.ptr(rope), L.ptr(q_out), L.ptr(k_out), L.ptr(vt_out), L.dt(k_out), M, tokens, tokens_pad, D, hd, hdp, q_scale)


def embed_rows(table, ids, out, *, T, pos=None, type0=None):
    """out[m] = table[ids[m]] (+ pos[m % T]) (+ type0): token embedding lookup of the text encoders (fp32)."""
    assert table.dtype == torch.float32 and ids.dtype == torch.int64 and out.dtype == torch.float32
    _call("ma3_embed_rows", L.ptr(table), table.shape[0], L.ptr(ids), L.ptr(pos), L.ptr(type0), L.ptr(out), ids.numel(), T,
          table.shape[1])
    return out


def cast(x, out):
    _call("ma3_cast", L.ptr(x), L.dt(x), L.ptr(out), L.dt(out), x.numel())
    return out


_filter_set = set()   # device ordinals whose __constant__ taps have been uploaded


def act1d(x, out, alpha, beta, logscale=True):
    """Fused Activation1d on channels-last x [B, T, C]."""
    if x.device.index not in _filter_set:
        taps = (C.c_float * 12)(*kaiser_sinc_taps())
        with torch.cuda.device(x.device):
            _call("ma3_act1d_set_filter", taps)
        _filter_set.add(x.device.index)
    B, T, Cc = x.shape
    _call("ma3_act1d", L.ptr(x), L.dt(x), L.ptr(out), L.dt(out), L.ptr(alpha), L.ptr(beta), B, T, Cc, int(logscale),
          work=float(B) * T * Cc * (x.element_size() + out.element_size()))
    return out


def kaiser_sinc_taps(cutoff=0.25, half_width=0.3, kernel_size=12):
    """Host-side filter design, run once (vocoder/bigvgan/alias_free_torch/filter.py:28-57)."""
    import math
    half = kernel_size // 2
    A = 2.285 * (half - 1) * math.pi * (4 * half_width) + 7.95
    beta = 0.1102 * (A - 8.7) if A > 50.0 else (0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0) if A >= 21.0 else 0.0)
    win = torch.kaiser_window(kernel_size, periodic=False, beta=beta, dtype=torch.float64)
    t = torch.arange(-half, half, dtype=torch.float64) + 0.5 if kernel_size % 2 == 0 else \
        torch.arange(kernel_size, dtype=torch.float64) - half
    f = 2 * cutoff * win * torch.sinc(2 * cutoff * t)
    f = f / f.sum()
    return [float(v) for v in f]


def alloc_vt(*lead, hd, hdp, tokens_pad, device, dtype=torch.bfloat16):
    """Zeroed V^T buffer [*lead, hdp, tokens_pad] in the layout ma3_attention expects: rows [0, hd) are written by
    the QKV GEMM epilogue, row hd (when hd < hdp) holds ones so that the P.V MMA also produces the softmax row sums
    (include/ma3_b200.h, ma3_attention), the remaining pad rows stay zero."""
    vt = torch.zeros(*lead, hdp, tokens_pad, device=device, dtype=dtype)
    if hd < hdp:
        vt[..., hd, :] = 1
    return vt


def attention(q, k, vt, ky, vyt, gate, out, *, hd):
    """q,k [NS,H,T,hdp]; vt [NS,H,hdp,Tp]; ky [NS,H,L,hdp]; vyt [NS,H,hdp,Lp]; out [NS,T,H*hd].
    vt / vyt must come from alloc_vt (ones in row hd)."""
    NS, H, T, hdp = q.shape
    Tp = vt.shape[-1]
    Lc = ky.shape[2] if ky is not None else 0
    Lp = vyt.shape[-1] if vyt is not None else 0
    _call("ma3_attention", L.ptr(q), L.ptr(k), L.ptr(vt), L.ptr(ky), L.ptr(vyt), L.ptr(gate), L.ptr(out), L.dt(q), NS, H,
          T, Tp, Lc, Lp, hd, hdp, work=4.0 * NS * H * T * (T + Lc) * hd)
    return out


def adaln_input(temb, cap, out, S, N, ts_s, ts_n):
    """out[s*N + n] = silu(temb[s*ts_s + n*ts_n] + cap[n]) as 16-bit GEMM operand."""
    _call("ma3_adaln_input", L.ptr(temb), L.ptr(cap), L.ptr(out), L.dt(out), S, N, cap.shape[-1], ts_s, ts_n)
    return out
