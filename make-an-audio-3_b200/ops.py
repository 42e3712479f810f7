"""Thin Python wrappers over the C ABI (one function per entry point).  Tensors must be CUDA tensors; every
wrapper enqueues on the current torch stream and returns immediately."""
import ctypes as C

import torch

from . import lib as L


def declare(lib):
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    protos = {
    }
    for name, args in protos.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int


# ---------------------------------------------------------------------------------------------------- tap-GEMM
def gemm(a, b, *, M, N, K, batch=1, a_rows=None, a_ld=None, a_batch_stride=0, b_rows=None, b_ld=None,
         b_batch_stride=0, taps=((0, 0),), epi=L.EPI_STORE, out=None, out_ld=None, out_batch_stride=0,
         out_row_mul=1, out_row_off=0, bias=None, bias_per_row=False, res=None, res_ld=None, res_batch_stride=0,
         alpha=1.0, accumulate=False, gate=None, rows_per_sample=0, q_out=None, k_out=None, vt_out=None, rope=None,
         model_dim=0, head_dim=0, head_dim_pad=0, tokens=0, tokens_pad=0, q_scale=1.0, tile_n=0):
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
    d.rows_per_sample = rows_per_sample
    if q_out is not None:
        d.q_out, d.k_out, d.vt_out = q_out.data_ptr(), k_out.data_ptr(), vt_out.data_ptr()
        d.rope = rope.data_ptr()
    d.model_dim, d.head_dim, d.head_dim_pad = model_dim, head_dim, head_dim_pad
    d.tokens, d.tokens_pad = tokens, tokens_pad
    d.q_scale = q_scale
    d.tile_n = tile_n
    L.check(lib.ma3_gemm(C.byref(d), L.stream_ptr()), "ma3_gemm")
    return out
