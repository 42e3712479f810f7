"""ctypes binding of libma3b200.so (the C ABI declared in include/ma3_b200.h).

There is no fallback: if the shared library is missing or the device is not sm_100, every op raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# MA3_LIB overrides the library path (used by tools/ to A/B kernel variants); the default is the in-tree build
LIB_PATH = os.environ.get("MA3_LIB") or os.path.join(_HERE, "csrc", "libma3b200.so")

F32, BF16, F16 = 0, 1, 2
EPI_STORE, EPI_GATE_RES, EPI_SWIGLU, EPI_QKV_ROPE = 0, 1, 2, 3
MAX_TAPS = 16

_DT = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}


class Ma3Error(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("a_rows", C.c_int64), ("a_ld", C.c_int64), ("a_batch_stride", C.c_int64),
        ("b", C.c_void_p), ("b_rows", C.c_int64), ("b_ld", C.c_int64), ("b_batch_stride", C.c_int64),
        ("dtype", C.c_int32), ("batch", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("taps", C.c_int32),
        ("a_shift", C.c_int32 * MAX_TAPS), ("b_row", C.c_int32 * MAX_TAPS),
        ("epi", C.c_int32),
        ("out", C.c_void_p), ("out_dtype", C.c_int32),
        ("out_ld", C.c_int64), ("out_batch_stride", C.c_int64),
        ("out_row_mul", C.c_int32), ("out_row_off", C.c_int32),
        ("bias", C.c_void_p), ("bias_per_row", C.c_int32),
        ("res", C.c_void_p), ("res_dtype", C.c_int32),
        ("res_ld", C.c_int64), ("res_batch_stride", C.c_int64),
        ("alpha", C.c_float), ("act", C.c_int32), ("accumulate", C.c_int32),
        ("gate", C.c_void_p), ("gate_ld", C.c_int64), ("gate_batch_stride", C.c_int64), ("rows_per_sample", C.c_int32),
        ("q_out", C.c_void_p), ("k_out", C.c_void_p), ("vt_out", C.c_void_p),
        ("rope", C.c_void_p),
        ("model_dim", C.c_int32), ("head_dim", C.c_int32), ("head_dim_pad", C.c_int32),
        ("tokens", C.c_int32), ("tokens_pad", C.c_int32),
        ("q_scale", C.c_float), ("first_section", C.c_int32),
        ("tile_n", C.c_int32), ("cta_group", C.c_int32), ("stream_k", C.c_int32),
        ("norm_out", C.c_void_p), ("norm_w", C.c_void_p), ("ss_out", C.c_void_p),
        ("row_ss", C.c_void_p), ("ss_cols", C.c_int32), ("ss_dim", C.c_int32), ("ss_eps", C.c_float),
        ("col_bias2", C.c_void_p), ("col_bias2_ld", C.c_int64),
    ]


_lib = None


def load():
    """Load the shared library (no CUDA calls are made by loading)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Ma3Error(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for this path)")
        lib = C.CDLL(LIB_PATH)
        lib.ma3_version.restype = C.c_int
        lib.ma3_check_device.restype = C.c_int
        lib.ma3_launch_count.restype = C.c_int64
        lib.ma3_last_error.restype = C.c_char_p
        lib.ma3_gemm.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
        lib.ma3_gemm.restype = C.c_int
        _declare_ops(lib)
        _lib = lib
    return _lib


def _declare_ops(lib):
    """argtypes of the non-GEMM entry points (kept next to their wrappers in ops.py)."""
    from . import ops  # noqa: F401  (ops registers its prototypes on import)
    ops.declare(lib)


def check(rc, what):
    if rc != 0:
        msg = load().ma3_last_error().decode("utf-8", "replace")
        raise Ma3Error(f"{what} failed (rc={rc}): {msg}")


def require_device():
    lib = load()
    if not torch.cuda.is_available():
        raise Ma3Error("CUDA device required: ma3_b200 has no CPU fallback")
    check(lib.ma3_check_device(), "ma3_check_device")
    return lib


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(load().ma3_launch_count())


PDL_MAX_ROWS = 2048   # token rows (batch x latent frames) up to which programmatic dependent launch pays (see ma3_set_pdl)


def auto_pdl(rows):
    """Switch programmatic dependent launch for the launches (and graph captures) that follow: on for small, latency-bound
    batches, off for the power-capped large ones (measured both ways).  MA3_PDL=0|1 overrides."""
    load().ma3_set_pdl(1 if rows <= PDL_MAX_ROWS else 0)


def dt(t):
    return _DT[t.dtype]


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
