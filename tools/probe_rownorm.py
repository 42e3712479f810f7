"""GPU probe: phase timeline (clock64) of CTA 0 of the row-owning GEMM + launch time by CUDA events."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L
lib = L.require_device()
lib.ma3_debug_set_gemm_trace.argtypes = [ctypes.c_void_p]
dev = "cuda"; bf = torch.bfloat16
N, T, D, F = 16, 312, 1152, 3072
M = N * T
mod = torch.randn(N, 3 * D, device=dev)
h = torch.randn(M, D, device=dev)
u = torch.empty(M, D, device=dev, dtype=bf)
names = ["entry", "setup done", "first stage full", "half of k", "all MMAs done", "phase A done", "phase B done",
         "cluster sync done", "exit"]
for K in (D, F):
    a = torch.randn(M, K, device=dev).to(bf); w = (torch.randn(D, K, device=dev) / K ** 0.5).to(bf)
    run = lambda: ops.gemm_rownorm(a, w, h, mod[:, :D], rows_per_sample=T, wn=mod[:, D:2 * D], shift=mod[:, 2 * D:], u_out=u)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run()
    e1.record(); torch.cuda.synchronize()
    print(f"K={K}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch (back to back, L2-warm)")
    tr = torch.zeros(256, dtype=torch.int64, device=dev)
    tr[32] = 2 ** 62
    lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(tr.data_ptr()))
    run(); torch.cuda.synchronize()
    lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(0))
    g = tr.cpu()[32:36].tolist()
    print(f"   all CTAs: first entry -> last exit {(g[1] - g[0]) / 1e3:.1f} us; CTA 0: {(g[3] - g[2]) / 1e3:.1f} us for "
          f"{int(tr[8] - tr[0])} clk = {(int(tr[8] - tr[0])) / max(1, g[3] - g[2]):.2f} GHz")
    t = tr.cpu()[:9].tolist()
    for i, n in enumerate(names):
        print(f"   {n:20s} {t[i] - t[0]:8d} clk" + (f"  (+{t[i] - t[i - 1]})" if i else ""))
