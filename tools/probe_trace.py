"""GPU probe: per-tile pipeline timeline of CTA 0 of one GEMM launch (clock64 stamps)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L
lib = L.require_device()
lib.ma3_debug_set_gemm_trace.argtypes = [ctypes.c_void_p]
dev = "cuda"; bf = torch.bfloat16
M, N = 4992, 6144
for K, tn, N in ((1152, 192, 3456), (1152, 256, 6144), (1152, 192, 1152), (3072, 192, 1152)):
    a = torch.randn(M, K, device=dev).to(bf); b = (torch.randn(N, K, device=dev) / K ** .5).to(bf)
    out = torch.empty(M, N, device=dev, dtype=bf)
    ops.gemm(a, b, M=M, N=N, K=K, out=out, tile_n=tn, cta_group=1); torch.cuda.synchronize()
    tr = torch.zeros(256, dtype=torch.int64, device=dev)
    lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(tr.data_ptr()))
    ops.gemm(a, b, M=M, N=N, K=K, out=out, tile_n=tn, cta_group=1); torch.cuda.synchronize()
    lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(0))
    t = tr.cpu().view(16, 16)
    base = int(t[0, 0])
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); ops.gemm(a, b, M=M, N=N, K=K, out=out, tile_n=tn, cta_group=1); e1.record(); torch.cuda.synchronize()
    ent, setup, done = int(t[15, 0]) - base, int(t[15, 1]) - base, int(t[15, 2]) - base
    print(f"N={N} K={K}: event time {e0.elapsed_time(e1)*1e3:.1f} us; CTA0 entry {ent} setup_done {setup} all_done {done} clocks ({(done-ent)/1.9e3:.1f} us @1.9GHz)")
    print(f"K={K} tile_n={tn}: clocks relative to first event; per tile: mma[wait_tempty, start, issued] epi[ready, tfull, done]")
    for i in range(8):
        r = [int(v) - base if int(v) else -1 for v in t[i, :11]]
        print(f"  tile {i}: mma {r[0]:7d} {r[1]:7d} {r[2]:7d}   epi {r[4]:7d} {r[5]:7d} {r[6]:7d}   epi_busy {r[6]-r[5]:6d}  chunk0: ld {r[9]-r[8]:5d} rest {r[10]-r[9]:5d}")
