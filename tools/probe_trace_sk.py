import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L
lib = L.require_device()
lib.ma3_debug_set_gemm_trace.argtypes = [ctypes.c_void_p]
lib.ma3_debug_set_gemm_mode.argtypes = [ctypes.c_int]
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lib.ma3_debug_set_gemm_mode(mode); print('debug mode', mode)
dev = "cuda"; bf = torch.bfloat16
Ns, T, D, F = 16, 312, 1152, 3072
M = Ns * T
u = torch.randn(M, D, device=dev).to(bf); mid = torch.randn(M, F, device=dev).to(bf)
wo = (torch.randn(D, D, device=dev) / D ** .5).to(bf); w2 = (torch.randn(D, F, device=dev) / F ** .5).to(bf)
h = torch.randn(M, D, device=dev); mod = torch.randn(Ns, D, device=dev) * 0.1
cases = {"wo": lambda **kw: ops.gemm(u, wo, M=M, N=D, K=D, epi=L.EPI_GATE_RES, out=h, gate=mod, rows_per_sample=T, **kw),
         "w2": lambda **kw: ops.gemm(mid, w2, M=M, N=D, K=F, epi=L.EPI_GATE_RES, out=h, gate=mod, rows_per_sample=T, **kw)}
for name, fn in cases.items():
    for kw in (dict(cta_group=1, tile_n=192, stream_k=-1), dict(cta_group=2, tile_n=192, stream_k=-1), dict(cta_group=1, tile_n=192, stream_k=1)):
        for _ in range(3): fn(**kw)
        torch.cuda.synchronize()
        tr = torch.zeros(256, dtype=torch.int64, device=dev)
        lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(tr.data_ptr()))
        fn(**kw); torch.cuda.synchronize()
        lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(0))
        t = tr.cpu().view(16, 16); base = int(t[0, 0])
        print(name, kw, "entry", int(t[15, 0]) - base, "setup", int(t[15, 1]) - base, "all_done", int(t[15, 2]) - base)
        for i in range(6):
            if int(t[i, 2]) == 0: continue
            r = [int(v) - base if int(v) else -1 for v in t[i, :11]]
            print(f"  item {i}: mma wait_tempty {r[0]:7d} start {r[1]:7d} issued {r[2]:7d} (mainloop {r[2]-r[1]:6d}) | epi ready {r[4]:7d} tfull {r[5]:7d} done {r[6]:7d} (busy {r[6]-r[5]:6d})")
