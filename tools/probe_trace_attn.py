"""GPU probe: per-KV-tile timeline of CTA 0 of the attention kernel (clock64 stamps)."""
import sys, os, ctypes, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L
lib = L.require_device()
lib.ma3_debug_set_gemm_trace.argtypes = [ctypes.c_void_p]
dev = "cuda"; bf = torch.bfloat16
N, T, Lc, D, H = 16, 312, 154, 1152, 16
hd = D // H; hdp = 128; Tp = 312; Lp = 160
q = torch.randn(N, H, T, hdp, device=dev).to(bf); k = torch.randn_like(q); vt = torch.randn(N, H, hdp, Tp, device=dev).to(bf); vt[:, :, hd:] = 0; vt[:, :, hd] = 1
ky = torch.randn(N, H, Lc, hdp, device=dev).to(bf); vyt = torch.randn(N, H, hdp, Lp, device=dev).to(bf); vyt[:, :, hd:] = 0; vyt[:, :, hd] = 1
gate = torch.randn(H, device=dev); att = torch.empty(N * T, D, device=dev, dtype=bf)
ops.attention(q, k, vt, ky, vyt, gate, att, hd=hd); torch.cuda.synchronize()
tr = torch.zeros(256, dtype=torch.int64, device=dev)
lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(tr.data_ptr()))
ops.attention(q, k, vt, ky, vyt, gate, att, hd=hd); torch.cuda.synchronize()
lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(0))
t = tr.cpu().view(16, 16)
base = int(t[0, 0])
print("softmax warp0: [wait_s start, s ready, pass1 done, o ready, o accumulated, pass2 done, arrived]  mma: [wait_p start, p ready, pv issued]")
for i in range(8):
    r = [int(v) - base if int(v) else -1 for v in t[i, :11]]
    print(f" tile {i}: sm {r[0]:6d} {r[1]:6d} {r[2]:6d} {r[3]:6d} {r[4]:6d} {r[5]:6d} {r[6]:6d} | mma {r[8]:6d} {r[9]:6d} {r[10]:6d} | pass1 {r[2]-r[1]:5d} acc {r[4]-r[2]:5d} pass2 {r[5]-r[4]:5d}")
