"""One vocoder conv shape, a few launches (ncu target): C T k B [res] [acc]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L
L.require_device()
C, Tt, kk, B = (int(v) for v in sys.argv[1:5])
res_on, acc_on = "res" in sys.argv[5:], "acc" in sys.argv[5:]
x = torch.randn(B, Tt, C, device="cuda").half()
y = torch.zeros(B, Tt, C, device="cuda", dtype=torch.float16)
w = (torch.randn(kk * C, C, device="cuda") / (C * kk) ** .5).half()
bias = torch.zeros(C, device="cuda")
taps = [(j - kk // 2, j * C) for j in range(kk)]
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for _ in range(4):
    flush.zero_()
    ops.gemm(x, w, M=Tt, N=C, K=C, batch=B, a_rows=Tt, a_batch_stride=Tt * C, b_rows=kk * C, taps=taps, out=y,
             out_batch_stride=Tt * C, bias=bias, res=x if res_on else None, alpha=1 / 3 if acc_on else 1.0,
             accumulate=acc_on)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
