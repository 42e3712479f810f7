"""GPU probe: vocoder-stage kernels (fused Activation1d and the conv tap-GEMMs) at the bench shapes, 8 clips."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L
dev = "cuda"
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _bench import bench
B = 8
only_act = len(sys.argv) > 1 and sys.argv[1] == "act"
for (C, Tt, kk) in [(768, 2496, 11), (384, 9984, 7), (192, 19968, 7), (96, 39936, 7), (48, 79872, 7), (32, 159744, 7)]:
    x = torch.randn(B, Tt, C, device=dev).half(); y = torch.empty(B, Tt, C, device=dev, dtype=torch.float16)
    al = torch.zeros(C, device=dev)
    bench(f"act1d C{C} T{Tt}", lambda: ops.act1d(x, y, al, al), bytes_=B * Tt * C * 4)
    if only_act: continue
    w = (torch.randn(kk * C, C, device=dev) / (C * kk) ** .5).half(); bias = torch.zeros(C, device=dev)
    taps = [(j - kk // 2, j * C) for j in range(kk)]
    bench(f"conv k{kk} C{C} T{Tt} +res", lambda: ops.gemm(x, w, M=Tt, N=C, K=C, batch=B, a_rows=Tt, a_batch_stride=Tt * C, b_rows=kk * C, taps=taps, out=y, out_batch_stride=Tt * C, bias=bias, res=x), flops=2.0 * B * Tt * C * C * kk, bytes_=B * Tt * C * 2 * 3)
    bench(f"conv k{kk} C{C} T{Tt} +res+acc", lambda: ops.gemm(x, w, M=Tt, N=C, K=C, batch=B, a_rows=Tt, a_batch_stride=Tt * C, b_rows=kk * C, taps=taps, out=y, out_batch_stride=Tt * C, bias=bias, res=x, alpha=1 / 3, accumulate=True), flops=2.0 * B * Tt * C * C * kk, bytes_=B * Tt * C * 2 * 4)
    bench(f"conv k{kk} C{C} T{Tt}", lambda: ops.gemm(x, w, M=Tt, N=C, K=C, batch=B, a_rows=Tt, a_batch_stride=Tt * C, b_rows=kk * C, taps=taps, out=y, out_batch_stride=Tt * C, bias=bias), flops=2.0 * B * Tt * C * C * kk, bytes_=B * Tt * C * 2 * 2)
