"""GPU probe: tile shape / CTA group of the generic tap-GEMM for the 192- and 384-channel vocoder convolutions
(8 clips): automatic choice against forced (tile_n, cta_group)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _bench import bench
L.require_device()
B = 8
for (C, Tt) in [(192, 19968), (384, 9984)]:
    x = torch.randn(B, Tt, C, device="cuda").half(); y = torch.empty(B, Tt, C, device="cuda", dtype=torch.float16)
    bias = torch.zeros(C, device="cuda")
    for kk in (3, 7, 11):
        w = (torch.randn(kk * C, C, device="cuda") / (C * kk) ** .5).half()
        taps = [(j - kk // 2, j * C) for j in range(kk)]
        for (tn, cg) in [(0, 0), (C, 1), (C, 2), (C // 2, 1), (C // 2, 2), (64, 1), (128, 1), (128, 2)]:
            if tn > 256 or (tn and C % tn):
                continue
            try:
                bench(f"conv k{kk} C{C} +res tile_n={tn} cg={cg}",
                      lambda: ops.gemm(x, w, M=Tt, N=C, K=C, batch=B, a_rows=Tt, a_batch_stride=Tt * C, b_rows=kk * C, taps=taps,
                                       out=y, out_batch_stride=Tt * C, bias=bias, res=x, tile_n=tn, cta_group=cg),
                      flops=2.0 * B * Tt * C * C * kk, n=3)
            except Exception as e:
                print("skip", tn, cg, str(e)[:80])
