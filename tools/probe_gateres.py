"""GPU probe: the two gated-residual GEMMs of a DiT block (wo, w2) over tile shapes and work splits."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L
from _bench import bench

dev = "cuda"
N, T, D, F = 16, 312, 1152, 3072
if len(sys.argv) > 1 and sys.argv[1] == "M":
    N, D, F = 32, 768, 2048
M = N * T; bf = torch.bfloat16
torch.manual_seed(0)
h = torch.randn(M, D, device=dev)
mod = torch.randn(N, 6 * D, device=dev) * 0.1
wo = (torch.randn(D, D, device=dev) / D ** .5).to(bf); w2 = (torch.randn(D, F, device=dev) / F ** .5).to(bf)
att = torch.randn(M, D, device=dev).to(bf); mid = torch.randn(M, F, device=dev).to(bf)
cases = {
    "wo": (lambda **kw: ops.gemm(att, wo, M=M, N=D, K=D, epi=L.EPI_GATE_RES, out=h, gate=mod[:, :D], rows_per_sample=T, **kw), 2.0 * M * D * D),
    "w2": (lambda **kw: ops.gemm(mid, w2, M=M, N=D, K=F, epi=L.EPI_GATE_RES, out=h, gate=mod[:, :D], rows_per_sample=T, **kw), 2.0 * M * D * F),
}
for name, (fn, fl) in cases.items():
    bench(f"{name} auto", lambda: fn(), flops=fl, n=3)
    for tn in (96, 128, 192, 256):
        for cg in (1, 2):
            if cg == 2 and tn % 32: continue
            for sk in (-1, 1):
                bench(f"{name} tile_n={tn} cta_group={cg} stream_k={sk}", lambda: fn(tile_n=tn, cta_group=cg, stream_k=sk), flops=fl, n=3)
