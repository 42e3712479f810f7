"""GPU probe: GEMM time vs K (slope = per-chunk cost, intercept = per-tile overhead) for both cta_group modes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L
dev = "cuda"; bf = torch.bfloat16
M, N = 4992, 6144
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
def bench(fn, n=8):
    for _ in range(2): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(n):
        flush.sum()   # evict with CLEAN lines (a dirty flush makes every output store pay a write-back)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n * 1e3
out = torch.empty(M, N, device=dev, dtype=bf)
for tn in (256, 128):
    for cg in (1, 2):
        res = []
        for K in (64, 576, 1152, 2304, 4608):
            a = torch.randn(M, K, device=dev).to(bf); b = (torch.randn(N, K, device=dev) / K ** .5).to(bf)
            us = bench(lambda: ops.gemm(a, b, M=M, N=N, K=K, out=out, tile_n=tn, cta_group=cg))
            res.append((K, us, 2.0 * M * N * K / us / 1e6))
        print(f"tile_n={tn} cta_group={cg}: " + "  ".join(f"K={k}: {u:.1f}us {t:.0f}TF" for k, u, t in res), flush=True)
K = 1152
a = torch.randn(M, K, device=dev).to(bf); b = (torch.randn(N, K, device=dev) / K ** .5).to(bf)
print("cublas K=1152: %.1f us" % bench(lambda: torch.matmul(a, b.t())))
K = 4608
a = torch.randn(M, K, device=dev).to(bf); b = (torch.randn(N, K, device=dev) / K ** .5).to(bf)
print("cublas K=4608: %.1f us" % bench(lambda: torch.matmul(a, b.t())))
