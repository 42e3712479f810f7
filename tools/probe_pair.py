"""GPU probe: RMSNorm-mod alone vs right after the gated-residual GEMM that wrote its input (graph replay, sustained):
what reading freshly reduced-into data costs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L
L.require_device()
dev = "cuda"; bf = torch.bfloat16
N, T, D, F = 16, 312, 1152, 3072
M = N * T
h = torch.randn(M, D, device=dev); u = torch.empty(M, D, device=dev, dtype=bf)
mod = torch.randn(N, 6 * D, device=dev) * 0.1; wn = torch.randn(D, device=dev)
wo = (torch.randn(D, D, device=dev) / D ** .5).to(bf); att = torch.randn(M, D, device=dev).to(bf)
w13 = (torch.randn(2 * F, D, device=dev) / D ** .5).to(bf); mid = torch.empty(M, F, device=dev, dtype=bf)
h2 = torch.randn(M, D, device=dev)
def rms(src=h): ops.rmsnorm_modulate(src, wn, u, mod=mod, shift_off=0, scale_off=D, rows_per_sample=T)
def gemm_wo(): ops.gemm(att, wo, M=M, N=D, K=D, epi=L.EPI_GATE_RES, out=h, gate=mod[:, 2 * D:3 * D], rows_per_sample=T)
def gemm_w13(): ops.gemm(u, w13, M=M, N=2 * F, K=D, epi=L.EPI_SWIGLU, out=mid, out_ld=F)
def timeit(name, fn, reps=200):
    for _ in range(3): fn()
    g = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(10): fn()
    torch.cuda.synchronize()
    for _ in range(20): g.replay()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * 10)
    print(f"{name:60s} {us:7.1f} us", flush=True)
    return us
a = timeit("rmsnorm alone (input static)", rms)
b = timeit("wo gemm alone", gemm_wo)
c = timeit("wo gemm -> rmsnorm(h)", lambda: (gemm_wo(), rms()))
d = timeit("wo gemm -> rmsnorm(other static tensor)", lambda: (gemm_wo(), rms(h2)))
e = timeit("w13 gemm alone", gemm_w13)
f = timeit("rmsnorm -> w13 gemm (reads the fresh u)", lambda: (rms(), gemm_w13()))
print(f"rmsnorm after the GEMM that wrote h: {c - b:.1f} us (alone {a:.1f}); after the GEMM but on a static tensor: {d - b:.1f} us")
print(f"w13 GEMM on a fresh operand: {f - a:.1f} us (alone {e:.1f})")
