"""GPU probe: (tile_n, cta_group) sweep of the four DiT GEMMs at the bench shapes (graph-replay timing), to calibrate
the tile cost model in csrc/gemm.cu."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L
from _bench import bench

dev = "cuda"
N, T, D, H, F = 16, 312, 1152, 16, 3072
if len(sys.argv) > 1 and sys.argv[1] == "M":
    N, D, H, F = 32, 768, 32, 2048
hd = D // H; hdp = 64 if hd <= 64 else 128
M = N * T; bf = torch.bfloat16
torch.manual_seed(0)
h = torch.randn(M, D, device=dev); u = torch.randn(M, D, device=dev).to(bf)
mod = torch.randn(N, 6 * D, device=dev) * 0.1
wqkv = (torch.randn(3 * D, D, device=dev) / D ** .5).to(bf); wo = (torch.randn(D, D, device=dev) / D ** .5).to(bf)
w13 = (torch.randn(2 * F, D, device=dev) / D ** .5).to(bf); w2 = (torch.randn(D, F, device=dev) / F ** .5).to(bf)
q = torch.zeros(N, H, T, hdp, device=dev, dtype=bf); k = torch.zeros_like(q)
vt = ops.alloc_vt(N, H, hd=hd, hdp=hdp, tokens_pad=T, device=dev)
att = torch.randn(M, D, device=dev).to(bf); mid = torch.randn(M, F, device=dev).to(bf)
ang = torch.outer(torch.arange(1000, device=dev).float(), 1.0 / (10000 ** (torch.arange(0, hd, 2, device=dev).float() / hd)))
rope = torch.stack([ang.cos(), ang.sin()], -1).contiguous()
cases = {
    "qkv": (lambda **kw: ops.gemm(u, wqkv, M=M, N=3 * D, K=D, epi=L.EPI_QKV_ROPE, q_out=q, k_out=k, vt_out=vt, rope=rope, model_dim=D, head_dim=hd, head_dim_pad=hdp, tokens=T, tokens_pad=T, q_scale=0.1, **kw), 2.0 * M * 3 * D * D),
    "wo": (lambda **kw: ops.gemm(att, wo, M=M, N=D, K=D, epi=L.EPI_GATE_RES, out=h, gate=mod[:, :D], rows_per_sample=T, **kw), 2.0 * M * D * D),
    "w13": (lambda **kw: ops.gemm(u, w13, M=M, N=2 * F, K=D, epi=L.EPI_SWIGLU, out=mid, out_ld=F, **kw), 2.0 * M * 2 * F * D),
    "w2": (lambda **kw: ops.gemm(mid, w2, M=M, N=D, K=F, epi=L.EPI_GATE_RES, out=h, gate=mod[:, :D], rows_per_sample=T, **kw), 2.0 * M * D * F),
}
for name, (fn, fl) in cases.items():
    bench(f"{name} auto", lambda: fn(), flops=fl, n=3)
    for tn in (128, 192, 256):
        for cg in (1, 2):
            bench(f"{name} tile_n={tn} cta_group={cg}", lambda: fn(tile_n=tn, cta_group=cg), flops=fl, n=3)
