"""GPU probe: per-kernel timing of one XL DiT block at the bench shapes (N=16, T=312, L=154)."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L

dev = "cuda"
N, T, Lc, D, H, F = 16, 312, 154, 1152, 16, 3072
if len(sys.argv) > 1 and sys.argv[1] == "M":
    N, D, H, F = 32, 768, 32, 2048
hd = D // H; hdp = 64 if hd <= 64 else 128
M = N * T; Tp = (T + 7) // 8 * 8; Lp = (Lc + 7) // 8 * 8
bf = torch.bfloat16
torch.manual_seed(0)
h = torch.randn(M, D, device=dev); u = torch.empty(M, D, device=dev, dtype=bf)
mod = torch.randn(N, 6 * D, device=dev) * 0.1
wn = torch.randn(D, device=dev)
wqkv = (torch.randn(3 * D, D, device=dev) / D ** .5).to(bf); wo = (torch.randn(D, D, device=dev) / D ** .5).to(bf)
w13 = (torch.randn(2 * F, D, device=dev) / D ** .5).to(bf); w2 = (torch.randn(D, F, device=dev) / F ** .5).to(bf)
q = torch.zeros(N, H, T, hdp, device=dev, dtype=bf); k = torch.zeros_like(q); vt = ops.alloc_vt(N, H, hd=hd, hdp=hdp, tokens_pad=Tp, device=dev)
ky = torch.randn(N, H, Lc, hdp, device=dev).to(bf); vyt = torch.randn(N, H, hdp, Lp, device=dev).to(bf); vyt[:, :, hd:] = 0; vyt[:, :, hd] = 1
gate = torch.randn(H, device=dev); att = torch.empty(M, D, device=dev, dtype=bf); mid = torch.empty(M, F, device=dev, dtype=bf)
ang = torch.outer(torch.arange(1000, device=dev).float(), 1.0 / (10000 ** (torch.arange(0, hd, 2, device=dev).float() / hd)))
rope = torch.stack([ang.cos(), ang.sin()], -1).contiguous()
qs = math.log2(math.e) / math.sqrt(hd)

from _bench import bench

bench("rmsnorm_modulate", lambda: ops.rmsnorm_modulate(h, wn, u, mod=mod, shift_off=0, scale_off=D, rows_per_sample=T), bytes_=M * D * 6)
bench("qkv gemm + rope", lambda: ops.gemm(u, wqkv, M=M, N=3 * D, K=D, epi=L.EPI_QKV_ROPE, q_out=q, k_out=k, vt_out=vt, rope=rope, model_dim=D, head_dim=hd, head_dim_pad=hdp, tokens=T, tokens_pad=Tp, q_scale=qs), flops=2.0 * M * 3 * D * D)
bench("attention self+cross", lambda: ops.attention(q, k, vt, ky, vyt, gate, att, hd=hd), flops=4.0 * N * H * T * (T + Lc) * hd)
bench("wo gemm + gate_res", lambda: ops.gemm(att, wo, M=M, N=D, K=D, epi=L.EPI_GATE_RES, out=h, gate=mod[:, 2 * D:3 * D], rows_per_sample=T), flops=2.0 * M * D * D)
bench("w13 gemm + swiglu", lambda: ops.gemm(u, w13, M=M, N=2 * F, K=D, epi=L.EPI_SWIGLU, out=mid, out_ld=F), flops=2.0 * M * 2 * F * D)
bench("w2 gemm + gate_res", lambda: ops.gemm(mid, w2, M=M, N=D, K=F, epi=L.EPI_GATE_RES, out=h, gate=mod[:, 5 * D:6 * D], rows_per_sample=T), flops=2.0 * M * D * F)
for cg in (1, 2):
    bench(f"qkv cta_group={cg}", lambda: ops.gemm(u, wqkv, M=M, N=3 * D, K=D, epi=L.EPI_QKV_ROPE, q_out=q, k_out=k, vt_out=vt, rope=rope, model_dim=D, head_dim=hd, head_dim_pad=hdp, tokens=T, tokens_pad=Tp, q_scale=qs, cta_group=cg), flops=2.0 * M * 3 * D * D)
    bench(f"w13 cta_group={cg}", lambda: ops.gemm(u, w13, M=M, N=2 * F, K=D, epi=L.EPI_SWIGLU, out=mid, out_ld=F, cta_group=cg), flops=2.0 * M * 2 * F * D)
    for sk in (-1, 1):
        for tn in (128, 192, 256):
            bench(f"wo cta_group={cg} stream_k={sk} tile_n={tn}", lambda: ops.gemm(att, wo, M=M, N=D, K=D, epi=L.EPI_GATE_RES, out=h, gate=mod[:, 2 * D:3 * D], rows_per_sample=T, cta_group=cg, stream_k=sk, tile_n=tn), flops=2.0 * M * D * D)
            bench(f"w2 cta_group={cg} stream_k={sk} tile_n={tn}", lambda: ops.gemm(mid, w2, M=M, N=D, K=F, epi=L.EPI_GATE_RES, out=h, gate=mod[:, 5 * D:6 * D], rows_per_sample=T, cta_group=cg, stream_k=sk, tile_n=tn), flops=2.0 * M * D * F)
sys.exit(0)
for tn in (128, 192, 256):
    bench(f"w13 swiglu tile_n={tn}", lambda: ops.gemm(u, w13, M=M, N=2 * F, K=D, epi=L.EPI_SWIGLU, out=mid, out_ld=F, tile_n=tn), flops=2.0 * M * 2 * F * D)
    bench(f"qkv rope tile_n={tn}", lambda: ops.gemm(u, wqkv, M=M, N=3 * D, K=D, epi=L.EPI_QKV_ROPE, q_out=q, k_out=k, vt_out=vt, rope=rope, model_dim=D, head_dim=hd, head_dim_pad=hdp, tokens=T, tokens_pad=Tp, q_scale=qs, tile_n=tn), flops=2.0 * M * 3 * D * D)
o32 = torch.empty(M, 2 * F, device=dev, dtype=bf)
for tn in (128, 192, 256):
    bench(f"w13 plain tile_n={tn}", lambda: ops.gemm(u, w13, M=M, N=2 * F, K=D, out=o32, tile_n=tn), flops=2.0 * M * 2 * F * D)
bench("w13 gemm plain store bf16", lambda: ops.gemm(u, w13, M=M, N=2 * F, K=D, out=o32), flops=2.0 * M * 2 * F * D)
bench("cublas w13", lambda: torch.matmul(u, w13.t()), flops=2.0 * M * 2 * F * D)
bench("cublas wo", lambda: torch.matmul(att, wo.t()), flops=2.0 * M * D * D)
if len(sys.argv) > 2: sys.exit(0)
# vocoder-like convs, 8 clips
B = 8
for (C, Tt, kk) in [(768, 2496, 11), (384, 9984, 7), (192, 19968, 7), (96, 39936, 7), (48, 79872, 7), (32, 159744, 7)]:
    x = torch.randn(B, Tt, C, device=dev).half(); w = (torch.randn(kk * C, C, device=dev) / (C * kk) ** .5).half()
    y = torch.empty(B, Tt, C, device=dev, dtype=torch.float16); bias = torch.zeros(C, device=dev)
    taps = [(j - kk // 2, j * C) for j in range(kk)]
    bench(f"conv k{kk} C{C} T{Tt}", lambda: ops.gemm(x, w, M=Tt, N=C, K=C, batch=B, a_rows=Tt, a_batch_stride=Tt * C, b_rows=kk * C, taps=taps, out=y, out_batch_stride=Tt * C, bias=bias, res=x), flops=2.0 * B * Tt * C * C * kk, bytes_=B * Tt * C * 2 * 3)
    al = torch.zeros(C, device=dev)
    bench(f"act1d C{C} T{Tt}", lambda: ops.act1d(x, y, al, al), bytes_=B * Tt * C * 4)
