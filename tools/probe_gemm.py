"""GPU probe: tap-GEMM vs torch on a few shapes (run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L

torch.manual_seed(0)
dev = "cuda"

def check(name, got, ref, tol=2e-2):
    err = (got.float() - ref.float()).abs().max().item()
    den = ref.float().abs().max().item() + 1e-9
    ok = err / den < tol
    print(f"{'OK ' if ok else 'BAD'} {name}: max_abs_err={err:.4g} rel={err/den:.4g}", flush=True)
    return ok

def bench(fn, flops, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"    {ms*1e3:.1f} us  {flops/ms/1e9:.1f} TFLOP/s", flush=True)

ok = True
for (M, N, K, tn) in [(128, 128, 64, 0), (128, 256, 128, 0), (256, 64, 256, 0), (1000, 1152, 1152, 0), (4992, 3456, 1152, 0),
                      (4992, 1152, 3072, 192), (4992, 6144, 1152, 256), (4992, 1152, 1152, 128), (300, 80, 96, 0), (300, 48, 48, 0)]:
    for dt in (torch.bfloat16, torch.float16):
        a = torch.randn(M, K, device=dev).to(dt)
        b = (torch.randn(N, K, device=dev) / K ** 0.5).to(dt)
        out = torch.empty(M, N, device=dev, dtype=torch.float32)
        ops.gemm(a, b, M=M, N=N, K=K, out=out, tile_n=tn)
        torch.cuda.synchronize()
        ref = a.float() @ b.float().t()
        ok &= check(f"store f32 M{M} N{N} K{K} {dt}", out, ref, 2e-3)
    if M >= 1000:
        bench(lambda: ops.gemm(a, b, M=M, N=N, K=K, out=out, tile_n=tn), 2.0 * M * N * K)
        o16 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        bench(lambda: ops.gemm(a, b, M=M, N=N, K=K, out=o16, tile_n=tn), 2.0 * M * N * K)
        bench(lambda: torch.matmul(a, b.t()), 2.0 * M * N * K)

# bias + residual + bf16 out + accumulate
M, N, K = 500, 384, 384
a = torch.randn(M, K, device=dev).bfloat16(); b = (torch.randn(N, K, device=dev) / K ** .5).bfloat16()
bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev).bfloat16()
out = torch.randn(M, N, device=dev).bfloat16(); old = out.clone()
ops.gemm(a, b, M=M, N=N, K=K, out=out, bias=bias, res=res, alpha=0.5, accumulate=True)
ref = (a.float() @ b.float().t() + bias + res.float()) * 0.5 + old.float()
ok &= check("store bias+res+alpha+acc bf16", out, ref, 1e-2)

# conv1d k=3 dilation 2 via taps, batched, channels-last
B, T, Ci, Co, k, dil = 3, 300, 128, 192, 3, 2
pad = (k * dil - dil) // 2
x = torch.randn(B, T, Ci, device=dev).bfloat16()
w = (torch.randn(Co, Ci, k, device=dev) / (Ci * k) ** .5).bfloat16()
wp = w.permute(2, 0, 1).contiguous().view(k * Co, Ci)
bias = torch.randn(Co, device=dev)
y = torch.empty(B, T, Co, device=dev, dtype=torch.bfloat16)
ops.gemm(x, wp, M=T, N=Co, K=Ci, batch=B, a_rows=T, a_batch_stride=T * Ci, b_rows=k * Co,
         taps=[(j * dil - pad, j * Co) for j in range(k)], out=y, out_batch_stride=T * Co, bias=bias)
ref = torch.nn.functional.conv1d(x.float().transpose(1, 2), w.float(), bias, padding=pad, dilation=dil).transpose(1, 2)
ok &= check("conv1d k3 d2", y, ref, 1e-2)

# gate-residual
M, N, K, T = 624, 768, 768, 312
a = torch.randn(M, K, device=dev).bfloat16(); b = (torch.randn(N, K, device=dev) / K ** .5).bfloat16()
h = torch.randn(M, N, device=dev); h0 = h.clone(); gate = torch.randn(2, N, device=dev)
ops.gemm(a, b, M=M, N=N, K=K, epi=L.EPI_GATE_RES, out=h, gate=gate, rows_per_sample=T)
ref = h0 + gate.repeat_interleave(T, 0) * (a.float() @ b.float().t())
ok &= check("gate_res", h, ref, 1e-3)

# swiglu
F = 2048
w1 = (torch.randn(F, K, device=dev) / K ** .5).bfloat16(); w3 = (torch.randn(F, K, device=dev) / K ** .5).bfloat16()
w13 = torch.stack([w1, w3], 1).reshape(2 * F, K).contiguous()
o = torch.empty(M, F, device=dev, dtype=torch.bfloat16)
ops.gemm(a, w13, M=M, N=2 * F, K=K, epi=L.EPI_SWIGLU, out=o, out_ld=F)
ref = torch.nn.functional.silu(a.float() @ w1.float().t()) * (a.float() @ w3.float().t())
ok &= check("swiglu", o, ref, 1e-2)

# qkv rope
D, H = 768, 32; hd = D // H; hdp = 64; Tp = 320; Ns = 2
wqkv = (torch.randn(3 * D, K, device=dev) / K ** .5).bfloat16()
ang = torch.outer(torch.arange(T, device=dev).float(), 1.0 / (10000 ** (torch.arange(0, hd, 2, device=dev).float() / hd)))
rope = torch.stack([ang.cos(), ang.sin()], -1).contiguous()
q = torch.zeros(Ns, H, T, hdp, device=dev, dtype=torch.bfloat16); kk = torch.zeros_like(q)
vt = ops.alloc_vt(Ns, H, hd=hd, hdp=hdp, tokens_pad=Tp, device=dev)
ops.gemm(a, wqkv, M=M, N=3 * D, K=K, epi=L.EPI_QKV_ROPE, q_out=q, k_out=kk, vt_out=vt, rope=rope,
         model_dim=D, head_dim=hd, head_dim_pad=hdp, tokens=T, tokens_pad=Tp, q_scale=0.5)
full = (a.float() @ wqkv.float().t()).view(Ns, T, 3, H, hd)
def rot(x):
    xc = torch.view_as_complex(x.reshape(*x.shape[:-1], -1, 2).contiguous())
    fc = torch.polar(torch.ones_like(ang), ang)[None, :, None, :]
    return torch.view_as_real(xc * fc).flatten(3)
ok &= check("rope q", q[..., :hd], 0.5 * rot(full[:, :, 0]).permute(0, 2, 1, 3), 1e-2)
ok &= check("rope k", kk[..., :hd], rot(full[:, :, 1]).permute(0, 2, 1, 3), 1e-2)
ok &= check("vt", vt[:, :, :hd, :T], full[:, :, 2].permute(0, 2, 3, 1), 1e-2)
ok &= bool((q[..., hd:] == 0).all() and (vt[:, :, hd + 1:] == 0).all() and (vt[:, :, hd] == 1).all())
print("ALL OK" if ok else "SOME BAD")
