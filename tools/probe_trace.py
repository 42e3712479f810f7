"""GPU probe: per-tile pipeline timeline of CTA 0 of one GEMM launch (clock64 stamps), for the DiT epilogues."""
import sys, os, ctypes, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L
lib = L.require_device()
lib.ma3_debug_set_gemm_trace.argtypes = [ctypes.c_void_p]
lib.ma3_debug_set_gemm_mode.argtypes = [ctypes.c_int]
dev = "cuda"; bf = torch.bfloat16
Ns, T, D, H, F = 16, 312, 1152, 16, 3072
M = Ns * T; hd = D // H; hdp = 128
u = torch.randn(M, D, device=dev).to(bf); mid = torch.randn(M, F, device=dev).to(bf)
wqkv = (torch.randn(3 * D, D, device=dev) / D ** .5).to(bf); wo = (torch.randn(D, D, device=dev) / D ** .5).to(bf)
w13 = (torch.randn(2 * F, D, device=dev) / D ** .5).to(bf); w2 = (torch.randn(D, F, device=dev) / F ** .5).to(bf)
h = torch.randn(M, D, device=dev); mod = torch.randn(Ns, D, device=dev) * 0.1
q = torch.zeros(Ns, H, T, hdp, device=dev, dtype=bf); k = torch.zeros_like(q)
vt = ops.alloc_vt(Ns, H, hd=hd, hdp=hdp, tokens_pad=T, device=dev)
ang = torch.outer(torch.arange(1000, device=dev).float(), 1.0 / (10000 ** (torch.arange(0, hd, 2, device=dev).float() / hd)))
rope = torch.stack([ang.cos(), ang.sin()], -1).contiguous()
o16 = torch.empty(M, 2 * F, device=dev, dtype=bf)
cases = {
    "qkv_rope": lambda **kw: ops.gemm(u, wqkv, M=M, N=3 * D, K=D, epi=L.EPI_QKV_ROPE, q_out=q, k_out=k, vt_out=vt, rope=rope, model_dim=D, head_dim=hd, head_dim_pad=hdp, tokens=T, tokens_pad=T, q_scale=0.1, **kw),
    "wo_gate_res": lambda **kw: ops.gemm(u, wo, M=M, N=D, K=D, epi=L.EPI_GATE_RES, out=h, gate=mod, rows_per_sample=T, **kw),
    "w13_swiglu": lambda **kw: ops.gemm(u, w13, M=M, N=2 * F, K=D, epi=L.EPI_SWIGLU, out=mid, out_ld=F, **kw),
    "w2_gate_res": lambda **kw: ops.gemm(mid, w2, M=M, N=D, K=F, epi=L.EPI_GATE_RES, out=h, gate=mod, rows_per_sample=T, **kw),
    "w13_store": lambda **kw: ops.gemm(u, w13, M=M, N=2 * F, K=D, out=o16, **kw),
}
def mainloop(fn, kw):
    tr = torch.zeros(256, dtype=torch.int64, device=dev)
    lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(tr.data_ptr()))
    fn(**kw); torch.cuda.synchronize()
    lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(0))
    t = tr.cpu().view(16, 16)
    ml = [int(t[i, 2] - t[i, 1]) for i in range(1, 8) if int(t[i, 2])]
    eb = [int(t[i, 6] - t[i, 5]) for i in range(1, 8) if int(t[i, 6])]
    return sum(ml) / max(1, len(ml)), sum(eb) / max(1, len(eb))

print("mainloop clocks per tile (tiles 1..): mode 0 normal / 1 no TMA / 2 no MMA; epilogue busy clocks")
for name in ("w13_swiglu", "w13_store", "qkv_rope", "w2_gate_res"):
    for kw in (dict(cta_group=1, tile_n=256), dict(cta_group=2, tile_n=256), dict(cta_group=1, tile_n=192), dict(cta_group=1, tile_n=128), dict(cta_group=2, tile_n=128)):
        if "gate" in name: kw["stream_k"] = -1
        fn = cases[name]
        res = []
        for mode in (0, 1, 2):
            lib.ma3_debug_set_gemm_mode(mode)
            fn(**kw); torch.cuda.synchronize()
            res.append(mainloop(fn, kw))
        lib.ma3_debug_set_gemm_mode(0)
        iters = (F if "w2" in name else D) // 64
        print(f"{name:12s} {str(kw):55s} per k-iter: normal {res[0][0]/iters:6.0f}  noTMA {res[1][0]/iters:6.0f}  noMMA {res[2][0]/iters:6.0f} | epi busy {res[0][1]:6.0f} {res[1][1]:6.0f} {res[2][1]:6.0f}")
