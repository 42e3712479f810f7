"""GPU probe: CTA-0 pipeline timeline (clock64) of the four DiT GEMMs at the bench shape with the auto-selected tiles."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L
lib = L.require_device()
lib.ma3_debug_set_gemm_trace.argtypes = [ctypes.c_void_p]
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
cases = {
    "qkv_rope": lambda **kw: ops.gemm(u, wqkv, M=M, N=3 * D, K=D, epi=L.EPI_QKV_ROPE, q_out=q, k_out=k, vt_out=vt, rope=rope, model_dim=D, head_dim=hd, head_dim_pad=hdp, tokens=T, tokens_pad=T, q_scale=0.1, **kw),
    "w13_swiglu": lambda **kw: ops.gemm(u, w13, M=M, N=2 * F, K=D, epi=L.EPI_SWIGLU, out=mid, out_ld=F, **kw),
    "w2_gate_res": lambda **kw: ops.gemm(mid, w2, M=M, N=D, K=F, epi=L.EPI_GATE_RES, out=h, gate=mod, rows_per_sample=T, **kw),
}
kws = [dict()] + [eval("dict(%s)" % a) for a in sys.argv[1:]]
for name, fn in cases.items():
    for kw in kws:
        for _ in range(3): fn(**kw)
        torch.cuda.synchronize()
        tr = torch.zeros(256, dtype=torch.int64, device=dev); tr[250] = 2 ** 62
        lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(tr.data_ptr()))
        fn(**kw); torch.cuda.synchronize()
        lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(0))
        t = tr.cpu().view(16, 16); base = int(t[0, 0])
        print(name, kw, "entry", int(t[15, 0]) - base, "setup", int(t[15, 1]) - base, "all_done", int(t[15, 2]) - base, "span_us", (int(t[15, 11]) - int(t[15, 10])) / 1e3)
        for i in range(8):
            if int(t[i, 2]) == 0: continue
            r = [int(v) - base if int(v) else -1 for v in t[i, :11]]
            print(f"  item {i}: mma wait_tempty {r[0]:7d} start {r[1]:7d} issued {r[2]:7d} (mainloop {r[2]-r[1]:6d}) | epi ready {r[4]:7d} tfull {r[5]:7d} done {r[6]:7d} (busy {r[6]-r[5]:6d}) | chunk0: tmem_ld {r[9]-r[8]:5d} rest {r[10]-r[9]:5d}")
