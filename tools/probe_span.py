"""GPU probe: launch span (first CTA entry -> last CTA exit, %globaltimer) of one GEMM launch inside a replayed CUDA graph
of ten identical launches, against the graph's per-launch time: the difference is what a kernel boundary costs."""
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
    "wo_gate_res": lambda **kw: ops.gemm(u, wo, M=M, N=D, K=D, epi=L.EPI_GATE_RES, out=h, gate=mod, rows_per_sample=T, **kw),
    "w13_swiglu": lambda **kw: ops.gemm(u, w13, M=M, N=2 * F, K=D, epi=L.EPI_SWIGLU, out=mid, out_ld=F, **kw),
    "w2_gate_res": lambda **kw: ops.gemm(mid, w2, M=M, N=D, K=F, epi=L.EPI_GATE_RES, out=h, gate=mod, rows_per_sample=T, **kw),
}
for name, fn in cases.items():
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tr = torch.zeros(256, dtype=torch.int64, device=dev)
    g = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for i in range(10):
                lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(tr.data_ptr() if i == 5 else 0))
                fn()
            lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(0))
    torch.cuda.synchronize()
    spans = []
    for rep in range(4):
        tr.zero_(); tr[250] = 2 ** 62
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        t = tr.cpu()
        spans.append((e0.elapsed_time(e1) * 100, (int(t[251]) - int(t[250])) / 1e3, int(t[242]) - int(t[240])))
    per, span, clk = spans[-1]
    print(f"{name:12s} graph per-launch {per:6.1f} us | in-kernel span {span:6.1f} us | boundary {per - span:5.1f} us | CTA0 life {clk} clk ({clk / span / 1e3:.2f} GHz if it spanned the launch)")
