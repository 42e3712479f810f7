"""GPU probe: CTA-0 pipeline timeline and launch span of ONE GEMM launch inside the real, replayed sampling graph (XL,
8 prompts): the trace pointer is baked into the N-th launch of the chosen epilogue during capture, so the numbers are
taken at the step's clocks, L2 state and neighbours -- to compare with the isolated tools/probe_timeline.py."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as Bn
from ma3_b200 import ops, lib as L
from ma3_b200.pipeline import MODEL_CONFIGS, build_random_pipeline
lib = L.require_device()
lib.ma3_debug_set_gemm_trace.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda", 0)
cfg = MODEL_CONFIGS["XL"]
pipe = build_random_pipeline("XL", vocoder_h=dict(Bn.BIGVGAN_H), seed=0, device=dev, use_graph=True)
B, T, Lc, Cd = 8, Bn.T_LATENT, Bn.L_CTX, cfg["context_dim"]
cond = torch.randn(B, Lc, Cd, device=dev); unc = torch.randn(B, Lc, Cd, device=dev); x0 = torch.randn(B, 20, T, device=dev)
which = {"qkv": L.EPI_QKV_ROPE, "swiglu": L.EPI_SWIGLU, "gate": L.EPI_GATE_RES}
if len(sys.argv) > 1: which = {k: v for k, v in which.items() if k in sys.argv[1:]}
target_n = 300   # the 300th launch of that epilogue in the captured step (a mid-step DiT block)
real_gemm = ops.gemm
for name, epi in which.items():
    for kfilter in ((None,) if name != "gate" else (1152, 3072)):
        tr = torch.zeros(256, dtype=torch.int64, device=dev)
        count = [0]
        def traced(*a, **kw):
            hit = kw.get("epi", L.EPI_STORE) == epi and (kfilter is None or kw.get("K") == kfilter) and torch.cuda.is_current_stream_capturing()
            if hit:
                count[0] += 1
                if count[0] == target_n:
                    lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(tr.data_ptr()))
                    try:
                        return real_gemm(*a, **kw)
                    finally:
                        lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(0))
            return real_gemm(*a, **kw)
        ops.gemm = traced
        pipe.sampler._graphs, pipe._tail = {}, {}
        pipe.generate(cond, unc, x0, scale=Bn.GUIDANCE, timesteps=Bn.N_POINTS)      # capture
        torch.cuda.synchronize()
        for _ in range(3):
            tr.zero_(); tr[250] = 2 ** 62
            pipe.generate(cond, unc, x0, scale=Bn.GUIDANCE, timesteps=Bn.N_POINTS)  # replay at steady clocks
        torch.cuda.synchronize()
        ops.gemm = real_gemm
        t = tr.cpu().view(16, 16); base = int(t[0, 0])
        span = (int(t[15, 11]) - int(t[15, 10])) / 1e3
        life = int(t[15, 2]) - int(t[15, 0])
        print(f"{name} K={kfilter}: launch span {span:.1f} us | CTA0 life {life} clk -> {life / max(span, 1e-9) / 1e3:.2f} GHz if CTA0 spanned the launch")
        for i in range(8):
            if int(t[i, 2]) == 0: continue
            r = [int(v) - base if int(v) else -1 for v in t[i, :11]]
            print(f"  item {i}: mma wait_tempty {r[0]:7d} start {r[1]:7d} issued {r[2]:7d} (mainloop {r[2]-r[1]:6d}) | epi ready {r[4]:7d} tfull {r[5]:7d} done {r[6]:7d} (busy {r[6]-r[5]:6d})")
