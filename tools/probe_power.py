"""GPU probe: board power and SM clock while ONE kernel of the XL DiT block runs back to back for ~1.5 s (CUDA graph of
20 launches replayed in a loop; NVML sampled every 10 ms from a thread), plus a cuBLAS bf16 GEMM of the SwiGLU shape as
the yardstick.  Energy per launch = power x time: where the joules of the power-capped step go."""
import sys, os, math, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, pynvml
from ma3_b200 import ops, lib as L
L.require_device()
pynvml.nvmlInit(); hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = "cuda"
N, T, Lc, D, H, F = 16, 312, 154, 1152, 16, 3072
hd = D // H; hdp = 128; M = N * T; Tp = T; Lp = (Lc + 7) // 8 * 8
bf = torch.bfloat16
torch.manual_seed(0)
h = torch.randn(M, D, device=dev); u = torch.randn(M, D, device=dev).to(bf)
mod = torch.randn(N, 6 * D, device=dev) * 0.1; wn = torch.randn(D, device=dev)
wqkv = (torch.randn(3 * D, D, device=dev) / D ** .5).to(bf); wo = (torch.randn(D, D, device=dev) / D ** .5).to(bf)
w13 = (torch.randn(2 * F, D, device=dev) / D ** .5).to(bf); w2 = (torch.randn(D, F, device=dev) / F ** .5).to(bf)
q = torch.randn(N, H, T, hdp, device=dev).to(bf); k = torch.randn(N, H, T, hdp, device=dev).to(bf); vt = ops.alloc_vt(N, H, hd=hd, hdp=hdp, tokens_pad=Tp, device=dev)
ky = torch.randn(N, H, Lc, hdp, device=dev).to(bf); vyt = torch.randn(N, H, hdp, Lp, device=dev).to(bf); vyt[:, :, hd:] = 0; vyt[:, :, hd] = 1
gate = torch.randn(H, device=dev); att = torch.randn(M, D, device=dev).to(bf); mid = torch.randn(M, F, device=dev).to(bf)
ang = torch.outer(torch.arange(1000, device=dev).float(), 1.0 / (10000 ** (torch.arange(0, hd, 2, device=dev).float() / hd)))
rope = torch.stack([ang.cos(), ang.sin()], -1).contiguous(); qs = math.log2(math.e) / math.sqrt(hd)
o16 = torch.empty(M, 2 * F, device=dev, dtype=bf)
cases = [
    ("idle", None, 0),
    ("rmsnorm_modulate", lambda: ops.rmsnorm_modulate(h, wn, u, mod=mod, shift_off=0, scale_off=D, rows_per_sample=T), 0),
    ("qkv gemm + rope", lambda: ops.gemm(u, wqkv, M=M, N=3 * D, K=D, epi=L.EPI_QKV_ROPE, q_out=q, k_out=k, vt_out=vt, rope=rope, model_dim=D, head_dim=hd, head_dim_pad=hdp, tokens=T, tokens_pad=Tp, q_scale=qs), 2.0 * M * 3 * D * D),
    ("attention", lambda: ops.attention(q, k, vt, ky, vyt, gate, att, hd=hd), 4.0 * N * H * T * (T + Lc) * hd),
    ("wo gemm + gate_res", lambda: ops.gemm(att, wo, M=M, N=D, K=D, epi=L.EPI_GATE_RES, out=h, gate=mod[:, 2 * D:3 * D], rows_per_sample=T), 2.0 * M * D * D),
    ("w13 gemm + swiglu", lambda: ops.gemm(u, w13, M=M, N=2 * F, K=D, epi=L.EPI_SWIGLU, out=mid, out_ld=F), 2.0 * M * 2 * F * D),
    ("w2 gemm + gate_res", lambda: ops.gemm(mid, w2, M=M, N=D, K=F, epi=L.EPI_GATE_RES, out=h, gate=mod[:, 5 * D:6 * D], rows_per_sample=T), 2.0 * M * D * F),
    ("cuBLAS w13 shape", lambda: torch.matmul(u, w13.t(), out=o16), 2.0 * M * 2 * F * D),
    ("cuBLAS 8192^3", None, 2.0 * 8192 ** 3),
]
A8 = torch.randn(8192, 8192, device=dev).to(bf); B8 = torch.randn(8192, 8192, device=dev).to(bf); C8 = torch.empty(8192, 8192, device=dev, dtype=bf)
cases[-1] = ("cuBLAS 8192^3", lambda: torch.matmul(A8, B8, out=C8), 2.0 * 8192 ** 3)

def sample(stop, out):
    while not stop.is_set():
        out.append((pynvml.nvmlDeviceGetPowerUsage(hnd) / 1e3, pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM)))
        time.sleep(0.01)

print(f"{'kernel':22s} {'us/launch':>10s} {'W':>7s} {'SM MHz':>7s} {'mJ/launch':>10s} {'TFLOP/s':>8s} {'pJ/flop':>8s}")
for name, fn, flops in cases:
    if fn is None:
        s = []; stop = threading.Event(); th = threading.Thread(target=sample, args=(stop, s)); th.start(); time.sleep(1.0); stop.set(); th.join()
        print(f"{name:22s} {'':>10s} {sorted(x[0] for x in s)[len(s)//2]:7.0f} {sorted(x[1] for x in s)[len(s)//2]:7.0f}")
        continue
    for _ in range(3): fn()
    g = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=st):
            for _ in range(20): fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    reps = max(10, int(1.5e3 / max(e0.elapsed_time(e1), 1e-3)))
    s = []; stop = threading.Event(); th = threading.Thread(target=sample, args=(stop, s)); th.start()
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    us = e0.elapsed_time(e1) * 1e3 / (reps * 20)
    tail = s[len(s) // 3:]                      # the power reading lags: use the last two thirds
    w = sorted(x[0] for x in tail)[len(tail) // 2]; mhz = sorted(x[1] for x in tail)[len(tail) // 2]
    line = f"{name:22s} {us:10.1f} {w:7.0f} {mhz:7.0f} {w * us * 1e-3:10.2f}"
    if flops: line += f" {flops / us / 1e6:8.0f} {w * us * 1e-6 / flops * 1e12:8.2f}"
    print(line, flush=True)
    time.sleep(0.5)
