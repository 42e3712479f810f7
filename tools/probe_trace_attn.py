"""GPU probe: per-KV-tile timeline of CTA 0 of the attention kernel (clock64 stamps)."""
import sys, os, ctypes, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L
lib = L.require_device()
lib.ma3_debug_set_gemm_trace.argtypes = [ctypes.c_void_p]
dev = "cuda"; bf = torch.bfloat16
N, T, Lc, D, H = 16, int(os.environ.get('T', 312)), 154, 1152, 16
hd = D // H; hdp = 128; Tp = (T + 7) // 8 * 8; Lp = 160
q = (torch.randn(N, H, T, hdp, device=dev) * float(os.environ.get('QSCALE', 1.4427 / math.sqrt(hd)))).to(bf); k = torch.randn_like(q); vt = torch.randn(N, H, hdp, Tp, device=dev).to(bf); vt[:, :, hd:] = 0; vt[:, :, hd] = 1
ky = torch.randn(N, H, Lc, hdp, device=dev).to(bf); vyt = torch.randn(N, H, hdp, Lp, device=dev).to(bf); vyt[:, :, hd:] = 0; vyt[:, :, hd] = 1
gate = torch.randn(H, device=dev); att = torch.empty(N * T, D, device=dev, dtype=bf)
ops.attention(q, k, vt, ky, vyt, gate, att, hd=hd); torch.cuda.synchronize()
v3 = os.environ.get('MA3_ATTN_VER', '3') != '2'
QT = (((T + 127) // 128) + 2) // 3 if v3 else (T + 127) // 128
NCTA = QT * H * N
tr = torch.zeros(256 + 4 * NCTA, dtype=torch.int64, device=dev)
lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(tr.data_ptr()))
ops.attention(q, k, vt, ky, vyt, gate, att, hd=hd); torch.cuda.synchronize()
lib.ma3_debug_set_gemm_trace(ctypes.c_void_p(0))
rec = tr.cpu()[256:].view(NCTA, 4)
t = tr.cpu()[:256].view(16, 16)
base = int(t[0, 0])
print("softmax warp0: [wait_s start, s ready, pass1 done, o ready, o accumulated, pass2 done, arrived]  mma: [wait_p start, p ready, pv issued]")
for i in range(12 if v3 else 8):
    r = [int(v) - base if int(v) else -1 for v in t[i, :16]]
    if v3:
        print(f" tile {i}: wait_s {r[0]:6d} s_ready {r[1]:6d} s_freed {r[7]:6d} max {r[2]:6d} resc {r[3]:6d} exp {r[4]:6d} stP {r[5]:6d} arrived {r[6]:6d} | mma ctx0: S {r[8]:6d} PV {r[9]:6d}  ctx2: S {r[10]:6d} PV {r[11]:6d} | exp phase ctx0 [{r[3]:6d},{r[4]:6d}] ctx1 [{r[12]:6d},{r[13]:6d}] ctx2 [{r[14]:6d},{r[15]:6d}]")
        continue
    print(f" tile {i}: sm {r[0]:6d} {r[1]:6d} {r[2]:6d} {r[3]:6d} {r[4]:6d} {r[5]:6d} {r[6]:6d} | mma {r[8]:6d} {r[9]:6d} {r[10]:6d} | pass1 {r[2]-r[1]:5d} acc {r[4]-r[2]:5d} pass2 {r[5]-r[4]:5d}")

import collections
t0 = int(rec[:, 1].min()); t1 = int(rec[:, 2].max())
dur = (rec[:, 2] - rec[:, 1]).float()
clk = rec[:, 3].float()
print(f"CTAs {NCTA}: kernel span {(t1 - t0) / 1e3:.1f} us; CTA duration mean {dur.mean() / 1e3:.2f} us (min {dur.min() / 1e3:.2f}, max {dur.max() / 1e3:.2f}); mean clocks {clk.mean():.0f} -> {clk.mean() / dur.mean():.2f} GHz")
for qt in range(QT):
    sel = torch.arange(NCTA) % QT == qt
    print(f"  q-tile {qt}: mean duration {dur[sel].mean() / 1e3:.2f} us, mean clocks {clk[sel].mean():.0f}")
per_sm = collections.defaultdict(list)
for i in range(NCTA):
    per_sm[int(rec[i, 0])].append((int(rec[i, 1]) - t0, int(rec[i, 2]) - t0))
busy = []
for sm, iv in per_sm.items():
    busy.append(sum(b - a for a, b in iv))
print(f"SMs used {len(per_sm)}; CTAs per SM min {min(len(v) for v in per_sm.values())} max {max(len(v) for v in per_sm.values())}; mean CTA-time per SM {sum(busy) / len(busy) / 1e3:.1f} us (x2 slots -> {sum(busy) / len(busy) / 2e3:.1f} us if perfectly packed)")
sm0 = sorted(per_sm[0])
import time
for _ in range(3): ops.attention(q, k, vt, ky, vyt, gate, att, hd=hd)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.attention(q, k, vt, ky, vyt, gate, att, hd=hd)
e1.record(); torch.cuda.synchronize()
print(f"back-to-back launches: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
print("SM 0 timeline (us):", [(round(a / 1e3, 1), round(b / 1e3, 1)) for a, b in sm0])
