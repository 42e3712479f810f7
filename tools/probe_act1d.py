"""GPU probe: fused Activation1d at the six BigVGAN stage shapes of the bench batch (8 clips), TMA-staged kernel
(default) against the first-generation kernel (ma3_debug_set_act_version(1)); GB/s are algorithmic (read + write)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _bench import bench
lib = L.require_device()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for (C, Tt) in [(768, 2496), (384, 9984), (192, 19968), (96, 39936), (48, 79872), (32, 159744)]:
    x = torch.randn(B, Tt, C, device="cuda").half()
    y = torch.empty(B, Tt, C, device="cuda", dtype=torch.float16)
    al = torch.zeros(C, device="cuda")
    for ver in (0, 1):
        lib.ma3_debug_set_act_version(ver)
        bench(f"act1d v{ver} B{B} C{C} T{Tt}", lambda: ops.act1d(x, y, al, al), bytes_=B * Tt * C * 4)
    lib.ma3_debug_set_act_version(0)
