"""One Activation1d shape, a few launches (the ncu target of tools/gpu_call.sh): C T B [version]."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ma3_b200 import ops, lib as L
lib = L.require_device()
C, Tt, B = (int(v) for v in sys.argv[1:4])
lib.ma3_debug_set_act_version(int(sys.argv[4]) if len(sys.argv) > 4 else 0)
x = torch.randn(B, Tt, C, device="cuda").half()
y = torch.empty_like(x)
al = torch.zeros(C, device="cuda")
flush = torch.empty(64 * 1024 * 1024, device="cuda")
for _ in range(4):
    flush.zero_()
    ops.act1d(x, y, al, al)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()))
