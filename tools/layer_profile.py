"""Per-launch timing of the tensor-core convolutions of ONE generator forward (optimisation aid, not a bench)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moonsuperresolution_b200 import GauGAN, _lib
from moonsuperresolution_b200 import weights as W

I = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
G = int(sys.argv[3]) if len(sys.argv) > 3 else 1
w = W.random_init("spade", I, seed=0)
m = GauGAN(I, B, precision="bf16", weights=w, max_groups=G)
n = B * G
src = (torch.rand((n, I, I, 2), device="cuda") - 0.5).contiguous()
out = torch.empty((n, I, I), device="cuda")
eps = torch.randn((n, 256), device="cuda")
for _ in range(3):
    m.forward_device(src, out, eps, G)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    m.forward_device(src, out, eps, G)
e1.record(); torch.cuda.synchronize()
print("forward ms (avg of 5, n=%d): %.3f  -> %.1f TFLOP/s model" % (n, e0.elapsed_time(e1) / 5, 702.61e9 * n / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e12 if I == 512 else 0))
_lib.profile_enable(True)
m.forward_device(src, out, eps, G)
fam = _lib.profile_read()
rec = _lib.profile_records("conv_tc")
other = {f: _lib.profile_records(f) for f in ("stats", "mask_conv", "elementwise", "dense")}
_lib.profile_enable(False)
tot = sum(v["ms"] for v in fam.values())
for k, v in fam.items():
    if v["launches"]:
        print("%-12s %8.3f ms  %5.1f%%  launches %d" % (k, v["ms"], 100 * v["ms"] / tot, v["launches"]))
print("conv_tc launches in order: idx ms GFLOP TFLOP/s")
for k, (ms, wk) in enumerate(rec):
    print("%3d %8.4f %9.2f %8.1f" % (k, ms, wk / 1e9, wk / ms / 1e9 if ms > 0 else 0))
for f, rs in other.items():
    print(f + " launches in order: idx ms MB GB/s")
    for k, (ms, wk) in enumerate(rs):
        print("%3d %8.4f %9.2f %8.1f" % (k, ms, wk / 1e6, wk / ms / 1e6 if ms > 0 else 0))
