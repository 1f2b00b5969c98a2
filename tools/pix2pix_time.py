"""pix2pix-256 forward time at the cfg2 call shape (optimisation aid): MSR_TC_MASK=0 restores the im2col form of block 1."""
import os, sys
sys.path.insert(0, os.getcwd())
import torch
import moonsuperresolution_b200 as msr
from moonsuperresolution_b200 import weights as W
B, G = 16, 8
w = W.random_init("pix2pix", 256, seed=0)
m = msr.Pix2Pix(B, precision="bf16", weights=w, max_groups=G)
n = B * G
src = (torch.rand((n, 256, 256, 2), device="cuda") - 0.5).contiguous()
out = torch.empty((n, 256, 256), device="cuda")
for _ in range(5):
    m.forward_device(src, out, None, G)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    m.forward_device(src, out, None, G)
e1.record(); torch.cuda.synchronize()
print("pix2pix-256 forward, n=%d: %.4f ms, launches %d" % (n, e0.elapsed_time(e1) / 50, m.last_launch_count))
