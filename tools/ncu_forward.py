"""Two GauGAN forwards at the bench's call shape (ncu target: skip the first forward's launches, capture the second)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moonsuperresolution_b200 import GauGAN
from moonsuperresolution_b200 import weights as W

I = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
G = int(sys.argv[3]) if len(sys.argv) > 3 else 8
m = GauGAN(I, B, precision="bf16", weights=W.random_init("spade", I, seed=0), max_groups=G)
n = B * G
src = torch.rand((n, I, I, 2), device="cuda") - 0.5
eps = torch.randn((n, 256), device="cuda")
out = torch.empty((n, I, I), device="cuda")
for _ in range(2):
    m.forward_device(src, out, eps, G)
    torch.cuda.synchronize()
print("launches per forward", m.last_launch_count)
