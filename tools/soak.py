"""Bit-identity soak of the generator at the bench call shape: the same GauGAN-512 forward (batch 16 x 8 groups) ITER
times; every output must equal the first bit for bit.  The tcgen05 kernel synchronises its TMA / MMA / epilogue roles with
hand-written mbarrier protocols (relaxed arrives on the accumulator hand-back): a protocol hole shows up as a rare
mismatch, not as a crash."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moonsuperresolution_b200 import GauGAN
from moonsuperresolution_b200 import weights as W

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 50
I, B, G = 512, 16, 8
m = GauGAN(I, B, precision="bf16", weights=W.random_init("spade", I, seed=0, perturb_affine=True), max_groups=G)
gen = torch.Generator(device="cuda").manual_seed(1)
src = torch.rand((B * G, I, I, 2), device="cuda", generator=gen) - 0.5
src[5 * B + 9:6 * B] = 0
eps = torch.randn((B * G, 256), device="cuda", generator=gen)
out = torch.empty((B * G, I, I), device="cuda")
first, bad = None, 0
for k in range(iters):
    out.fill_(float("nan"))
    m.forward_device(src, out, eps, G)
    torch.cuda.synchronize()
    h = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()
    if first is None:
        first = h
        assert torch.isfinite(out).all()
    elif h != first:
        bad += 1
        print("iteration", k, "differs:", h[:16], "vs", first[:16], flush=True)
print(f"soak: {iters} forwards of GauGAN-512 (128 patches each), sha256 {first[:16]}..., mismatches: {bad}")
sys.exit(1 if bad else 0)
