"""Stage-by-stage error of the CUDA generator against the fp64 oracle (accuracy diagnostic, GPU box only)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.nn.functional as F
from moonsuperresolution_b200 import CNNSpade, GauGAN
from moonsuperresolution_b200 import weights as W
from oracle import generator as OG

arch = sys.argv[1] if len(sys.argv) > 1 else "cnn"
I = int(sys.argv[2]) if len(sys.argv) > 2 else 64
B = int(sys.argv[3]) if len(sys.argv) > 3 else 3
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 12
w = W.random_init(arch, I, seed=seed, perturb_affine=True)
rng = np.random.default_rng(1)
x = rng.uniform(-0.5, 0.5, (B, I, I, 2)).astype(np.float32)
x[-1] = 0
eps = rng.standard_normal((B, 256)).astype(np.float32)
dt = torch.float64
with torch.no_grad():
    src = torch.from_numpy(x).to(dt).permute(0, 3, 1, 2).contiguous()
    mean, var = OG.encoder(src, w)
    latent = mean + torch.exp(0.5 * var) * torch.from_numpy(eps).to(dt) if arch == "spade" else mean + var
    sw = I // 64
    h = latent @ OG._t(w["gen.dense.kernel"], dt) + OG._t(w["gen.dense.bias"], dt)
    stages = {"latent": latent.numpy(), "x0": h.numpy()}
    h = h.reshape(B, sw, sw, 1024).permute(0, 3, 1, 2)
    for k in range(1, 7):
        h = OG.residual_block(h, src, w, f"gen.rb{k}")
        stages[f"rb{k}.out"] = h.permute(0, 2, 3, 1).contiguous().numpy()
        h = h.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    out = OG.conv2d_same(F.leaky_relu(h, 0.2), w["gen.out.kernel"], w["gen.out.bias"])
    stages["out"] = out.permute(0, 2, 3, 1).contiguous().numpy()
for precision in ("fp32", "bf16"):
    cls = GauGAN if arch == "spade" else CNNSpade
    m = cls(I, B, precision=precision, weights=w)
    got = m(x, eps=eps)
    print(f"== {precision}")
    for name, ref in stages.items():
        try:
            a = m.read_activation(name)
        except Exception as e:
            print(f"{name:10s} (not available)")
            continue
        a = a.reshape(ref.shape).astype(np.float64)
        err = np.abs(a - ref)
        print(f"{name:10s} ref_rms {np.sqrt((ref**2).mean()):9.4f} ref_max {np.abs(ref).max():9.4f}  err_max {err.max():.3e} err_rms {np.sqrt((err**2).mean()):.3e}  rel_rms {np.sqrt((err**2).mean())/np.sqrt((ref**2).mean()):.3e}")
