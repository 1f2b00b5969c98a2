"""One tensor-core layer of the generator in isolation (optimisation aid for ncu captures, not a bench).

    python tools/one_layer.py mask   [n r]        # SPADE mask conv: K = 64 im2col GEMM -> relu -> bf16 (128 columns)
    python tools/one_layer.py conv128 [n r]       # rb6.conv_1: 3x3, 256 -> 128, fp32 out
    python tools/one_layer.py gb [n r]            # rb6.spade_1 gamma|beta conv with the fused SPADE epilogue (C = 256)
    python tools/one_layer.py maskk [n r]         # the same layer with the operand tile built in the kernel, csrc/mask_tc.cu
    python tools/one_layer.py phase [n r]         # last layer (4x4 conv of the x2-upsampled tensor), csrc/phase_tc.cu
Prints the CUDA-event time of the timed launches and the achieved TFLOP/s / GB/s.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moonsuperresolution_b200 import _lib

kind = sys.argv[1] if len(sys.argv) > 1 else "mask"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
r = int(sys.argv[3]) if len(sys.argv) > 3 else 256
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
L = _lib.lib()
st = _lib.stream_ptr()
g = torch.Generator(device="cuda").manual_seed(0)


def rnd(*shape, dtype=torch.bfloat16, scale=1.0):
    return (torch.randn(shape, generator=g, device="cuda") * scale).to(dtype).contiguous()


if kind == "mask":
    x, w, b = rnd(n, r, r, 64), rnd(128, 64, scale=0.1), rnd(128, dtype=torch.float32)
    y = torch.empty((n, r, r, 128), dtype=torch.bfloat16, device="cuda")
    run = lambda: _lib.check(L.msr_op_conv_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), None, y.data_ptr(), n, r, 64, 128,
                                              1, 1, 0, 1, 0.2, None, st), "mask")
    flops, byts = 2.0 * n * r * r * 128 * 64, n * r * r * (128.0 + 256.0)
elif kind == "conv128":
    x, w, b = rnd(n, r, r, 256), rnd(128, 9 * 256, scale=0.02), rnd(128, dtype=torch.float32)
    y = torch.empty((n, r, r, 128), dtype=torch.float32, device="cuda")
    run = lambda: _lib.check(L.msr_op_conv3x3_bf16(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), n, r, 256, 128,
                                                   st), "conv128")
    flops, byts = 2.0 * n * r * r * 128 * 9 * 256, n * r * r * (512.0 + 512.0)
elif kind == "maskk":
    src = (torch.rand((n, 2 * r, 2 * r, 2), generator=g, device="cuda") - 0.5).contiguous()
    h_w = (np.random.default_rng(0).standard_normal((3, 3, 2, 128)) / 4).astype(np.float32)
    h_b = np.zeros(128, np.float32)
    y = torch.empty((n, r, r, 128), dtype=torch.bfloat16, device="cuda")
    run = lambda: _lib.check(L.msr_op_mask_tc(src.data_ptr(), 2 * r, h_w.ctypes.data, h_b.ctypes.data, y.data_ptr(), n, r, st),
                             "maskk")
    flops, byts = 2.0 * n * r * r * 128 * 18, n * r * r * (8.0 + 256.0)
elif kind == "phase":
    cin = 128
    x = rnd(n, r, r, cin)
    w4 = torch.zeros((4, 3, 3, cin))
    for q in range(4):
        for ty in range(3):
            for tx in range(3):
                if (q >> 1 == 0 or ty >= 1) and (q & 1 == 0 or tx >= 1):
                    w4[q, ty, tx] = torch.randn(cin) * 0.02
    h_w4 = w4.reshape(4, 9 * cin).to(torch.bfloat16).contiguous().view(torch.int16).numpy()
    b = rnd(1, dtype=torch.float32)
    y = torch.empty((n, 2 * r, 2 * r), dtype=torch.float32, device="cuda")
    run = lambda: _lib.check(L.msr_op_phase_tc(x.data_ptr(), h_w4.ctypes.data, b.data_ptr(), y.data_ptr(), n, r, cin, 0, st),
                             "phase")
    flops, byts = 2.0 * n * (2 * r) ** 2 * 16 * cin, n * r * r * (2.0 * cin + 16.0)
else:
    C = 256
    a, w, b = rnd(n, r, r, 128), rnd(2 * C, 1152, scale=0.03), rnd(2 * C, dtype=torch.float32)
    xs = rnd(n, r // 2, r // 2, C, dtype=torch.float32)
    mean, rstd = rnd(n // 16, C, dtype=torch.float32), torch.ones((n // 16, C), device="cuda")
    y = torch.empty((n, r, r, C), dtype=torch.bfloat16, device="cuda")
    run = lambda: _lib.check(L.msr_op_spade_tc(a.data_ptr(), w.data_ptr(), b.data_ptr(), xs.data_ptr(), 1, mean.data_ptr(),
                                               rstd.data_ptr(), 16, y.data_ptr(), n, r, C, st), "gb")
    flops, byts = 2.0 * n * r * r * 2 * C * 1152, n * r * r * (256.0 + 2 * C) + n * r * r / 4 * C * 4
for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{kind} n={n} r={r}: {ms:.4f} ms  {flops / ms / 1e9:.1f} TFLOP/s  {byts / ms / 1e6:.1f} GB/s (algorithmic)")
