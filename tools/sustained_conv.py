"""Sustained (>= 1 s, power-capped) throughput of the tcgen05 conv kernel per layer shape, next to cuBLAS (torch.matmul
bf16) on the equivalent GEMM shape.  Optimisation aid."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moonsuperresolution_b200 import _lib

L = _lib.lib()
shapes = [  # n, r, cin, cout
    (16, 64, 1024, 512), (16, 128, 512, 256), (16, 256, 128, 512), (16, 256, 256, 128), (16, 256, 128, 128),
    (16, 32, 1024, 1024), (16, 256, 128, 2048 // 4),
]
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
for (n, r, cin, cout) in shapes:
    x = (torch.randn((n, r, r, cin), device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn((cout, 9 * cin), device="cuda") * 0.02).to(torch.bfloat16)
    b = torch.zeros(cout, device="cuda")
    y = torch.empty((n, r, r, cout), device="cuda")
    st = _lib.stream_ptr()
    flops = 2.0 * n * r * r * cout * 9 * cin

    def launch():
        _lib.check(L.msr_op_conv_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), None, n, r, cin, cout, 9, 1, 1,
                                    0, 0.2, None, st))
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); launch(); e1.record(); torch.cuda.synchronize()
    t1 = e0.elapsed_time(e1) * 1e-3
    iters = max(3, int(secs / t1))
    e0.record()
    for _ in range(iters):
        launch()
    e1.record(); torch.cuda.synchronize()
    ours = flops * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12
    # cuBLAS on the same GEMM shape: (M x K) @ (K x N)
    M, K, N = n * r * r, 9 * cin, cout
    try:
        A = torch.randn((M, K), device="cuda", dtype=torch.bfloat16)
        Bm = torch.randn((K, N), device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            torch.matmul(A, Bm)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            torch.matmul(A, Bm)
        e1.record(); torch.cuda.synchronize()
        cub = flops * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12
        del A, Bm
    except Exception as ex:
        cub = float("nan")
    print(f"n={n} r={r} cin={cin} cout={cout}  M={M} K={K} N={N}: burst {flops / t1 / 1e12:7.1f}  sustained {ours:7.1f} TFLOP/s   cuBLAS sustained {cub:7.1f}", flush=True)
