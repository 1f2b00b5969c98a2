"""Where do the producer / MMA issuer / epilogue threads of the tcgen05 conv kernel wait?  (optimisation aid)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moonsuperresolution_b200 import _lib
L = _lib.lib()
shapes = [(16, 64, 1024, 512, 9), (16, 256, 256, 128, 9), (16, 256, 128, 128, 9), (16, 256, 128, 512, 9), (16, 256, 64, 128, 1)]
dbg = torch.zeros((148 * 8,), dtype=torch.int64, device="cuda")
for (n, r, cin, cout, taps) in shapes:
    x = (torch.randn((n, r, r, cin), device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn((cout, taps * cin), device="cuda") * 0.02).to(torch.bfloat16)
    b = torch.zeros(cout, device="cuda")
    y = torch.empty((n, r, r, cout), device="cuda")
    st = _lib.stream_ptr()
    def launch():
        _lib.check(L.msr_op_conv_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), None, n, r, cin, cout, taps, 1,
                                    1 if taps == 9 else 0, 0, 0.2, None, st))
    L.msr_debug_tc_counters(None)
    for _ in range(20):
        launch()
    torch.cuda.synchronize()
    dbg.zero_()
    L.msr_debug_tc_counters(dbg.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); launch(); e1.record(); torch.cuda.synchronize()
    L.msr_debug_tc_counters(None)
    d = dbg.view(148, 8).double().cpu()
    act = d[:, 1] > 0
    lead = d[:, 4] > 0
    ms = e0.elapsed_time(e1)
    flops = 2.0 * n * r * r * cout * taps * cin
    print(f"n={n} r={r} cin={cin} cout={cout} taps={taps}: {ms:.3f} ms {flops/ms/1e9:.0f} TFLOP/s | producer: wait_empty {d[act,0].mean():.0f} of {d[act,1].mean():.0f} cyc"
          f" | mma: wait_full {d[lead,2].mean():.0f} wait_acc {d[lead,3].mean():.0f} of {d[lead,4].mean():.0f} cyc | epi wait_acc {d[act,5].mean():.0f}", flush=True)

# fused SPADE epilogue (the real gamma | beta layers): rb6 spade_1 (C = 256, x at r/2), rb5 spade_1 (C = 512)
for (n, r, C) in [(16, 256, 256), (16, 128, 512), (16, 256, 128)]:
    a = torch.randn((n, r, r, 128), device="cuda").clamp_(min=0).to(torch.bfloat16)
    w = (torch.randn((2 * C, 1152), device="cuda") * 0.03).to(torch.bfloat16)
    b = torch.zeros(2 * C, device="cuda")
    x = torch.randn((n, r // 2, r // 2, C), device="cuda")
    mean = torch.zeros((1, C), device="cuda"); rstd = torch.ones((1, C), device="cuda")
    out = torch.empty((n, r, r, C), device="cuda", dtype=torch.bfloat16)
    st = _lib.stream_ptr()
    def launch():
        _lib.check(L.msr_op_spade_tc(a.data_ptr(), w.data_ptr(), b.data_ptr(), x.data_ptr(), 1, mean.data_ptr(),
                                     rstd.data_ptr(), n, out.data_ptr(), n, r, C, st))
    L.msr_debug_tc_counters(None)
    for _ in range(10):
        launch()
    torch.cuda.synchronize()
    dbg.zero_()
    L.msr_debug_tc_counters(dbg.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); launch(); e1.record(); torch.cuda.synchronize()
    L.msr_debug_tc_counters(None)
    d = dbg.view(148, 8).double().cpu()
    act = d[:, 1] > 0; lead = d[:, 4] > 0
    ms = e0.elapsed_time(e1)
    flops = 2.0 * n * r * r * 2 * C * 1152
    print(f"SPADE n={n} r={r} C={C}: {ms:.3f} ms {flops/ms/1e9:.0f} TFLOP/s | producer: wait_empty {d[act,0].mean():.0f} of {d[act,1].mean():.0f} cyc"
          f" | mma: wait_full {d[lead,2].mean():.0f} wait_acc {d[lead,3].mean():.0f} of {d[lead,4].mean():.0f} cyc | epi wait_acc {d[act,5].mean():.0f}", flush=True)
