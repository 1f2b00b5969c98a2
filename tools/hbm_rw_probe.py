import torch
def t(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
N = 2 * 1024**3
a = torch.empty(N, dtype=torch.uint8, device="cuda")
b = torch.empty(N, dtype=torch.uint8, device="cuda")
af = a.view(torch.float32)
ms = t(lambda: a.zero_()); print("memset zero_ 2 GiB: %.3f ms  %.0f GB/s write" % (ms, N / ms / 1e6))
ms = t(lambda: af.fill_(1.5)); print("fill_ f32 2 GiB: %.3f ms  %.0f GB/s write" % (ms, N / ms / 1e6))
ms = t(lambda: b.copy_(a)); print("copy 2 GiB: %.3f ms  %.0f GB/s read+write" % (ms, 2 * N / ms / 1e6))
ms = t(lambda: af.sum()); print("sum f32 2 GiB: %.3f ms  %.0f GB/s read" % (ms, N / ms / 1e6))
ms = t(lambda: torch.cuda.current_stream().synchronize() or torch.cuda._sleep(0) or af.max()); print("max f32 2 GiB: %.3f ms  %.0f GB/s read" % (ms, N / ms / 1e6))
big = af.view(-1, 4096)
row = torch.randn(4096, device="cuda")
ms = t(lambda: torch.add(row.expand_as(big), 1.0, out=big)); print("broadcast row + 1 -> 2 GiB (non-constant write, tiny read): %.3f ms  %.0f GB/s write" % (ms, N / ms / 1e6))
col = torch.randn(big.shape[0], 1, device="cuda")
ms = t(lambda: torch.add(col.expand_as(big), 1.0, out=big)); print("broadcast col + 1 -> 2 GiB: %.3f ms  %.0f GB/s write" % (ms, N / ms / 1e6))
h = af[: N // 8]
ms = t(lambda: torch.neg(h, out=af[N // 8: N // 4])); print("neg 1 GiB -> 1 GiB: %.3f ms  %.0f GB/s read+write" % (ms, N / 2 / ms / 1e6))
