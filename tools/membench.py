import torch, time
n = 20 * 1024**3 // 8
x = torch.empty(n, dtype=torch.float64, device="cuda")
y = torch.empty(n // 2, dtype=torch.float64, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.fill_(1.5)); print("fill 20GiB: %.3f ms  %.0f GB/s" % (ms, x.numel() * 8 / ms / 1e6))
ms = t(lambda: x.zero_()); print("memset 20GiB: %.3f ms  %.0f GB/s" % (ms, x.numel() * 8 / ms / 1e6))
ms = t(lambda: y.copy_(x[: n // 2])); print("copy 10GiB->10GiB: %.3f ms  %.0f GB/s (r+w)" % (ms, 2 * y.numel() * 8 / ms / 1e6))
ms = t(lambda: x.sum()); print("read 20GiB: %.3f ms  %.0f GB/s" % (ms, x.numel() * 8 / ms / 1e6))
