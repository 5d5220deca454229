import torch
x = torch.empty(1 << 30, dtype=torch.float32, device="cuda").normal_()   # 4 GiB
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.sum())
print("read-only sum 4 GiB: %.3f ms -> %.0f GB/s" % (ms, x.numel() * 4 / ms / 1e6))
y = torch.empty_like(x)
ms = t(lambda: y.copy_(x))
print("copy 4 GiB: %.3f ms -> %.0f GB/s (read+write)" % (ms, 2 * x.numel() * 4 / ms / 1e6))
ms = t(lambda: y.zero_())
print("write-only 4 GiB: %.3f ms -> %.0f GB/s" % (ms, x.numel() * 4 / ms / 1e6))
