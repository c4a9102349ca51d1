"""read-only / write-only / copy HBM bandwidth with library kernels (context for the read-stream kernels: gate_wgrad, int_bwd_a_stream)"""
import torch
n = 1 << 30   # 4 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device="cuda").normal_()
b = torch.empty_like(a)
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: a.sum());           print("read-only  (sum over 4 GiB):   %.3f ms  %.0f GB/s" % (ms, 4 * n / ms / 1e6))
ms = t(lambda: torch.max(a));      print("read-only  (max over 4 GiB):   %.3f ms  %.0f GB/s" % (ms, 4 * n / ms / 1e6))
ms = t(lambda: b.fill_(1.0));      print("write-only (fill 4 GiB):       %.3f ms  %.0f GB/s" % (ms, 4 * n / ms / 1e6))
ms = t(lambda: b.copy_(a));        print("copy       (4 GiB -> 4 GiB):   %.3f ms  %.0f GB/s read+write" % (ms, 8 * n / ms / 1e6))
