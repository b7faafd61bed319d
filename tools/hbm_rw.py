"""Read-only / write-only / copy HBM bandwidth with plain torch ops (context for the K2 and attention rooflines)."""
import torch
n = 2 << 30                                   # 8 GiB of fp32
x = torch.empty(n, dtype=torch.float32, device="cuda")
y = torch.empty(n, dtype=torch.float32, device="cuda")
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
gb = n * 4 / 1e9
print("write-only (fill_)   %.0f GB/s" % (gb / timed(lambda: x.fill_(1.0))))
print("read-only  (sum)     %.0f GB/s" % (gb / timed(lambda: x.sum())))
print("copy (read + write)  %.0f GB/s" % (2 * gb / timed(lambda: y.copy_(x))))
# 1 read : 4 write, like the projection GEMM (reads 1 KB, writes 4 KB per row)
z = torch.empty(n // 4, dtype=torch.float32, device="cuda")
print("1:4 read:write (repeat_interleave-like expand copy) %.0f GB/s" % (1.25 * gb / timed(lambda: x.view(4, -1).copy_(z.view(1, -1).expand(4, -1)))))
