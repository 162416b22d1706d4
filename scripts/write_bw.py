"""GPU: pure-write and pure-read HBM bandwidth next to the copy figure of MEASURED_PEAKS.json (what a write-dominated
kernel such as imu_fc1_mma_kernel can be compared with)."""
import torch
x = torch.empty(1 << 31, dtype=torch.uint8, device="cuda")      # 2 GiB
y = torch.empty(1 << 31, dtype=torch.uint8, device="cuda")
def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
xf = x.view(torch.float32)
ms = timed(lambda: x.zero_()); print(f"memset (write only)   {x.numel() / ms / 1e6:8.1f} GB/s")
ms = timed(lambda: xf.fill_(1.5)); print(f"fill kernel (write)   {x.numel() / ms / 1e6:8.1f} GB/s")
ms = timed(lambda: y.copy_(x)); print(f"copy (read + write)   {2 * x.numel() / ms / 1e6:8.1f} GB/s")
ms = timed(lambda: xf.sum()); print(f"sum reduction (read)  {x.numel() / ms / 1e6:8.1f} GB/s")
