"""Write-only and read-only HBM bandwidth reference points (torch fill / sum) for buffers far larger than L2."""
import torch
n = 1310720000 // 2
x = torch.empty(n, dtype=torch.bfloat16, device="cuda")
for name, fn, nbytes in [("fill (write-only) 1.31 GB", lambda: x.zero_(), n * 2), ("sum (read-only) 1.31 GB", lambda: x.view(torch.int32).sum(), n * 2),
                         ("copy 0.65 GB -> 0.65 GB", lambda: x[: n // 2].copy_(x[n // 2:]), n * 2)]:
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name:32s}: {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:7.0f} GB/s")
