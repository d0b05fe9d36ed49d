"""Write-only HBM bandwidth on this GPU (the Gram kernels only write): torch fill_ and cudaMemset over 16 GB."""
import json
import torch

x = torch.empty(2 * 1024 ** 3, dtype=torch.float64, device="cuda:0")      # 16 GiB
out = {}
for name, fn in (("fill_kernel", lambda: x.fill_(1.0)), ("memset", lambda: x.zero_())):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[name + "_GBps"] = x.numel() * 8 / (best * 1e-3) / 1e9
print(json.dumps(out))
