"""what a plain device copy of one C2-sized plane achieves on this GPU (torch b.copy_(a), CUDA events, rotating buffers)"""
import torch
for n_mib, rot in ((64, 6), (256, 4), (2048, 2)):
    n = n_mib * 1024 * 1024 // 4
    a = [torch.randn(n, device="cuda") for _ in range(rot)]
    b = [torch.empty(n, device="cuda") for _ in range(rot)]
    for i in range(5): b[i % rot].copy_(a[i % rot])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 40
    e0.record()
    for i in range(iters): b[i % rot].copy_(a[i % rot])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"copy {n_mib} MiB -> {n_mib} MiB: {ms*1e3:.1f} us  {2*n*4/ms/1e6:.0f} GB/s (read+write)")
