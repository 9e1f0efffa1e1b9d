import torch, time
x = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for name, (a, b) in {"H2D": (d, x), "D2H": (x, d)}.items():
    for _ in range(2): a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): a.copy_(b, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(name, "pinned 256 MiB x10:", round(10 * x.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1), "GB/s")
