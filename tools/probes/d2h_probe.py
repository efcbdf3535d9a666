import torch, time
for mb in (8, 64):
    n = mb * 1024 * 1024 // 8
    d = torch.zeros(n, dtype=torch.float64, device='cuda')
    h = torch.empty(n, dtype=torch.float64).pin_memory()
    for _ in range(3): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"D2H {mb} MiB pinned: {ms*1e3:.1f} us -> {n*8/ms/1e6:.1f} GB/s")
    e0.record()
    for _ in range(20): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"H2D {mb} MiB pinned: {ms*1e3:.1f} us -> {n*8/ms/1e6:.1f} GB/s")
