import torch, time
for mb in (34, 134, 168, 512):
    h = torch.empty(mb * 1024 * 1024 // 8, dtype=torch.float64).pin_memory()
    d = torch.empty_like(h, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e.record(); torch.cuda.synchronize()
    t = s.elapsed_time(e) / 10
    print(f"H2D pinned {mb} MiB: {t:.3f} ms = {h.numel()*8/t/1e6:.1f} GB/s")
