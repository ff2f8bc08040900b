"""device-to-device copy bandwidth at a few sizes (context for the HBM roofline of short kernels)"""
import torch
for n in (4096 * 4096, 4 * 4096 * 4096, 16 * 4096 * 4096):
    a = torch.rand(n, dtype=torch.float64, device="cuda"); b = torch.empty_like(a)
    for _ in range(5): b.copy_(a)
    torch.cuda.synchronize()
    best = 1e9; tot = 0
    for _ in range(20):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); b.copy_(a); e.record(); torch.cuda.synchronize()
        t = s.elapsed_time(e); best = min(best, t); tot += t
    print(f"copy {n*8/1e6:8.1f} MB -> same: best {best*1e3:7.1f} us ({16*n/best/1e6:7.1f} GB/s), mean {tot/20*1e3:7.1f} us")
    # write-only (fill)
    for _ in range(5): b.fill_(1.0)
    torch.cuda.synchronize(); best = 1e9
    for _ in range(20):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); b.fill_(2.0); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    print(f"fill {n*8/1e6:8.1f} MB: best {best*1e3:7.1f} us ({8*n/best/1e6:7.1f} GB/s)")
