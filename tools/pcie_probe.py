import torch, time
for mb in (1, 4, 16, 64):
    n = mb * 1024 * 1024 // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory(); d = torch.empty(n, dtype=torch.float32, device="cuda")
    for direction in ("h2d", "d2h"):
        for _ in range(3):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{direction} {mb:3d} MiB: {ms*1e3:8.1f} us  {mb*1.048576/ms:6.1f} GB/s")
# bidirectional overlap
n = 16 * 1024 * 1024 // 4
h1 = torch.empty(n).pin_memory(); h2 = torch.empty(n).pin_memory(); d1 = torch.empty(n, device="cuda"); d2 = torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print(f"bidirectional 16+16 MiB: {dt*1e6:.1f} us -> {2*16*1.048576/(dt*1e3):.1f} GB/s total")
import os; print("cpus", os.cpu_count())
