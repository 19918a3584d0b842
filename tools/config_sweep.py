"""Per-GPU timings of the five BASELINE.json configs (development record; the contract bench is bench.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leaf_pytorch_b200 as L
from leaf_pytorch_b200.streaming import forward_chunked


def timeit(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def batch(B, T, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, 1, T, generator=g).clamp_(-4, 4) / 4).cuda()


rows = []
with torch.no_grad():
    fe = L.Leaf().cuda()
    x = batch(4, 16000); ms = timeit(lambda: fe(x)); rows.append(("cfg1 F=40 B=4 x 1 s fwd", 4, ms))
    x = batch(256, 16000); ms = timeit(lambda: fe(x)); rows.append(("cfg2 F=40 B=256 x 1 s fwd", 256, ms))
    x = batch(64, 160000); ms = timeit(lambda: fe(x)); rows.append(("cfg4 F=40 B=64 x 10 s fwd (per-GPU shard of 512)", 640, ms))
    fe64 = L.Leaf(n_filters=64).cuda()
    x = batch(8, 960000)
    ms = timeit(lambda: forward_chunked(fe64, x, chunk_frames=1000), iters=5)
    rows.append(("cfg5 F=64 B=8 x 60 s fwd, 10 s chunks with carried PCEN state (per-GPU shard of 64)", 480, ms))
    ms = timeit(lambda: fe64(x), iters=5)
    rows.append(("cfg5 same, un-chunked", 480, ms))
fe80 = L.Leaf(n_filters=80).cuda()
x = batch(1024, 16000)
G = torch.randn(1024, 80, 100, generator=torch.Generator().manual_seed(1235)).cuda()


def step():
    fe80.zero_grad(set_to_none=True)
    fe80(x).backward(G)


ms = timeit(step, iters=5); rows.append(("cfg3 F=80 B=1024 x 1 s fwd+bwd", 1024, ms))
with torch.no_grad():
    ms = timeit(lambda: fe80(x), iters=5); rows.append(("cfg3 forward only", 1024, ms))
for name, secs, ms in rows:
    print(f"{name:90s} {ms:9.3f} ms  {secs / (ms * 1e-3):12.0f} audio-s/s")
