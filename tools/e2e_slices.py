"""Development helper: throughput of the pipelined host path (HostPipeline) against the number of H2D slices / depth."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leaf_pytorch_b200 as L

B, T, F = 256, 16000, 40
fe = L.Leaf().cuda()
n_frames = fe.num_frames(T)
g = torch.Generator().manual_seed(1)
hosts = [(torch.randn(B, 1, T, generator=g).clamp_(-4, 4) / 4).pin_memory() for _ in range(6)]
steps = 60
for depth in (2, 3):
    for n_slices in (1, 2, 4, 8, 16):
        pipe = L.HostPipeline(fe, B, T, depth=depth, n_slices=n_slices)
        outs = [torch.empty((B, F, n_frames), dtype=torch.float32).pin_memory() for _ in range(depth)]
        for i in range(depth):
            pipe.result(pipe.submit(hosts[i], outs[i]))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pend = []
        for i in range(steps):
            pend.append(pipe.submit(hosts[i % len(hosts)], outs[i % depth]))
            if len(pend) >= depth:
                pipe.result(pend.pop(0))
        while pend:
            pipe.result(pend.pop(0))
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        pipe.close()
        print(f"depth {depth} n_slices {n_slices:2d}: {dt*1e3:.4f} ms/step  {B*T/16000/dt:9.0f} audio-s/s  H2D {B*T*4/dt/1e9:.1f} GB/s")
