"""Development A/B timing helper (not the contract bench): interleaves the conv kernels round by round so that
clock / power drift hits every variant alike.  Env: F, B, T, ALGOS (comma list), ROUNDS, ITERS."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leaf_pytorch_b200 as L


def main():
    F = int(os.environ.get("F", 40)); B = int(os.environ.get("B", 256)); T = int(os.environ.get("T", 16000))
    algos = os.environ.get("ALGOS", "tc_full,tc").split(",")
    rounds = int(os.environ.get("ROUNDS", 8)); iters = int(os.environ.get("ITERS", 5))
    g = torch.Generator().manual_seed(1234)
    xs = [(torch.randn(B, 1, T, generator=g).clamp_(-4, 4) / 4).cuda() for _ in range(4)]
    fes = {}
    for a in algos:
        fes[a] = L.Leaf(n_filters=F, algo=a).cuda()
    times = {a: [] for a in algos}
    with torch.no_grad():
        for a in algos:
            for _ in range(3):
                fes[a](xs[0])
        torch.cuda.synchronize()
        for r in range(rounds):
            for a in algos:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(iters):
                    fes[a](xs[i % len(xs)])
                e1.record()
                torch.cuda.synchronize()
                times[a].append(e0.elapsed_time(e1) / iters)
    for a in algos:
        ms = statistics.median(times[a])
        print(f"F={F} B={B} T={T} algo={a:8s} median {ms:7.4f} ms  min {min(times[a]):7.4f}  {B*T/16000/(ms*1e-3):10.0f} audio-s/s")


if __name__ == "__main__":
    main()
