"""Development timing helper (not the contract bench): forward of cfg2 with each conv kernel."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import leaf_pytorch_b200 as L
import leaf_pytorch_b200.functional as LF


def timeit(fn, warm=5, iters=20):
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


def main():
    F = int(os.environ.get("F", 40)); B = int(os.environ.get("B", 256)); T = int(os.environ.get("T", 16000))
    algos = os.environ.get("ALGOS", "fp32,tc").split(",")
    g = torch.Generator().manual_seed(1234)
    x = (torch.randn(B, 1, T, generator=g).clamp_(-4, 4) / 4).cuda()
    fe = L.Leaf(n_filters=F).cuda()
    outs = {}
    for algo in algos:
        fe.algo = algo
        with torch.no_grad():
            try:
                outs[algo] = fe(x)
                torch.cuda.synchronize()
            except Exception as ex:
                print(algo, "FAILED:", ex); continue
            ms = timeit(lambda: fe(x))
        print(f"algo={algo:5s} F={F} B={B} T={T}: {ms:8.3f} ms/forward  {B*T/16000/(ms*1e-3):12.0f} audio-s/s  "
              f"{2*2*F*fe.spec.K*T*B/(ms*1e-3)/1e12:7.2f} TFLOP/s(alg)")
    names = list(outs)
    for i in range(len(names)):
        for j in range(i + 1, len(names)):
            a, b = outs[names[i]].cpu().numpy(), outs[names[j]].cpu().numpy()
            d = np.abs(a - b)
            print("%s vs %s: max|d| %.3e  max rel %.3e  finite=%s" % (names[j], names[i], d.max(), (d / np.maximum(np.abs(a), 1e-3)).max(), np.isfinite(b).all()))


if __name__ == "__main__":
    main()
