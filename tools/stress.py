"""Stress / soak: thousands of launches over random shapes; every result must reproduce bit for bit and nothing
may trap or hang (mbarrier / cluster protocol soak).  Exit code != 0 on any mismatch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import leaf_pytorch_b200 as L

secs = float(os.environ.get("SECS", 40))
rng = np.random.default_rng(0)
fes = {F: L.Leaf(n_filters=F).cuda() for F in (40, 64, 80)}
shapes = [(int(rng.integers(1, 300)), int(rng.integers(1, 40000))) for _ in range(12)] + [(256, 16000), (1, 1), (7, 1023), (3, 1025)]
cases = []
for i, (B, T) in enumerate(shapes):
    F = (40, 64, 80)[i % 3]
    x = torch.randn(B, 1, T, generator=torch.Generator().manual_seed(i)).cuda()
    with torch.no_grad():
        ref = fes[F](x).clone()
    cases.append((F, x, ref))
xg = torch.randn(8, 1, 9000, generator=torch.Generator().manual_seed(99)).cuda()
G = torch.randn(8, 40, fes[40].num_frames(9000), generator=torch.Generator().manual_seed(98)).cuda()
fes[40].zero_grad(); (fes[40](xg) * G).sum().backward(); gref = [p.grad.clone() for p in fes[40].parameters()]
t0 = time.time(); n = 0; bad = 0
while time.time() - t0 < secs:
    for F, x, ref in cases:
        with torch.no_grad():
            out = fes[F](x)
        if not torch.equal(out, ref): bad += 1
        n += 1
    fes[40].zero_grad(set_to_none=True); (fes[40](xg) * G).sum().backward()
    if not all(torch.equal(a.grad, b) for a, b in zip(fes[40].parameters(), gref)): bad += 1
    xh = cases[-4][1].cpu().pin_memory()
    if not torch.equal(fes[40].forward_host(xh, n_slices=5), cases[-4][2].cpu()): bad += 1
    n += 2
torch.cuda.synchronize()
print(f"stress: {n} launches in {time.time() - t0:.1f}s, mismatches: {bad}")
sys.exit(1 if bad else 0)
