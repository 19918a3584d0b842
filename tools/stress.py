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
extra = {"fp32": L.Leaf(n_filters=12, algo="fp32").cuda(), "long": torch.randn(3, 1, 200000, generator=torch.Generator().manual_seed(5)).cuda(),
         "raw": torch.randn(5, 1, 9000, generator=torch.Generator().manual_seed(6)).cuda() * 0.5,
         "lens": torch.tensor([9000, 100, 6000, 5999, 7777]), "hx": (torch.randn(16, 1, 8000, generator=torch.Generator().manual_seed(7)) / 4).pin_memory()}
with torch.no_grad():
    extra["prep_ref"] = fes[40].forward_prepared(extra["raw"], 6000, raw_lengths=extra["lens"]).clone()
    extra["pipe_ref"] = fes[40](extra["hx"].cuda()).cpu()
extra["pipe"] = L.HostPipeline(fes[40], 16, 8000, depth=2, n_slices=3)
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
    # round-2 paths: training forward at F=80 (5 groups), generic FP32 backward with the waveform gradient, long rows
    # (block-per-row PCEN), clips prepared on the fly, host pipeline with several batches in flight
    if n % 7 == 0:
        f80 = fes[80]; f80.zero_grad(set_to_none=True)
        x8 = cases[2][1][:4] if cases[2][1].shape[0] >= 4 else cases[2][1]
        f80(x8).sum().backward()
        if not all(torch.isfinite(p.grad).all() for p in f80.parameters()): bad += 1
        xr = xg[:2].clone().requires_grad_(True)
        fe32 = extra["fp32"]; fe32.zero_grad(set_to_none=True); fe32(xr).sum().backward()
        if not torch.isfinite(xr.grad).all(): bad += 1
        with torch.no_grad():
            o1 = fes[64](extra["long"]); o2 = fes[64](extra["long"])
            if not torch.equal(o1, o2): bad += 1
            p1 = fes[40].forward_prepared(extra["raw"], 6000, raw_lengths=extra["lens"])
            if not torch.equal(p1, extra["prep_ref"]): bad += 1
        pipe = extra["pipe"]
        tk = [pipe.submit(extra["hx"]) for _ in range(3)]
        for t in tk:
            if not torch.equal(pipe.result(t), extra["pipe_ref"]): bad += 1
        n += 8
torch.cuda.synchronize()
print(f"stress: {n} launches in {time.time() - t0:.1f}s, mismatches: {bad}")
sys.exit(1 if bad else 0)
