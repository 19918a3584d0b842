"""Development helper: how much of the parity tolerance (|d| <= 1e-4 |ref| + 1e-5) each conv kernel uses on the golden
vectors of the real reference: max over elements of |d| / (1e-4 |ref| + 1e-5), per case."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tests.cases import CASES
from tests.util import load_golden
from tests.test_forward_gpu import build, algo_available

worst = {}
worst_rel = {}
for c in CASES:
    if c.grads or c.T > 200000:
        continue
    case, x, prm, z = load_golden(c.name)
    row = []
    for algo in ("fp32", "tc_full", "tc"):
        if not algo_available(case, algo):
            row.append("   n/a             ")
            continue
        fe = build(case, prm, algo)
        with torch.no_grad():
            out = fe(x.cuda()).cpu().numpy().astype(np.float64)
        ref = z["out"].astype(np.float64)
        s = float(np.max(np.abs(out - ref) / (1e-4 * np.abs(ref) + 1e-5)))
        worst[algo] = max(worst.get(algo, 0.0), s)
        big = np.abs(ref) >= 1e-2 * np.abs(ref).max()
        rel = float(np.max(np.abs(out - ref)[big] / np.abs(ref)[big])) if np.any(big) else 0.0
        worst_rel[algo] = max(worst_rel.get(algo, 0.0), rel)
        row.append(f"{s:6.3f} (rel {rel:.1e})")
    print(f"{c.name:22s} fp32 {row[0]}  tc_full {row[1]}  tc {row[2]}")
print("worst share of the 1e-4*|ref| + 1e-5 tolerance:", {k: round(v, 3) for k, v in worst.items()})
print("worst pure relative error on elements with |ref| >= 1% of the case's peak (tolerance 1e-4):",
      {k: float(f"{v:.2e}") for k, v in worst_rel.items()})
