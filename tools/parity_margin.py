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
for c in CASES:
    if c.grads or c.T > 200000:
        continue
    case, x, prm, z = load_golden(c.name)
    row = []
    for algo in ("fp32", "tc_full", "tc"):
        if not algo_available(case, algo):
            row.append("   n/a")
            continue
        fe = build(case, prm, algo)
        with torch.no_grad():
            out = fe(x.cuda()).cpu().numpy().astype(np.float64)
        ref = z["out"].astype(np.float64)
        s = float(np.max(np.abs(out - ref) / (1e-4 * np.abs(ref) + 1e-5)))
        worst[algo] = max(worst.get(algo, 0.0), s)
        row.append(f"{s:6.3f}")
    print(f"{c.name:22s} fp32 {row[0]}  tc_full {row[1]}  tc {row[2]}")
print("worst:", {k: round(v, 3) for k, v in worst.items()})
