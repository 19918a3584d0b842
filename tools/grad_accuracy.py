"""Gradient accuracy of the backward modes against the golden reference gradients (dev helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.cases import CASES, make_grad_out
from tests.test_forward_gpu import build
from tests.util import load_golden, scaled_err
SD = {"kernel": "_complex_conv._kernel", "pool_w": "_pooling.weights", "pool_b": "_pooling._bias", "alpha": "_compression.alpha",
      "delta": "_compression.delta", "root": "_compression.root", "ema_w": "_compression.ema._weights"}
for c in [c for c in CASES if c.grads]:
    case, x, prm, z = load_golden(c.name)
    fe = build(case, prm, "auto")
    out = fe(x.cuda())
    G = torch.from_numpy(make_grad_out(tuple(out.shape), case.seed)).cuda()
    (out * G).sum().backward()
    named = dict(fe.named_parameters())
    print(c.name, "LEAFK_BWD_FULL=" + os.environ.get("LEAFK_BWD_FULL", "0"),
          {k: f"{scaled_err(named[sk].grad.cpu().numpy().reshape(-1), z['grad_' + k].reshape(-1)):.1e}" for k, sk in SD.items()})
