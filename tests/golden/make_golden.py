#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the REAL reference.

Runs only in the build container, where /root/reference is mounted; the GPU box never runs
this (it reads the committed .npz files).  Usage:

    python tests/golden/make_golden.py            # all cases
    python tests/golden/make_golden.py cfg1_default T161

For every case in tests/cases.py the unmodified ``leaf_pytorch.frontend.Leaf`` is built with
the case's constructor arguments, its parameters are overwritten with the case's parameter
set via ``load_state_dict``, and ``forward`` (plus, for gradient cases, ``backward`` of
``sum(out*G)``) is run on CPU in float32.  Stored per case: the parameters, the input (or its
sha256 when large), the output, the floored pooled energies ``p`` and the gradients.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from tests.cases import CASES, CASES_BY_NAME, make_signal, perturb_params, make_grad_out, sha256  # noqa: E402

from leaf_pytorch.frontend import Leaf as RefLeaf  # noqa: E402  (the reference)

SD_KEYS = {
    "kernel": "_complex_conv._kernel",
    "pool_w": "_pooling.weights",
    "pool_b": "_pooling._bias",
    "alpha": "_compression.alpha",
    "delta": "_compression.delta",
    "root": "_compression.root",
    "ema_w": "_compression.ema._weights",
}


def build_reference(case):
    with contextlib.redirect_stdout(io.StringIO()):
        ref = RefLeaf(n_filters=case.F, sample_rate=case.sr, window_len=case.wlen,
                      window_stride=case.wstride, init_min_freq=case.min_freq,
                      init_max_freq=case.max_freq, pcen_compression=case.compression,
                      use_legacy_complex=case.legacy)
    return ref


def default_params(ref, case):
    sd = ref.state_dict()
    out = {}
    for k, sk in SD_KEYS.items():
        if sk in sd:
            out[k] = sd[sk].detach().reshape(-1, 2).numpy().copy() if k == "kernel" \
                else sd[sk].detach().reshape(-1).numpy().copy()
    if not case.compression:      # no PCEN submodule: give the perturbation something to overwrite
        F = case.F
        out.update(alpha=np.full(F, 0.96, np.float32), delta=np.full(F, 2.0, np.float32),
                   root=np.full(F, 2.0, np.float32), ema_w=np.full(F, 0.04, np.float32))
    return out


def load_params(ref, prm, case):
    sd = {}
    for k, sk in SD_KEYS.items():
        if not case.compression and sk.startswith("_compression"):
            continue
        v = torch.from_numpy(prm[k])
        if k == "pool_w":
            v = v.reshape(1, 1, -1, 1)
        sd[sk] = v
    ref.load_state_dict(sd)


def run_case(case):
    ref = build_reference(case)
    init = default_params(ref, case)
    prm = perturb_params(init, case.params, case.K, case.seed)
    load_params(ref, prm, case)
    x = make_signal(case.signal, case.B, case.T, case.seed)
    xt = torch.from_numpy(x)
    store = {"init_kernel": init["kernel"], "x_sha256": np.array(sha256(x))}
    for k, v in prm.items():
        store["prm_" + k] = v
    if case.store_x:
        store["x"] = x
    if case.grads:
        out = ref(xt)
        G = make_grad_out(tuple(out.shape), case.seed)
        (out * torch.from_numpy(G)).sum().backward()
        for k, sk in SD_KEYS.items():
            prmt = dict(ref.named_parameters())[sk]
            store["grad_" + k] = prmt.grad.detach().reshape(prm[k].shape).numpy().copy()
        out = out.detach()
    else:
        with torch.no_grad():
            out = ref(xt)
    with torch.no_grad():          # floored pooled energies, via the reference's own sub-modules
        p = ref._pooling(ref._activation(ref._complex_conv(xt)))
        p = torch.maximum(p, torch.tensor(1e-5))
    store["out"] = out.numpy()
    if case.T <= 20000:
        store["p"] = p.numpy()
    path = os.path.join(HERE, case.name + ".npz")
    np.savez_compressed(path, **store)
    print(f"{case.name:20s} F={case.F} K={case.K} H={case.H} x{tuple(x.shape)} -> out{tuple(out.shape)} "
          f"{os.path.getsize(path) / 1024:.0f} KiB")


def main(argv):
    torch.manual_seed(0)
    names = argv or [c.name for c in CASES]
    for n in names:
        run_case(CASES_BY_NAME[n])


if __name__ == "__main__":
    main(sys.argv[1:])
