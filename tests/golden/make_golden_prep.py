#!/usr/bin/env python
"""Golden vectors for the on-the-fly clip preparation, generated with the REFERENCE'S OWN transform
classes.  utilities/data/raw_transforms.py cannot be imported here (it needs `augment`, `torch_audiomentations`, ...),
so the three classes the evaluation / training pipelines use -- PadToSize, CenterCrop, RandomCrop
(raw_transforms.py:121-183) -- are cut out of the unmodified source file with `ast` and executed as they are; the
reference Leaf then runs on the prepared batch.  PeakNormalization wraps torch_audiomentations (not installed): its rule
(divide by max|x| when it exceeds 1) is applied by hand here and said so in the fixture.

    python tests/golden/make_golden_prep.py      ->  tests/golden/prep_eval.npz, tests/golden/prep_train.npz
"""
import ast
import contextlib
import io
import os
import random
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from leaf_pytorch.frontend import Leaf as RefLeaf  # noqa: E402  (the reference)

SRC = "/root/reference/utilities/data/raw_transforms.py"
tree = ast.parse(open(SRC).read())
ns = {"torch": torch, "np": np, "random": random}
for node in tree.body:
    if isinstance(node, ast.ClassDef) and node.name in ("PadToSize", "CenterCrop", "RandomCrop"):
        exec(compile(ast.Module(body=[node], type_ignores=[]), SRC, "exec"), ns)
PadToSize, CenterCrop = ns["PadToSize"], ns["CenterCrop"]

N_SAMPLES = 8000
LENS = [8000, 12001, 5000, 8001, 11000, 1, 7999, 3]


def peak_normalize(t):      # torch_audiomentations.PeakNormalization(apply_to="only_too_loud_sounds"), restated
    peak = t.abs().max()
    return t / peak if bool(peak > 1.0) else t


def main():
    rng = np.random.Generator(np.random.PCG64(2024))
    raws = [(rng.standard_normal(n) * (0.2 if i % 2 else 0.6)).astype(np.float32) for i, n in enumerate(LENS)]
    with contextlib.redirect_stdout(io.StringIO()):
        ref = RefLeaf()
    pad = np.zeros((len(LENS), max(LENS)), np.float32)
    for i, r in enumerate(raws):
        pad[i, :len(r)] = r
    # evaluation pipeline: leaf_supervised_transforms(is_train=False) = PadToSize(size,'wrap'), CenterCrop, PeakNormalization
    ev = []
    for r in raws:
        t = torch.from_numpy(r).reshape(1, -1)
        t = CenterCrop(N_SAMPLES)(PadToSize(N_SAMPLES, "wrap")(t))
        ev.append(peak_normalize(t.reshape(-1)))
    ev = torch.stack(ev).unsqueeze(1)
    # training pipeline up to the crop: PadToSize(size,'constant') then RandomCrop -- the crop offsets a caller's RNG drew
    starts = []
    tr = []
    for r in raws:
        t = PadToSize(N_SAMPLES, "constant")(torch.from_numpy(r).reshape(1, -1))
        s0 = int(rng.integers(0, t.shape[1] - N_SAMPLES + 1))
        starts.append(s0)
        tr.append(peak_normalize(t[:, s0:s0 + N_SAMPLES].reshape(-1)))
    tr = torch.stack(tr).unsqueeze(1)
    with torch.no_grad():
        out_ev, out_tr = ref(ev), ref(tr)
    common = dict(raw=pad, lengths=np.array(LENS, np.int32), n_samples=np.int32(N_SAMPLES))
    np.savez_compressed(os.path.join(HERE, "prep_eval.npz"), prepared=ev.numpy(), out=out_ev.numpy(), **common)
    np.savez_compressed(os.path.join(HERE, "prep_train.npz"), prepared=tr.numpy(), out=out_tr.numpy(),
                        starts=np.array(starts, np.int32), **common)
    print("eval  prepared", tuple(ev.shape), "peak", float(ev.abs().max()), "out", tuple(out_ev.shape))
    print("train prepared", tuple(tr.shape), "starts", starts)


if __name__ == "__main__":
    main()
