"""Helpers shared by the parity tests (golden loading, error metrics)."""
from __future__ import annotations

import os
from typing import Dict

import numpy as np
import torch

from tests.cases import CASES_BY_NAME, Case, make_signal, sha256

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PRM_KEYS = ("kernel", "pool_w", "pool_b", "alpha", "delta", "root", "ema_w")


def load_golden(name: str):
    """-> (case, x (B,1,T) float32 tensor, params dict of tensors, npz)"""
    case: Case = CASES_BY_NAME[name]
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    x = z["x"] if "x" in z.files else make_signal(case.signal, case.B, case.T, case.seed)
    assert sha256(x) == str(z["x_sha256"]), f"input of case {name} does not reproduce"
    prm: Dict[str, torch.Tensor] = {k: torch.from_numpy(z["prm_" + k]) for k in PRM_KEYS}
    return case, torch.from_numpy(x), prm, z


def rel_err(got, want, floor: float = 0.0) -> float:
    """max |got-want| / max(|want|, floor)"""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return float(np.max(np.abs(got - want) / np.maximum(np.abs(want), floor))) if want.size else 0.0


def scaled_err(got, want) -> float:
    """max |got-want| / max |want|  (error relative to the tensor's scale)"""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    return float(np.max(np.abs(got - want)) / max(np.max(np.abs(want)), 1e-30)) if want.size else 0.0
