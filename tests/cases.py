"""Shared, deterministic definitions of the parity cases (inputs + parameter sets).

Everything here is generated from numpy's PCG64 streams, so the same arrays are rebuilt
bit-for-bit on the CPU container (where the golden vectors are made from the real reference)
and on the GPU box (where /root/reference does not exist).  Large inputs are therefore not
stored in tests/golden/: only their sha256, the parameters and the reference outputs are.
"""
from __future__ import annotations

import hashlib
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np


@dataclass(frozen=True)
class Case:
    name: str
    F: int = 40
    sr: int = 16000
    wlen: float = 25.0
    wstride: float = 10.0
    B: int = 2
    T: int = 16000
    params: str = "default"        # default | perturbed | perturbed2
    signal: str = "randn"          # randn | speech | quiet | pcm16 | loud | zeros | impulse
    seed: int = 0
    legacy: bool = False
    compression: bool = True
    grads: bool = False
    min_freq: float = 60.0
    max_freq: float = 7800.0
    store_x: bool = True
    tags: tuple = field(default_factory=tuple)

    @property
    def K(self) -> int:
        return int(self.sr * self.wlen // 1000 + 1)

    @property
    def H(self) -> int:
        return int(self.sr * self.wstride // 1000)


CASES: List[Case] = [
    # BASELINE.json configs[0]: default Leaf, batch 4 x 1 s
    Case("cfg1_default", B=4, T=16000, seed=0),
    Case("cfg1_legacy", B=2, T=16000, seed=0, legacy=True),
    Case("perturbed_F40", B=2, T=16000, params="perturbed", seed=1),
    Case("perturbed2_F40", B=2, T=8000, params="perturbed2", seed=2),
    Case("speech_default", B=2, T=8000, signal="speech", seed=3),
    Case("quiet_perturbed", B=2, T=8000, signal="quiet", params="perturbed", seed=4),
    Case("pcm16_default", B=2, T=8000, signal="pcm16", seed=5),
    Case("loud_default", B=1, T=8000, signal="loud", seed=6),
    Case("zeros_default", B=1, T=4000, signal="zeros", seed=7),
    Case("impulse_perturbed", B=1, T=4000, signal="impulse", params="perturbed", seed=8),
    # ragged / edge lengths
    Case("T1", B=2, T=1, seed=10),
    Case("T159", B=2, T=159, params="perturbed", seed=11),
    Case("T161", B=2, T=161, seed=12),
    Case("T400", B=3, T=400, params="perturbed", seed=13),
    Case("T1024", B=1, T=1024, seed=14),
    Case("T1025", B=1, T=1025, params="perturbed", seed=15),
    Case("T15999", B=1, T=15999, seed=16),
    Case("T16001", B=1, T=16001, params="perturbed", seed=17),
    # other filter counts / window geometries
    Case("F64", F=64, B=2, T=4000, seed=20),
    Case("F80", F=80, B=2, T=4000, params="perturbed", seed=21),
    Case("F8", F=8, B=2, T=2000, params="perturbed", seed=22),
    Case("sr22050_evenK", sr=22050, B=2, T=5000, seed=23, max_freq=11000.0),
    Case("sr8000", sr=8000, B=2, T=4000, seed=24, max_freq=3800.0),
    Case("win32_hop8", wlen=32.0, wstride=8.0, B=1, T=6000, params="perturbed", seed=25),
    Case("win10_hop10", wlen=10.0, wstride=10.0, B=1, T=6000, seed=26),
    Case("nopcen", B=2, T=4000, compression=False, params="perturbed", seed=27),
    # long clips (inputs regenerated from the seed, not stored)
    Case("long10s", B=1, T=160000, seed=30, store_x=False),
    Case("long60s_F64", F=64, B=1, T=960000, seed=31, store_x=False, tags=("chunked",)),
    # gradients
    Case("grad_default", B=2, T=4000, seed=40, grads=True),
    Case("grad_perturbed", B=2, T=4000, params="perturbed", seed=41, grads=True),
    Case("grad_F80", F=80, B=1, T=2000, params="perturbed2", seed=42, grads=True),
    Case("grad_evenK", sr=22050, B=1, T=3000, params="perturbed", seed=43, grads=True, max_freq=11000.0),
]

CASES_BY_NAME: Dict[str, Case] = {c.name: c for c in CASES}


def make_signal(kind: str, B: int, T: int, seed: int) -> np.ndarray:
    """(B,1,T) float32 waveform."""
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    if kind == "randn":
        x = np.clip(rng.standard_normal((B, 1, T)), -4, 4) / 4
    elif kind in ("speech", "quiet", "pcm16", "loud"):
        t = np.arange(T, dtype=np.float64) / 16000.0
        x = np.zeros((B, 1, T))
        for b in range(B):
            for _ in range(6):
                f0 = rng.uniform(80, 3500)
                amp = rng.uniform(0.05, 1.0)
                dec = rng.uniform(0.5, 30.0)
                ph = rng.uniform(0, 2 * math.pi)
                x[b, 0] += amp * np.exp(-dec * t) * np.sin(2 * math.pi * f0 * t + ph)
            x[b, 0] += 0.01 * rng.standard_normal(T)
            x[b, 0] /= max(np.abs(x[b, 0]).max(), 1e-9)          # peak normalisation (raw_transforms.py:334-344)
        if kind == "quiet":
            x *= 1e-3
        elif kind == "pcm16":
            x = np.round(x * 32767.0) / 32768.0
        elif kind == "loud":
            x = np.round(x * 32767.0)
    elif kind == "zeros":
        x = np.zeros((B, 1, T))
    elif kind == "impulse":
        x = np.zeros((B, 1, T))
        x[:, 0, T // 3] = 1.0
        x[:, 0, 0] = -0.5
        x[:, 0, T - 1] = 0.25
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(x.astype(np.float32))


def perturb_params(default: Dict[str, np.ndarray], kind: str, K: int, seed: int) -> Dict[str, np.ndarray]:
    """Parameter sets that exercise every clamp / min / max branch."""
    if kind == "default":
        return {k: v.copy() for k, v in default.items()}
    rng = np.random.Generator(np.random.PCG64(2000 + seed))
    F = default["kernel"].shape[0]
    p = {k: v.copy() for k, v in default.items()}
    kern = p["kernel"].astype(np.float64)
    kern[:, 0] *= rng.uniform(0.9, 1.1, F)
    kern[:, 1] *= rng.uniform(0.7, 1.4, F)
    # push a few filters past each clamp bound
    kern[0, 0] = -0.05
    kern[F - 1, 0] = math.pi + 0.2
    kern[1, 1] = 0.5                                   # below 4*sqrt(2 ln2)/pi
    kern[F // 2, 1] = K * 0.6                          # above K*sqrt(2 ln2)/pi
    p["kernel"] = kern.astype(np.float32)
    p["pool_w"] = rng.uniform(0.0, 0.7, F).astype(np.float32)     # crosses 2/K and 0.5
    p["pool_w"][2] = 0.0
    if kind == "perturbed":
        p["pool_b"] = rng.uniform(0.0, 1.5, F).astype(np.float32)
    else:                                              # perturbed2: tiny / negative bias -> floor branch
        p["pool_b"] = rng.uniform(-0.02, 0.05, F).astype(np.float32)
    p["alpha"] = rng.uniform(0.8, 1.1, F).astype(np.float32)
    p["delta"] = rng.uniform(0.5, 3.0, F).astype(np.float32)
    p["root"] = rng.uniform(0.9, 3.0, F).astype(np.float32)
    p["ema_w"] = rng.uniform(-0.05, 0.5, F).astype(np.float32)
    p["ema_w"][3 % F] = 1.2                            # above the clamp
    return p


def make_grad_out(shape, seed: int) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(3000 + seed))
    return rng.standard_normal(shape).astype(np.float32)


def sha256(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
