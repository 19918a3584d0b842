"""ctypes loader of oracle/_build/libleaforacle.so (the plain-C restatement; test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libleaforacle.so")


def load():
    if not os.path.isfile(LIB):
        subprocess.run(["make", "-s", "-C", HERE], check=True)
    lib = C.CDLL(LIB)
    fp = C.POINTER(C.c_float)
    lib.leaf_oracle_forward.restype = C.c_int
    lib.leaf_oracle_forward.argtypes = [fp, C.c_int, C.c_int, fp, fp, fp, fp, fp, fp, fp, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, fp, fp]
    lib.leaf_oracle_num_frames.restype = C.c_int
    lib.leaf_oracle_num_frames.argtypes = [C.c_int] * 3
    return lib


def forward(x, prm, K, H, compression=True, use_double=False):
    """x (B,1,T) float32 numpy; prm dict of numpy arrays -> (out, p) float32 (B,F,N)."""
    lib = load()
    x = np.ascontiguousarray(x, np.float32)
    B, _, T = x.shape
    F = prm["kernel"].shape[0]
    N = lib.leaf_oracle_num_frames(T, K, H)
    out = np.empty((B, F, N), np.float32)
    p = np.empty((B, F, N), np.float32)
    arrs = {k: (None if prm.get(k) is None else np.ascontiguousarray(prm[k], np.float32)) for k in
            ("kernel", "pool_w", "pool_b", "alpha", "delta", "root", "ema_w")}

    def ptr(a):
        return None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))
    rc = lib.leaf_oracle_forward(ptr(x), B, T, ptr(arrs["kernel"]), ptr(arrs["pool_w"]), ptr(arrs["pool_b"]),
                                 ptr(arrs["alpha"]), ptr(arrs["delta"]), ptr(arrs["root"]), ptr(arrs["ema_w"]),
                                 F, K, H, int(compression), int(use_double), ptr(out), ptr(p))
    if rc != 0:
        raise MemoryError("leaf_oracle_forward failed")
    return out, p
