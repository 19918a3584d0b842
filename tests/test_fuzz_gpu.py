"""Randomised geometries: the CUDA path (both conv kernels) against the oracle on the same seeded inputs,
at sizes the oracle finishes in a fraction of a second.  Covers odd/even K, hop > / < K, F not a multiple of
the channel-group sizes, clips shorter than one tile, batch sizes that leave CTA pairs half empty."""
import numpy as np
import pytest
import torch

from oracle import leaf_oracle as O
from tests.cases import perturb_params
from tests.test_forward_gpu import assert_close

pytestmark = pytest.mark.gpu


def random_case(seed):
    rng = np.random.Generator(np.random.PCG64(5000 + seed))
    F = int(rng.choice([1, 3, 8, 12, 20, 33, 40, 47, 64]))
    K = int(rng.integers(9, 700))
    H = int(rng.integers(max(1, K // 5), 2 * K))
    B = int(rng.integers(1, 6))
    T = int(rng.integers(1, 6000))
    return F, K, H, B, T, rng


@pytest.mark.parametrize("seed", range(24))
def test_random_geometry_matches_oracle(seed):
    import leaf_pytorch_b200.functional as LF
    F, K, H, B, T, rng = random_case(seed)
    base = {
        "kernel": np.stack([np.sort(rng.uniform(0.02, 3.0, F)), rng.uniform(2.0, min(K, 300) * 0.3, F)], 1).astype(np.float32),
        "pool_w": np.full(F, 0.4, np.float32), "pool_b": np.ones(F, np.float32),
        "alpha": np.full(F, 0.96, np.float32), "delta": np.full(F, 2.0, np.float32),
        "root": np.full(F, 2.0, np.float32), "ema_w": np.full(F, 0.04, np.float32)}
    prm_np = perturb_params(base, "perturbed" if seed % 2 else "default", K, seed) if F >= 4 else base
    prm = {k: torch.from_numpy(v) for k, v in prm_np.items()}
    x = torch.from_numpy((np.clip(rng.standard_normal((B, 1, T)), -4, 4) / 4).astype(np.float32))
    ref = O.forward_f32(x, prm, K, H).numpy()
    for algo in ("fp32", "tc"):
        if algo == "tc" and not LF.tc_supported(F, K, H):
            continue
        spec = LF.LeafSpec(F=F, K=K, H=H, compression=True, algo=algo)
        p = [prm[k].cuda() for k in ("kernel", "pool_w", "pool_b", "alpha", "delta", "root", "ema_w")]
        out, _ = LF.forward_raw(spec, x.cuda(), *p)
        torch.cuda.synchronize()
        assert_close(out.cpu().numpy(), ref, f"seed {seed} F={F} K={K} H={H} B={B} T={T} {algo}")


@pytest.mark.parametrize("seed", range(10))
def test_random_geometry_gradients_match_oracle_autograd(seed):
    """Parameter AND waveform gradients on random geometries.  Seeds >= 6 use hops the tensor-core kernels do not
    cover (more than 5 frames per 8 samples): forward on the FP32 kernel, backward on the generic FP32 kernel -- the
    reference trains any geometry (train.py:258)."""
    import leaf_pytorch_b200.functional as LF
    from tests.util import scaled_err
    rng = np.random.Generator(np.random.PCG64(7000 + seed))
    F = int(rng.choice([4, 8, 20, 40]))
    K = int(rng.choice([101, 201, 256, 401]))
    H = int(rng.choice([K // 3 + 1, K // 2, 160]))
    B, T = int(rng.integers(1, 4)), int(rng.integers(500, 3000))
    if seed >= 6:
        K = int(rng.choice([64, 101, 401]))
        H = int(rng.choice([5, K // 10, K // 7]))
        T = int(rng.integers(300, 1500))
        assert not LF.train_supported(LF.LeafSpec(F=F, K=K, H=H))
    base = {
        "kernel": np.stack([np.sort(rng.uniform(0.05, 2.9, F)), rng.uniform(3.0, K * 0.2, F)], 1).astype(np.float32),
        "pool_w": rng.uniform(0.1, 0.45, F).astype(np.float32), "pool_b": rng.uniform(0.2, 1.2, F).astype(np.float32),
        "alpha": rng.uniform(0.8, 0.99, F).astype(np.float32), "delta": rng.uniform(1.0, 3.0, F).astype(np.float32),
        "root": rng.uniform(1.2, 3.0, F).astype(np.float32), "ema_w": rng.uniform(0.02, 0.4, F).astype(np.float32)}
    prm = {k: torch.from_numpy(v) for k, v in base.items()}
    x = torch.from_numpy((np.clip(rng.standard_normal((B, 1, T)), -4, 4) / 4).astype(np.float32))
    N = O.num_frames(T, K, H)
    G = torch.from_numpy(rng.standard_normal((B, F, N)).astype(np.float32))
    want = O.grads_f32(x, prm, K, H, G, with_input=True)
    leaves = {k: v.clone().cuda().requires_grad_(True) for k, v in prm.items()}
    xg = x.cuda().requires_grad_(seed % 2 == 0)
    spec = LF.LeafSpec(F=F, K=K, H=H, compression=True, algo="auto")
    out = LF.leaf_forward(spec, xg, leaves["kernel"], leaves["pool_w"], leaves["pool_b"], leaves["alpha"],
                          leaves["delta"], leaves["root"], leaves["ema_w"])
    (out * G.cuda()).sum().backward()
    for k in O.PARAM_KEYS:
        err = scaled_err(leaves[k].grad.cpu().numpy().reshape(-1), want[k].numpy().reshape(-1))
        assert err < 1e-3, (k, err, F, K, H, B, T)
    if xg.requires_grad:
        err = scaled_err(xg.grad.cpu().numpy().reshape(-1), want["x"].numpy().reshape(-1))
        assert err < 1e-3, ("x", err, F, K, H, B, T)


@pytest.mark.parametrize("F,K,H", [(8, 1601, 480), (12, 1345, 400), (20, 2001, 700)])
def test_long_windows_match_oracle(F, K, H):
    """Windows of 84-125 ms: 3-4 copy chunks per producer thread, many k-steps per zone, small channel groups."""
    import leaf_pytorch_b200.functional as LF
    assert LF.tc_supported(F, K, H)
    rng = np.random.Generator(np.random.PCG64(K))
    B, T = 2, 9000
    prm_np = {
        "kernel": np.stack([np.sort(rng.uniform(0.02, 3.0, F)), rng.uniform(3.0, 250.0, F)], 1).astype(np.float32),
        "pool_w": rng.uniform(0.2, 0.5, F).astype(np.float32), "pool_b": rng.uniform(0.0, 1.0, F).astype(np.float32),
        "alpha": np.full(F, 0.96, np.float32), "delta": np.full(F, 2.0, np.float32),
        "root": np.full(F, 2.0, np.float32), "ema_w": np.full(F, 0.04, np.float32)}
    prm = {k: torch.from_numpy(v) for k, v in prm_np.items()}
    x = torch.from_numpy((np.clip(rng.standard_normal((B, 1, T)), -4, 4) / 4).astype(np.float32))
    ref = O.forward_f32(x, prm, K, H).numpy()
    p = [prm[k].cuda() for k in ("kernel", "pool_w", "pool_b", "alpha", "delta", "root", "ema_w")]
    for algo in ("tc", "tc_full", "fp32"):
        spec = LF.LeafSpec(F=F, K=K, H=H, compression=True, algo=algo)
        out, _ = LF.forward_raw(spec, x.cuda(), *p)
        torch.cuda.synchronize()
        assert_close(out.cpu().numpy(), ref, f"long window F={F} K={K} H={H} {algo}")
