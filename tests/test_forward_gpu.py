"""GPU parity: the CUDA path (through the C ABI) against the golden vectors of the real reference
and against the oracle on the same seeded inputs.

Tolerance (stated by BASELINE.json north_star: 1e-4 relative, fp32):  |got-ref| <= 1e-4*|ref| + 1e-5.
The absolute term is needed because the reference itself is ~1e-5 (absolute) away from exact
arithmetic where u^q - delta^q cancels (see tests/test_oracle_golden.py).
"""
import numpy as np
import pytest
import torch

from tests.cases import CASES
from tests.util import load_golden, scaled_err

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-4, 1e-5
ALGOS = ["fp32", "tc", "tc_full"]     # tc: support-pruned k-steps (default), tc_full: every tap
SMALL = [c.name for c in CASES if c.T <= 20000 and not c.grads]


def build(case, prm, algo, device="cuda"):
    import leaf_pytorch_b200 as L
    fe = L.Leaf(n_filters=case.F, sample_rate=case.sr, window_len=case.wlen, window_stride=case.wstride,
                init_min_freq=case.min_freq, init_max_freq=case.max_freq, pcen_compression=case.compression,
                use_legacy_complex=case.legacy, algo=algo)
    sd = {"_complex_conv._kernel": prm["kernel"], "_pooling.weights": prm["pool_w"].reshape(1, 1, -1, 1),
          "_pooling._bias": prm["pool_b"]}
    if case.compression:
        sd.update({"_compression.alpha": prm["alpha"], "_compression.delta": prm["delta"],
                   "_compression.root": prm["root"], "_compression.ema._weights": prm["ema_w"]})
    fe.load_state_dict(sd)
    return fe.to(device)


def algo_available(case, algo):
    if not algo.startswith("tc"):
        return True
    import leaf_pytorch_b200.functional as LF
    return LF.tc_supported(case.F, case.K, case.H)


def assert_close(got, ref, what):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, what
    assert np.all(np.isfinite(got)) or not np.all(np.isfinite(ref)), f"{what}: non-finite output"
    d = np.abs(got - ref)
    lim = RTOL * np.abs(ref) + ATOL
    worst = np.unravel_index(np.argmax(d - lim), d.shape)
    assert np.all(d <= lim), (f"{what}: |d|={d[worst]:.3e} at {worst} (ref {ref[worst]:.6e}, got {got[worst]:.6e}); "
                              f"scaled err {scaled_err(got, ref):.3e}")
    # north_star's own form -- pure relative 1e-4 -- wherever it applies: every element whose reference value is not in
    # the cancellation regime (|ref| >= 1e-2 * max|ref|; below that u^q - delta^q has lost digits in the reference itself)
    big = np.abs(ref) >= 1e-2 * max(np.abs(ref).max(), 1e-30)
    if np.any(big) and np.all(np.isfinite(ref)):
        rel = float(np.max(d[big] / np.abs(ref[big])))
        assert rel <= RTOL, f"{what}: pure relative error {rel:.3e} on the {int(big.sum())} elements with |ref| >= 1% of the peak"


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("name", SMALL)
def test_forward_matches_reference_golden(name, algo):
    case, x, prm, z = load_golden(name)
    if not algo_available(case, algo):
        pytest.skip("tensor-core kernel does not cover this geometry (falls back to fp32 under algo=auto)")
    fe = build(case, prm, algo)
    with torch.no_grad():
        out = fe(x.cuda())
    torch.cuda.synchronize()
    assert_close(out.cpu().numpy(), z["out"], f"{name}/{algo}")


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("name", ["cfg1_default", "perturbed2_F40", "T161", "sr22050_evenK", "F80"])
def test_pooled_energies_match(name, algo):
    """the saved, floored pooled energies (input of PCEN) against the reference's own sub-modules"""
    import leaf_pytorch_b200.functional as LF
    case, x, prm, z = load_golden(name)
    if not algo_available(case, algo):
        pytest.skip("geometry not covered by the tensor-core kernel")
    fe = build(case, prm, algo)
    out, p = LF.forward_raw(fe.spec, x.cuda(), *fe._param_tuple(), save_p=True)
    torch.cuda.synchronize()
    got, ref = p.cpu().numpy().astype(np.float64), z["p"].astype(np.float64)
    assert np.all(np.abs(got - ref) <= 1e-4 * np.abs(ref) + 2e-6)


@pytest.mark.parametrize("algo", ALGOS)
def test_long_clip_10s(algo):
    case, x, prm, z = load_golden("long10s")
    fe = build(case, prm, algo)
    with torch.no_grad():
        out = fe(x.cuda())
    assert_close(out.cpu().numpy(), z["out"], f"long10s/{algo}")


@pytest.mark.parametrize("algo", ALGOS)
def test_deterministic_and_batch_invariant(algo):
    case, x, prm, z = load_golden("cfg1_default")
    fe = build(case, prm, algo)
    xg = x.cuda()
    with torch.no_grad():
        a = fe(xg)
        b = fe(xg)
        c = fe(xg[1:3].contiguous())
    assert torch.equal(a, b)
    assert torch.equal(a[1:3], c)


def test_auto_uses_available_kernel_and_counts_launches():
    import leaf_pytorch_b200 as L
    case, x, prm, z = load_golden("cfg1_default")
    fe = build(case, prm, "auto")
    L.launch_count(reset=True)
    with torch.no_grad():
        out = fe(x.cuda())
    torch.cuda.synchronize()
    assert L.launch_count() >= 3
    assert_close(out.cpu().numpy(), z["out"], "auto")


def test_cpu_tensor_fails_loudly():
    import leaf_pytorch_b200 as L
    fe = L.Leaf()
    with pytest.raises(L.LeafNativeError):
        fe(torch.zeros(1, 1, 1600))


@pytest.mark.parametrize("n_slices", [1, 3, 8])
def test_forward_host_pipelined_equals_device_forward(n_slices):
    """leafk_forward_host (pinned host in/out; one persistent launch gated by per-slice ready flags that follow
    the H2D copies) must give exactly what the device-resident forward gives."""
    case, x, prm, z = load_golden("cfg1_default")
    fe = build(case, prm, "auto")
    xb = torch.cat([x, x.flip(0), 0.5 * x], dim=0)           # 12 clips
    with torch.no_grad():
        want = fe(xb.cuda()).cpu()
    xh = xb.pin_memory()
    for _ in range(3):                                       # buffers are reused across calls
        got = fe.forward_host(xh, n_slices=n_slices)
        assert torch.equal(got, want)
    assert_close(got[:4].numpy(), z["out"], "forward_host")


def test_forward_host_fp32_algo_uses_sliced_path():
    case, x, prm, z = load_golden("cfg1_default")
    fe = build(case, prm, "fp32")
    got = fe.forward_host(x.pin_memory(), n_slices=2)
    assert_close(got.numpy(), z["out"], "forward_host/fp32")


def test_exact_power_of_two_scaling_of_energies():
    """Size-independent property at BASELINE configs[1] size (256 x 1 s): with PCEN off and zero bias the pooled
    energies are homogeneous of degree 2, and because the kernel's per-tile scaling is an exact power of two,
    out(4x) == 16*out(x) bit for bit."""
    import leaf_pytorch_b200 as L
    g = torch.Generator().manual_seed(7)
    x = (torch.randn(256, 1, 16000, generator=g).clamp_(-4, 4) / 4).cuda()
    fe = L.Leaf(pcen_compression=False).cuda()
    with torch.no_grad():
        fe._pooling._bias.zero_()
        a = fe(x)
        b = fe(4.0 * x)
    assert torch.equal(b, 16.0 * a)
    assert float(a.min()) >= 1e-5


def test_batch_permutation_equivariance_full_size():
    import leaf_pytorch_b200 as L
    g = torch.Generator().manual_seed(8)
    x = (torch.randn(256, 1, 16000, generator=g).clamp_(-4, 4) / 4).cuda()
    perm = torch.randperm(256, generator=g).cuda()
    fe = L.Leaf().cuda()
    with torch.no_grad():
        assert torch.equal(fe(x)[perm], fe(x[perm].contiguous()))


def test_forward_is_cuda_graph_capturable():
    """The 3-launch forward can be captured into a CUDA graph and replayed (no host-side state, no sync,
    all scratch from the caching allocator)."""
    case, x, prm, z = load_golden("cfg1_default")
    fe = build(case, prm, "auto")
    xg = x.cuda()
    static_x = xg.clone()
    with torch.no_grad():
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                fe(static_x)
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static_out = fe(static_x)
        static_x.copy_(0.5 * xg)
        graph.replay()
        torch.cuda.synchronize()
        want = fe(0.5 * xg)
    assert torch.equal(static_out, want)


@pytest.mark.parametrize("algo", ALGOS)
def test_int16_pcm_input_equals_float_path(algo):
    """LEAFK_INPUT_S16: 16-bit PCM converted in the kernel as s/32768 (exact in fp32), so the result must be
    bit-identical to feeding the converted float waveform -- forward, host-pipelined forward and gradients."""
    case, x, prm, z = load_golden("pcm16_default")
    fe = build(case, prm, algo)
    pcm = torch.round(x * 32768.0).clamp_(-32768, 32767).to(torch.int16)
    xf = pcm.to(torch.float32) / 32768.0
    assert torch.equal(xf, x)                                  # the golden input is already PCM-quantised
    with torch.no_grad():
        a = fe(pcm.cuda())
        b = fe(xf.cuda())
    assert torch.equal(a, b)
    assert_close(a.cpu().numpy(), z["out"], f"pcm16/{algo}")
    got = fe.forward_host(pcm.pin_memory(), n_slices=2)
    assert torch.equal(got, a.cpu())
    if algo == "tc":
        G = torch.randn(a.shape, generator=torch.Generator().manual_seed(3)).cuda()
        grads = []
        for inp in (pcm.cuda(), xf.cuda()):
            fe.zero_grad(set_to_none=True)
            (fe(inp) * G).sum().backward()
            grads.append([p.grad.clone() for p in fe.parameters()])
        for ga, gb in zip(*grads):
            assert torch.equal(ga, gb)


def test_empty_batch_and_bad_shapes():
    import leaf_pytorch_b200 as L
    fe = L.Leaf().cuda()
    out = fe(torch.zeros(0, 1, 16000, device="cuda"))
    assert tuple(out.shape) == (0, 40, 100)
    for bad in (torch.zeros(2, 2, 800, device="cuda"), torch.zeros(2, 800, device="cuda"),
                torch.zeros(2, 1, 0, device="cuda")):
        with pytest.raises(ValueError):
            fe(bad)
    with pytest.raises(TypeError):
        fe(torch.zeros(2, 1, 800, device="cuda", dtype=torch.float64))
    # non-contiguous input is accepted (made contiguous), like any torch op would
    x = torch.randn(3, 1, 4000, device="cuda")
    xt = x.expand(3, 1, 4000).transpose(0, 2).contiguous().transpose(0, 2)
    assert not xt.is_contiguous()
    with torch.no_grad():
        assert torch.equal(fe(xt), fe(x))


def test_maximum_batch_config3_size_properties():
    """BASELINE configs[2] size (80 filters, 1024 x 1 s): no oracle at this size; check the size-independent
    properties: every clip equals the same clip processed in a small batch, output finite and above the PCEN floor."""
    import leaf_pytorch_b200 as L
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(1024, 1, 16000, generator=g).clamp_(-4, 4) / 4).cuda()
    fe = L.Leaf(n_filters=80).cuda()
    with torch.no_grad():
        big = fe(x)
        idx = [0, 1, 511, 512, 1022, 1023]
        small = fe(x[idx].contiguous())
    assert tuple(big.shape) == (1024, 80, 100)
    assert torch.equal(big[idx], small)
    assert torch.isfinite(big).all()


@pytest.mark.parametrize("depth,dtype", [(1, torch.float32), (2, torch.float32), (3, torch.float32), (2, torch.int16)])
def test_host_pipeline_batches_in_flight(depth, dtype):
    """HostPipeline (leafk_forward_host_async): several batches in flight, every result identical to the
    device-resident forward of the same input, whatever the submit/collect interleaving."""
    import leaf_pytorch_b200 as L
    case, x, prm, z = load_golden("cfg1_default")
    fe = build(case, prm, "auto")
    B, T = 6, 9000
    g = torch.Generator().manual_seed(21)
    xs = [(torch.randn(B, 1, T, generator=g).clamp_(-4, 4) / 4) for _ in range(7)]
    if dtype == torch.int16:
        xs = [(v * 32767).round().to(torch.int16) for v in xs]
    with torch.no_grad():
        want = [fe(v.cuda()).cpu() for v in xs]
    hosts = [v.pin_memory() for v in xs]
    pipe = L.HostPipeline(fe, B, T, depth=depth, n_slices=3, input_dtype=dtype)
    tickets, got = [], []
    for i, h in enumerate(hosts):
        tickets.append(pipe.submit(h))
        if len(tickets) == depth:                         # keep `depth` batches in flight
            got.append(pipe.result(tickets.pop(0)).clone())
    while tickets:
        got.append(pipe.result(tickets.pop(0)).clone())
    pipe.close()
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert torch.equal(a, b)


def test_benchmark_config_subsample_against_oracle():
    """BASELINE configs[1] exactly as bench.py runs it (256 x 1 s, default Leaf, bench's synthetic generator): a
    subsample of the clips is recomputed by the CPU oracle and must agree within the parity tolerance."""
    import leaf_pytorch_b200 as L
    from oracle import leaf_oracle as O
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(256, 1, 16000, generator=g).clamp_(-4, 4) / 4
    fe = L.Leaf().cuda()
    with torch.no_grad():
        out = fe(x.cuda()).cpu()
    idx = [0, 37, 101, 128, 200, 255]
    prm = O.params_from_state_dict({k: v.detach().cpu() for k, v in fe.state_dict().items()})
    ref = O.forward_f32(x[idx], prm, 401, 160).numpy()
    assert_close(out[idx].numpy(), ref, "cfg2 subsample vs oracle")


def _schedule(fe, x):
    import leaf_pytorch_b200.functional as LF
    return LF.tc_schedule(fe.spec, x, *[None if q is None else q.detach() for q in fe._param_tuple()])


def test_support_pruning_schedule_default_init():
    """Default mel initialisation: the narrow (high-frequency) filters are skipped on the outer k-steps, every
    channel runs on the middle k-steps, the schedule is nested/unimodal, and tc_full runs everything."""
    import leaf_pytorch_b200 as L
    x = (torch.randn(2, 1, 4000, generator=torch.Generator().manual_seed(0)).clamp_(-4, 4) / 4).cuda()
    sch = _schedule(L.Leaf(algo="tc").cuda(), x)
    assert sch["n_groups"] == 1 and sch["channels_per_group"] == 80 and sch["n_ksteps"] == 26
    na = sch["active"][0]
    assert max(na) == 80 and na[12] == 80 and min(na) >= 16 and all(v % 16 == 0 for v in na)
    peak = na.index(80)
    assert all(na[i] <= na[i + 1] for i in range(peak)) and all(na[i] >= na[i + 1] for i in range(peak, 25))
    na3 = sch["active_all_products"][0]
    assert all(a3 <= a for a3, a in zip(na3, na)) and na3[12] == 80 and min(na3) >= 0
    assert 0.4 < sch["executed_fraction"] < 0.85
    full = _schedule(L.Leaf(algo="tc_full").cuda(), x)
    assert full["executed_fraction"] == 1.0
    # two channel groups (F=80): the width-sorted filters are dealt round-robin, so both groups prune alike
    s80 = _schedule(L.Leaf(n_filters=80, algo="tc").cuda(), x)
    assert s80["n_groups"] == 2
    a0, a1 = s80["active"]
    assert sum(abs(u - v) for u, v in zip(a0, a1)) <= 16 * 6


def test_support_pruning_worst_cases_against_oracle():
    """Inputs and parameters chosen against the pruning: widths at and beyond both clamps (1.5 .. 150 samples), no
    pooling bias (nothing hides a relative error of the energies), and (a) a full-scale tone far outside most pass
    bands, (b) a single impulse (the output frames are the squared filter taps themselves, tails included)."""
    import leaf_pytorch_b200.functional as LF
    from oracle import leaf_oracle as O
    F, K, H, T = 40, 401, 160, 4000
    rng = np.random.Generator(np.random.PCG64(77))
    sig = np.concatenate([[0.5, 1.499, 2.0, 3.0, 140.0, 150.3, 400.0], rng.uniform(2.0, 60.0, F - 7)]).astype(np.float32)
    prm = {"kernel": torch.from_numpy(np.stack([rng.uniform(0.0, 3.14, F).astype(np.float32), sig], 1)),
           "pool_w": torch.full((F,), 0.4), "pool_b": torch.zeros(F),
           "alpha": torch.full((F,), 0.96), "delta": torch.full((F,), 2.0), "root": torch.full((F,), 2.0),
           "ema_w": torch.full((F,), 0.04)}
    t = torch.arange(T, dtype=torch.float32)
    tone = torch.sin(2.9 * t).reshape(1, 1, T)
    imp = torch.zeros(1, 1, T)
    imp[0, 0, 1777] = 1.0
    noise = (torch.randn(1, 1, T, generator=torch.Generator().manual_seed(5)).clamp_(-4, 4) / 4) * 1e-3
    x = torch.cat([tone, imp, tone + noise], 0)
    ref = O.forward_f32(x, prm, K, H).numpy()
    p = [prm[k].cuda() for k in ("kernel", "pool_w", "pool_b", "alpha", "delta", "root", "ema_w")]
    outs = {}
    for algo in ("tc", "tc_full"):
        spec = LF.LeafSpec(F=F, K=K, H=H, compression=True, algo=algo)
        out, _ = LF.forward_raw(spec, x.cuda(), *p)
        torch.cuda.synchronize()
        outs[algo] = out.cpu().numpy()
        assert_close(outs[algo], ref, f"pruning worst case / {algo}")
    sch = LF.tc_schedule(LF.LeafSpec(F=F, K=K, H=H, compression=True, algo="tc"), x.cuda(), *p)
    assert sch["executed_fraction"] < 0.95


def _expected_schedule(sigma, F, K, cg, n_groups, c1=5.5, c3=3.7):
    """Host restatement of k0's pruning schedule (k1_tc_layout.cuh): per group and 16-tap k-step the channels that
    run (inside ceil(5.5 sigma), filters rounded up to 8) and those that run all three products (inside
    ceil(3.7 sigma), then made constant per zone of constant na1 and side of the middle k-step)."""
    import math
    Kp = (K + 15) // 16 * 16
    ks, FG, Fp, kc = Kp // 16, cg // 2, (cg // 2) * n_groups, K // 2
    key = [float(sigma[i]) if i < F else -1.0 for i in range(Fp)]
    rank = [sum(1 for j in range(Fp) if key[j] < key[i] or (key[j] == key[i] and j < i)) for i in range(Fp)]

    def rng(s, c):
        if s < 0:
            return 1, 0
        R = min(K, int(math.ceil(np.float32(c) * np.float32(s))))
        return max(0, kc - R) // 16, min(K - 1, kc + R) // 16
    sf = (ks - 1) // 2
    na1 = [[0] * ks for _ in range(n_groups)]
    na3 = [[0] * ks for _ in range(n_groups)]
    for g in range(n_groups):
        members = [i for i in range(Fp) if rank[i] % n_groups == g]
        for s in range(ks):
            c1n = sum(1 for i in members if rng(key[i], c1)[0] <= s <= rng(key[i], c1)[1])
            c3n = sum(1 for i in members if max(rng(key[i], c3)[0], rng(key[i], c1)[0]) <= s <= min(rng(key[i], c3)[1], rng(key[i], c1)[1]))
            nf = min(FG, (c1n + 7) // 8 * 8)
            nf3 = min(nf, (c3n + 7) // 8 * 8)
            if s == sf:
                nf = nf3 = FG
            na1[g][s], na3[g][s] = 2 * nf, 2 * nf3
        al = list(na3[g])
        for s in range(ks):
            if na1[g][s] == cg:
                al[s] = cg
            else:
                side = range(0, sf) if s < sf else range(sf + 1, ks)
                al[s] = max([na3[g][t] for t in side if na1[g][t] == na1[g][s]] + [0])
        na3[g] = al
    return na1, na3


@pytest.mark.parametrize("F,seed", [(40, 0), (40, 1), (80, 2), (64, 3), (20, 4), (47, 5)])
def test_pruning_schedule_matches_host_restatement(F, seed):
    import leaf_pytorch_b200 as L
    import leaf_pytorch_b200.functional as LF
    fe = L.Leaf(n_filters=F, algo="tc").cuda()
    rng = np.random.Generator(np.random.PCG64(900 + seed))
    if seed:                                              # seed 0: the default mel initialisation
        with torch.no_grad():
            k = fe._complex_conv._kernel
            k[:, 1] = torch.from_numpy(rng.uniform(1.0, 120.0, F).astype(np.float32)).to(k.device)
    x = (torch.randn(2, 1, 3000, generator=torch.Generator().manual_seed(seed)).clamp_(-4, 4) / 4).cuda()
    sch = LF.tc_schedule(fe.spec, x, *[None if q is None else q.detach() for q in fe._param_tuple()])
    sig = fe._complex_conv._kernel.detach()[:, 1].cpu().numpy().astype(np.float32)
    lo, hi = np.float32(4 * np.sqrt(2 * np.log(2.0)) / np.pi), np.float32(401 * np.sqrt(2 * np.log(2.0)) / np.pi)
    sig = np.clip(sig, lo, hi)
    na1, na3 = _expected_schedule(sig, F, 401, sch["channels_per_group"], sch["n_groups"])
    assert sch["active"] == na1
    assert sch["active_all_products"] == na3
