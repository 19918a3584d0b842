"""GPU parity of the parameter gradients (leafk_backward through autograd) against the reference's own
autograd (golden vectors) -- SURVEY 8c-v.  Tolerance: max|d| <= 1e-3 * max|g| per parameter tensor
(the reference's f32 gradients are themselves ~1e-5 from an f64 evaluation of the same graph)."""
import numpy as np
import pytest
import torch

from tests.cases import CASES, make_grad_out
from tests.test_forward_gpu import build
from tests.util import load_golden, scaled_err

pytestmark = pytest.mark.gpu

GRAD_CASES = [c.name for c in CASES if c.grads]
SD = {"kernel": "_complex_conv._kernel", "pool_w": "_pooling.weights", "pool_b": "_pooling._bias",
      "alpha": "_compression.alpha", "delta": "_compression.delta", "root": "_compression.root",
      "ema_w": "_compression.ema._weights"}


@pytest.mark.parametrize("name", GRAD_CASES)
def test_param_grads_match_reference_autograd(name):
    case, x, prm, z = load_golden(name)
    fe = build(case, prm, "auto")
    out = fe(x.cuda())
    G = torch.from_numpy(make_grad_out(tuple(out.shape), case.seed)).cuda()
    (out * G).sum().backward()
    torch.cuda.synchronize()
    named = dict(fe.named_parameters())
    worst = {}
    for k, sk in SD.items():
        got = named[sk].grad.detach().cpu().numpy().reshape(-1)
        want = z["grad_" + k].reshape(-1)
        assert np.all(np.isfinite(got)), k
        worst[k] = scaled_err(got, want)
    assert all(v < 1e-3 for v in worst.values()), worst
    # the typical agreement is two orders better than the bound; keep an eye on it
    assert np.median(list(worst.values())) < 1e-4, worst


def test_backward_is_deterministic_and_clamped_params_get_zero_grad():
    case, x, prm, z = load_golden("grad_perturbed")
    fe = build(case, prm, "auto")
    G = torch.from_numpy(make_grad_out(z["out"].shape, case.seed)).cuda()
    grads = []
    for _ in range(2):
        fe.zero_grad(set_to_none=True)
        (fe(x.cuda()) * G).sum().backward()
        grads.append([p.grad.clone() for p in fe.parameters()])
    for a, b in zip(*grads):
        assert torch.equal(a, b)
    gk = fe._complex_conv._kernel.grad.cpu().numpy()
    want = z["grad_kernel"]
    assert np.array_equal(gk == 0, want == 0)            # same entries gated by the clamps


def test_no_pcen_backward():
    """Leaf(pcen_compression=False): gradients of sum(p*G) for the three conv/pool parameters against the
    oracle's autograd (the golden 'nopcen' case stores no gradients)."""
    from oracle import leaf_oracle as O
    case, x, prm, z = load_golden("nopcen")
    fe = build(case, prm, "auto")
    out = fe(x.cuda())
    G = torch.from_numpy(make_grad_out(tuple(out.shape), 99))
    (out * G.cuda()).sum().backward()
    leaves = {k: prm[k].clone().requires_grad_(True) for k in ("kernel", "pool_w", "pool_b")}
    full = dict(prm); full.update(leaves)
    ref = O._forward(x, full, case.K, case.H, False, False)
    (ref * G).sum().backward()
    named = dict(fe.named_parameters())
    for k in leaves:
        got = named[SD[k]].grad.detach().cpu().numpy().reshape(-1)
        assert scaled_err(got, leaves[k].grad.numpy().reshape(-1)) < 1e-3, k


def test_waveform_gradient_matches_oracle_autograd():
    """dL/dx (SURVEY A.2 last bullet; leafk.h grad_x) against autograd over the oracle, together with the parameter
    gradients of the same backward."""
    from oracle import leaf_oracle as O
    case, x, prm, z = load_golden("grad_default")
    for algo in ("auto", "fp32"):
        fe = build(case, prm, algo)
        xg = x.cuda().requires_grad_(True)
        out = fe(xg)
        G = torch.from_numpy(make_grad_out(tuple(out.shape), case.seed))
        (out * G.cuda()).sum().backward()
        want = O.grads_f32(x, prm, case.K, case.H, G, with_input=True)
        assert scaled_err(xg.grad.cpu().numpy().reshape(-1), want["x"].numpy().reshape(-1)) < 1e-3, algo
        named = dict(fe.named_parameters())
        for k, sk in SD.items():
            assert scaled_err(named[sk].grad.cpu().numpy().reshape(-1), want[k].numpy().reshape(-1)) < 1e-3, (algo, k)


def test_generic_backward_equals_tensor_core_backward():
    """algo='fp32' trains through the generic FP32 backward (the path of geometries the tensor-core training kernel
    does not cover); on a covered geometry both must give the same gradients."""
    case, x, prm, z = load_golden("grad_perturbed")
    G = torch.from_numpy(make_grad_out(z["out"].shape, case.seed)).cuda()
    res = {}
    for algo in ("auto", "fp32"):
        fe = build(case, prm, algo)
        (fe(x.cuda()) * G).sum().backward()
        res[algo] = {k: p.grad.detach().cpu().numpy().reshape(-1) for k, p in fe.named_parameters()}
    for k in res["auto"]:
        assert scaled_err(res["fp32"][k], res["auto"][k]) < 2e-4, k
        assert scaled_err(res["fp32"][k], z["grad_" + [a for a, b in SD.items() if b == k][0]].reshape(-1)) < 1e-3, k


def test_training_forward_output_equals_inference_forward():
    """The features of the training forward (unpruned y bank next to the derivative banks) agree with the inference
    forward to fp32 class, and nothing is saved under no_grad."""
    case, x, prm, z = load_golden("grad_default")
    fe = build(case, prm, "auto")
    out_t = fe(x.cuda())
    assert out_t.grad_fn is not None
    with torch.no_grad():
        out_i = fe(x.cuda())
    assert out_i.grad_fn is None
    d = (out_t.detach() - out_i).abs().max().item()
    assert d < 5e-6, d
    fe.requires_grad_(False)
    assert fe(x.cuda()).grad_fn is None


def test_c_abi_leafk_backward_from_waveform_and_saved_energies():
    """include/leafk.h leafk_backward: gradients from x + the pooled energies saved by leafk_forward only (the entry
    point a non-PyTorch host would call).  On a tensor-core geometry it re-runs the training forward inside its
    workspace; the result must equal what the autograd path (leafk_forward_train + leafk_backward_saved) gives."""
    import ctypes as C
    import leaf_pytorch_b200.functional as LF
    from leaf_pytorch_b200 import _native as N
    case, x, prm, z = load_golden("grad_default")
    fe = build(case, prm, "auto")
    xd = x.cuda()
    G = torch.from_numpy(make_grad_out(z["out"].shape, case.seed)).cuda()
    (fe(xd) * G).sum().backward()
    want = [p.grad.detach().reshape(-1).clone() for p in fe._param_tuple()]
    # the same through the raw C ABI
    L = N.lib()
    spec = fe.spec
    prm_t = [p.detach() for p in fe._param_tuple()]
    out, saved_p = LF.forward_raw(spec, xd, *prm_t, save_p=True)
    cfg = spec.config(xd.dtype)
    ps, keep = LF._params_struct(spec, *prm_t, xd.device)
    grads_t = [torch.empty(p.numel(), device="cuda") for p in prm_t]
    gs = N.Grads(*[g.data_ptr() for g in grads_t])
    B, _, T = xd.shape
    gx = torch.empty_like(xd)
    nbytes = L.leafk_backward_workspace_bytes(C.byref(cfg), B, T)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    rc = L.leafk_backward(C.byref(cfg), C.byref(ps), C.c_void_p(xd.data_ptr()), B, T, C.c_void_p(G.data_ptr()),
                          C.c_void_p(saved_p.data_ptr()), C.byref(gs), C.c_void_p(gx.data_ptr()), C.c_void_p(ws.data_ptr()),
                          nbytes, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    N.check(rc, "leafk_backward")
    torch.cuda.synchronize()
    for got, w in zip(grads_t, want):
        # same kernels; the saved energies here come from the (pruned) inference forward, those of the autograd path
        # from the training forward: fp32-class differences
        assert scaled_err(got.cpu().numpy(), w.cpu().numpy()) < 1e-5
    from oracle import leaf_oracle as O
    ref = O.grads_f32(x, prm, case.K, case.H, G.cpu(), with_input=True)
    assert scaled_err(gx.cpu().numpy().reshape(-1), ref["x"].numpy().reshape(-1)) < 1e-3
