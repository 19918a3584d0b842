"""GPU tests of the pieces around the fused path: on-the-fly clip preparation (crop / pad / peak normalisation,
SURVEY 8f rank 3), the optional stages the reference only declares (pre-emphasis, mean/variance normalisation, filter
sorting) and the caller glue towards the backbone.  Oracles: the CPU restatements
in oracle/leaf_oracle.py and plain torch ops on the CPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as TF

from tests.test_forward_gpu import assert_close
from tests.util import scaled_err

pytestmark = pytest.mark.gpu


def oracle_prm(fe):
    from oracle import leaf_oracle as O
    return O.params_from_state_dict({k: v.detach().cpu() for k, v in fe.state_dict().items()})


@pytest.mark.parametrize("dtype", [torch.float32, torch.int16])
@pytest.mark.parametrize("pad_mode", ["edge", "min", "wrap", "zero"])
def test_prepared_forward_equals_cpu_transforms_then_forward(dtype, pad_mode):
    """Raw clips of unequal length, some too loud: forward_prepared (crop / pad / peak-normalise inside the kernels'
    staging) == the reference's CPU transforms followed by the plain forward."""
    import leaf_pytorch_b200 as L
    from oracle import leaf_oracle as O
    n = 8000
    lens = [8000, 12001, 5000, 8001, 16000, 1, 7999, 3]
    g = torch.Generator().manual_seed(17)
    raw = torch.zeros(len(lens), 1, max(lens))
    for b, ln in enumerate(lens):
        raw[b, 0, :ln] = torch.randn(ln, generator=g) * (0.2 if b % 2 else 0.6)          # every other clip exceeds 1
    if dtype == torch.int16:
        raw = (raw.clamp(-1, 1) * 32767).round().to(torch.int16)
    as_float = raw.to(torch.float32) / (32768.0 if dtype == torch.int16 else 1.0)
    prepared = torch.stack([O.prepare_clip(as_float[b, 0, :ln].numpy(), n, "center", pad_mode) for b, ln in enumerate(lens)]).unsqueeze(1)
    if dtype == torch.float32:
        assert float(prepared.abs().max()) <= 1.0 and any(float(as_float[b].abs().max()) > 1 for b in range(len(lens)))
    for algo in ("auto", "fp32"):
        fe = L.Leaf(algo=algo).cuda()
        with torch.no_grad():
            got = fe.forward_prepared(raw.cuda(), n, raw_lengths=torch.tensor(lens), starts="center", pad_mode=pad_mode)
            same = fe(prepared.cuda())
        assert torch.equal(got, same), (algo, (got - same).abs().max().item())           # same arithmetic, no copy of the batch
        ref = O.forward_f32(prepared, oracle_prm(fe), 401, 160).numpy()
        assert_close(got.cpu().numpy(), ref, f"prepared clips {dtype} {pad_mode} {algo}")


@pytest.mark.parametrize("which", ["prep_eval", "prep_train"])
def test_prepared_forward_matches_reference_transforms_golden(which):
    """Golden vectors made with the reference's own PadToSize / CenterCrop / RandomCrop classes and the reference Leaf
    (tests/golden/make_golden_prep.py): the kernels' on-the-fly preparation gives the same features."""
    import os
    import leaf_pytorch_b200 as L
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", which + ".npz"))
    n, lens = int(z["n_samples"]), torch.from_numpy(z["lengths"])
    raw = torch.from_numpy(z["raw"]).unsqueeze(1).cuda()
    for algo in ("auto", "fp32"):
        fe = L.Leaf(algo=algo).cuda()
        with torch.no_grad():
            if which == "prep_eval":      # PadToSize(size, 'wrap') [= replicate], CenterCrop, PeakNormalization
                got = fe.forward_prepared(raw, n, raw_lengths=lens, starts="center", pad_mode="edge")
            else:                         # PadToSize(size, 'constant') [= clip minimum], RandomCrop offsets, PeakNormalization
                pad_front = (n - lens).clamp(min=0) // 2
                starts = torch.where(lens < n, -pad_front, torch.from_numpy(z["starts"])).to(torch.int32)
                got = fe.forward_prepared(raw, n, raw_lengths=lens, starts=starts.cuda(), pad_mode="min")
            same = fe(torch.from_numpy(z["prepared"]).cuda())
        assert torch.equal(got, same), (which, algo)
        assert_close(got.cpu().numpy(), z["out"], f"{which} {algo}")


def test_random_crop_offsets_and_training_through_prepared_clips():
    import leaf_pytorch_b200 as L
    import leaf_pytorch_b200.functional as LF
    from oracle import leaf_oracle as O
    n, Traw = 6000, 9000
    g = torch.Generator().manual_seed(4)
    raw = torch.randn(5, 1, Traw, generator=g) * 0.3
    starts = torch.tensor([0, 3000, 1234, 2999, 17], dtype=torch.int32)
    prepared = torch.stack([O.prepare_clip(raw[b, 0].numpy(), n, int(starts[b]), "zero") for b in range(5)]).unsqueeze(1)
    fe = L.Leaf().cuda()
    prep = LF.prepare_clips(fe.spec, raw.cuda(), n, None, starts.cuda(), "zero", True)
    out = LF.leaf_forward(fe.spec, raw.cuda(), *fe._param_tuple(), prep=prep)
    G = torch.randn(out.shape, generator=g)
    (out * G.cuda()).sum().backward()
    want = O.grads_f32(prepared, oracle_prm(fe), 401, 160, G)
    names = {"kernel": "_complex_conv._kernel", "pool_w": "_pooling.weights", "pool_b": "_pooling._bias",
             "alpha": "_compression.alpha", "delta": "_compression.delta", "root": "_compression.root",
             "ema_w": "_compression.ema._weights"}
    named = dict(fe.named_parameters())
    for k, sk in names.items():
        assert scaled_err(named[sk].grad.cpu().numpy().reshape(-1), want[k].numpy().reshape(-1)) < 1e-3, k


def torch_restatement(x, prm, K, H, preemp_w=None, instnorm=False, sort=False):
    """CPU restatement with the optional stages (plain torch ops around the oracle's forward)."""
    from oracle import leaf_oracle as O
    if preemp_w is not None:
        x = TF.conv1d(TF.pad(x, (0, 1)), preemp_w.reshape(1, 1, 2))
    if sort:
        prm = dict(prm)
        order = torch.argsort(prm["kernel"].detach()[:, 0].clamp(0.0, float(np.pi)), stable=True)
        prm["kernel"] = prm["kernel"][order]
    out = O.forward_on_device(x, prm, K, H)
    if instnorm:
        out = TF.instance_norm(out, eps=1e-5)
    return out


@pytest.mark.parametrize("preemp,instnorm,sort", [(True, False, False), (False, True, False), (False, False, True), (True, True, True)])
def test_optional_stages_forward_and_gradients(preemp, instnorm, sort):
    import leaf_pytorch_b200 as L
    torch.manual_seed(3)
    fe = L.Leaf(n_filters=24, preemp=preemp, mean_var_norm=instnorm, sort_filters=sort)
    with torch.no_grad():
        k = fe._complex_conv._kernel
        k[:, 0] = k[torch.randperm(24), 0]                     # unsorted centre frequencies
        if preemp:
            fe._preemp.weight.copy_(torch.tensor([[[-0.9, 1.05]]]))
    fe = fe.cuda()
    g = torch.Generator().manual_seed(8)
    x = torch.randn(3, 1, 5000, generator=g).clamp_(-4, 4) / 4
    prm = {k: (None if v is None else v.clone().requires_grad_(True)) for k, v in oracle_prm(fe).items()}
    pw = fe._preemp.weight.detach().cpu().clone().requires_grad_(True) if preemp else None
    ref = torch_restatement(x, prm, 401, 160, pw, instnorm, sort)
    out = fe(x.cuda())
    if instnorm:
        got, want = out.detach().cpu().numpy(), ref.detach().numpy()
        assert np.abs(got - want).max() < 2e-3 * max(1.0, np.abs(want).max())       # normalised features, unit scale
    else:
        assert_close(out.detach().cpu().numpy(), ref.detach().numpy(), f"optional stages {preemp} {instnorm} {sort}")
    G = torch.randn(out.shape, generator=g)
    (out * G.cuda()).sum().backward()
    (ref * G).sum().backward()
    names = {"kernel": "_complex_conv._kernel", "pool_w": "_pooling.weights", "pool_b": "_pooling._bias",
             "alpha": "_compression.alpha", "delta": "_compression.delta", "root": "_compression.root",
             "ema_w": "_compression.ema._weights"}
    named = dict(fe.named_parameters())
    for k, sk in names.items():
        err = scaled_err(named[sk].grad.cpu().numpy().reshape(-1), prm[k].grad.numpy().reshape(-1))
        assert err < 2e-3, (k, err)
    if preemp:
        err = scaled_err(fe._preemp.weight.grad.cpu().numpy().reshape(-1), pw.grad.numpy().reshape(-1))
        assert err < 2e-3, ("preemp", err)


class TinyBackbone(torch.nn.Module):
    """Stand-in for the reference's 2-D CNN backbones (models/classifier.py:12 get_classifier): takes (B,1,F,N)."""

    def __init__(self, n_classes=7):
        super().__init__()
        self.c1 = torch.nn.Conv2d(1, 8, 3, padding=1)
        self.c2 = torch.nn.Conv2d(8, 16, 3, stride=2, padding=1)
        self.fc = torch.nn.Linear(16, n_classes)

    def forward(self, x):
        x = torch.relu(self.c1(x))
        x = torch.relu(self.c2(x))
        return self.fc(x.mean(dim=(2, 3)))


def test_classifier_logits_with_our_frontend_equal_logits_with_the_reference_frontend():
    """SURVEY 8c fixture (vii): Classifier.forward = frontend -> unsqueeze(1) -> backbone (reference
    models/classifier.py:14-18).  The same backbone on the oracle's features (the reference frontend, CPU) and on our
    features gives the same logits; the CUDA-graph replay of frontend + backbone gives them too."""
    import leaf_pytorch_b200 as L
    from leaf_pytorch_b200.classifier import Classifier
    from oracle import leaf_oracle as O
    torch.manual_seed(11)
    backbone = TinyBackbone()
    clf = Classifier(L.Leaf(), backbone).cuda().eval()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(6, 1, 16000, generator=g).clamp_(-4, 4) / 4
    with torch.no_grad():
        feats = O.forward_f32(x, oracle_prm(clf.features), 401, 160)
        want = TinyBackbone().eval()
        want.load_state_dict({k: v.cpu() for k, v in clf.model.state_dict().items()})
        ref_logits = want(feats.unsqueeze(1))
        old = torch.backends.cudnn.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        try:
            logits = clf(x.cuda())
            graphed = clf.capture(x.cuda())
            replay = graphed(x.cuda()).clone()
            x2 = torch.randn(6, 1, 16000, generator=g).clamp_(-4, 4) / 4
            replay2 = graphed(x2.cuda()).clone()
            direct2 = clf(x2.cuda())
        finally:
            torch.backends.cudnn.allow_tf32 = old
    assert tuple(logits.shape) == (6, 7)
    assert scaled_err(logits.cpu().numpy(), ref_logits.numpy()) < 1e-4
    assert torch.equal(replay, logits) and torch.equal(replay2, direct2)
    # training through the wrapper: gradients reach the frontend parameters
    clf.train()
    clf(x.cuda()).sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in clf.features.parameters())


def test_prepared_forward_applies_mean_var_norm():
    """forward_prepared runs the same stages as forward: with mean_var_norm=True the features are normalised."""
    import leaf_pytorch_b200 as L
    import leaf_pytorch_b200.functional as LF
    g = torch.Generator().manual_seed(5)
    raw = (torch.randn(4, 1, 9000, generator=g) * 0.7).cuda()
    lens = torch.tensor([9000, 3000, 6000, 7001])
    plain, normed = L.Leaf().cuda(), L.Leaf(mean_var_norm=True).cuda()
    normed.load_state_dict(plain.state_dict())
    with torch.no_grad():
        a = plain.forward_prepared(raw, 6000, raw_lengths=lens)
        b = normed.forward_prepared(raw, 6000, raw_lengths=lens)
        assert torch.equal(b, LF.instance_norm(a, 1e-5))
        assert b.mean(dim=2).abs().max().item() < 1e-4


def test_clip_minimum_padding_propagates_nan_like_torch_min():
    """PadToSize(mode='constant') pads with x.min(); torch.min returns NaN for a clip holding a NaN, so the padding --
    and with it every frame of that clip -- is NaN, as when the transform runs on the CPU first."""
    import leaf_pytorch_b200 as L
    from oracle import leaf_oracle as O
    n = 6000
    raw = torch.randn(2, 1, 4000, generator=torch.Generator().manual_seed(3)) * 0.3
    raw[1, 0, 3999] = float("nan")
    prepared = torch.stack([O.prepare_clip(raw[b, 0].numpy(), n, "center", "min", peak_normalize=False) for b in range(2)]).unsqueeze(1)
    assert torch.isnan(prepared[1, 0, 0]) and torch.isfinite(prepared[0]).all()
    fe = L.Leaf().cuda()
    with torch.no_grad():
        got = fe.forward_prepared(raw.cuda(), n, pad_mode="min", peak_normalize=False).cpu()
        same = fe(prepared.cuda()).cpu()
    assert torch.isfinite(got[0]).all() and torch.isnan(got[1]).all()
    assert torch.equal(torch.isnan(got), torch.isnan(same)) and torch.equal(got[0], same[0])
