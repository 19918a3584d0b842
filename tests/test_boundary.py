"""The drop-in boundary on CPU (no GPU, no kernels run): constructor surface, parameter names and
shapes, initial values, error behaviour, C-ABI symbols.  Mirrors what SURVEY 8b lists."""
import contextlib
import ctypes
import io
import os
import re
import sys

import numpy as np
import pytest
import torch

import leaf_pytorch_b200 as L
from leaf_pytorch_b200 import _native
from tests.cases import CASES_BY_NAME

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXPECTED_KEYS = {
    "_complex_conv._kernel": lambda F: (F, 2),
    "_pooling.weights": lambda F: (1, 1, F, 1),
    "_pooling._bias": lambda F: (F,),
    "_compression.alpha": lambda F: (F,),
    "_compression.delta": lambda F: (F,),
    "_compression.root": lambda F: (F,),
    "_compression.ema._weights": lambda F: (F,),
}


@pytest.mark.parametrize("F", [8, 40, 64, 80])
def test_state_dict_layout(F):
    fe = L.Leaf(n_filters=F)
    sd = fe.state_dict()
    assert set(sd) == set(EXPECTED_KEYS)
    for k, shp in EXPECTED_KEYS.items():
        assert tuple(sd[k].shape) == shp(F), k
    assert sum(p.numel() for p in fe.parameters()) == 8 * F
    assert len(list(fe.buffers())) == 0


def test_default_values_and_attributes():
    fe = L.Leaf()
    assert fe._preemp is None and fe._instance_norm is None
    assert float(fe._maximum_val) == pytest.approx(1e-5)
    assert fe._complex_conv._kernel_size == 401 and fe._complex_conv._filters == 40
    assert fe._complex_conv.use_legacy_complex is False
    assert fe._pooling.kernel_size == 401 and fe._pooling.strides == 160 and fe._pooling.in_channels == 40
    assert fe._compression._floor == pytest.approx(1e-12)
    assert torch.all(fe._pooling.weights == 0.4) and torch.all(fe._pooling._bias == 1.0)
    assert torch.allclose(fe._compression.alpha, torch.full((40,), 0.96))
    assert torch.all(fe._compression.delta == 2.0) and torch.all(fe._compression.root == 2.0)
    assert torch.allclose(fe._compression.ema._weights, torch.full((40,), 0.04))
    assert fe.num_frames(16000) == 100 and fe.num_frames(16001) == 101 and fe.num_frames(161) == 2


@pytest.mark.parametrize("name", ["cfg1_default", "F64", "F80", "F8", "sr22050_evenK", "sr8000"])
def test_initial_gabor_parameters_bit_identical_to_reference(name):
    c = CASES_BY_NAME[name]
    fe = L.Leaf(n_filters=c.F, sample_rate=c.sr, window_len=c.wlen, window_stride=c.wstride,
                init_min_freq=c.min_freq, init_max_freq=c.max_freq)
    want = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))["init_kernel"]
    assert np.array_equal(fe._complex_conv._kernel.detach().numpy(), want)


def test_error_behaviour_matches_reference():
    # preemp / mean_var_norm / sort_filters: the reference only declares them (NotImplementedError, frontend.py:40-41,
    # 62-63, convolution.py:74-75); here they are implemented with the original LEAF's semantics and
    # add no state unless switched on
    assert set(L.Leaf(mean_var_norm=True, sort_filters=True).state_dict()) == set(L.Leaf().state_dict())
    assert set(L.Leaf(preemp=True).state_dict()) - set(L.Leaf().state_dict()) == {"_preemp.weight"}
    assert L.Leaf(preemp=True)._preemp.weight.detach().reshape(-1).tolist() == pytest.approx([-0.97, 1.0])
    assert L.Leaf()._preemp is None and L.Leaf()._instance_norm is None
    with pytest.raises(ValueError):
        L.Leaf(initializer="nonsense")
    with pytest.raises(NotImplementedError):
        L.get_frontend({"frontend": {"name": "mel"}, "audio_config": {}})
    fe = L.Leaf()
    with pytest.raises(L.LeafNativeError):           # no silent CPU path
        fe(torch.zeros(2, 1, 800))
    with pytest.raises(RuntimeError):                # stages are fused, not callable alone
        fe._complex_conv(torch.zeros(1, 1, 800))


def test_string_initializers_consume_rng_like_reference():
    for name in ("random", "xavier_normal", "kaiming_normal"):
        torch.manual_seed(3)
        a = L.Leaf(n_filters=8, initializer=name)._complex_conv._kernel.detach().clone()
        torch.manual_seed(3)
        t = torch.randn(8, 2)
        if name == "xavier_normal":
            t = torch.nn.init.xavier_normal_(t)
        elif name == "kaiming_normal":
            t = torch.nn.init.kaiming_normal_(t)
        assert torch.equal(a, t)
    fe = L.Leaf(n_filters=4, initializer=lambda shape: torch.full(shape, 0.5))
    assert torch.all(fe._complex_conv._kernel == 0.5)


def test_get_frontend_reads_reference_config_keys():
    cfg = {"frontend": {"name": "leaf", "default_args": True, "use_legacy_complex": True},
           "audio_config": {"sample_rate": 16000}}
    fe = L.get_frontend(cfg)
    assert fe._complex_conv.use_legacy_complex is True and fe.spec.F == 40
    cfg = {"frontend": {"name": "LEAF", "n_filters": 64, "min_freq": 60.0, "max_freq": 7800.0, "pcen_compress": False},
           "audio_config": {"sample_rate": 16000, "window_len": 25.0, "window_stride": 10.0}}
    fe = L.get_frontend(cfg)
    assert fe.spec.F == 64 and fe._compression is None and fe.spec.compression is False
    assert set(fe.state_dict()) == {"_complex_conv._kernel", "_pooling.weights", "_pooling._bias"}


def test_c_abi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "leafk.h")).read()
    declared = set(re.findall(r"\b(leafk_[a-z_0-9]+)\s*\(", header))
    declared -= {"leafk_params", "leafk_grads", "leafk_config"}
    assert declared, "no prototypes found"
    lib = ctypes.CDLL(_native.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in include/leafk.h but not exported"
    assert set(_native.SYMBOLS) <= declared
    L_ = _native.lib()
    assert L_.leafk_version() == 200
    assert L_.leafk_num_frames(16000, 401, 160) == 100
    lo, hi = ctypes.c_int(), ctypes.c_int()
    L_.leafk_same_padding(552, ctypes.byref(lo), ctypes.byref(hi))
    assert (lo.value, hi.value) == (275, 276)
    cfg = _native.Config(40, 401, 160, 1e-12, 1e-5, 1, 0)
    assert L_.leafk_workspace_bytes(ctypes.byref(cfg), 256, 100) > 0
    assert L_.leafk_tc_supported(40, 401, 160) == 1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "leaf_pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle" not in src.replace("leaf_oracle", "oracle") or fn == "never", (
                    f"{fn} mentions the oracle: the product path must not depend on test infrastructure")


@pytest.mark.reference
def test_state_dict_round_trips_with_the_real_reference():
    sys.path.insert(0, "/root/reference")
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            from leaf_pytorch.frontend import Leaf as RefLeaf
            ref = RefLeaf(n_filters=40)
        ours = L.Leaf(n_filters=40)
        assert list(ref.state_dict()) == list(ours.state_dict())
        for k, v in ref.state_dict().items():
            assert tuple(v.shape) == tuple(ours.state_dict()[k].shape)
            assert torch.equal(v, ours.state_dict()[k]), k           # identical initial values
        ours.load_state_dict(ref.state_dict())
        ref.load_state_dict(ours.state_dict())
    finally:
        sys.path.remove("/root/reference")
        for m in [m for m in sys.modules if m.startswith("leaf_pytorch.") or m == "leaf_pytorch"]:
            del sys.modules[m]


@pytest.mark.reference
def test_reference_classifier_accepts_our_frontend():
    """models.classifier.Classifier (reference models/classifier.py:7-18) builds unchanged when
    leaf_pytorch.get_frontend is ours, and its features sub-module has the reference's keys."""
    sys.path.insert(0, "/root/reference")
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import models.classifier as MC
            orig = MC.get_frontend
            MC.get_frontend = L.get_frontend
            try:
                cfg = {"frontend": {"name": "leaf", "default_args": True, "use_legacy_complex": True},
                       "audio_config": {"sample_rate": 16000},
                       "model": {"arch": "resnet", "num_classes": 35, "model_depth": 18, "pool": "avgpool",
                                 "type": "multiclass"}}
                clf = MC.Classifier(cfg)
            finally:
                MC.get_frontend = orig
        assert isinstance(clf.features, L.Leaf)
        assert {k for k in clf.state_dict() if k.startswith("features.")} == {"features." + k for k in EXPECTED_KEYS}
    finally:
        sys.path.remove("/root/reference")
        for m in [m for m in sys.modules if m.split(".")[0] in ("leaf_pytorch", "models")]:
            del sys.modules[m]


def test_plain_c_client_compiles_links_and_runs_host_calls(tmp_path):
    """examples/c_abi_example.c: a C program (no Python, no torch) builds against include/leafk.h + libleafk.so and
    runs the host-only entry points; without a GPU it stops there."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None or not os.path.isdir("/usr/local/cuda/include"):
        pytest.skip("needs gcc and the CUDA headers")
    exe = tmp_path / "c_abi_example"
    libdir = os.path.dirname(_native.LIB_PATH)
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                    os.path.join(ROOT, "examples", "c_abi_example.c"), "-L", libdir, "-lleafk",
                    "-L/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{libdir}", "-Wl,-rpath,/usr/local/cuda/lib64",
                    "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert "libleafk version 200, 100 frames per clip" in out


@pytest.mark.reference
def test_install_patches_and_restores_the_reference():
    from leaf_pytorch_b200 import integration
    sys.path.insert(0, "/root/reference")
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            import models.classifier as MC
            import leaf_pytorch.frontend as RF
            ref_leaf, ref_factory = RF.Leaf, MC.get_frontend
            done = integration.install()
            assert done["models.classifier.get_frontend"] and done["leaf_pytorch.frontend.Leaf"]
            cfg = {"frontend": {"name": "leaf", "default_args": True}, "audio_config": {"sample_rate": 16000},
                   "model": {"arch": "resnet", "num_classes": 35, "model_depth": 18, "pool": "avgpool", "type": "multiclass"}}
            clf = MC.Classifier(cfg)
            assert isinstance(clf.features, L.Leaf) and RF.Leaf is L.Leaf
            integration.uninstall()
            assert RF.Leaf is ref_leaf and MC.get_frontend is ref_factory
    finally:
        sys.path.remove("/root/reference")
        for m in [m for m in sys.modules if m.split(".")[0] in ("leaf_pytorch", "models")]:
            del sys.modules[m]


def test_bench_reference_arm_prints_contract_json():
    """bench.py --impl reference runs on CPU (the oracle port) and prints one JSON line with the contract keys."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], check=True, capture_output=True, text=True, timeout=600).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "audio_seconds_per_second" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True
    assert "workload" in line["config"] and line["vs_baseline"] is None
    # both arms print the SAME config object for the same --config / world size (the driver compares them key by key)
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.workload_config(2, 1)
    assert set(bench.CONFIGS) == {1, 2, 3, 4, 5}


def test_bench_reference_arm_training_config():
    """--impl reference --config 3: forward + backward of the CPU port on a stated sample of the workload's clips."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "3", "--steps", "1",
                          "--warmup", "1"], check=True, capture_output=True, text=True, timeout=600).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["config"]["config_id"] == 3 and line["config"]["pass"] == "train"
    assert line["step_clips"] == 64 and "forward+backward" in line["cpu_baseline"]["sample"] and line["value"] > 0


def test_header_constants_match_the_python_binding():
    """Every #define of include/leafk.h that the ctypes layer mirrors has the same value there, and the Config
    structure has the fields of leafk_config in the same order."""
    import re
    from leaf_pytorch_b200 import _native as N
    text = open(os.path.join(ROOT, "include", "leafk.h")).read()
    defs = {m.group(1): int(m.group(2).strip("()")) for m in re.finditer(r"#define\s+(LEAFK_\w+)\s+(\(?-?\d+\)?)", text)}
    assert defs["LEAFK_ALGO_AUTO"] == N.ALGO_AUTO and defs["LEAFK_ALGO_FP32"] == N.ALGO_FP32 and defs["LEAFK_ALGO_TC"] == N.ALGO_TC
    assert defs["LEAFK_TC_NOPRUNE"] == N.TC_NOPRUNE
    assert defs["LEAFK_REUSE_BANKS"] == N.REUSE_BANKS
    assert N.ALGOS["tc_full"] == defs["LEAFK_ALGO_TC"] | defs["LEAFK_TC_NOPRUNE"]
    flags = [defs["LEAFK_TC_NOPRUNE"], defs["LEAFK_REUSE_BANKS"]]
    assert all(f > 15 and f & (f - 1) == 0 for f in flags) and len(set(flags)) == 2     # distinct bits above the kernel choice
    body = re.search(r"typedef struct leafk_config \{(.*?)\} leafk_config;", text, re.S).group(1)
    fields = re.findall(r"^\s*(?:int|float|const leafk_clip_prep\*)\s+(\w+);", body, re.M)
    assert fields == [n for n, _ in N.Config._fields_]


def test_host_side_planning_calls_work_without_a_gpu():
    """Planning entry points of the C ABI are pure host code: workspace size (monotone in clips and frames, covers the
    per-clip completion counters), kernel coverage, frame count; the config flags reach the C structure."""
    import torch
    import leaf_pytorch_b200.functional as LF
    from leaf_pytorch_b200 import _native as N
    spec = LF.LeafSpec(F=40, K=401, H=160)
    w1, w2, w3 = LF.workspace_bytes(spec, 4, 100), LF.workspace_bytes(spec, 256, 100), LF.workspace_bytes(spec, 256, 1000)
    assert 0 < w1 < w2 < w3
    assert LF.workspace_bytes(spec, 1 << 16, 100) - LF.workspace_bytes(spec, 1 << 15, 100) >= 4 * (1 << 15)   # counters + partial sums
    assert LF.tc_supported(40, 401, 160) and LF.tc_supported(80, 401, 160) and LF.tc_supported(8, 1601, 480)
    assert not LF.tc_supported(40, 401, 40)            # more than 5 frames per 8 samples: CUDA-core kernel
    assert spec.num_frames(16000) == 100 and spec.num_frames(15999) == 100 and spec.num_frames(16001) == 101
    cfg = LF.LeafSpec(F=40, K=401, H=160, algo="tc_full", out_dtype=torch.bfloat16).config(torch.int16, reuse_banks=True)
    assert cfg.algo == N.ALGO_TC | N.TC_NOPRUNE | N.REUSE_BANKS and cfg.input_format == 1 and cfg.output_format == 1
    # training path planning: the tensor-core training kernel covers the default and the 80-filter geometry, not hop 40
    assert LF.train_supported(spec) and LF.train_supported(LF.LeafSpec(F=80, K=401, H=160))
    assert not LF.train_supported(LF.LeafSpec(F=40, K=401, H=40)) and not LF.train_supported(LF.LeafSpec(F=40, K=401, H=160, algo="fp32"))
    c = spec.config()
    L = N.lib()
    import ctypes as C
    assert L.leafk_train_workspace_bytes(C.byref(c), 256, 16000) > 4 * 256 * 16 * 4 * 40 * 9
    assert L.leafk_backward_saved_workspace_bytes(C.byref(c), 256, 16000, 1) > L.leafk_backward_saved_workspace_bytes(C.byref(c), 256, 16000, 0) > 0
    assert L.leafk_backward_workspace_bytes(C.byref(c), 4, 16000) > 0


def test_kernel_plans_for_the_benchmark_and_edge_geometries():
    """leafk_describe_plan (host only): the channel grouping / training grouping the kernels use.  F <= 48 and the
    default window: one group; F = 49..64: one group of up to 128 channels (lean shared-memory plan) when the hop allows
    the fast tile-end path, else two groups; F = 80: two groups of 80; training: 16 filters x 6 channels per group."""
    import leaf_pytorch_b200.functional as LF
    p = LF.describe_plan(40, 401, 160)
    assert p == {"forward_groups": 1, "forward_channels_per_group": 80, "train_filters_per_group": 16, "frame_slots": 3}
    assert LF.describe_plan(64, 401, 160)["forward_groups"] == 1 and LF.describe_plan(64, 401, 160)["forward_channels_per_group"] == 128
    assert LF.describe_plan(56, 401, 160)["forward_channels_per_group"] == 112
    small_hop = LF.describe_plan(64, 201, 100)                       # hop < 124: no lean plan
    assert small_hop["forward_groups"] == 2 and small_hop["forward_channels_per_group"] == 64
    p80 = LF.describe_plan(80, 401, 160)
    assert (p80["forward_groups"], p80["forward_channels_per_group"]) == (2, 80)
    assert LF.describe_plan(40, 401, 40) == {"forward_groups": 0, "forward_channels_per_group": 0, "train_filters_per_group": 0, "frame_slots": 11}
    long_win = LF.describe_plan(8, 1601, 480)
    assert long_win["forward_groups"] == 1 and long_win["train_filters_per_group"] == 0   # 3 banks of 1601 taps exceed shared memory: generic backward
    many_slots = LF.describe_plan(20, 401, 100)                      # 5 frame slots: small groups, training with 8 filters per group
    assert many_slots["frame_slots"] == 5 and many_slots["forward_channels_per_group"] <= 64 and many_slots["train_filters_per_group"] == 8
    for F, K, H in ((40, 401, 160), (64, 401, 160), (80, 401, 160), (8, 1601, 480), (20, 401, 100), (40, 401, 40)):
        assert LF.tc_supported(F, K, H) == (LF.describe_plan(F, K, H)["forward_groups"] > 0)
        assert LF.train_supported(LF.LeafSpec(F=F, K=K, H=H)) == (LF.describe_plan(F, K, H)["train_filters_per_group"] > 0)
