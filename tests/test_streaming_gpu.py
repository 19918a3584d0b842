"""Chunked / streaming forward with carried PCEN state against the reference's un-chunked golden
output (BASELINE.json configs[4]: 60 s clips, 64 filters)."""
import numpy as np
import pytest
import torch

from tests.test_forward_gpu import assert_close, build
from tests.util import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("algo", ["tc", "fp32"])
def test_long60s_chunked_equals_reference_unchunked(algo):
    from leaf_pytorch_b200.streaming import forward_chunked
    case, x, prm, z = load_golden("long60s_F64")
    fe = build(case, prm, algo)
    xg = x.cuda()
    out = forward_chunked(fe, xg, chunk_frames=1000)          # 10 s chunks
    assert_close(out.cpu().numpy(), z["out"], f"long60s chunked/{algo}")
    if algo == "tc":
        with torch.no_grad():
            whole = fe(xg)
        # seams: chunked and un-chunked differ only by fp32 summation order inside tiles / scan segments
        assert float((whole - out).abs().max()) < 2e-6
        odd = forward_chunked(fe, xg, chunk_frames=337)
        assert float((whole - odd).abs().max()) < 2e-6


@pytest.mark.parametrize("name", ["cfg1_default", "perturbed_F40", "T16001", "sr22050_evenK"])
def test_chunked_small_chunks(name):
    from leaf_pytorch_b200.streaming import forward_chunked
    case, x, prm, z = load_golden(name)
    fe = build(case, prm, "auto")
    for chunk in (1, 7, 64):
        out = forward_chunked(fe, x.cuda(), chunk_frames=chunk)
        assert_close(out.cpu().numpy(), z["out"], f"{name} chunk={chunk}")


def test_stream_push_flush_equals_offline():
    from leaf_pytorch_b200.streaming import LeafStream
    case, x, prm, z = load_golden("cfg1_default")
    fe = build(case, prm, "auto")
    xg = x.cuda()
    st = LeafStream(fe, batch=xg.shape[0])
    outs = []
    pos = 0
    for blk in (1, 399, 1600, 37, 4000, 163, 6000, 10000):
        if pos >= xg.shape[2]:
            break
        outs.append(st.push(xg[:, :, pos:pos + blk]))
        pos += blk
    outs.append(st.flush())
    got = torch.cat(outs, dim=2)
    assert got.shape == tuple(z["out"].shape) or tuple(got.shape) == tuple(z["out"].shape)
    assert_close(got.cpu().numpy(), z["out"], "stream")


def test_window_validation_errors():
    import leaf_pytorch_b200 as L
    import leaf_pytorch_b200.functional as LF
    case, x, prm, z = load_golden("cfg1_default")
    fe = build(case, prm, "auto")
    p = [q.detach() for q in fe._param_tuple()]
    xg = x.cuda()
    with pytest.raises(L.LeafNativeError):       # window misses the halo the frames need
        LF.forward_window(fe.spec, xg[:, :, 8000:9000].contiguous(), 16000, 8000, 50, 10, *p)
    with pytest.raises(L.LeafNativeError):       # frame range outside the clip
        LF.forward_window(fe.spec, xg, 16000, 0, 95, 10, *p)
