"""Host-side multi-rank logic on CPU with the gloo backend, world_size 2 (no GPU needed)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import leaf_pytorch_b200 as L
        from leaf_pytorch_b200 import distributed as D
        torch.manual_seed(0)
        fe = L.Leaf(n_filters=8)
        # different parameters per rank -> broadcast makes them equal to rank 0's
        with torch.no_grad():
            for p in D.frontend_parameters(fe):
                p.add_(float(rank))
        D.broadcast_frontend_params(fe, src=0)
        ref = L.Leaf(n_filters=8)
        same = all(torch.equal(a, b) for a, b in zip(D.frontend_parameters(fe), D.frontend_parameters(ref)))
        # rank-dependent fake gradients; one parameter left without a gradient on rank 1
        for i, p in enumerate(D.frontend_parameters(fe)):
            if not (rank == 1 and i == 2):
                p.grad = torch.full_like(p, float((rank + 1) * (i + 1)))
        flat = D.allreduce_frontend_grads(fe, average=False)
        want = []
        for i, p in enumerate(D.frontend_parameters(fe)):
            v = 1.0 * (i + 1) + (0.0 if i == 2 else 2.0 * (i + 1))
            want.append(torch.full((p.numel(),), v))
        want = torch.cat(want)
        ok_sum = torch.equal(flat, want) and all(torch.equal(p.grad.reshape(-1), w) for p, w in zip(
            D.frontend_parameters(fe), torch.split(want, [p.numel() for p in D.frontend_parameters(fe)])))
        x = torch.arange(5 * 3, dtype=torch.float32).reshape(5, 1, 3)
        shard = D.shard_batch(x)
        lo, hi = D.shard_bounds(5, rank, world)
        ok_shard = torch.equal(shard, x[lo:hi])
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi))
        ret[rank] = (same, ok_sum, ok_shard, gathered, flat.numel())
    finally:
        dist.destroy_process_group()


def test_gloo_world2_allreduce_broadcast_shard():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        same, ok_sum, ok_shard, gathered, n = ret[r]
        assert same and ok_sum and ok_shard
        assert n == 8 * 8                      # 8*F floats in ONE message
        assert gathered == [(0, 3), (3, 5)]    # disjoint cover of the batch


def test_shard_bounds_cover_and_balance():
    from leaf_pytorch_b200.distributed import shard_bounds
    for n in (1, 7, 8, 256, 513):
        for w in (1, 2, 4, 8):
            parts = [shard_bounds(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def test_numa_binding_helpers_are_safe_without_a_gpu():
    """bind_to_gpu_numa_node never raises and changes nothing when the topology cannot be read (no GPU here)."""
    import os
    from leaf_pytorch_b200 import distributed as D
    assert D._parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert D._parse_cpulist("") == []
    before = os.sched_getaffinity(0)
    assert D.bind_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before


def test_frontend_parameters_are_the_modules_own_leaves():
    """With sort_filters the kernels see a sorted VIEW of the Gabor kernel; the all-reduce must address the leaf that owns
    the gradient, and the pre-emphasis weight belongs to the frontend's parameters too."""
    import leaf_pytorch_b200 as L
    import leaf_pytorch_b200.distributed as D
    fe = L.Leaf(n_filters=8, sort_filters=True, preemp=True)
    ps = D.frontend_parameters(fe)
    assert all(isinstance(p, torch.nn.Parameter) and p.is_leaf for p in ps)
    assert sum(p.numel() for p in ps) == 8 * 8 + 2
    assert [id(p) for p in ps] == [id(p) for p in fe.parameters()]
    for p in ps:
        p.grad = torch.full_like(p, 2.0)
    flat = D.allreduce_frontend_grads(fe)                  # no process group: identity, but exercises the packing
    assert flat.numel() == 66 and torch.all(flat == 2.0)


def test_side_entry_points_refuse_modules_with_optional_stages():
    """forward_host / HostPipeline / forward_chunked / LeafStream run the fused path only; with preemp or mean_var_norm
    they must refuse instead of silently skipping the stage."""
    import pytest
    import leaf_pytorch_b200 as L
    from leaf_pytorch_b200.streaming import forward_chunked, LeafStream
    x = torch.zeros(1, 1, 4000)
    for kw in ({"preemp": True}, {"mean_var_norm": True}):
        fe = L.Leaf(n_filters=8, **kw)
        with pytest.raises(NotImplementedError):
            fe.forward_host(x)
        with pytest.raises(NotImplementedError):
            L.HostPipeline(fe, 1, 4000)
        with pytest.raises(NotImplementedError):
            forward_chunked(fe, x)
        with pytest.raises(NotImplementedError):
            LeafStream(fe, 1)
    with pytest.raises(NotImplementedError):
        L.Leaf(n_filters=8, preemp=True).forward_prepared(x, 2000)
