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
