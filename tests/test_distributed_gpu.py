"""The path's only multi-GPU exchange on real GPUs (NCCL): a training step whose batch is sharded over two ranks --
forward + backward through the kernels on each rank, ONE all-reduce of the flat 8*F gradient buffer -- gives the
gradients of the same step on one GPU (reference train_xla.py:283 analogue).  Skips below 2 GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import leaf_pytorch_b200 as L
        from leaf_pytorch_b200 import distributed as D
        F, B, T = 40, 10, 12000
        g = torch.Generator().manual_seed(77)
        x = torch.randn(B, 1, T, generator=g).clamp_(-4, 4) / 4
        fe = L.Leaf(n_filters=F).to(dev)
        with torch.no_grad():                                  # rank-dependent parameters -> broadcast restores rank 0's
            for p in D.frontend_parameters(fe):
                p.add_(0.01 * rank)
        D.broadcast_frontend_params(fe, src=0)
        G = torch.randn(B, F, fe.num_frames(T), generator=g)
        lo, hi = D.shard_bounds(B, rank, world)
        out = fe(D.shard_batch(x).to(dev))
        out.backward(G[lo:hi].to(dev))
        flat = D.allreduce_frontend_grads(fe, average=False).cpu()
        launches = L.launch_count()
        ok = True
        err = 0.0
        if rank == 0:
            one = L.Leaf(n_filters=F).to(dev)
            one.load_state_dict(fe.state_dict())
            one(x.to(dev)).backward(G.to(dev))
            want = torch.cat([p.grad.reshape(-1) for p in D.frontend_parameters(one)]).cpu()
            err = float((flat - want).abs().max() / want.abs().max())
            ok = err < 1e-5
        gathered = [None] * world
        dist.all_gather_object(gathered, flat.sum().item())
        ret[rank] = (ok, err, gathered[0] == gathered[1], launches)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_training_step_equals_single_gpu_gradients():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        ok, err, same, launches = ret[r]
        assert ok, f"rank {r}: sharded gradients differ from the single-GPU step by {err:.2e} of max|g|"
        assert same                                  # every rank holds the same reduced buffer
        assert launches >= 5                         # the kernels ran on this rank (training forward 3 + backward 2)
