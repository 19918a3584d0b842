"""Multi-GPU plumbing for the frontend: one process per GPU, clips sharded over ranks.

The forward path has NO collective: every clip is independent through the Gabor correlation, pooling
and PCEN, and the parameters (8*F floats) are replicated.  Training needs exactly one
exchange per step: the sum of the 8*F frontend parameter gradients.  The reference does this only on
TPU (xm.optimizer_step, reference train_xla.py:283); here it is one NCCL all-reduce of a single flat
buffer (1.28 KB at F=40) -- latency-bound, so the seven tensors are packed into one message.
Works with any torch.distributed backend (NCCL on the B200s, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers it allocates
    afterwards (first touch) and the threads that fill them are local to the GPU's PCIe root: with one process
    per GPU the host-to-device copies of the eight ranks otherwise cross the socket interconnect.  Linux only;
    returns the node, or None when the topology cannot be read (nothing is changed then)."""
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        with open(path) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def shard_bounds(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first (n_items % world) ranks get one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    """This rank's clips of a (B,1,T) batch (a view; no copy, no communication)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def frontend_parameters(leaf) -> List[torch.nn.Parameter]:
    """The module's own parameters in registration order: (pre-emphasis weight,) Gabor kernel, pooling width and bias,
    alpha, delta, root, smoother coefficient -- the leaves that own the gradients (not the sorted view of the kernel
    that sort_filters feeds to the kernels)."""
    return list(leaf.parameters())


def allreduce_frontend_grads(leaf, group=None, average: bool = True) -> Optional[torch.Tensor]:
    """Sum (or mean) the frontend parameter gradients over ranks with ONE all-reduce of a flat
    buffer; parameters without a gradient contribute zeros.  Returns the flat reduced buffer."""
    params = frontend_parameters(leaf)
    if not params:
        return None
    dev = params[0].device
    flat = torch.zeros(sum(p.numel() for p in params), dtype=torch.float32, device=dev)
    off = 0
    for p in params:
        if p.grad is not None:
            flat[off:off + p.numel()] = p.grad.detach().reshape(-1)
        off += p.numel()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        g = flat[off:off + p.numel()].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += p.numel()
    return flat


def broadcast_frontend_params(leaf, src: int = 0, group=None) -> None:
    """Replicate rank ``src``'s frontend parameters (one flat broadcast)."""
    params = frontend_parameters(leaf)
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    with torch.no_grad():
        for p in params:
            p.copy_(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
