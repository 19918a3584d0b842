"""Multi-rank probe of the host-buffer leg (development record): bare H2D, bare H2D+D2H, and the HostPipeline loop with
different depths / slice counts / pinned-memory kinds.  Launch with torchrun (one rank per GPU)."""
import ctypes, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
import leaf_pytorch_b200 as L

rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B, T, F = 256, 16000, 40
STEPS = 40
try:
    cudart = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
except OSError:
    cudart = None


def wc_pinned(shape, dtype=np.float32):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = ctypes.c_void_p()
    rc = cudart.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(4))       # cudaHostAllocWriteCombined
    assert rc == 0, rc
    arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(n,)).view(dtype).reshape(shape)
    return torch.from_numpy(arr)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def gather_max(v):
    t = torch.tensor([v], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


g = torch.Generator().manual_seed(1234 + rank)
src = [torch.randn(B, 1, T, generator=g).clamp_(-4, 4) / 4 for _ in range(4)]
hosts = [s.pin_memory() for s in src]
res = {}
# bare H2D / bare H2D + D2H
dst = torch.empty((B, 1, T), device=dev); dout = torch.empty((B, F, 100), device=dev); hout = torch.empty((B, F, 100)).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for both in (False, True):
    barrier(); t0 = time.perf_counter()
    for i in range(STEPS):
        with torch.cuda.stream(s1):
            dst.copy_(hosts[i % 4], non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                hout.copy_(dout, non_blocking=True)
    barrier(); dt = gather_max(time.perf_counter() - t0)
    res["bare_h2d+d2h" if both else "bare_h2d"] = {"ms_per_step": 1e3 * dt / STEPS, "h2d_gbs_total": world * STEPS * B * T * 4 / dt / 1e9}
fe = L.Leaf().to(dev)
outs = [torch.empty((B, F, 100)).pin_memory() for _ in range(4)]


def loop(hs, depth, slices):
    pipe = L.HostPipeline(fe, B, T, depth=depth, n_slices=slices)
    for i in range(depth):
        pipe.result(pipe.submit(hs[i % 4], outs[i % 4]))
    barrier(); t0 = time.perf_counter()
    pend = []
    for i in range(STEPS):
        pend.append(pipe.submit(hs[i % 4], outs[i % 4]))
        if len(pend) >= depth:
            pipe.result(pend.pop(0))
    while pend:
        pipe.result(pend.pop(0))
    barrier(); dt = gather_max(time.perf_counter() - t0)
    pipe.close()
    return {"ms_per_step": 1e3 * dt / STEPS, "audio_s_per_s": world * STEPS * B * T / 16000 / dt}


for depth, slices in ((2, 2), (3, 2), (3, 1), (4, 4), (2, 8)):
    res[f"pipe_depth{depth}_slices{slices}"] = loop(hosts, depth, slices)
if cudart is not None:
    wc = []
    for s in src:
        w = wc_pinned((B, 1, T)); w.copy_(s); wc.append(w)
    res["pipe_depth3_slices2_writecombined"] = loop(wc, 3, 2)
    res["pipe_depth2_slices2_writecombined"] = loop(wc, 2, 2)
    barrier(); t0 = time.perf_counter()
    for i in range(STEPS):
        dst.copy_(wc[i % 4], non_blocking=True)
    barrier(); dt = gather_max(time.perf_counter() - t0)
    res["bare_h2d_writecombined"] = {"ms_per_step": 1e3 * dt / STEPS, "h2d_gbs_total": world * STEPS * B * T * 4 / dt / 1e9}
if rank == 0:
    print(json.dumps({"world": world, **res}, indent=1))
if world > 1:
    dist.destroy_process_group()
