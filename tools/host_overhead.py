"""Host-side cost of one HostPipeline.submit / result (development probe): wall time spent inside the Python + driver
calls, against the device time per batch.  Launch with torchrun for several ranks."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import leaf_pytorch_b200 as L
rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B, T, F = 256, 16000, 40
fe = L.Leaf().to(dev)
hosts = [(torch.randn(B, 1, T).clamp_(-4, 4) / 4).pin_memory() for _ in range(4)]
outs = [torch.empty((B, F, 100)).pin_memory() for _ in range(4)]
pipe = L.HostPipeline(fe, B, T, depth=2, n_slices=2)
for i in range(4):
    pipe.result(pipe.submit(hosts[i % 4], outs[i % 4]))
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
N = 200
t_submit = t_result = 0.0
t0 = time.perf_counter()
prev = None
for i in range(N):
    a = time.perf_counter()
    cur = pipe.submit(hosts[i % 4], outs[i % 2])
    b = time.perf_counter()
    if prev is not None:
        pipe.result(prev)
    c = time.perf_counter()
    t_submit += b - a; t_result += c - b
    prev = cur
pipe.result(prev)
torch.cuda.synchronize()
tot = time.perf_counter() - t0
print(json.dumps({"rank": rank, "world": world, "ms_per_step": 1e3 * tot / N, "submit_us": 1e6 * t_submit / N, "result_wait_us": 1e6 * t_result / N}))
if world > 1:
    dist.destroy_process_group()
