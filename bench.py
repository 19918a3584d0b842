#!/usr/bin/env python
"""bench.py -- headline benchmark of the LEAF frontend hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1..5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic waveforms per GPU (weak scaling: every rank runs its
own batch; no collective on the forward path).  --config selects the BASELINE.json configuration (default 2 = the
one the metric is quoted on):
    1  default Leaf (F=40, K=401, hop 160), 4 x 1 s, forward                       (the reference's CPU-runnable case)
    2  default Leaf, 256 x 1 s per GPU, forward                                    (headline)
    3  F=80, 1024 x 1 s per GPU, forward + backward (7 parameter gradients); multi-GPU: + NCCL all-reduce of the grads
    4  F=40, 10 s clips, 64 per GPU (= 512 over 8 GPUs), forward
    5  F=64, 60 s clips, 8 per GPU (= 64 over 8 GPUs), forward in 10 s chunks with carried PCEN state

Printed JSON (rank 0, one line):
  value        audio-seconds per second, whole job, inputs resident in HBM, CUDA-event timed
  e2e          same metric from pinned HOST buffers: every step's input goes H2D and its result D2H inside the timed region
  roofline     the Gabor GEMM kernel alone: algorithmic FLOPs / its CUDA-event duration vs the measured dense bf16/fp16
               tensor peak in MEASURED_PEAKS.json (the path is tensor-bound: ~12.9 kFLOP per HBM byte), the HBM view the
               metric asks for, the PCEN kernel's HBM fraction, and the measured concurrent H2D ceiling of the box
  cpu_baseline the oracle port (same ATen CPU ops as the reference) on the host cores, bounded sample   (N=1 only)
  gpu_torch_baseline  the same torch ops on this GPU (cuDNN conv1d + eager ops + Python smoother loop), allow_tf32 on/off
--impl reference times the CPU port alone (the reference is pure Python/torch; /root/reference does not exist on the
GPU box, so its own file cannot be imported there -- see DESIGN.md).
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

SR = 16000
METRIC = "audio_seconds_per_second"
UNIT = "audio-s/s"
L2_BYTES = 126e6

CONFIGS = {
    1: dict(F=40, B=4, T=16000, mode="fwd", cpu_B=4, gpu_B=4,
            name="configs[0]: default Leaf (F=40, K=401, hop=160, 16 kHz), batch 4 x 1 s, forward only"),
    2: dict(F=40, B=256, T=16000, mode="fwd", cpu_B=256, gpu_B=64,
            name="configs[1]: default Leaf (F=40, K=401, hop=160, 16 kHz), batch 256 x 1 s per GPU, forward only"),
    3: dict(F=80, B=1024, T=16000, mode="train", cpu_B=64, gpu_B=64,
            name="configs[2]: 80 filters / 25 ms window / 10 ms hop, batch 1024 x 1 s per GPU, forward+backward (param grads)"),
    4: dict(F=40, B=64, T=160000, mode="fwd", cpu_B=8, gpu_B=8,
            name="configs[3]: AudioSet shape, 10 s @16 kHz clips, 40 filters, 64 clips per GPU (batch 512 over 8 GPUs), forward"),
    5: dict(F=64, B=8, T=960000, mode="chunked", chunk_frames=1000, cpu_B=1, gpu_B=1,
            name="configs[4]: long-form 60 s @16 kHz, 64 filters, 8 clips per GPU (batch 64 over 8 GPUs), "
                 "10 s chunks with carried PCEN state"),
}
K_TAPS, HOP = 401, 160


def synth_batch(B: int, T: int, seed: int) -> torch.Tensor:
    """SURVEY 8d synthetic input: clamp(randn,-4,4)/4, |x| <= 1 like the peak-normalised pipeline."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 1, T, generator=g).clamp_(-4, 4) / 4


def workload_config(cfg_id: int, world: int) -> dict:
    """The workload description both arms print (same keys, same values)."""
    c = CONFIGS[cfg_id]
    n_rot = n_rotate(c)
    return {"workload": c["name"], "config_id": cfg_id, "n_filters": c["F"], "taps": K_TAPS, "hop": HOP,
            "batch_per_gpu": c["B"], "samples_per_clip": c["T"], "pass": c["mode"],
            "parallelism": f"batch-sharded x{world}, no collective on the forward"
                           + (", NCCL all-reduce of the 8*F parameter gradients per step" if c["mode"] == "train" and world > 1 else ""),
            "l2": (f"{n_rot} rotating input batches ({n_rot * c['B'] * c['T'] * 4 / 1e6:.0f} MB > 126 MB L2)"
                   if n_rot * c["B"] * c["T"] * 4 > L2_BYTES else "L2 flushed between timed steps (256 MB write), per-step events")}


def n_rotate(c) -> int:
    per = c["B"] * c["T"] * 4
    return max(2, min(10, int(L2_BYTES * 1.3 / per) + 1))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), tensor=float(d["bf16_tflops"]),
                    tensor_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU / torch baselines
def oracle_params(F: int, device="cpu", requires_grad=False):
    from oracle import leaf_oracle as O
    import leaf_pytorch_b200 as L
    fe = L.Leaf(n_filters=F)
    prm = O.params_from_state_dict({k: v.detach() for k, v in fe.state_dict().items()})
    return {k: (None if v is None else v.clone().to(device).requires_grad_(requires_grad)) for k, v in prm.items()}


def torch_ops_step(c, x, prm, G):
    """One step of the reference's torch-op path (oracle restatement) on x's device: forward, or forward+backward."""
    from oracle import leaf_oracle as O
    if c["mode"] == "train":
        for v in prm.values():
            if v is not None:
                v.grad = None
        out = O.forward_on_device(x, prm, K_TAPS, HOP)
        out.backward(G)
        return out
    with torch.no_grad():
        return O.forward_on_device(x, prm, K_TAPS, HOP)


def cpu_port_throughput(cfg_id: int, min_runs: int = 3, budget_s: float = 10.0):
    """The oracle port of the reference CPU path, all host threads, bounded sample: at least
    `min_runs` passes over the sample, repeated until about `budget_s` seconds of CPU work are timed (cap 30 s)."""
    c = CONFIGS[cfg_id]
    torch.set_num_threads(os.cpu_count() or 1)
    Bs = min(c["B"], c["cpu_B"])
    prm = oracle_params(c["F"], "cpu", requires_grad=c["mode"] == "train")
    x = synth_batch(Bs, c["T"], 1234)
    G = torch.randn(Bs, c["F"], (c["T"] - 1) // HOP + 1, generator=torch.Generator().manual_seed(1235))
    torch_ops_step(c, x, prm, G)                                   # warm-up
    ts = []
    while len(ts) < 200:
        t0 = time.perf_counter()
        torch_ops_step(c, x, prm, G)
        ts.append(time.perf_counter() - t0)
        if (len(ts) >= min_runs and sum(ts) >= budget_s) or sum(ts) > 30:
            break
    t = statistics.median(ts)
    what = "forward+backward" if c["mode"] == "train" else "forward"
    return {"value": (Bs * c["T"] / SR) / t, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{Bs} of the {c['B']} clips, {what} (1 warm-up + {len(ts)} runs = {sum(ts):.1f} s of CPU work, median {t:.2f} s per run)"}


def gpu_torch_baseline(cfg_id: int, dev):
    """Stock torch ops on this GPU (what the reference runs on CUDA): cuDNN conv1d + eager ops + Python smoother loop."""
    c = CONFIGS[cfg_id]
    Bs = min(c["B"], c["gpu_B"])
    res = {"sample": f"{Bs} of the {c['B']} clips per step (the torch path materialises the (B,2F,T) activation: "
                     f"{Bs * 2 * c['F'] * c['T'] * 4 / 1e9:.1f} GB here), same pass as the bench", "unit": UNIT}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        prm = oracle_params(c["F"], dev, requires_grad=c["mode"] == "train")
        x = synth_batch(Bs, c["T"], 1234).to(dev)
        G = torch.randn(Bs, c["F"], (c["T"] - 1) // HOP + 1, generator=torch.Generator().manual_seed(1235)).to(dev)
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch_ops_step(c, x, prm, G)
            torch.cuda.synchronize(dev)
            n = 3
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                torch_ops_step(c, x, prm, G)
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / n
            res["allow_tf32_true" if tf32 else "allow_tf32_false"] = {"value": Bs * c["T"] / SR / (ms * 1e-3), "ms_per_step": ms}
    except Exception as ex:                                          # e.g. out of memory: record, do not fail the bench
        res["error"] = f"{type(ex).__name__}: {str(ex)[:200]}"
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        torch.cuda.empty_cache()
    return res


def run_reference(args, rank: int, world: int, emit):
    """--impl reference: the reference's CPU implementation of the path (oracle port; kind 'port')."""
    if rank != 0:
        return
    c = CONFIGS[args.config]
    steps = max(1, args.steps)
    torch.set_num_threads(os.cpu_count() or 1)
    Bs = min(c["B"], c["cpu_B"])
    prm = oracle_params(c["F"], "cpu", requires_grad=c["mode"] == "train")
    x = synth_batch(Bs, c["T"], 1234)
    G = torch.randn(Bs, c["F"], (c["T"] - 1) // HOP + 1, generator=torch.Generator().manual_seed(1235))
    for _ in range(max(1, min(args.warmup, 2))):
        torch_ops_step(c, x, prm, G)
    t0 = time.perf_counter()
    n_done = 0
    for _ in range(steps):
        torch_ops_step(c, x, prm, G)
        n_done += 1
        if time.perf_counter() - t0 > 120:                        # keep the whole arm within minutes
            break
    dt = time.perf_counter() - t0
    val = n_done * Bs * c["T"] / SR / dt
    cores = torch.get_num_threads()
    what = "forward+backward" if c["mode"] == "train" else "forward"
    sample = (f"each step = {Bs} of the {c['B']} clips of the workload, {what} on the host cores (CPU throughput is flat in "
              f"batch); {n_done} steps timed")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": n_done,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / n_done, "step_clips": Bs, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.config, world),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def k1_source_sha16() -> str:
    h = hashlib.sha256()
    for name in ("k1_tc_kernel.cuh", "k1_tc.cu", "k1_tc_layout.cuh", "tc_ptx.cuh", "k0_banks.cu"):
        with open(os.path.join(ROOT, "leaf_pytorch_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algo", default="auto", choices=["auto", "tc", "tc_full", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    # stdout carries exactly ONE line (the JSON): anything a library prints there (NCCL's version banner, ...)
    # goes to stderr instead
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return

    import torch.distributed as dist
    import leaf_pytorch_b200 as L
    import leaf_pytorch_b200.functional as LF
    from leaf_pytorch_b200.streaming import forward_chunked
    from leaf_pytorch_b200.distributed import bind_to_gpu_numa_node, allreduce_frontend_grads

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None     # before any pinned allocation
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    c = CONFIGS[args.config]
    B, T, F, K, H, mode = c["B"], c["T"], c["F"], K_TAPS, HOP, c["mode"]
    W = max(3, args.warmup)
    steps = max(1, args.steps)
    fe = L.Leaf(n_filters=F, algo=args.algo).to(dev)
    if mode != "train":
        fe.requires_grad_(False)
    n_frames = fe.num_frames(T)
    NR = n_rotate(c)
    flush_l2 = NR * B * T * 4 <= L2_BYTES
    xs_host = [synth_batch(B, T, 1234 + 17 * i + 1000 * rank).pin_memory() for i in range(NR)]
    xs = [x.to(dev) for x in xs_host]
    G = torch.randn(B, F, n_frames, generator=torch.Generator().manual_seed(1235)).to(dev) if mode == "train" else None
    flush_buf = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev) if flush_l2 else None
    params = [p for p in fe.parameters()]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(x):
        """one pass of the hot path over one device-resident batch"""
        if mode == "fwd":
            with torch.no_grad():
                return fe(x)
        if mode == "chunked":
            return forward_chunked(fe, x, chunk_frames=c["chunk_frames"])
        fe.zero_grad(set_to_none=True)
        out = fe(x)
        out.backward(G)
        if world > 1:
            allreduce_frontend_grads(fe)                   # the only collective of the path (train_xla.py:283 analogue)
        return out

    def timed_steps(n, first):
        """CUDA-event time of n steps in ms (L2 flushed between steps when the rotating inputs fit in L2)"""
        if not flush_l2:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(n):
                step(xs[(first + i) % NR])
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1)
        evs = []
        for i in range(n):
            flush_buf.fill_(float(i))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); step(xs[(first + i) % NR]); b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)

    # ------------------------------------------------------------- device-resident throughput
    for i in range(W):
        step(xs[i % NR])
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    LF.launch_count(reset=True)
    barrier()
    ms_total = timed_steps(steps, W)
    barrier()
    launches = LF.launch_count()

    # ------------------------------------------------------------- end to end from pinned host buffers
    out_elems = B * F * n_frames
    if mode == "fwd" and LF.tc_supported(F, K, H) and args.algo != "fp32":
        def e2e_loop(hosts, dtype, out_dtype=torch.float32):
            fe.out_dtype = out_dtype
            pipe = L.HostPipeline(fe, B, T, depth=2, n_slices=2, input_dtype=dtype)
            fe.out_dtype = torch.float32
            outs = [torch.empty((B, F, n_frames), dtype=out_dtype).pin_memory() for _ in range(2)]
            pipe.result(pipe.submit(hosts[0], outs[0]))
            pipe.result(pipe.submit(hosts[1 % len(hosts)], outs[1]))
            barrier()
            t_start = time.perf_counter()
            prev = pipe.submit(hosts[0], outs[0])
            for i in range(1, steps):
                cur = pipe.submit(hosts[i % len(hosts)], outs[i % 2])
                pipe.result(prev)
                prev = cur
            pipe.result(prev)
            barrier()
            dt = time.perf_counter() - t_start
            pipe.close()
            return dt
        e2e_s = e2e_loop(xs_host, torch.float32)
        pcm_host = [(xh * 32767.0).round().to(torch.int16).pin_memory() for xh in xs_host[:min(4, NR)]]
        pcm_s = e2e_loop(pcm_host, torch.int16)
        compact_s = e2e_loop(pcm_host, torch.int16, torch.bfloat16)
        e2e_api = ("HostPipeline.submit/result -> leafk_forward_host_async: pinned host in/out, 2 batches in flight, "
                   "H2D in 2 slices with ready flags feeding one persistent launch per batch")
        d2h_bytes = out_elems * 4
    else:
        # training step / chunked long clips: double-buffered uploads on a copy stream, the step on the current stream,
        # the result (gradients / features) read back to pinned memory every step
        copy_s = torch.cuda.Stream(device=dev)
        d2h_s = torch.cuda.Stream(device=dev)
        dbuf = [torch.empty((B, 1, T), dtype=torch.float32, device=dev) for _ in range(2)]
        up_done = [torch.cuda.Event() for _ in range(2)]
        used = [torch.cuda.Event() for _ in range(2)]
        if mode == "train":
            res_host = torch.empty(sum(p.numel() for p in params), dtype=torch.float32).pin_memory()
            d2h_bytes = res_host.numel() * 4
        else:
            res_host = torch.empty((B, F, n_frames), dtype=torch.float32).pin_memory()
            d2h_bytes = out_elems * 4
        cur_s = torch.cuda.current_stream(dev)

        def upload(i):
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(used[i % 2])
                dbuf[i % 2].copy_(xs_host[i % NR], non_blocking=True)
                up_done[i % 2].record(copy_s)

        def e2e_loop_generic(n):
            ring = [None, None, None]          # results stay alive until their read-back has finished (no allocator churn)
            d2h_done = [torch.cuda.Event() for _ in range(3)]
            upload(0)
            for i in range(n):
                if i + 1 < n:
                    upload(i + 1)
                cur_s.wait_event(up_done[i % 2])
                if ring[i % 3] is not None:
                    cur_s.wait_event(d2h_done[i % 3])    # the buffer about to be released has been read back
                out = step(dbuf[i % 2])
                used[i % 2].record(cur_s)
                res = torch.cat([p.grad.reshape(-1) for p in params]) if mode == "train" else out
                ring[i % 3] = res
                d2h_s.wait_stream(cur_s)                 # the read-back overlaps the next step
                with torch.cuda.stream(d2h_s):
                    res_host.copy_(res, non_blocking=True)
                    d2h_done[i % 3].record(d2h_s)
            torch.cuda.synchronize()
        e2e_loop_generic(4)
        barrier()
        t_start = time.perf_counter()
        e2e_loop_generic(steps)
        barrier()
        e2e_s = time.perf_counter() - t_start
        pcm_s = compact_s = None
        e2e_api = ("Leaf module on double-buffered uploads: pinned host batch -> device on a copy stream while the previous "
                   "step runs, the step, its result (7 parameter gradients / features) -> pinned host every step")
    clocks = sampler.stop() if rank == 0 else None

    # ------------------------------------------------------------- concurrent bare copy ceilings (all ranks at once)
    # (a) the step's input H2D alone; (b) the step's traffic mix: input H2D and result D2H on two streams at the same
    # time, no kernels.  (b) is the ceiling of the end-to-end leg on this box: on the 4/8-GPU VMs the two directions
    # share the host side and the H2D rate drops when results flow back (profiles/r02_e2e_probe_n4.json).
    big = xs_host[0]
    dst = torch.empty_like(big, device=dev)
    d2h_elems = max(1, d2h_bytes // 4)
    dsrc = torch.empty(d2h_elems, dtype=torch.float32, device=dev)
    hdst = torch.empty(d2h_elems, dtype=torch.float32).pin_memory()
    reps = max(4, int(256e6 / (big.numel() * 4)))
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    copy_ms = []
    for both in (False, True):
        dst.copy_(big, non_blocking=True)
        barrier()
        t_start = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(s_in):
                dst.copy_(big, non_blocking=True)
            if both:
                with torch.cuda.stream(s_out):
                    hdst.copy_(dsrc, non_blocking=True)
        barrier()
        copy_ms.append((time.perf_counter() - t_start) * 1e3)
    h2d_ms, mix_ms = copy_ms
    del dst, dsrc, hdst

    # ------------------------------------------------------------- per-kernel durations (roofline)
    LF.profile_begin()
    for i in range(steps):
        step(xs[i % NR])
    torch.cuda.synchronize()
    n_prof, ms_k0, ms_k1, ms_k2 = LF.profile_end()
    calls_per_step = n_prof / steps if steps else 1.0
    k1_cyc = k1_ns = 0
    sched = None
    tc_used = args.algo != "fp32" and LF.tc_supported(F, K, H)
    if tc_used and mode != "train":
        with torch.no_grad():
            prm_t = [None if q is None else q.detach() for q in fe._param_tuple()]
            probe_x = xs[0] if mode != "chunked" else xs[0][:, :, :160000].contiguous()
            for i in range(3):
                fe(probe_x)
            k1_cyc, k1_ns = LF.k1_clock_probe(fe.spec, probe_x, *prm_t)
            sched = LF.tc_schedule(fe.spec, probe_x, *prm_t)

    vals = torch.tensor([ms_total, e2e_s * 1e3, (pcm_s or 0.0) * 1e3, h2d_ms, mix_ms, (compact_s or 0.0) * 1e3],
                        dtype=torch.float64, device=dev)
    if world > 1:
        gathered = [torch.zeros_like(vals) for _ in range(world)]
        dist.all_gather(gathered, vals)
        per_rank = torch.stack(gathered).cpu()
    else:
        per_rank = vals.cpu().unsqueeze(0)
    ms_total, e2e_ms, pcm_ms, h2d_ms, mix_ms, compact_ms = [float(v) for v in per_rank.max(dim=0).values]

    if rank == 0:
        audio_s_step = world * B * T / SR
        value = audio_s_step * steps / (ms_total * 1e-3)
        e2e_value = audio_s_step * steps / (e2e_ms * 1e-3)
        peaks = measured_peaks()
        passes = 3.0 if mode == "train" else 1.0               # backward = 2 x forward
        flops_alg = passes * 2.0 * (2 * F) * K * T * B          # conv term, per step and GPU
        bytes_alg = 4.0 * B * T + 4.0 * B * F * n_frames + 32.0 * F
        if mode == "train":
            bytes_alg += 4.0 * B * F * n_frames + 16.0 * B * F * n_frames     # grad_out + saved p, Q_mu, Q_sigma, Q_poolw
        k1_ms_step = ms_k1 * calls_per_step
        k1_s = k1_ms_step * 1e-3
        long_run = ms_total > 200.0                             # held for >0.2 s the board sits at its power cap
        peak = peaks["tensor_sustained"] if long_run else peaks["tensor"]
        if not tc_used:
            kernel, bound, exec_mult, exec_frac = "k1_fp32_kernel", "fp32-fma", 1.0, 1.0
        elif mode == "train":
            kernel = "k1_tc_kernel<96,3,1,26> on CTA pairs (training forward: y, dy/dmu, dy/dsigma banks + 4 pooled quantities)"
            bound, exec_frac = "tensor", 1.0
            exec_mult = 3.0 * ((K + 15) // 16 * 16) / K          # 3 banks are the 3 algorithmic passes; x 3 fp16 products
        else:
            cg = sched["channels_per_group"] if sched else 80
            kernel = f"k1_tc_kernel<{cg},3,0,0> on CTA pairs (Gabor Toeplitz GEMM + modulus + pooling partials)"
            bound = "tensor"
            exec_frac = sched["executed_fraction"] if sched else 1.0
            exec_mult = 3.0 * ((K + 15) // 16 * 16) / K * exec_frac
        n_elem = B * F * n_frames
        roofline = {
            "kernel": kernel, "bound": bound,
            "achieved": flops_alg / k1_s / 1e12, "peak": peak, "unit": "TFLOP/s", "frac": flops_alg / k1_s / 1e12 / peak,
            "peak_source": f"{peaks['source']} MEASURED_PEAKS.json " + ("bf16_tflops_sustained (timed region > 0.2 s: the board "
                           "sits at its power cap, see clocks)" if long_run else "bf16_tflops (burst; fp16 runs at the same rate)"),
            "frac_of_burst_peak": flops_alg / k1_s / 1e12 / peaks["tensor"],
            "frac_of_sustained_peak": flops_alg / k1_s / 1e12 / peaks["tensor_sustained"],
            "algorithmic_flops_per_step": flops_alg,
            "algorithmic_note": "2*(2F)*K*T*B per correlation pass" + (" x 3 (backward = 2 x forward, SURVEY 8d)" if mode == "train" else ""),
            "executed_tflops": exec_mult / passes * flops_alg / k1_s / 1e12,
            "executed_frac": exec_mult / passes * flops_alg / k1_s / 1e12 / peak,
            "executed_note": "tensor work actually issued: 3 fp16 products per fp32 product (hi/lo split), taps padded 401->416"
                             + (", times the share of (channel, k-step) pairs the support pruning keeps" if mode != "train" else
                                "; the training kernel runs every tap of the three banks"),
            "pruning": None if not sched else {"executed_fraction": exec_frac, "active_channels_per_kstep": sched["active"],
                                               "all_three_products_channels_per_kstep": sched["active_all_products"],
                                               "channels_per_group": sched["channels_per_group"]},
            "k1_ms": k1_ms_step, "k0_ms": ms_k0 * calls_per_step, "k2_ms": ms_k2 * calls_per_step,
            "kernel_launches_profiled": n_prof, "forward_calls_per_step": calls_per_step,
            "k1_sm_cycles": k1_cyc, "k1_sm_mhz_effective": (1e3 * k1_cyc / k1_ns) if k1_ns else None,
            "traffic": None,
            "hbm": {"algorithmic_bytes": bytes_alg, "achieved": bytes_alg / k1_s / 1e9, "peak": peaks["hbm"],
                    "unit": "GB/s", "frac": bytes_alg / k1_s / 1e9 / peaks["hbm"],
                    "note": "path is compute-bound by ~1000x (SURVEY 8d): HBM fraction is reported because "
                            "the metric names it, the tensor fraction is the binding one"},
            "k2": {"kernel": "k2_pcen_kernel (assembly + floor + EMA scan + PCEN)", "bound": "hbm",
                   "algorithmic_bytes": 8.0 * n_elem, "achieved": 8.0 * n_elem / (ms_k2 * calls_per_step * 1e-3) / 1e9,
                   "peak": peaks["hbm"], "unit": "GB/s",
                   "frac": 8.0 * n_elem / (ms_k2 * calls_per_step * 1e-3) / 1e9 / peaks["hbm"],
                   "note": "instruction-bound on three accurate exp2f(y*log2f(x)) per element; launched under K1's tail "
                           "(per-clip completion counters), so its event interval also counts the wait for K1's last tiles"},
            "h2d_ceiling_gbs": world * reps * big.numel() * 4 / (h2d_ms * 1e-3) / 1e9,
            "h2d_ceiling_note": f"all {world} rank(s) copying a pinned {big.numel() * 4 / 1e6:.0f} MB batch H2D {reps} x at the same time "
                                "(bare copy_, no kernels), aggregate",
            "copy_ceiling_ms_per_step": mix_ms / reps,
            "copy_ceiling_value": audio_s_step / (mix_ms / reps * 1e-3),
            "copy_ceiling_note": "bare copies of ONE step's traffic mix per rank -- input H2D and result D2H on two streams at the "
                                 "same time, all ranks concurrently, no kernels: the most audio-s/s any end-to-end path from "
                                 "host buffers can reach on this box",
        }
        prof_path = os.path.join(ROOT, "profiles", "r02_k1_traffic.json")
        if os.path.isfile(prof_path):
            try:
                with open(prof_path) as f:
                    tr = json.load(f)
                ent = tr.get(str(args.config))
                if ent and ent.get("k1_source_sha16") == k1_source_sha16():
                    roofline["traffic"] = ent.get("dram_bytes_per_launch")
                    roofline["traffic_source"] = ent.get("source")
                else:
                    roofline["traffic_source"] = "no ncu capture of this config for the current kernel sources"
            except (OSError, ValueError):
                pass
        e2e_h2d_gbs = world * B * T * 4 * steps / (e2e_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": W,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.config, world),
            "implementation": {
                "algo": ("tc_full" if args.algo == "tc_full" else "tc") if tc_used else "fp32",
                "arithmetic": "fp16 hi/lo split operands (3 products), fp32 accumulate; ~2^-21 relative; inference forward: taps "
                              "beyond 5.5 sigma of a filter skipped per 16-tap step (< 2.7e-7 of its peak), beyond 3.7 sigma only "
                              "the main product; training forward: every tap",
                "training": None if mode != "train" else "leafk_forward_train pools the gradient bilinear forms; the backward "
                                                         "(leafk_backward_saved) runs no correlation"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * B * T * 4,
                    "d2h_bytes_per_step": world * d2h_bytes, "ms_per_step": e2e_ms / steps, "api": e2e_api,
                    "h2d_gbs": e2e_h2d_gbs, "h2d_frac_of_ceiling": e2e_h2d_gbs / roofline["h2d_ceiling_gbs"],
                    "frac_of_copy_ceiling": min(1.0, (mix_ms / reps) / (e2e_ms / steps)),
                    "bound": "copies (host side of the box)" if (mix_ms / reps) > 1.05 * (ms_total / steps) else "kernels"},
            "per_rank": {"ms_per_step": [float(v) / steps for v in per_rank[:, 0]],
                         "e2e_ms_per_step": [float(v) / steps for v in per_rank[:, 1]],
                         "bare_h2d_gbs": [reps * big.numel() * 4 / (float(v) * 1e-3) / 1e9 for v in per_rank[:, 3]],
                         "numa_node_rank0": numa_node},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if pcm_s is not None:
            line["e2e_pcm16"] = {"value": audio_s_step * steps / (pcm_ms * 1e-3), "unit": UNIT,
                                 "h2d_bytes_per_step": world * B * T * 2, "d2h_bytes_per_step": world * d2h_bytes,
                                 "ms_per_step": pcm_ms / steps,
                                 "note": "same pipelined loop with int16 PCM host buffers (LEAFK_INPUT_S16, s/32768 in the kernel)"}
            line["e2e_pcm16_bf16"] = {"value": audio_s_step * steps / (compact_ms * 1e-3), "unit": UNIT,
                                      "h2d_bytes_per_step": world * B * T * 2, "d2h_bytes_per_step": world * d2h_bytes // 2,
                                      "ms_per_step": compact_ms / steps,
                                      "note": "int16 PCM in, bf16 features out (LEAFK_OUTPUT_BF16, written by the PCEN kernel): half "
                                              "the bytes of the float32 interface in both directions; NOT the headline e2e (the "
                                              "reference interface is float32) -- what a serving loop fed from audio files gets"}
        if world == 1:
            del xs, xs_host
            torch.cuda.empty_cache()
            if not args.no_torch_baseline:
                line["gpu_torch_baseline"] = gpu_torch_baseline(args.config, dev)
            if not args.no_cpu_baseline:
                line["cpu_baseline"] = cpu_port_throughput(args.config)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
