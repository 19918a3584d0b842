#!/usr/bin/env python
"""bench.py -- headline benchmark of the LEAF frontend hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of Leaf.forward over one batch of synthetic waveforms:
BASELINE.json configs[1] = default Leaf (40 filters, 401 taps, hop 160, 16 kHz), batch 256 x 1 s,
forward only, per GPU (weak scaling: every rank runs its own 256 clips, no collective on the path).

Printed JSON (rank 0, one line):
  value      audio-seconds per second, whole job, inputs resident in HBM, CUDA-event timed
  e2e        same metric from pinned HOST buffers through HostPipeline (2 batches in flight): every step's
             input goes H2D and its result D2H inside the timed region; e2e_sync = one blocking call per batch
  roofline   K1 (Gabor GEMM + pooling) alone: algorithmic FLOPs / its CUDA-event duration vs the
             measured dense bf16/fp16 tensor peak in MEASURED_PEAKS.json (the path is tensor-bound:
             ~12.9 kFLOP per HBM byte), plus the HBM view the metric asks for
  cpu_baseline  the oracle port (same ATen CPU ops as the reference) on the host cores, bounded sample
--impl reference times that CPU port alone (the reference is pure Python/torch; /root/reference does
not exist on the GPU box, so its own file cannot be imported there -- see DESIGN.md).
"""
from __future__ import annotations

import argparse
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
CFG = dict(F=40, K=401, H=160, B=256, T=16000)        # BASELINE.json configs[1]
WORKLOAD = "configs[1]: default Leaf (F=40, K=401, hop=160, 16 kHz), batch 256 x 1 s per GPU, forward only"
METRIC = "audio_seconds_per_second"
UNIT = "audio-s/s"
N_ROTATE = 10                                          # 10 x 16.4 MB inputs > 126 MB L2


def synth_batch(B: int, T: int, seed: int) -> torch.Tensor:
    """SURVEY 8d synthetic input: clamp(randn,-4,4)/4, |x| <= 1 like the peak-normalised pipeline."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 1, T, generator=g).clamp_(-4, 4) / 4


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
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for name, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port_throughput(B_sample: int, T: int, runs: int = 3):
    """The oracle port of the reference CPU path, all host threads, bounded sample."""
    from oracle import leaf_oracle as O
    import leaf_pytorch_b200 as L
    torch.set_num_threads(os.cpu_count() or 1)
    fe = L.Leaf(n_filters=CFG["F"])
    prm = O.params_from_state_dict({k: v.detach() for k, v in fe.state_dict().items()})
    x = synth_batch(B_sample, T, 1234)
    with torch.no_grad():
        O.forward_f32(x, prm, CFG["K"], CFG["H"])                 # warm-up
        ts = []
        for _ in range(runs):
            t0 = time.perf_counter()
            O.forward_f32(x, prm, CFG["K"], CFG["H"])
            ts.append(time.perf_counter() - t0)
    t = statistics.median(ts)
    return (B_sample * T / SR) / t, t, torch.get_num_threads()


def run_reference(args, rank: int, world: int, emit):
    """--impl reference: the reference's CPU implementation of the path (oracle port; kind 'port')."""
    if rank != 0:
        return
    B_s = 32
    steps = max(1, args.steps)
    from oracle import leaf_oracle as O
    import leaf_pytorch_b200 as L
    torch.set_num_threads(os.cpu_count() or 1)
    fe = L.Leaf(n_filters=CFG["F"])
    prm = O.params_from_state_dict({k: v.detach() for k, v in fe.state_dict().items()})
    x = synth_batch(B_s, CFG["T"], 1234)
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 2))):
            O.forward_f32(x, prm, CFG["K"], CFG["H"])
        t0 = time.perf_counter()
        n_done = 0
        for _ in range(steps):
            O.forward_f32(x, prm, CFG["K"], CFG["H"])
            n_done += 1
            if time.perf_counter() - t0 > 120:                    # keep the whole arm within minutes
                break
        dt = time.perf_counter() - t0
    val = n_done * B_s * CFG["T"] / SR / dt
    cores = torch.get_num_threads()
    sample = f"{B_s} of the 256 clips per step (CPU throughput is flat in batch), {n_done} steps"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": n_done,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / n_done, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--algo", default="auto", choices=["auto", "tc", "tc_full", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from leaf_pytorch_b200.distributed import bind_to_gpu_numa_node
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None     # before any pinned allocation
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, T, F, K, H = CFG["B"], CFG["T"], CFG["F"], CFG["K"], CFG["H"]
    W = max(3, args.warmup)
    steps = max(1, args.steps)
    fe = L.Leaf(n_filters=F, algo=args.algo).to(dev)
    n_frames = fe.num_frames(T)
    xs_host = [synth_batch(B, T, 1234 + 17 * i + 1000 * rank).pin_memory() for i in range(N_ROTATE)]
    xs = [x.to(dev) for x in xs_host]
    out_host = torch.empty((B, F, n_frames), dtype=torch.float32).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------- device-resident throughput
    with torch.no_grad():
        for i in range(W):
            fe(xs[i % N_ROTATE])
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        LF.launch_count(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(steps):
            fe(xs[(W + i) % N_ROTATE])
        e1.record()
        barrier()
        ms_total = e0.elapsed_time(e1)
        launches = LF.launch_count()

        # --------------------------------------------------------- end to end from host buffers
        # (a) synchronous call per step (latency of one batch): Leaf.forward_host
        for i in range(2):
            fe.forward_host(xs_host[i % N_ROTATE], out_host)
        barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            fe.forward_host(xs_host[(2 + i) % N_ROTATE], out_host)     # returns with the result on the host
        barrier()
        sync_s = time.perf_counter() - t0

        # (b) serving loop with two batches in flight: HostPipeline.submit / result.  Every step still copies
        # its own input H2D from pinned memory and reads its own result D2H; the copies of neighbouring steps
        # overlap the kernels.
        def pipelined(hosts, dtype):
            pipe = L.HostPipeline(fe, B, T, depth=2, n_slices=2, input_dtype=dtype)
            outs = [torch.empty((B, F, n_frames), dtype=torch.float32).pin_memory() for _ in range(2)]
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
        e2e_s = pipelined(xs_host, torch.float32)
        # same with 16-bit PCM host buffers converted in the kernel (extra, not the headline)
        pcm_host = [(xh * 32767.0).round().to(torch.int16).pin_memory() for xh in xs_host[:4]]
        pcm_s = pipelined(pcm_host, torch.int16)
        clocks = sampler.stop() if rank == 0 else None

        # --------------------------------------------------------- per-kernel durations (roofline)
        LF.profile_begin()
        for i in range(steps):
            fe(xs[i % N_ROTATE])
        torch.cuda.synchronize()
        n_prof, ms_k0, ms_k1, ms_k2 = LF.profile_end()
        # effective SM clock of K1: cycles and nanoseconds counted inside the kernel, right after a hot loop
        k1_cyc = k1_ns = 0
        sched = None
        if args.algo != "fp32" and LF.tc_supported(F, K, H):
            for i in range(steps):
                fe(xs[i % N_ROTATE])
            prm_t = [None if q is None else q.detach() for q in fe._param_tuple()]
            k1_cyc, k1_ns = LF.k1_clock_probe(fe.spec, xs[0], *prm_t)
            sched = LF.tc_schedule(fe.spec, xs[0], *prm_t)

    times = torch.tensor([ms_total, e2e_s * 1e3, pcm_s * 1e3, sync_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, pcm_ms, sync_ms = float(times[0]), float(times[1]), float(times[2]), float(times[3])

    if rank == 0:
        audio_s_step = world * B * T / SR
        value = audio_s_step * steps / (ms_total * 1e-3)
        e2e_value = audio_s_step * steps / (e2e_ms * 1e-3)
        peaks = measured_peaks()
        flops_alg = 2.0 * (2 * F) * K * T * B                  # conv term, per launch
        bytes_alg = 4.0 * B * T + 4.0 * B * F * n_frames + 32.0 * F
        k1_s = ms_k1 * 1e-3
        algo_used = "tc" if (args.algo != "fp32" and LF.tc_supported(F, K, H)) else "fp32"
        algo_name = ("tc_full" if args.algo == "tc_full" else "tc") if algo_used == "tc" else "fp32"
        exec_frac = sched["executed_fraction"] if sched else 1.0   # support pruning: share of (channel, k-step) pairs run
        exec_mult = 3.0 * ((K + 15) // 16 * 16) / K * exec_frac if algo_used == "tc" else 1.0
        roofline = {
            "kernel": "k1_tc_kernel<80,3,0,0> on CTA pairs (Gabor Toeplitz GEMM + modulus + pooling partials)" if algo_used == "tc"
            else "k1_fp32_kernel",
            "bound": "tensor" if algo_used == "tc" else "fp32-fma",
            "achieved": flops_alg / k1_s / 1e12, "peak": peaks["tensor"], "unit": "TFLOP/s",
            "frac": flops_alg / k1_s / 1e12 / peaks["tensor"],
            "peak_source": f"{peaks['source']} MEASURED_PEAKS.json bf16_tflops (burst; fp16 runs at the same rate)",
            "executed_tflops": exec_mult * flops_alg / k1_s / 1e12,
            "executed_frac": exec_mult * flops_alg / k1_s / 1e12 / peaks["tensor"],
            "executed_note": "3 fp16 products per fp32 product (hi/lo split), taps padded 401->416, times the share of "
                             "(channel, k-step) pairs the support pruning keeps (active channels per k-step below)",
            "pruning": None if not sched else {"executed_fraction": exec_frac, "active_channels_per_kstep": sched["active"],
                                               "all_three_products_channels_per_kstep": sched["active_all_products"],
                                               "channels_per_group": sched["channels_per_group"]},
            "k1_ms": ms_k1, "k0_ms": ms_k0, "k2_ms": ms_k2, "launches_profiled": n_prof,
            "k1_sm_cycles": k1_cyc, "k1_sm_mhz_effective": (1e3 * k1_cyc / k1_ns) if k1_ns else None,
            "traffic": None,
            "hbm": {"algorithmic_bytes": bytes_alg, "achieved": bytes_alg / k1_s / 1e9, "peak": peaks["hbm"],
                    "unit": "GB/s", "frac": bytes_alg / k1_s / 1e9 / peaks["hbm"],
                    "note": "path is compute-bound by ~1000x (SURVEY 8d): HBM fraction is reported because "
                            "the metric names it, the tensor fraction is the binding one"},
        }
        prof_path = os.path.join(ROOT, "profiles", "r01_k1_traffic.json")
        if os.path.isfile(prof_path):
            try:
                with open(prof_path) as f:
                    roofline["traffic"] = json.load(f).get("dram_bytes_per_launch")
            except (OSError, ValueError):
                pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": W,
            "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "algo": algo_name, "n_filters": F, "taps": K, "hop": H,
                       "batch_per_gpu": B, "samples_per_clip": T, "parallelism": f"batch-sharded x{world}, no collective",
                       "l2": f"{N_ROTATE} rotating input batches ({N_ROTATE * B * T * 4 / 1e6:.0f} MB > 126 MB L2)",
                       "arithmetic": "fp16 hi/lo split operands (3 products), fp32 accumulate; ~2^-21 relative; "
                                     "taps beyond 5.5 sigma of a filter skipped per 16-tap step (< 2.7e-7 of its peak), "
                                     "beyond 3.7 sigma only the main product"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * B * T * 4,
                    "d2h_bytes_per_step": world * B * F * n_frames * 4, "ms_per_step": e2e_ms / steps,
                    "api": "HostPipeline.submit/result -> leafk_forward_host_async: pinned host in/out, 2 batches in flight, "
                           "H2D in 2 slices with ready flags feeding one persistent launch per batch"},
            "e2e_sync": {"value": audio_s_step * steps / (sync_ms * 1e-3), "unit": UNIT, "ms_per_step": sync_ms / steps,
                         "api": "Leaf.forward_host (one synchronous call per batch: H2D slices + flags, kernels, D2H)"},
            "e2e_pcm16": {"value": audio_s_step * steps / (pcm_ms * 1e-3), "unit": UNIT,
                          "h2d_bytes_per_step": world * B * T * 2, "d2h_bytes_per_step": world * B * F * n_frames * 4,
                          "ms_per_step": pcm_ms / steps,
                          "note": "same pipelined loop with int16 PCM host buffers (LEAFK_INPUT_S16, s/32768 in the kernel)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, t, cores = cpu_port_throughput(32, T)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"32 of the {B} clips (1 warm-up + 3 runs, median {t:.2f} s)"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
