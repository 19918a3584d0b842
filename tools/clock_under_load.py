"""Sample SM clock / power with nvidia-smi while a loop of forwards (or training steps) runs (development record)."""
import sys, os, subprocess, threading, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leaf_pytorch_b200 as L

F = int(os.environ.get("F", 80)); B = int(os.environ.get("B", 1024)); T = int(os.environ.get("T", 16000))
TRAIN = int(os.environ.get("TRAIN", 1)); SECS = float(os.environ.get("SECS", 3))
x = (torch.randn(B, 1, T, generator=torch.Generator().manual_seed(1234)).clamp_(-4, 4) / 4).cuda()
fe = L.Leaf(n_filters=F).cuda()
G = torch.randn(B, F, fe.num_frames(T), generator=torch.Generator().manual_seed(1235)).cuda()
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits",
                         "-lms", "50", "-i", "0"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append((time.time(), l.strip())) for l in proc.stdout], daemon=True).start()


def step():
    if TRAIN:
        fe.zero_grad(set_to_none=True)
        fe(x).backward(G)
    else:
        with torch.no_grad():
            fe(x)


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.time(); n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < SECS:
    for _ in range(10):
        step()
    n += 10
    torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
t1 = time.time()
time.sleep(0.2); proc.terminate()
load = [r for ts, r in rows if t0 + 0.3 < ts < t1]
clk = [float(r.split(",")[0]) for r in load]; pw = [float(r.split(",")[1]) for r in load]
cap = sum("Active" in r for r in load)
print(f"F={F} B={B} train={TRAIN}: {e0.elapsed_time(e1)/n:.3f} ms/step over {n} steps; SM clock median {statistics.median(clk):.0f} MHz "
      f"(min {min(clk):.0f}, max {max(clk):.0f}), power median {statistics.median(pw):.0f} W, sw_power_cap in {cap}/{len(load)} samples")
