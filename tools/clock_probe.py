"""SM clock and board power while the forward runs back to back for a few seconds (is K1 power-limited?)."""
import os, subprocess, sys, time, threading, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leaf_pytorch_b200 as L

secs = float(os.environ.get("SECS", 3))
g = torch.Generator().manual_seed(1)
x = (torch.randn(256, 1, 16000, generator=g).clamp_(-4, 4) / 4).cuda()
fe = L.Leaf(algo=os.environ.get("ALGO", "tc")).cuda()
rows = []
proc = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,temperature.gpu",
                         "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [rows.append(l.strip()) for l in proc.stdout], daemon=True).start()
with torch.no_grad():
    for _ in range(20):
        fe(x)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.perf_counter() - t0 < secs:
        for _ in range(50):
            fe(x)
        n += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
time.sleep(0.1); proc.terminate()
busy = rows[len(rows) // 3:-2]
clk = [float(r.split(",")[0]) for r in busy]; pw = [float(r.split(",")[1]) for r in busy]
cap = sum("Active" in r and "Not" not in r.split(",")[2] for r in busy)
print(f"LEAFK_DEBUG={os.environ.get('LEAFK_DEBUG','0')}: {ms:.4f} ms/forward sustained over {secs}s; SM clock median {statistics.median(clk):.0f} MHz "
      f"(min {min(clk):.0f}, max {max(clk):.0f}), power median {statistics.median(pw):.0f} W (max {max(pw):.0f}), sw_power_cap active in {cap}/{len(busy)} samples")
