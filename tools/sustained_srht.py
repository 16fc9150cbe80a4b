"""Time 30 back-to-back SRHT launches individually (sustained vs burst) with clocks."""
import os, sys, subprocess, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rla4mor_b200 as rb
m, n, k = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 2 ** 24, 4000
x = torch.empty(m, n, dtype=torch.float64, device="cuda")
for lo in range(0, m, 32): x[lo:lo + 32].normal_()
plan = rb.get_plan(n, k, 0, torch.float64, x.device)
y = plan.apply(x); torch.cuda.synchronize()
lines = []
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,temperature.gpu", "--format=csv,noheader", "-lms", "20"], stdout=subprocess.PIPE, text=True)
th = threading.Thread(target=lambda: [lines.append(l.strip()) for l in p.stdout], daemon=True); th.start()
time.sleep(0.3)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(31)]
ev[0].record()
for i in range(30):
    plan.apply(x, out=y); ev[i + 1].record()
torch.cuda.synchronize(); time.sleep(0.2); p.terminate()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(30)]
byts = m * n * 8 + m * k * 8
print("ms per launch:", " ".join(f"{t:.1f}" for t in ts))
print("GB/s first 3:", [round(byts / t / 1e6) for t in ts[:3]], "last 3:", [round(byts / t / 1e6) for t in ts[-3:]])
print("clock samples:"); print("\n".join(lines[::3][:40]))
