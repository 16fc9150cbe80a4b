"""Developer experiment: sustained (back-to-back) SRHT timing on the full C3 block for one
RLA_SRHT_OPT mask (read by the library at first use), with SM clock / power samples.

    RLA_SRHT_OPT=7 python tools/exp_srht_opt.py [m] [launches]
"""
import os, sys, subprocess, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rla4mor_b200 as rb

m = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = int(sys.argv[2]) if len(sys.argv) > 2 else 24
n, k = 2 ** 24, 4000
x = torch.empty(m, n, dtype=torch.float64, device="cuda")
for lo in range(0, m, 32):
    x[lo:lo + 32].normal_()
plan = rb.get_plan(n, k, 0, torch.float64, x.device)
y = plan.apply(x)
torch.cuda.synchronize()
time.sleep(2.0)                                   # start from an idle (cool, uncapped) GPU
lines = []
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "20"],
                     stdout=subprocess.PIPE, text=True)
th = threading.Thread(target=lambda: [lines.append(l.strip()) for l in p.stdout], daemon=True)
th.start()
time.sleep(0.2)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(L + 1)]
ev[0].record()
for i in range(L):
    plan.apply(x, out=y)
    ev[i + 1].record()
torch.cuda.synchronize()
n_load = len(lines)
time.sleep(0.1)
p.terminate()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(L)]
byts = m * n * 8 + m * k * 8 + n + 4 * k
tail = ts[L // 2:]
mean_tail = sum(tail) / len(tail)
clk = [float(l.split(",")[0]) for l in lines[5:n_load] if "," in l]
pw = [float(l.split(",")[1]) for l in lines[5:n_load] if "," in l]
print(f"opt={os.environ.get('RLA_SRHT_OPT', 'default')} m={m} first={ts[0]:.2f} best={min(ts):.2f} "
      f"sustained(mean of last {len(tail)})={mean_tail:.2f} ms -> {byts / mean_tail / 1e6:.0f} GB/s "
      f"frac={byts / mean_tail / 1e6 / 6549.8:.3f} (best {byts / min(ts) / 1e6 / 6549.8:.3f}) "
      f"sm_mhz median={sorted(clk)[len(clk) // 2] if clk else 0:.0f} min={min(clk) if clk else 0:.0f} "
      f"power max={max(pw) if pw else 0:.0f} W", flush=True)
