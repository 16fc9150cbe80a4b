"""Developer probe: SM clock / power / throttle reasons under (a) a pure HBM read stream,
(b) back-to-back SRHT launches on the full C3 block.  Prints per-phase medians."""
import os, sys, subprocess, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rla4mor_b200 as rb

Q = "clocks.sm,clocks.mem,power.draw,power.draw.instant,power.limit,enforced.power.limit,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu,temperature.memory"
lines = []
p = subprocess.Popen(["nvidia-smi", f"--query-gpu={Q}", "--format=csv,noheader,nounits", "-lms", "25"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append((time.time(), l.strip())) for l in p.stdout], daemon=True).start()

def phase(name, fn, secs):
    torch.cuda.synchronize(); time.sleep(1.5)
    t0 = time.time(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        fn(); n += 1
        if n % 4 == 0: torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize(); t1 = time.time()
    ms = e0.elapsed_time(e1) / n
    rows = [l.split(", ") for (t, l) in lines if t0 + 0.5 < t < t1]
    def med(i):
        v = sorted(float(r[i]) for r in rows if r[i].replace(".", "").isdigit())
        return v[len(v) // 2] if v else -1
    print(f"{name}: {ms:.2f} ms/iter over {n} iters; sm_mhz med {med(0):.0f}; power.draw med {med(2):.0f} inst {med(3):.0f} limit {med(4):.0f}/{med(5):.0f}; "
          f"sw_power_cap {sum(r[6] == 'Active' for r in rows)}/{len(rows)} hw_slow {sum(r[7] == 'Active' for r in rows)} sw_therm {sum(r[8] == 'Active' for r in rows)} temp {med(9):.0f} mem {med(10):.0f}", flush=True)
    return ms

m, n, k = 1024, 2 ** 24, 4000
x = torch.empty(m, n, dtype=torch.float64, device="cuda")
for lo in range(0, m, 32): x[lo:lo + 32].normal_()
plan = rb.get_plan(n, k, 0, torch.float64, x.device)
y = plan.apply(x)
nb = x.numel() * 8
ms = phase("read stream (torch.sum over 137 GB)", lambda: x.sum(), 4.0)
print(f"   -> {nb / ms / 1e6:.0f} GB/s")
half = x[:512]
dst = torch.empty(256, n, dtype=torch.float64, device="cuda") if torch.cuda.mem_get_info()[0] > 40e9 else None
if dst is not None:
    ms = phase("copy 34 GB -> 34 GB", lambda: dst.copy_(x[:256]), 4.0)
    print(f"   -> {2 * dst.numel() * 8 / ms / 1e6:.0f} GB/s read+write")
    del dst
for opt in (sys.argv[1:] or ["1"]):
    os.environ["RLA_SRHT_OPT"] = opt
    ms = phase(f"srht opt={opt}", lambda: plan.apply(x, out=y), 4.0)
    print(f"   -> {nb / ms / 1e6:.0f} GB/s = {nb / ms / 1e6 / 6549.8:.3f} of 6549.8")
p.terminate()
