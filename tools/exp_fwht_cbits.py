"""Developer experiment: per-pass times of the full FWHT for cluster sizes of the strided pass."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rla4mor_b200 as rb
from rla4mor_b200.srht import _fwht_device
for m, d in ((64, 24), (8, 27)):
    a = torch.randn(m, 2 ** d, dtype=torch.float64, device="cuda"); out = torch.empty_like(a)
    ref = None
    for cs in (0, 2, 4, 8):
        os.environ["RLA_FWHT_CLUSTER"] = str(cs)
        os.environ.pop("RLA_FWHT_TIMING", None)
        _fwht_device(a, 1.0, out=out); torch.cuda.synchronize()
        if ref is None: ref = out.clone()
        else: assert torch.equal(ref, out), cs
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); _fwht_device(a, 1.0, out=out); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        rw = 2 * a.numel() * 8
        print(f"({m}, 2^{d}) cluster={cs}: best {min(ts):.3f} ms = {rw / min(ts) / 1e6 / 6549.8:.2f} of a single read+write pass at HBM peak", flush=True)
        os.environ["RLA_FWHT_TIMING"] = "1"
        _fwht_device(a, 1.0, out=out); torch.cuda.synchronize()
    del a, out
