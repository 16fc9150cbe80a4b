for b in 8 4 2; do RLA_JACOBI_B=$b python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from rla4mor_b200 import reductor_ops as ops
from rla4mor_b200.rangefinder import sketch_svd
S = torch.randn(256, 1024, dtype=torch.float64, device="cuda") / 32
Q, R = ops.gram_schmidt(S)
for name, f in (("preconditioned", lambda: sketch_svd(S, qr=(Q, R))), ("direct", lambda: ops.svd_jacobi(S, want_v=True))):
    f(); torch.cuda.synchronize(); ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(os.environ["RLA_JACOBI_B"], name, "%.2f ms" % min(ts), [int(v) for v in ops.svd_jacobi.last_info.tolist()])
PY
done
