for rpc in 8 4 2; do RLA_GS_RPC=$rpc python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from rla4mor_b200 import reductor_ops as ops
import numpy as np
from oracle import reductor_oracle as ro
for (r,k) in ((256,1024),(64,1000),(128,2000)):
    S = torch.randn(r, k, dtype=torch.float64, device="cuda")
    ops.gram_schmidt(S); torch.cuda.synchronize()
    ts=[]
    for _ in range(5):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); Q,R=ops.gram_schmidt(S); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    Qo,Ro=ro.gram_schmidt(S.cpu().numpy()) if r<=128 else (None,None)
    err = float(np.linalg.norm(R.cpu().numpy()-Ro)/np.linalg.norm(Ro)) if Ro is not None else -1
    print(os.environ["RLA_GS_RPC"], (r,k), "%.3f ms"%min(ts), "orth %.1e"%float((Q@Q.T-torch.eye(r,device="cuda",dtype=torch.float64)).norm()), "R err %.1e"%err)
PY
done
