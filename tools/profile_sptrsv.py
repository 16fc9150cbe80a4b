"""One LU solve (n = 1e6 FEM, 64 right-hand sides) for an ncu launch list of the sparse
triangular-solve kernels:
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:trsv -c 1200 --csv \
        --log-file gpurun_out/launches_sptrsv.csv python tools/profile_sptrsv.py
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rla4mor_b200 as rb
from rla4mor_b200.factorization import InverseLuOperator

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ex = np.ones(nx)
T = sp.diags([-ex[:-1], 2 * ex, -ex[:-1]], [-1, 0, 1])
A = (sp.kron(sp.eye(nx), T) + sp.kron(T, sp.eye(nx)) + sp.eye(nx * nx)).tocsc()
op = rb.MatrixOperator(A, source_id="S", range_id="S")
inv = InverseLuOperator(op, symetric=True)
V = op.source.from_numpy(torch.randn(64, nx * nx, dtype=torch.float64, device="cuda"))
W = inv.apply(V)
torch.cuda.synchronize()
print("lu solve ok:", float((op.apply(W).data - V.data).norm() / V.data.norm()))
