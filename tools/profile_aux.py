"""One pass over the kernels around the sketch, for an ncu launch list: range finder (sketch,
Gram-Schmidt, T = R^-1, block Jacobi SVD) and the device LU solve in front of a sketch.

    ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv \
        --log-file gpurun_out/launches_aux.csv python tools/profile_aux.py
"""
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rla4mor_b200 as rb
from rla4mor_b200.factorization import InverseLuOperator
from rla4mor_b200.rangefinder import sketched_range_finder

m, n, k = 256, 2 ** 21, 1024
U = torch.randn(m, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    res = sketched_range_finder(U, n, k, 0, "srht")
torch.cuda.synchronize()
print("range finder ok: s[0] / s[-1] =", float(res["s"][0] / res["s"][-1]))
nx = 400
ex = np.ones(nx)
T = sp.diags([-ex[:-1], 2 * ex, -ex[:-1]], [-1, 0, 1])
A = (sp.kron(sp.eye(nx), T) + sp.kron(T, sp.eye(nx)) + sp.eye(nx * nx)).tocsc()
op = rb.MatrixOperator(A, source_id="S", range_id="S")
inv = InverseLuOperator(op, symetric=True)
V = op.source.from_numpy(torch.randn(64, nx * nx, dtype=torch.float64, device="cuda"))
for _ in range(2):
    W = inv.apply(V)
torch.cuda.synchronize()
print("lu solve ok:", float((op.apply(W).data - V.data).norm() / V.data.norm()))
