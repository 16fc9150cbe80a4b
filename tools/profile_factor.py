"""Gram-Schmidt and block Jacobi SVD of a 256 x 1024 sketch, for an ncu capture:
    ncu --set full --import-source on --clock-control none -k regex:'gs_grid|jacobi_block' -c 2 \
        -o gpurun_out/factor_r01 python tools/profile_factor.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rla4mor_b200 import reductor_ops as ops
from rla4mor_b200.rangefinder import sketch_svd

S = torch.randn(256, 1024, dtype=torch.float64, device="cuda") / 32.0
Q, R = ops.gram_schmidt(S)
U, s, W = sketch_svd(S, qr=(Q, R))
torch.cuda.synchronize()
print("ok", float(s[0]), float(s[-1]))
