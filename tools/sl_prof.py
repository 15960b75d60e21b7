#!/usr/bin/env python3
"""Launch the shared-memory test-hook kernel in its steady-state modes (64 products / 64 Cholesky + inverse per launch)
for an ncu capture: python tools/sl_prof.py [T]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hdpgpc_b200 import ops
T = int(sys.argv[1]) if len(sys.argv) > 1 else 90
rng = np.random.default_rng(0)
A = torch.from_numpy(rng.standard_normal((T, T))).cuda()
S = A @ A.T + T * torch.eye(T, device="cuda", dtype=torch.float64)
for rep in range(2):
    for op in (20, 22):
        a = S.clone() if op == 22 else A.clone()
        b = A.clone(); c = torch.zeros((3, T, T), dtype=torch.float64, device="cuda")
        ops.la_op(op, a, b, c, T=T)
torch.cuda.synchronize()
print("done")
