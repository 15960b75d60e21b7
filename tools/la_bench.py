#!/usr/bin/env python3
"""Time the CTA-level routines (one CTA) at a given T through the hgp_la_op test hook."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hdpgpc_b200 import ops
T = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rng = np.random.default_rng(0)
A = torch.from_numpy(rng.standard_normal((T, T))).cuda()
S = A @ A.T + T * torch.eye(T, device="cuda", dtype=torch.float64)
names = {0: "gemm_nn", 1: "gemm_tn", 2: "gemm_nt", 4: "chol", 5: "trsm_lower", 6: "trsm_lower_trans", 7: "lu_factor+solve", 8: "symmetrize", 9: "transpose"}
if T <= 92:      # shared-memory routines (hgp_smem_la.cuh); "empty" = launch + event overhead of this harness
    names.update({10: "sl_gemm_nn", 11: "sl_gemm_tn", 12: "sl_gemm_nt", 14: "sl_cholinv", 15: "sl_spd_solve(cholinv+2 products)", 99: "empty",
                  20: "sl_gemm x64 (operands cached)", 21: "sl_gemm x64 (one operand miss each)", 22: "sl_cholinv x64",
                  23: "gemv x64", 24: "axpby x64"})
res = {"T": T}
for op, nm in names.items():
    best = 1e9
    for rep in range(3):
        a = (S.clone() if op in (4, 5, 6, 7, 14, 22) else A.clone()); b = A.clone(); c = torch.zeros_like(A)
        if op in (15, 21): c = torch.zeros((3, T, T), dtype=torch.float64, device="cuda")
        if op == 15: b = S.clone()
        if op in (5, 6): a = torch.linalg.cholesky(S)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.la_op(op, a, b, c, T=T); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[nm + "_ms"] = round(best, 4)
print(json.dumps(res))
