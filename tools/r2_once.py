#!/usr/bin/env python3
"""One R2 workload (4096 factors, T = 256, 65536 pairs) through hgp_score_groups, for ncu captures.  usage: python tools/r2_once.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hdpgpc_b200 import ops
dev = "cuda"
T, S, M = 256, 4096, 16
A = torch.randn((T, T), dtype=torch.float64, device=dev); base = A @ A.T / T + torch.eye(T, dtype=torch.float64, device=dev)
Sig = torch.empty((S, T, T), dtype=torch.float64, device=dev)
for s0 in range(0, S, 512):
    e = torch.randn((512, T, 1), dtype=torch.float64, device=dev); Sig[s0:s0 + 512] = base[None] + e @ e.transpose(1, 2)
_, W, _ = ops.cholinv_batched(Sig)
N = S
Y = torch.randn((N, T), dtype=torch.float64, device=dev); mu = torch.randn((S, T), dtype=torch.float64, device=dev)
fos = torch.arange(S, dtype=torch.int32, device=dev)
so = torch.randint(0, S, (N, M), dtype=torch.int32, device=dev)
for mp in (16, 32, None):
    plan = ops.group_plan(so, fos, max_pairs=mp)
    for _ in range(3):
        ops.score_groups(Y, mu, W, so, fos, plan)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(5):
        e0.record(); ops.score_groups(Y, mu, W, so, fos, plan); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("max_pairs", mp, "->", plan["max_pairs"], "chunks", plan["n_chunks"], "grouped_ms", best, "pairs/s", N * M / best * 1e3)
