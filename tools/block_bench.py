#!/usr/bin/env python3
"""Time the block path (hgp_score_blocks, beats longer than 256 samples) on one lead plane.
usage: python tools/block_bench.py [beats] [T] [clusters] [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hdpgpc_b200 import ops, synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 400
M = int(sys.argv[3]) if len(sys.argv) > 3 else 64
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
wl = synthetic.make_workload(B, T=T, L=1, M=M, seed=1234, device="cuda")
eng = synthetic.build_engine(wl)
tb = eng.leads[0]
assert tb.use_tiles and tb.block_path
best = 1e9
for _ in range(reps + 1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.score_blocks(tb.Y, tb.nu, tb.W, tb.state_of, tb.factor_of_cluster, out=eng.q[0])
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
out = eng.sweep()
acc = float((out["z"].cpu() == torch.from_numpy(wl["labels"])).double().mean())
print(json.dumps({"beats": B, "T": T, "clusters": M, "blocks_ms": best, "beats_per_s_plane": B / (best * 1e-3),
                  "tflops_triangular": B * M * (T * T + 3 * T) / (best * 1e-3) / 1e12, "label_accuracy": acc}))
