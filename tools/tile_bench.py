#!/usr/bin/env python3
"""Time the tile kernel alone at the cfg4 shape (one lead plane) for the library named by HGP_LIB.
usage: HGP_LIB=path python tools/tile_bench.py [beats] [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hdpgpc_b200 import ops, synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
T, L, M = 256, 1, 64
wl = synthetic.make_workload(B, T=T, L=L, M=M, seed=1234, device="cuda")
eng = synthetic.build_engine(wl)
tb = eng.leads[0]
res = {"lib": os.environ.get("HGP_LIB", "default"), "beats": B}
for fused in (False, True):
    kw = dict(mu_sm=tb.mu_sm, snr_state_of=tb.snr_state_of, snr_out=eng.snr[0]) if fused else {}
    for _ in range(3):
        ops.score_tiles(tb.Y, tb.nu, tb.Wpacked, tb.state_of, tb.factor_of_cluster, out=eng.q[0], tile_state=tb.tile_state, **kw)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.score_tiles(tb.Y, tb.nu, tb.Wpacked, tb.state_of, tb.factor_of_cluster, out=eng.q[0], tile_state=tb.tile_state, **kw)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res["fused_ms" if fused else "plain_ms"] = best
    res["fused_tflops" if fused else "plain_tflops"] = B * M * (T * T + 3 * T) / (best * 1e-3) / 1e12
print(json.dumps(res))
