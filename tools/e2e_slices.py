#!/usr/bin/env python3
"""Time EStepEngine.sweep_from_host at the cfg4 shape for several slice schedules (n_slices, growth) next to the
device-resident sweep.  usage: python tools/e2e_slices.py [beats] [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hdpgpc_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
wl = synthetic.make_workload(B, T=256, L=2, M=64, seed=1234, device="cuda")
Y_host = wl["Y"].cpu().pin_memory()
eng = synthetic.build_engine(wl)


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


res = {"beats": B, "resident_ms": timed(eng.sweep)}
for n, g in [(8, 1.0), (4, 1.0), (2, 1.0), (1, 1.0), (3, 2.0), (4, 2.0), (5, 2.0), (3, 3.0), (4, 3.0), (5, 3.0), (3, 4.0), (4, 4.0),
             (6, 2.0), (3, 6.0)]:
    res[f"n{n}_g{g:g}_ms"] = timed(lambda: eng.sweep_from_host(Y_host, n_slices=n, growth=g))
print(json.dumps(res))
