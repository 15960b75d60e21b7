#!/usr/bin/env python3
"""Where one device-resident E-step sweep at the cfg4 shape spends its time: CUDA events between the stages of
EStepEngine.sweep (tile kernel, SNR kernel, exception list per lead plane; lead weights; HMM; statistics), averaged
over `reps` sweeps, next to the event-timed whole sweep.  usage: python tools/stage_times.py [beats] [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hdpgpc_b200 import ops, synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
wl = synthetic.make_workload(B, T=256, L=2, M=64, seed=1234, device="cuda")
eng = synthetic.build_engine(wl)
marks = []


def mark(name):
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    marks.append((name, ev))


def sweep(stamp):
    m = mark if stamp else (lambda name: None)
    m("start")
    for ld, tb in enumerate(eng.leads):
        ops.score_tiles(tb.Y, tb.nu, tb.Wpacked, tb.state_of, tb.factor_of_cluster, out=eng.q[ld], tile_state=tb.tile_state)
        m(f"tiles{ld}")
        tb.snr(eng.snr[ld])
        m(f"snr{ld}")
        tb.score_exceptions(eng.q[ld])
        m(f"pairs{ld}")
    qbar, e, w, flags = ops.lead_weights(eng.q, eng.snr, eng.lead_w)
    m("lead_weights")
    hm = ops.hmm_smooth(e, eng.pi, eng.PiT, eng.Pi, eng.Pc, workspace=eng._hmm_ws)
    m("hmm")
    eng.statistics(qbar, hm)
    m("stats")


for _ in range(3):
    sweep(False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    eng.sweep()
e1.record()
torch.cuda.synchronize()
res = {"beats": B, "sweep_ms": e0.elapsed_time(e1) / reps}
for _ in range(reps):
    sweep(True)
torch.cuda.synchronize()
acc = {}
for (n0, a), (n1, b) in zip(marks[:-1], marks[1:]):
    if n1 != "start":
        acc[n1] = acc.get(n1, 0.0) + a.elapsed_time(b) / reps
res["stages_ms"] = acc
res["stages_sum_ms"] = sum(acc.values())
print(json.dumps(res))
