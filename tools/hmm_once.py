#!/usr/bin/env python3
"""One HMM smoothing at the cfg5 shape (2 M beats, K = 128) for ncu captures / timing.  usage: python tools/hmm_once.py [N] [K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hdpgpc_b200 as hb
from hdpgpc_b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 128
rng = np.random.default_rng(0)
tt = rng.gamma(1.0, 1.0, size=(K + 1, K + 1)) + np.eye(K + 1) * 20.0
st = rng.gamma(1.0, 1.0, size=K + 1)
startPi, _ = hb.hdp.expected_log_pi(tt, st, K)
pi, PiT, Pi, Pc = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in hb.hdp.hmm_operands(tt, startPi, K)]
lab = torch.randint(0, K, (N,), device="cuda")
q = torch.randn((N, K), dtype=torch.float64, device="cuda") - 50.0
q[torch.arange(N, device="cuda"), lab] += 30.0
_, e, _, _ = ops.lead_weights(q.reshape(1, N, K), None, torch.ones((N, 1), dtype=torch.float64, device="cuda"))
for _ in range(2):
    hm = ops.hmm_smooth(e, pi, PiT, Pi, Pc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); hm = ops.hmm_smooth(e, pi, PiT, Pi, Pc); e1.record(); torch.cuda.synchronize()
print("hmm_smooth_ms", e0.elapsed_time(e1), "rounds", hm.rounds, "acc", float((hm.z == lab).double().mean()))
