#!/usr/bin/env python3
"""Secondary measurements (reported beside the headline): chain replay in member-steps/s and
q_lat in members/s at T=256, next to the CPU oracle on the same chain.
usage: [HGP_CHAIN_PIPELINE=0|1|4] python tools/chain_bench.py [n_chains] [members_per_chain] [T] [nocpu]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hdpgpc_b200 as hb
from hdpgpc_b200 import synthetic, ops

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n_mem = int(sys.argv[2]) if len(sys.argv) > 2 else 24
T = int(sys.argv[3]) if len(sys.argv) > 3 else 256
M = n_chains
N = n_chains * n_mem
wl = synthetic.make_workload(N, T=T, L=1, M=M, seed=11, device="cuda")
Y = ops.pack_leads(wl["Y"])
lab = torch.from_numpy(wl["labels"]).cuda()
resp = torch.nn.functional.one_hot(lab, M).double()
x = np.arange(T, dtype=np.float64)
kern = (300.0, 1.2, 1e-3)
def fresh():
    return [[hb.GPI_model.fresh(x, kern, 0.5, 0.5, free_deg=5) for _ in range(M)]]
models = fresh()
hb.full_pass_weighted_batch(models, Y, resp)          # warm-up (also pages the library in)
torch.cuda.synchronize()
models = fresh()
descs = []
for m in range(M):
    d = models[0][m]._chain_prepare(Y[0], resp[:, m])
    if d is not None: descs.append(d)
steps = sum(d["n_members"] for d in descs)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); ops.chain_run(descs, T); e1.record(); torch.cuda.synchronize()
ms_chain = e0.elapsed_time(e1)
for d, m in zip(descs, [m for m in range(M) if models[0][m] is not None]):
    pass
res = {"T": T, "chains": len(descs), "member_steps": steps, "chain_ms": ms_chain,
       "member_steps_per_s": steps / (ms_chain * 1e-3), "longest_chain": max(d["n_members"] for d in descs),
       "flops_per_step_36T3": 36.0 * T ** 3, "achieved_tflops": steps * 36.0 * T ** 3 / (ms_chain * 1e-3) / 1e12}
# q_lat on one finished chain set
gi = 0
for m in range(M):
    d = descs[gi] if gi < len(descs) else None
gp_list = []
mi = 0
for m in range(M):
    dd = models[0][m]._chain_prepare  # noqa
for d, m in zip(descs, [mm for mm in range(M) if int((lab == mm).sum()) > 0]):
    models[0][m]._chain_finish(d); gp_list.append(models[0][m])
torch.cuda.synchronize()
e0.record()
for gp in gp_list: gp.compute_q_lat_all(Y[0])
e1.record(); torch.cuda.synchronize()
res["qlat_ms"] = e0.elapsed_time(e1); res["qlat_members_per_s"] = steps / (res["qlat_ms"] * 1e-3)
res["pipeline"] = os.environ.get("HGP_CHAIN_PIPELINE", "auto")
if len(sys.argv) > 4 and sys.argv[4] == "nocpu":
    print(json.dumps(res))
    sys.exit(0)
# CPU oracle on one chain of the same data
from oracle import hdpgpc_oracle as O
torch.set_num_threads(os.cpu_count())
m0 = int(np.argmax(np.bincount(wl["labels"], minlength=M)))
og = O.OracleGP(x, kern, 0.5, 0.5, free_deg=5)
Yc = wl["Y"][:, :, 0].cpu().numpy()
t0 = time.perf_counter()
og.full_pass_weighted(Yc, (wl["labels"] == m0).astype(float), fitted_kernel=kern)
dt = time.perf_counter() - t0
res["cpu_member_steps_per_s"] = og.N / dt; res["cpu_cores"] = os.cpu_count(); res["cpu_chain_members"] = og.N
print(json.dumps(res))
