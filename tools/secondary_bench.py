#!/usr/bin/env python3
"""Secondary measurements of SURVEY.md section 8d (reported beside the headline E-step figure, never instead of it):
alignment fits/s (a14), hyper-fits (a10), per-state-covariance scoring R2 (a1/a2), inducing-grid scoring R3 (a3),
MNIW log-likelihoods (a13).  Chain replay and q_lat have their own script (tools/chain_bench.py).
CUDA events on torch's current stream (the stream every wrapper launches on), warm-up first, one JSON line out.
usage: python tools/secondary_bench.py [out.json]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hdpgpc_b200 as hb
from hdpgpc_b200 import ops, synthetic

dev = "cuda"
res = {}


def timed(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


rng = np.random.default_rng(0)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

# ---- a14: alignment.  N beats x R representative beats, 50 Adam steps each
for T, N, R in ((90, 20000, 8), (256, 20000, 8)):
    wl = synthetic.make_workload(N, T=T, L=1, M=R, seed=3, device=dev)
    Y = ops.pack_leads(wl["Y"])[0]
    x = cu(np.arange(T, dtype=np.float64))
    Ym = Y[:R].contiguous()
    scale = cu(np.full(N, 1.0 / 128))
    ms = timed(lambda: ops.warp_fit_batched(x, Y, Ym, 0.3, 200.0, 1e-3, 8, 5e-2, 50, grad_scale=scale), reps=3)
    fac, logdet = ops.warp_prior_factor(x, 1.0, 1.0, 0.3 + 1e-6)
    xw, yw, _, _ = ops.warp_fit_batched(x, Y, Ym, 0.3, 200.0, 1e-3, 8, 5e-2, 50, grad_scale=scale)
    ms_prior = timed(lambda: ops.warp_prior_score(fac, logdet, xw.reshape(-1, T)), reps=3)
    res[f"warp_T{T}"] = {"fits": N * R, "fit_ms": ms, "fits_per_s": N * R / (ms * 1e-3), "adam_steps": 50,
                         "prior_score_ms": ms_prior, "prior_scores_per_s": N * R / (ms_prior * 1e-3)}
    del wl, Y, xw, yw

# ---- a10: hyper-fit, 5 candidate beats at once (n_explore_steps = 5), real iteration counts
for T in (90, 256):
    x = np.arange(T, dtype=np.float64)
    Ys = 120.0 * np.exp(-0.5 * ((x[None, :] - rng.uniform(0.3 * T, 0.7 * T, size=(5, 1))) / (0.06 * T)) ** 2) + rng.normal(size=(5, T)) * 3.0
    t0 = time.perf_counter()
    out = ops.hyperfit_batched(cu(x), cu(Ys), (0.5, 60.0)).cpu().numpy()
    dt = time.perf_counter() - t0
    res[f"hyperfit_T{T}"] = {"fits": 5, "wall_s": dt, "iterations": [int(v) for v in out[:, 5]],
                             "ms_per_iteration": 1e3 * dt / max(1.0, float(out[:, 5].max()))}

# ---- R2: per-state covariances (estimation_limit=None): one Cholesky + inverse per state, pair kernel
T, S, M = 256, 4096, 16
Sig = torch.empty((S, T, T), dtype=torch.float64, device=dev)
base = torch.randn((T, T), dtype=torch.float64, device=dev)
base = base @ base.T / T + torch.eye(T, dtype=torch.float64, device=dev)
for s0 in range(0, S, 512):
    e = torch.randn((512, T, 1), dtype=torch.float64, device=dev)
    Sig[s0:s0 + 512] = base[None] + e @ e.transpose(1, 2)
ms_chol = timed(lambda: ops.chol_batched(Sig), reps=2)
Lf, info = ops.chol_batched(Sig)
ms_inv = timed(lambda: ops.tri_inverse_batched(Lf), reps=2)
ms_fused = timed(lambda: ops.cholinv_batched(Sig), reps=2)
W = ops.tri_inverse_batched(Lf)
N = S
Y = torch.randn((N, T), dtype=torch.float64, device=dev)
mu = torch.randn((S, T), dtype=torch.float64, device=dev)
fos = torch.arange(S, dtype=torch.int32, device=dev)
hbm = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6553.3) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6553.3
r2 = {"states": S, "chol_ms": ms_chol, "tri_inverse_ms": ms_inv, "cholinv_fused_ms": ms_fused,
      "cholinv_tflops": S * 2 * T ** 3 / 3 / (ms_fused * 1e-3) / 1e12, "hbm_peak_gbs": hbm}
# (a) round 1's workload: every pair draws a random state;  (b) the members-per-state distribution of a real fit: the
# state index of a cluster advances only at its own members, so the beats between two members share one state
lab = torch.randint(0, M, (N,), device=dev)
onehot = torch.nn.functional.one_hot(lab, M)
first_state = torch.cumsum(torch.cat([torch.zeros(1, dtype=torch.long, device=dev), torch.bincount(lab, minlength=M)[:-1]]), 0)
runs = (torch.cumsum(onehot, 0) - onehot).clamp_min(0) + first_state[None, :]         # time-indexed state of (beat, cluster)
for tag, state_of in (("random_states", torch.randint(0, S, (N, M), dtype=torch.int32, device=dev)),
                      ("time_indexed_states", runs.clamp_max(S - 1).to(torch.int32).contiguous())):
    ms_pairs = timed(lambda: ops.score_pairs(Y, mu, W, state_of, fos), reps=2)
    plan = ops.group_plan(state_of, fos)
    ms_grp = timed(lambda: ops.score_groups(Y, mu, W, state_of, fos, plan), reps=3)
    n_fac = int(torch.unique(state_of).numel())
    bytes_min = n_fac * T * (T + 1) / 2 * 8 + N * T * 8            # every factor used once (lower triangle) + the beats
    r2[tag] = {"pairs": N * M, "factors_used": n_fac, "chunks": plan["n_chunks"], "pair_kernel_ms": ms_pairs,
               "grouped_ms": ms_grp, "speedup": ms_pairs / ms_grp, "pairs_per_s": N * M / (ms_grp * 1e-3),
               "tflops": N * M * (T * T + 3 * T) / (ms_grp * 1e-3) / 1e12,
               "algorithmic_gbs": bytes_min / (ms_grp * 1e-3) / 1e9, "hbm_frac": bytes_min / (ms_grp * 1e-3) / 1e9 / hbm}
res["R2_per_state_cov_T256"] = r2
del Sig, Lf, W

# ---- R3: inducing grid (x_train != x_basis): kernel matrices + Cholesky + projection per (beat, state) item
nb, nx, items = 128, 256, 1024
xb = cu(np.arange(0, nx, 2, dtype=np.float64))
xp = cu(np.arange(nx, dtype=np.float64)[None, :] + rng.uniform(-0.3, 0.3, size=(items, nx)))
A = torch.randn((nb, nb), dtype=torch.float64, device=dev)
Sg = (A @ A.T / nb + torch.eye(nb, dtype=torch.float64, device=dev))[None].contiguous()
mu = torch.randn((4, nb), dtype=torch.float64, device=dev)
mi = torch.randint(0, 4, (items,), dtype=torch.int32, device=dev)
si = torch.zeros(items, dtype=torch.int32, device=dev)
ms = timed(lambda: ops.pred_dist_inducing(xb, xp, mu, mi, Sg, si, (300.0, 8.0, 0.5)), reps=2)
f, cov, info = ops.pred_dist_inducing(xb, xp, mu, mi, Sg, si, (300.0, 8.0, 0.5))
ms2 = timed(lambda: ops.cholinv_batched(cov), reps=2)
flops_item = nx ** 3 / 3 + 4 * nb * nb * nx + 4 * nb * nx * nx
res["R3_inducing_nb128_nx256"] = {"items": items, "pred_dist_ms": ms, "chol_inverse_ms": ms2,
                                  "items_per_s": items / ((ms + ms2) * 1e-3),
                                  "tflops": items * flops_item / ((ms + ms2) * 1e-3) / 1e12,
                                  "frac_of_dgemm_35.4": items * flops_item / ((ms + ms2) * 1e-3) / 1e12 / 35.4,
                                  "bound": "FP64 tensor (per item: Cholesky of K_bb, two triangular solves, three products at D = 256)"}

# ---- a13: MNIW log-likelihood of 128 clusters' parameters
T, J = 256, 256
Sg = torch.randn((J, T, T), dtype=torch.float64, device=dev)
Sg = Sg @ Sg.transpose(1, 2) / T + torch.eye(T, dtype=torch.float64, device=dev)
Mm = torch.randn((J, T, T), dtype=torch.float64, device=dev) * 0.1
eye = torch.eye(T, dtype=torch.float64, device=dev)[None].contiguous()
ar = torch.arange(J, dtype=torch.int32, device=dev)
z = torch.zeros(J, dtype=torch.int32, device=dev)
ms = timed(lambda: ops.mniw_loglik_batched(Mm, ar, Sg, ar, eye, z, eye, z, Sg, ar), reps=2)
res["mniw_T256"] = {"items": J, "ms": ms, "items_per_s": J / (ms * 1e-3),
                    "tflops": J * (T ** 3 / 3 + T ** 3 / 3 + 2 * T ** 3 + T ** 3 + T ** 3) / (ms * 1e-3) / 1e12}

line = json.dumps(res)
print(line)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(line + "\n")
