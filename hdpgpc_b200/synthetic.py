"""Seeded synthetic E-step workloads of the shape named in BASELINE.json (regime
R1 "shared Sigma per (cluster, lead), state-indexed means"): MIT-BIH-like beats from drifting
Gaussian-bump templates, sticky-Markov labels, one full SPD observation covariance per (cluster,
lead), one emission mean per cluster state, and the reference's rule for which state scores which
beat (GPI_model.py:497-513).  Pure torch so the same generator feeds the GPU path (device="cuda")
and the CPU oracle / CPU baseline (device="cpu").  Synthetic-data plumbing, not product arithmetic.
"""
import math

import numpy as np
import torch

F64 = torch.float64


def _rbf(T, length, device):
    x = torch.arange(T, dtype=F64, device=device)
    return torch.exp(-0.5 * (x[:, None] - x[None, :]) ** 2 / length ** 2)


def make_workload(N, T=256, L=2, M=64, seed=1234, device="cpu", n_offset=0, N_total=None, jitter_first=True):
    """Returns a dict with beats Y[N,T,L], per-lead tables (mu, Sigma, add_diag, factor_of_state,
    state_of, snr_state_of), HDP pseudo-counts and the true labels.

    n_offset / N_total let a rank generate its contiguous slice [n_offset, n_offset+N) of a longer
    sequence (labels and cluster tables are generated for the whole sequence from the seed so that
    every rank sees the same clusters; beat noise is drawn per slice)."""
    N_total = N if N_total is None else N_total
    rng = np.random.default_rng(seed)
    dev = torch.device(device)
    x = np.arange(T, dtype=np.float64)
    # templates: three Gaussian bumps per cluster, MIT-BIH-like amplitudes
    a = rng.uniform(-300, 300, size=(M, 3))
    c = rng.uniform(0.15 * T, 0.85 * T, size=(M, 3))
    w = rng.uniform(4, 20, size=(M, 3)) * (T / 256.0)
    templ = np.einsum("kj,kjt->kt", a, np.exp(-0.5 * ((x[None, None, :] - c[:, :, None]) / w[:, :, None]) ** 2))
    # sticky Markov labels over the whole sequence
    u = rng.uniform(size=N_total)
    jump = rng.integers(0, M, size=N_total)
    labels = np.empty(N_total, dtype=np.int64)
    labels[0] = jump[0]
    stay = u < 0.9
    for t in range(1, N_total):   # O(N) scalar loop on the host; 2M steps take about a second
        labels[t] = labels[t - 1] if stay[t] else jump[t]
    amp = np.array([1.0, 0.6, 0.8, 0.5, 0.7, 0.9, 0.4, 0.3])[:L]
    # observation covariances: MNIW-like scale recursion (GPI_model.py:1336)
    nu, n_e = 5.0, 48
    Lk = np.linalg.cholesky(0.49 * (np.eye(T) + np.exp(-0.5 * (x[:, None] - x[None, :]) ** 2 / (8.0 * T / 256.0) ** 2))
                            + 1e-9 * np.eye(T))
    Sigma = np.empty((L, M, T, T))
    for ld in range(L):
        for m in range(M):
            E = Lk @ rng.standard_normal((T, n_e))
            Sigma[ld, m] = (nu * 0.5 * np.eye(T) + E @ E.T) / (nu + n_e)
    gen = torch.Generator(device=dev)
    gen.manual_seed(seed * 7919 + n_offset)
    lab_t = torch.from_numpy(labels).to(dev)
    Ld = torch.linalg.cholesky(_rbf(T, 10.0 * T / 256.0, dev) + 1e-8 * torch.eye(T, dtype=F64, device=dev))
    templ_t = torch.from_numpy(templ).to(dev)

    # latent drifting templates f[n] for the WHOLE sequence would need N_total*T*L doubles; instead each
    # slice integrates the drift of its own members from a per-(cluster, slice) offset that depends only
    # on the seed and the member count before the slice (deterministic function of labels).
    sl = slice(n_offset, n_offset + N)
    lab_s = lab_t[sl]
    order = torch.argsort(lab_s, stable=True)                  # members grouped by cluster, time-ordered
    counts_before = torch.zeros(M, dtype=torch.int64, device=dev)
    if n_offset > 0:
        counts_before = torch.bincount(lab_t[:n_offset], minlength=M)
    Y = torch.empty((N, T, L), dtype=F64, device=dev)
    out_leads = []
    seg_counts = torch.bincount(lab_s, minlength=M)
    seg_start = torch.cumsum(seg_counts, 0) - seg_counts
    pos_sorted = torch.arange(N, device=dev) - seg_start[lab_s[order]]      # member position within the slice
    pos = torch.empty(N, dtype=torch.int64, device=dev)
    pos[order] = pos_sorted
    # state ids: cluster m owns rows [soff[m], soff[m] + 1 + count_m): row 0 = prior mean (zeros)
    soff = torch.cumsum(seg_counts + 1, 0) - (seg_counts + 1)
    S = int((seg_counts + 1).sum())
    n_idx = torch.arange(N, device=dev)
    # A.2: member -> own state (pos+1); non-member -> state of the closest earlier member (>= 1)
    state_of = torch.empty((N, M), dtype=torch.int32, device=dev)
    snr_state_of = torch.empty((N, M), dtype=torch.int32, device=dev)
    for m in range(M):
        idx_m = torch.nonzero(lab_s == m).flatten()
        nm = idx_m.numel()
        if nm == 0:
            state_of[:, m] = -1
            snr_state_of[:, m] = int(soff[m])
            continue
        p = torch.searchsorted(idx_m, n_idx, right=True)       # members at or before n
        is_mem = lab_s == m
        i_vals = torch.where(is_mem, p, torch.clamp(p - 1, min=0).clamp(min=1))
        state_of[:, m] = (soff[m] + i_vals).to(torch.int32)
        j = torch.where(p > 0, p - 1, torch.zeros_like(p))
        j = torch.minimum(torch.clamp(j, min=1), torch.full_like(j, nm))
        snr_state_of[:, m] = (soff[m] + j).to(torch.int32)
    first_mask = (pos == 0)
    for ld in range(L):
        inc = torch.randn((N, T), dtype=F64, device=dev, generator=gen) @ Ld.T       # drift increments
        drift = torch.zeros((N, T), dtype=F64, device=dev)
        srt = inc[order]
        cs = torch.cumsum(srt, 0)
        # cumulative sum restarted per cluster segment
        prev_tot = torch.zeros((M, T), dtype=F64, device=dev)
        has = seg_counts > 0
        st_idx = seg_start[has]
        prev_tot[has] = torch.where((st_idx > 0)[:, None], cs[(st_idx - 1).clamp(min=0)], torch.zeros_like(cs[st_idx]))
        drift_sorted = cs - prev_tot[lab_s[order]]
        # deterministic per-cluster offset standing in for the drift accumulated before the slice
        off = torch.sqrt(counts_before.to(F64))[:, None] * torch.from_numpy(
            np.random.default_rng(seed + 17 * ld).standard_normal((M, T))).to(dev) @ Ld.T
        drift[order] = drift_sorted + off[lab_s[order]]
        f = (templ_t[lab_s] + drift) * amp[ld]
        Y[:, :, ld] = f + math.sqrt(0.5) * torch.randn((N, T), dtype=F64, device=dev, generator=gen)
        mu = torch.zeros((S, T), dtype=F64, device=dev)
        mu[(soff[lab_s] + pos + 1)] = f
        Sig = torch.from_numpy(Sigma[ld]).to(dev)
        fos = torch.repeat_interleave(torch.arange(M, device=dev), seg_counts + 1).to(torch.int32)
        add = torch.zeros(M, dtype=F64, device=dev)
        st_of = state_of
        if jitter_first:
            # `first` rule (GPI_model.py:527-529): the first member of each cluster is scored under
            # Sigma + 1e-2 mean(diag Sigma_0) I -> a duplicate state row whose factor carries the jitter.
            # Every rank holds the same 2M factors (M shared + M jittered) so they can be broadcast as one table.
            Sig = torch.cat([Sig, Sig], dim=0)
            add = torch.cat([add, 1e-2 * 0.5 * torch.ones(M, dtype=F64, device=dev)])
            mem_first = torch.nonzero(first_mask & (counts_before[lab_s] == 0)).flatten()
            kf = mem_first.numel()
            if kf:
                cl = lab_s[mem_first]
                mu = torch.cat([mu, f[mem_first]], dim=0)
                fos = torch.cat([fos, (M + cl).to(torch.int32)])
                st_of = state_of.clone()
                st_of[mem_first, cl] = (S + torch.arange(kf, device=dev)).to(torch.int32)
        out_leads.append(dict(mu=mu, Sigma=Sig, add_diag=add, factor_of_state=fos, state_of=st_of,
                              mu_sm=mu[:S] if jitter_first else mu, snr_state_of=snr_state_of))
    # HDP pseudo-counts from the true label sequence, (M+1) x (M+1) like _calcThetaFull (GPI_HDP.py:400-422)
    tc = np.zeros((M + 1, M + 1))
    np.add.at(tc, (labels[:-1], labels[1:]), 1.0)
    transTheta = tc + 1.0 / (M + 1)
    startTheta = np.full(M + 1, 0.1 / (M + 1))
    startTheta[labels[0]] += 1.0
    return dict(Y=Y, leads=out_leads, transTheta=transTheta, startTheta=startTheta, labels=labels[sl.start:sl.stop],
                N=N, T=T, L=L, M=M)


def build_engine(wl, group=None, tile_path=None, sharded=None):
    """Upload-side of the sweep: factorise the covariances with the library and assemble LeadTables."""
    from . import ops
    from .hdp import EStepEngine, LeadTables
    from .model import LinAlgError
    Yp = ops.pack_leads(wl["Y"])
    leads = []
    for ld, tb in enumerate(wl["leads"]):
        _, W, info = ops.cholinv_batched(tb["Sigma"], add_diag=tb["add_diag"])     # as EStepEngine.update_states does
        if int(torch.count_nonzero(info)):
            raise LinAlgError("synthetic covariance not SPD")
        leads.append(LeadTables(Yp[ld], tb["mu"], W, tb["state_of"], tb["factor_of_state"], tb["mu_sm"],
                                tb["snr_state_of"], tile_path=tile_path))
    return EStepEngine(leads, wl["transTheta"], wl["startTheta"], group=group, sharded=sharded)
