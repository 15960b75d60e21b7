#!/usr/bin/env python3
"""E-step sweep benchmark (BASELINE.json metric: beats/sec per VI E-step sweep; % of FP64 roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--beats B]

One "step" = one E-step sweep over the synthetic cfg4 workload of BASELINE.json configs[3]
(100k beats x 256 samples x 2 leads, 64 clusters, regime R1 of SURVEY.md section 8d) per GPU:
emission scores q[N,M,L] (tensor-core tile kernel + pair kernel for the `first`-jitter states),
SNR lead statistic, lead weights, HMM forward/backward, arg-max responsibilities and the sufficient
statistics.  N>1 (torchrun, one rank per GPU): beats shard by contiguous time slice (weak scaling,
the same 100k beats per GPU), the shared whitening factors are broadcast from rank 0 each sweep,
HMM boundary messages are all-gathered and the statistics all-reduced over NCCL.

`value` times the sweep with every input resident in HBM (CUDA events, max over ranks); `e2e` times
the public call with the beats in pinned HOST memory (H2D of the beats + D2H of labels/statistics
inside the timed region); `roofline` is the tile kernel's achieved FP64 TFLOP/s (algorithmic FLOPs
of SURVEY section 8d: L*M*(T^2+3T) per beat) against the cuBLAS DGEMM rate measured in this run;
`cpu_baseline` / `--impl reference` time the CPU restatement of the reference loop (oracle port,
the reference itself cannot travel to the GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "beats/sec per VI E-step sweep"
UNIT = "beats/s"
CFG = dict(T=256, L=2, M=64, beats_per_gpu=100_000)
# --config cfg5: BASELINE.json configs[4], 2M beats x 256 samples x 128 clusters, one lead, the TOTAL fixed and sharded
# by contiguous time slice over the ranks (strong scaling); cfg4 (default, the headline) keeps 100k beats per GPU (weak)
CFG5 = dict(T=256, L=1, M=128, beats_total=2_000_000)
SCALING = "weak"


def workload_name(beats, T, L, M):
    if SCALING == "strong":
        return (f"synthetic {CFG5['beats_total']} beats x {T} samples x {L} lead, {M} clusters sharded over the GPUs, "
                f"{beats} beats per GPU (BASELINE.json configs[4], regime R1)")
    return f"synthetic {beats} beats x {T} samples x {L} leads, {M} clusters per GPU (BASELINE.json configs[3], regime R1)"


def f_beat(T, L, M):
    return L * M * (T * T + 3 * T)          # algorithmic FLOPs per beat per sweep


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference loop, timed on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_sweep_beats_per_s(n_beats, T, L, M, steps=1, seed=1234, outputs=None):
    import numpy as np
    import torch
    from hdpgpc_b200 import synthetic
    from oracle import hdpgpc_oracle as O      # bench.py's cpu_baseline / reference arm may execute oracle/
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    wl = synthetic.make_workload(n_beats, T=T, L=L, M=M, seed=seed, device="cpu")
    Ys = [wl["Y"][:, :, ld].numpy() for ld in range(L)]
    tabs = [{k: v.numpy() for k, v in tb.items()} for tb in wl["leads"]]
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        q = np.zeros((n_beats, M, L))
        snr = np.zeros((n_beats, M, L))
        for ld, tb in enumerate(tabs):
            fos = tb["factor_of_state"]
            q[:, :, ld] = O.score_states(Ys[ld], tb["mu"], tb["Sigma"], tb["state_of"], fos, tb["add_diag"][fos])
            snr[:, :, ld] = O.snr_states(Ys[ld], tb["mu_sm"], tb["snr_state_of"])
        r = O.estep_responsibilities(q, snr, wl["transTheta"], wl["startTheta"])
        times.append(time.perf_counter() - t0)
    if outputs is not None:
        outputs.update(q=q, snr=snr, z=r["z"], transStateCount=r["transStateCount"])
    return n_beats / (sum(times) / len(times)), cores, times


def reference_sweep_beats_per_s(n_beats, T, L, M, steps=1, seed=1234, outputs=None):
    """The same sweep by the UNMODIFIED reference (torch CPU), imported through oracle/refshim from /root/reference or from
    the copy tools/make_ref.sh stages under oracle/_ref: per (cluster, lead) `GPI_model.compute_sq_err_all`
    (GPI_model.py:488-547) and `GPI_HDP.compute_snr` (GPI_HDP.py:732-748) on reference model objects that hold the
    workload's cluster states, then the HMM block of `estimate_q_all` (GPI_HDP.py:2856-2862: weight_mean, LogLik, forward,
    backward, coupled_state_coef, _safe_exp) and the count lines (:890-892).  Returns None if the reference is not there."""
    import contextlib
    import io
    import numpy as np
    import torch
    from scipy.special import digamma
    from hdpgpc_b200 import synthetic
    from oracle import refshim                 # bench.py's cpu_baseline / reference arm may execute oracle/
    if refshim.find_reference() is None:
        return None
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    with contextlib.redirect_stdout(io.StringIO()):
        hdp = refshim.install()
    wl = synthetic.make_workload(n_beats, T=T, L=L, M=M, seed=seed, device="cpu")
    labels = wl["labels"]
    x_basis = np.atleast_2d(np.arange(T, dtype=np.float64)).T
    with contextlib.redirect_stdout(io.StringIO()):
        sw = hdp.GPI_HDP(x_basis, M=M, n_outputs=L, x_basis_warp=x_basis[::2], kernels=None, model_type='dynamic',
                         ini_lengthscale=3.0, bound_lengthscale=(1.0, 20.0), ini_gamma=0.5, ini_sigma=0.5,
                         ini_outputscale=300.0, noise_warp=0.05, bound_sigma=(1e-5, 1.0), bound_gamma=(1e-5, 1.0),
                         bound_noise_warp=(0.005, 0.01), verbose=False, hmm_switch=True, max_models=M + 1,
                         bayesian_params=True, inducing_points=False, estimation_limit=1, free_deg_MNIV=5)
    seg = np.bincount(labels, minlength=M)
    soff = np.cumsum(seg + 1) - (seg + 1)
    eye = torch.eye(T)
    for ld in range(L):
        tb = wl["leads"][ld]
        for m in range(M):
            gp = sw.gpmodels[ld][m]
            idx = np.nonzero(labels == m)[0]
            rows = tb["mu"][soff[m]:soff[m] + idx.size + 1]
            gp.f_star = [r.reshape(T, 1).clone() for r in rows]          # state i = mean after the i-th member (0: prior)
            gp.f_star_sm = list(gp.f_star)
            gp.cov_f_sm = [eye] * len(gp.f_star)                         # read (not used) by resample_latent_mean on the basis grid
            gp.cov_f = gp.cov_f_sm
            gp.C = [eye, eye]
            gp.Sigma = [0.5 * eye, tb["Sigma"][m].clone()]               # Sigma[0]: prior (first-member jitter), Sigma[-1]: shared
            gp.indexes = [int(i) for i in idx]
            gp.N = int(idx.size)
            gp.estimation_limit = 1                                      # every state uses C[-1], Sigma[-1] (GPI_model.py:646-651)
    sw.transTheta = torch.from_numpy(np.asarray(wl["transTheta"], dtype=np.float64))
    sw.startTheta = torch.from_numpy(np.asarray(wl["startTheta"], dtype=np.float64))
    x_trains = torch.from_numpy(np.array([x_basis] * n_beats))
    sw.x_train = x_trains
    Y = wl["Y"]
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            q = torch.zeros((n_beats, M, L))
            snr = torch.zeros((n_beats, M, L))
            for ld in range(L):
                for m in range(M):
                    gp = sw.gpmodels[ld][m]
                    q[:, m, ld] = gp.compute_sq_err_all(x_trains, Y[:, :, [ld]])
                    snr[:, m, ld] = sw.compute_snr(Y[:, :, ld], gp)
            tt, st_ = sw.transTheta, sw.startTheta
            dsum = torch.log(torch.sum(torch.exp(digamma(tt[:M, :M + 1])), axis=1) + 1e-5)
            transPi = digamma(tt[:M, :M]) - dsum[:, np.newaxis]
            startPi = digamma(st_[:M]) - torch.log(torch.sum(torch.exp(digamma(st_[:M + 1]))) + 1e-5)
            q_norm, _ = sw.LogLik(sw.weight_mean(q, snr))
            alpha, margprob = sw.forward(startPi, transPi, q_norm)
            beta = sw.backward(transPi, q_norm, margprob)
            logresp, _ = sw.LogLik(torch.log(alpha * beta), axis=1)
            logrespPair, _ = sw.LogLik(sw.coupled_state_coef(alpha, beta, transPi, q_norm, margprob), axis=1)
            resp = sw._safe_exp(logresp)
            respPair = sw._safe_exp(logrespPair)
            startStateCount = resp[0]
            transStateCount = torch.sum(respPair, axis=0)
        times.append(time.perf_counter() - t0)
    if outputs is not None:          # tests/test_bench_contract.py: the arm computes the workload's sweep, nothing else
        outputs.update(q=q.numpy(), snr=snr.numpy(), z=torch.argmax(resp, dim=1).numpy(),
                       transStateCount=transStateCount.numpy(), startStateCount=startStateCount.numpy())
    return n_beats / (sum(times) / len(times)), cores, times


def cpu_arm(n_ref, n_port, T, L, M, steps):
    """(beats/s, cores, per-step seconds, kind, sample description): the unmodified reference when it is reachable,
    otherwise the oracle port of its loop."""
    try:
        got = reference_sweep_beats_per_s(n_ref, T, L, M, steps=steps)
    except Exception as e:           # a broken staging must not lose the line: fall back to the port, say why
        got = None
        sys.stderr.write(f"reference arm: the unmodified reference failed ({type(e).__name__}: {e}); timing the port\n")
    if got is not None:
        bps, cores, times = got
        return bps, cores, times, "reference", (
            f"{n_ref} beats of the same workload per step: the UNMODIFIED reference (torch CPU, oracle/refshim) -- "
            "GPI_model.compute_sq_err_all + GPI_HDP.compute_snr per (cluster, lead), then weight_mean / LogLik / forward / "
            "backward / coupled_state_coef / _safe_exp and the counts")
    bps, cores, times = cpu_sweep_beats_per_s(n_port, T, L, M, steps=steps)
    return bps, cores, times, "port", (
        f"{n_port} beats of the same workload per step (oracle port of the reference loop: one Cholesky + cholesky_solve "
        "per distinct cluster state, python HMM loop); the reference package itself was not reachable")


def cpu_baseline_record(args, T, L, M):
    bps, cores, times, kind, sample = cpu_arm(args.cpu_ref_beats, args.cpu_beats, T, L, M, steps=1)
    return {"value": bps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample + f"; {times[0]:.1f} s"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    T, L, M = CFG["T"], CFG["L"], CFG["M"]
    # a bounded sample per step (~15 s of host work at the default), shrunk so that K steps still end within a few minutes
    shrink = lambda n: max(512, min(n, (8 * n // max(1, args.steps)) // 64 * 64))
    bps, cores, times, kind, sample = cpu_arm(shrink(args.cpu_ref_beats), shrink(args.cpu_beats), T, L, M,
                                              steps=max(1, args.steps))
    line = {
        "impl": "reference", "metric": METRIC, "value": bps, "unit": UNIT, "n_gpus": args.gpus, "steps": max(1, args.steps),
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": SCALING,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(CFG["beats_per_gpu"], T, L, M)},
        "cpu_baseline": {"value": bps, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": bps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and clock-event reasons DURING the timed region -- the fields of the profiling recipe's nvidia-smi
    clocks line, read in-process through NVML (nvidia-ml-py) every 50 ms.  Round 2 found that spawning `nvidia-smi -lms`
    next to the timed region is itself a disturbance: on a fresh box its start-up (NVML init, enumeration of 8 GPUs) can
    take longer than the second it was given and then holds the driver while the sweeps launch -- ms_per_step read 28.8
    to 40.3 ms from run to run with the tile kernel at 13.2 ms every time.  NVML is initialised before the warm-up and a
    query is a few microseconds; nvidia-smi stays as the fallback (started before the warm-up, first row awaited)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, torch, gpu_index):
        self.rows = []            # (t, sm_mhz, sm_max_mhz, power_w, [reasons])
        self.proc = None
        self.nvml = None
        self._stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            p = torch.cuda.get_device_properties(gpu_index)
            bus = f"{getattr(p, 'pci_domain_id', 0):08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
            self.h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self._sample()        # first query (lazy initialisation inside the library) well before the timed region
            self.rows.clear()
        except Exception:
            self.nvml = None
            self.gpu = gpu_index

    def _sample(self):
        n = self.nvml
        mhz = float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
        try:
            r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h))
        except Exception:
            r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        names = [("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap)]
        try:
            pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        except Exception:
            pw = 0.0
        self.rows.append((time.perf_counter(), mhz, self.max_mhz, pw, [nm for nm, bit in names if r & bit]))

    def _loop(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                pass
            self._stop.wait(0.05)

    def _read_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            c = [x.strip() for x in line.split(",")]
            if len(c) >= 9:
                try:
                    self.rows.append((time.perf_counter(), float(c[1]), float(c[2]), float(c[3]),
                                      [nm for nm, v in zip(names, c[5:9]) if v.lower().startswith("active")]))
                except ValueError:
                    pass

    def start(self):
        """Call BEFORE the warm-up sweeps; returns once sampling is under way."""
        if self.nvml is not None:
            self.th = threading.Thread(target=self._loop, daemon=True)
            self.th.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read_smi, daemon=True)
            self.th.start()
            t0 = time.perf_counter()
            while not self.rows and time.perf_counter() - t0 < 20.0:      # start-up over before anything is timed
                time.sleep(0.05)
        except Exception:
            self.proc = None

    def stop(self, t0, t1):
        self._stop.set()
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
        elif self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["neither NVML nor nvidia-smi available"]}
        else:
            self.th.join(timeout=1.0)
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or list(self.rows)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(r[1] for r in rows)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": rows[0][2], "reasons": sorted({n for r in rows for n in r[4]}),
                "samples": len(rows), "power_w_max": max(r[3] for r in rows),
                "source": "NVML in-process, 50 ms" if self.proc is None else "nvidia-smi -lms 200"}


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def measure_fp64_peak(torch, seconds=1.5):
    """cuBLAS DGEMM 8192^3: burst (best of 5) and sustained (back to back for `seconds`)."""
    n = 8192
    a = torch.randn((n, n), dtype=torch.float64, device="cuda")
    b = torch.randn((n, n), dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(3, int(seconds / (2.0 * n ** 3 / (best * 1e12))))
    e0.record()
    for _ in range(reps):
        torch.matmul(a, b, out=c)
    e1.record(); torch.cuda.synchronize()
    sustained = reps * 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
    del a, b, c
    return best, sustained


def stage_breakdown(torch, ops, eng, reps=3):
    """CUDA events between the stages of one device-resident sweep (averaged over `reps`)."""
    marks = []

    def mark(name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        marks.append((name, ev))

    for _ in range(reps):
        mark("start")
        for ld, tb in enumerate(eng.leads):
            ops.score_tiles(tb.Y, tb.nu, tb.Wpacked, tb.state_of, tb.factor_of_cluster, out=eng.q[ld], tile_state=tb.tile_state)
            mark("tiles")
            tb.snr(eng.snr[ld])
            mark("snr")
            tb.score_exceptions(eng.q[ld])
            mark("pairs")
        qbar, e, w, hm = eng.responsibilities()
        mark("lead_weights+hmm")
        eng.statistics(qbar, hm)
        mark("stats")
    torch.cuda.synchronize()
    acc = {}
    for (n0, a), (n1, b) in zip(marks[:-1], marks[1:]):
        if n1 != "start":
            acc[n1] = acc.get(n1, 0.0) + a.elapsed_time(b) / reps
    return {k: round(v, 4) for k, v in acc.items()}


def table_stage_breakdown(torch, ops, eng, raw, reps=3):
    """CUDA events between the kernels of the table build (EStepEngine.update_states: factors and inverse factors of all
    leads in one launch, then packing and whitening per lead)."""
    acc = {}
    Sig = torch.cat([t["Sigma"] for t in raw], dim=0)
    add = torch.cat([t["add_diag"] for t in raw]) if all(t.get("add_diag") is not None for t in raw) else None
    for rep in range(reps + 1):          # rep 0 is a warm-up: it allocates the outputs (cudaMalloc stalls between the events)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        Lf, W, info = ops.cholinv_batched(Sig, add_diag=add)
        ev[1].record()
        off, packed = 0, []
        for tb, t in zip(eng.leads, raw):
            F = t["Sigma"].shape[0]
            packed.append(ops.pack_factors(W[off:off + F]))
            off += F
        ev[2].record()
        off = 0
        for tb, t, Wp in zip(eng.leads, raw, packed):
            F = t["Sigma"].shape[0]
            if tb._whiten_plan is not None:
                ops.whiten_means_tiles(t["mu"], W[off:off + F], Wp, tb.factor_of_state, tb._whiten_plan)
            else:
                ops.whiten_means(t["mu"], W[off:off + F], tb.factor_of_state)
            off += F
        ev[3].record()
        torch.cuda.synchronize()
        if rep == 0:
            continue
        for k, name in enumerate(("chol+inverse", "pack_factors", "whiten_means")):
            acc[name] = acc.get(name, 0.0) + ev[k].elapsed_time(ev[k + 1]) / reps
    return {k: round(v, 4) for k, v in acc.items()}


def measure_config(args, torch, dist, hb, ops, synthetic, cfg, scaling, steps, warmup, rank, world, local, with_clocks,
                   with_e2e=True, h2d_gbs=None):
    """One workload: device-resident sweep (`value`), table build, end-to-end call from pinned host beats."""
    T, L, M, B = cfg["T"], cfg["L"], cfg["M"], cfg["beats_per_gpu"]
    wl = synthetic.make_workload(B, T=T, L=L, M=M, seed=1234, device="cuda", n_offset=rank * B, N_total=world * B)
    raw = wl["leads"]                                         # (mu, Sigma, add_diag, ...) per lead: the cluster states
    Y_host = wl["Y"].cpu().pin_memory() if with_e2e else None   # e2e leg: beats live in pinned host memory
    eng = synthetic.build_engine(wl, sharded=world > 1)       # rank r holds the r-th contiguous time slice
    assert all(tb.use_tiles for tb in eng.leads)

    if world > 1:   # a size mismatch in a broadcast hangs NCCL silently: check once, loudly
        sz = torch.tensor([tb.Wpacked.numel() for tb in eng.leads], dtype=torch.int64, device="cuda")
        allsz = [torch.empty_like(sz) for _ in range(world)]
        dist.all_gather(allsz, sz)
        if any(not torch.equal(a, sz) for a in allsz):
            raise hb.HgpError(f"factor tables differ across ranks: {[a.tolist() for a in allsz]}")

    # shared factors are broadcast from rank 0 each sweep when sharded (cluster parameters broadcast)
    def broadcast_tables():
        if world > 1:
            for tb in eng.leads:
                dist.broadcast(tb.Wpacked, src=0)

    tile_events = []

    def sweep(record=False):
        broadcast_tables()
        for ld, tb in enumerate(eng.leads):
            if record:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            ops.score_tiles(tb.Y, tb.nu, tb.Wpacked, tb.state_of, tb.factor_of_cluster, out=eng.q[ld], tile_state=tb.tile_state)
            if record:
                e1.record()
                tile_events.append((e0, e1))
            tb.snr(eng.snr[ld])
            if tb.pair_n is not None:
                ops.score_pairs(tb.Y, tb.mu, tb.W, tb.state_of, tb.factor_of_state, tb.pair_n, tb.pair_m, out=eng.q[ld])
        qbar, e, w, hm = eng.responsibilities()
        return eng.statistics(qbar, hm), hm

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    sampler = ClockSampler(torch, local) if (with_clocks and rank == 0) else None
    if sampler is not None:
        sampler.start()      # before the warm-up: whatever the sampler costs to start is over when the timed region begins
    for _ in range(max(3, warmup)):
        sweep()
    barrier()
    launches0 = ops.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        st, hm = sweep(record=True)
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = ops.launch_count() - launches0
    ms = ev0.elapsed_time(ev1) / steps
    tile_ms = sum(a.elapsed_time(b) for a, b in tile_events) / max(1, len(tile_events))
    clocks = sampler.stop(t_wall0, t_wall1) if sampler is not None else None
    ms_max = max_over_ranks(ms)
    acc = float((hm.z.cpu() == torch.from_numpy(wl["labels"])).double().mean())

    # ---- table build: new cluster states -> factors, whitened means, packed factors (what the reference re-derives
    #      inside every E-step; it is NOT part of `value`, it IS part of `value_incl_tables` and of `e2e`) ----
    for _ in range(2):
        eng.update_states(raw)
    barrier()
    n_tb = max(1, min(steps, 5))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_tb):
        eng.update_states(raw)
    e1.record()
    barrier()
    eng.check_tables()
    table_ms = max_over_ranks(e0.elapsed_time(e1) / n_tb)

    out = dict(cfg=cfg, B=B, ms=ms, ms_max=ms_max, tile_ms=tile_ms, launches=int(launches), clocks=clocks, acc=acc,
               table_ms=table_ms, eng_rounds=(eng.hmm_rounds, eng.boundary_rounds),
               means_mb=sum(tb.nu.numel() for tb in eng.leads) * 8 / 1e6,
               bytes_launch=B * T * 8 + B * M * 8 + B * M * 4 + eng.leads[0].mu.numel() * 8 + eng.leads[0].Wpacked.numel() * 8)
    if rank == 0:
        out["stages_ms"] = stage_breakdown(torch, ops, eng) if world == 1 else None
        out["table_stages_ms"] = table_stage_breakdown(torch, ops, eng, raw) if world == 1 else None
    barrier()

    # ---- end-to-end leg: cluster states + host beats in, labels + statistics out ----
    if with_e2e:
        # slice schedule from the two rates measured above (this rank's H2D rate with all ranks copying, the sweep time)
        out["slices"] = list(eng.tune_slices(h2d_gbs, ms_max, head_start_ms=table_ms)) if h2d_gbs else None
        def e2e_step():
            # the public end-to-end call: table build from (mu, Sigma), pinned host beats in (sliced H2D copies overlapped
            # with scoring), labels and statistics back on the host
            eng.update_states(raw)
            broadcast_tables()
            st_ = eng.sweep_from_host(Y_host)
            return st_["z_host"], st_["stats_host"]
        for _ in range(2):
            e2e_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n_e2e = max(1, min(steps, 5))
        for _ in range(n_e2e):
            z_host, stats_host = e2e_step()
        e1.record()
        barrier()
        out["e2e_ms"] = max_over_ranks(e0.elapsed_time(e1) / n_e2e)
        out["h2d"] = Y_host.numel() * 8
        out["d2h"] = z_host.numel() * 4 + stats_host.numel() * 8
    del eng, wl, raw, Y_host
    torch.cuda.empty_cache()
    return out


def bind_to_gpu_numa(torch, local):
    """Run this rank (and first-touch its pinned staging memory) on the CPU cores next to its GPU: with all ranks on one
    NUMA node the host-to-device copies of 8 ranks share one memory controller and one root complex (round 1:
    end-to-end weak-scaling efficiency 0.74 at 8 GPUs against 0.97 device-resident).  Returns a short description."""
    try:
        import pynvml
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(local)
        bus = f"{getattr(p, 'pci_domain_id', 0):08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = sorted(64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1)
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return {"gpu_pci": bus, "cpus_bound": len(allowed), "first_cpu": allowed[0], "last_cpu": allowed[-1]}
        return {"gpu_pci": bus, "cpus_bound": 0, "note": "GPU-local cores not in this process's cpuset"}
    except Exception as e:
        return {"error": f"{type(e).__name__}: {e}"[:200]}


def h2d_bandwidth(torch, nbytes=256 << 20):
    """GB/s of one pinned host -> device copy of this rank (CUDA events, best of 3)."""
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    best = 0.0
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    return best


def fit_record():
    """Wall time of the reference's own offline fit of MIT-BIH record 100 (hdpgpc/tests/test_offline.py settings, 2272
    beats) driven through the device path (hdpgpc_b200.integration), next to the CPU time of the same fit by the
    unmodified reference recorded with the fixture (tests/golden/fit_rec100_offline.npz, dev container).  None when the
    reference is not staged (tools/make_ref.sh)."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import numpy as np
        from oracle import refshim                      # checker-side infrastructure: locates / imports the reference
        if refshim.find_reference() is None:
            return None
        from test_reference_fit_gpu import run_on_device
        z = np.load(os.path.join(ROOT, "tests", "golden", "fit_rec100_offline.npz"))
        got, wall, launches, twins = run_on_device("rec100_offline")
        return {"rec100_offline_s": round(wall, 2), "cpu_s": round(float(z["cpu_fit_seconds"]), 1),
                "cpu_where": f"unmodified reference, dev container, {int(z['cpu_threads'])} threads (tests/golden/fit_rec100_offline.npz)",
                "identical_labels": bool(np.array_equal(got["labels"], z["labels"])), "clusters": int(got["M"]),
                "clusters_reference": int(z["M"]), "gpu_launches": int(launches)}
    except Exception as e:           # the fit record is an extra: never lose the bench line over it
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def run_ours(args):
    if os.environ.get("HGP_BENCH_WATCHDOG"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["HGP_BENCH_WATCHDOG"]), exit=True)
    import torch
    import torch.distributed as dist
    import hdpgpc_b200 as hb
    from hdpgpc_b200 import ops, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise hb.HgpError("bench.py needs a B200; hdpgpc_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(torch, local)          # before any pinned allocation (first touch decides the NUMA node)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hb.load_library()
    T, L, M = CFG["T"], CFG["L"], CFG["M"]
    B = args.beats or CFG["beats_per_gpu"]
    cfg = dict(T=T, L=L, M=M, beats_per_gpu=B)

    peak_burst = peak_sust = None
    if rank == 0 and not args.no_peak:
        peak_burst, peak_sust = measure_fp64_peak(torch)

    h2d_gbs = h2d_bandwidth(torch)
    if world > 1:      # all ranks copy at once: the figure that matters for the end-to-end leg
        dist.barrier()
        h2d_gbs = h2d_bandwidth(torch)
        t_ = torch.tensor([h2d_gbs], dtype=torch.float64, device="cuda")
        dist.all_reduce(t_, op=dist.ReduceOp.MIN)
        h2d_gbs = float(t_)
    r = measure_config(args, torch, dist, hb, ops, synthetic, cfg, SCALING, args.steps, args.warmup, rank, world, local, True,
                       h2d_gbs=h2d_gbs)
    ms, ms_max, tile_ms, table_ms, e2e_ms = r["ms"], r["ms_max"], r["tile_ms"], r["table_ms"], r["e2e_ms"]

    # ---- second leg: BASELINE.json configs[4] (2M beats x 256 x 128 clusters, strong scaling), 3 sweeps ----
    cfg5 = None
    if args.config == "cfg4" and not args.no_cfg5 and not args.beats:
        c5 = dict(T=CFG5["T"], L=CFG5["L"], M=CFG5["M"], beats_per_gpu=CFG5["beats_total"] // world)
        r5 = measure_config(args, torch, dist, hb, ops, synthetic, c5, "strong", 3, 3, rank, world, local, False, with_e2e=False)
        if rank == 0:
            fl5 = c5["beats_per_gpu"] * c5["M"] * (c5["T"] ** 2 + 3 * c5["T"])
            peak5 = peak_sust if peak_sust else 37.0
            cfg5 = {"workload": f"synthetic {CFG5['beats_total']} beats x {c5['T']} samples x {c5['L']} lead, {c5['M']} clusters, "
                                f"{c5['beats_per_gpu']} beats per GPU (BASELINE.json configs[4], strong scaling)",
                    "value": CFG5["beats_total"] / (r5["ms_max"] * 1e-3), "unit": UNIT, "ms_per_step": r5["ms_max"], "steps": 3,
                    "frac": (fl5 / (r5["tile_ms"] * 1e-3) / 1e12) / peak5,
                    "sweep_frac": (c5["L"] * fl5 / (r5["ms_max"] * 1e-3) / 1e12) / peak5,
                    "tile_kernel_ms": r5["tile_ms"], "table_build_ms": r5["table_ms"], "stages_ms": r5.get("stages_ms"),
                    "label_accuracy": r5["acc"], "hmm_repair_rounds": r5["eng_rounds"][0], "boundary_rounds": r5["eng_rounds"][1]}

    if rank == 0:
        flops_launch = B * M * (T * T + 3 * T)               # one tile-kernel launch = one lead plane
        achieved = flops_launch / (tile_ms * 1e-3) / 1e12
        peak = peak_sust if peak_sust else 37.0
        mp = {}
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = traffic_source = None   # dram__bytes_read + dram__bytes_write of one launch, from the committed ncu capture
        for tag in ("r02", "r01"):       # the latest committed capture of this kernel at this shape
            try:
                prof = json.load(open(os.path.join(ROOT, "profiles", f"{tag}_score_tiles_ncu.json")))
                if args.config == "cfg4" and B == CFG["beats_per_gpu"]:
                    traffic = prof["dram_bytes_per_launch"]
                    traffic_source = (f"ncu --set full capture of this kernel at this shape, profiles/{tag}_score_tiles_ncu.json "
                                      "(tools/gpu_round.sh; not re-measured in this run)")
                break
            except Exception:
                continue
        bytes_launch = r["bytes_launch"]
        cpu = None
        if not args.no_cpu and world == 1:          # the CPU baseline is reported by the single-GPU run only
            cpu = cpu_baseline_record(args, T, L, M)
        fit = None if (args.no_fit or world > 1) else fit_record()
        line = {
            "metric": METRIC, "value": world * B / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_max, "higher_is_better": True, "scaling": SCALING,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(B, T, L, M), "beats_total": world * B, "l2": f"inputs_exceed_l2 (beats {B * T * L * 8 / 1e6:.0f} MB + whitened means {r['means_mb']:.0f} MB per GPU vs 126 MB L2)",
                       "label_accuracy": r["acc"], "hmm_repair_rounds": r["eng_rounds"][0], "boundary_rounds": r["eng_rounds"][1]},
            "table_build_ms": table_ms, "table_build_share_of_sweep": table_ms / ms_max,
            "value_incl_tables": world * B / ((ms_max + table_ms) * 1e-3),
            "stages_ms": r.get("stages_ms"), "table_stages_ms": r.get("table_stages_ms"),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_source, "kernel": "score_tiles_kernel", "kernel_ms": tile_ms,
                         "kernel_share_of_step": L * tile_ms / ms, "flops_per_launch": flops_launch,
                         "peak_source": "cuBLAS DGEMM 8192^3 float64 measured in this run, sustained (MEASURED_PEAKS.json has no FP64 figure)" if peak_sust else "nominal",
                         "peak_burst": peak_burst,
                         "hbm": {"algorithmic_bytes_per_launch": bytes_launch, "achieved_gbs": bytes_launch / (tile_ms * 1e-3) / 1e9,
                                 "peak_gbs": mp.get("hbm_gbs"), "frac": (bytes_launch / (tile_ms * 1e-3) / 1e9) / mp["hbm_gbs"] if mp.get("hbm_gbs") else None}},
            "cpu_baseline": cpu,
            "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                    "ms_per_step": e2e_ms, "includes": "table build from (mu, Sigma) + H2D of the beats + sweep + D2H of labels / statistics",
                    "slices": r.get("slices")},
            "host": {"numa_binding_rank0": numa, "h2d_gbs_per_rank_min": h2d_gbs,
                     "h2d_note": "pinned host -> device, 256 MiB, all ranks copying at the same time, slowest rank"},
            "cfg5": cfg5,
            "fit": fit,
            "gpu_launches": r["launches"],
            "clocks": r["clocks"],
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--beats", type=int, default=0, help="beats per GPU (default: the cfg4 100000)")
    ap.add_argument("--cpu-beats", type=int, default=8192, help="beats in the bounded CPU sample (oracle port)")
    ap.add_argument("--cpu-ref-beats", type=int, default=1024, help="beats in the bounded CPU sample (unmodified reference)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-peak", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true", help="skip the short configs[4] leg (2M beats x 128 clusters)")
    ap.add_argument("--no-fit", action="store_true", help="skip the record-100 fit through the device path")
    ap.add_argument("--config", default="cfg4", choices=["cfg4", "cfg5"])
    args = ap.parse_args()
    if args.config == "cfg5":
        global SCALING
        SCALING = "strong"
        world = int(os.environ.get("WORLD_SIZE", "1"))
        CFG.update(T=CFG5["T"], L=CFG5["L"], M=CFG5["M"], beats_per_gpu=CFG5["beats_total"] // world)
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
