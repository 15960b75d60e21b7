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
def cpu_sweep_beats_per_s(n_beats, T, L, M, steps=1, seed=1234):
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
        O.estep_responsibilities(q, snr, wl["transTheta"], wl["startTheta"])
        times.append(time.perf_counter() - t0)
    return n_beats / (sum(times) / len(times)), cores, times


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    T, L, M = CFG["T"], CFG["L"], CFG["M"]
    # a bounded sample per step: ~12 s of host work at the default, shrunk so that K steps still end within a few minutes
    n = max(1024, min(args.cpu_beats, (10 * args.cpu_beats // max(1, args.steps)) // 64 * 64))
    for _ in range(max(args.warmup, 0) and 1):
        cpu_sweep_beats_per_s(min(n, 256), T, L, M, steps=1)
    bps, cores, times = cpu_sweep_beats_per_s(n, T, L, M, steps=max(1, args.steps))
    sample = f"{n} beats of the same workload per step (oracle port of the reference loop: one Cholesky + cholesky_solve per distinct cluster state, python HMM loop)"
    line = {
        "impl": "reference", "metric": METRIC, "value": bps, "unit": UNIT, "n_gpus": args.gpus, "steps": max(1, args.steps),
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": SCALING,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(CFG["beats_per_gpu"], T, L, M), "cpu_sample_beats": n},
        "cpu_baseline": {"value": bps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": bps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 and len(r) >= 9] or [r for (_, r) in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows)}


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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hb.load_library()
    T, L, M = CFG["T"], CFG["L"], CFG["M"]
    B = args.beats or CFG["beats_per_gpu"]
    wl = synthetic.make_workload(B, T=T, L=L, M=M, seed=1234, device="cuda", n_offset=rank * B, N_total=world * B)
    Y_host = wl["Y"].cpu().pin_memory()                      # e2e leg: beats live in pinned host memory
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

    peak_burst = peak_sust = None
    if rank == 0 and not args.no_peak:
        peak_burst, peak_sust = measure_fp64_peak(torch)

    for _ in range(max(3, args.warmup)):
        sweep()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = ops.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        st, hm = sweep(record=True)
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = ops.launch_count() - launches0
    ms = ev0.elapsed_time(ev1) / args.steps
    tile_ms = sum(a.elapsed_time(b) for a, b in tile_events) / max(1, len(tile_events))
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t)

    # ---- end-to-end leg: host beats in, labels + statistics out ----
    def e2e_step():
        # the public end-to-end call: pinned host beats in (sliced H2D copies overlapped with scoring), labels and
        # statistics back on the host
        broadcast_tables()
        st = eng.sweep_from_host(Y_host)
        return st["z_host"], st["stats_host"]
    for _ in range(2):
        e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_e2e = max(1, min(args.steps, 5))
    for _ in range(n_e2e):
        z_host, stats_host = e2e_step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / n_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t)
    h2d = Y_host.numel() * 8
    d2h = z_host.numel() * 4 + stats_host.numel() * 8

    acc = float((hm.z.cpu() == torch.from_numpy(wl["labels"])).double().mean())

    if rank == 0:
        flops_launch = B * M * (T * T + 3 * T)               # one tile-kernel launch = one lead plane
        achieved = flops_launch / (tile_ms * 1e-3) / 1e12
        peak = peak_sust if peak_sust else 37.0
        mp = {}
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        traffic = None      # dram__bytes_read + dram__bytes_write of one launch, from the committed ncu --set full capture
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "r01_score_tiles_ncu.json")))
            if args.config == "cfg4" and B == CFG["beats_per_gpu"]:
                traffic = prof["dram_bytes_per_launch"]
        except Exception:
            pass
        bytes_launch = B * T * 8 + B * M * 8 + B * M * 4 + eng.leads[0].mu.numel() * 8 + eng.leads[0].Wpacked.numel() * 8
        cpu = None
        if not args.no_cpu:
            bps, cores, times = cpu_sweep_beats_per_s(args.cpu_beats, T, L, M, steps=1)
            cpu = {"value": bps, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{args.cpu_beats} beats of the same workload, 1 sweep, {times[0]:.1f} s (oracle port of the reference loop)"}
        line = {
            "metric": METRIC, "value": world * B / (ms_max * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_max, "higher_is_better": True, "scaling": SCALING,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(B, T, L, M), "beats_total": world * B, "l2": f"inputs_exceed_l2 (beats {B * T * L * 8 / 1e6:.0f} MB + whitened means {sum(tb.nu.numel() for tb in eng.leads) * 8 / 1e6:.0f} MB per GPU vs 126 MB L2)",
                       "label_accuracy": acc, "hmm_repair_rounds": eng.hmm_rounds, "boundary_rounds": eng.boundary_rounds},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "score_tiles_kernel", "kernel_ms": tile_ms,
                         "kernel_share_of_step": L * tile_ms / ms, "flops_per_launch": flops_launch,
                         "peak_source": "cuBLAS DGEMM 8192^3 float64 measured in this run, sustained (MEASURED_PEAKS.json has no FP64 figure)" if peak_sust else "nominal",
                         "peak_burst": peak_burst,
                         "hbm": {"algorithmic_bytes_per_launch": bytes_launch, "achieved_gbs": bytes_launch / (tile_ms * 1e-3) / 1e9,
                                 "peak_gbs": mp.get("hbm_gbs"), "frac": (bytes_launch / (tile_ms * 1e-3) / 1e9) / mp["hbm_gbs"] if mp.get("hbm_gbs") else None}},
            "cpu_baseline": cpu,
            "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms},
            "gpu_launches": int(launches),
            "clocks": clocks,
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
    ap.add_argument("--cpu-beats", type=int, default=8192, help="beats in the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-peak", action="store_true")
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
