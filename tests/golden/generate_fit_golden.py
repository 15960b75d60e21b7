#!/usr/bin/env python3
"""Whole-fit fixtures from the UNMODIFIED reference (run in the dev container, where /root/reference is mounted):

    python tests/golden/generate_fit_golden.py rec100_offline rec102_warp rec100_online [rec100_limit30]

Each scenario runs one of the reference's own entry scripts' fits on the CPU (hdpgpc/tests/test_offline.py:37-82,
hdpgpc/tests/test_online.py:36-87 -- same hyper-parameters, lead 0) and stores what `tests/test_reference_fit_gpu.py`
compares the device-backed run of the SAME reference driver against: the label vector after every outer iteration
(`resp_assigned`), the cluster count `M`, the ELBO trace (`train_elbo`), cluster sizes, the kernels the hyper-fits
returned, and the CPU wall time of the fit.

    rec100_offline   BASELINE.json configs[0]: include_batch on the full record 100 (2272 beats, T = 90)
    rec102_warp      BASELINE.json configs[2]: include_batch(warp=True) on the first 320 beats of record 102
    rec100_online    BASELINE.json configs[1]: include_sample beat by beat, first 500 beats of record 100
    rec100_limit30   the benchmarked "R1" regime (finite estimation_limit, GPI_model.py:646-651, :1092-1099): include_batch
                     on the first 160 beats of record 100 with estimation_limit = 30

The gpytorch hyper-fit is replaced by oracle/hyperfit.py (parity unpinned for that sub-step, see oracle/refshim)."""
import contextlib
import io
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from fit_scenarios import SCENARIOS, run_fit, summarize  # noqa: E402


def main(names):
    from oracle import refshim
    hdp = refshim.install()
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    for name in names:
        sc = SCENARIOS[name]
        t0 = time.time()
        log = io.StringIO()
        with contextlib.redirect_stdout(log):
            sw, wall = run_fit(hdp, sc, data_dir="/root/reference/hdpgpc/data/mitbih")
        out = summarize(sw)
        out["cpu_fit_seconds"] = np.float64(wall)
        out["cpu_threads"] = np.int64(torch.get_num_threads())
        np.savez_compressed(os.path.join(HERE, f"fit_{name}.npz"), **out)
        tail = [ln for ln in log.getvalue().splitlines() if "ELBO + Nonlinear" in ln or "Group responsability" in ln]
        with open(os.path.join(HERE, f"fit_{name}.log.txt"), "w") as f:
            f.write("\n".join(tail[-40:]) + "\n")
        print(f"{name}: M={int(out['M'])} sizes={out['sizes'].tolist()} fit {wall:.1f}s (total {time.time() - t0:.1f}s)",
              flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(SCENARIOS))
