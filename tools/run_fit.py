#!/usr/bin/env python3
"""Run one whole-fit scenario (tests/fit_scenarios.py) of the UNMODIFIED reference driver on the device path and compare
it with the CPU fixture: python tools/run_fit.py rec100_offline [--verbose].  Writes gpurun_out/fit_<name>.json."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("--")]
    verbose = "--verbose" in sys.argv
    from test_reference_fit_gpu import run_on_device
    res = {}
    for name in names:
        t0 = time.time()
        got, wall, launches, twins = run_on_device(name, quiet=not verbose, profile=True)
        import hdpgpc_b200.integration as hgi
        stages = {k: [v[0], round(v[1], 3)] for k, v in sorted(hgi.seam_times.items(), key=lambda kv: -kv[1][1])}
        z = np.load(os.path.join(ROOT, "tests", "golden", f"fit_{name}.npz"))
        same = bool(np.array_equal(got["labels"], z["labels"])) and int(got["M"]) == int(z["M"])
        n_it = min(int(got["n_outer"]), int(z["n_outer"]))
        first_bad = None
        if "labels_per_iteration" in z.files and "labels_per_iteration" in got:
            for k in range(n_it):
                if not np.array_equal(got["labels_per_iteration"][k], z["labels_per_iteration"][k]):
                    first_bad = k
                    break
        el = min(len(got["elbo"]), len(z["elbo"]))
        res[name] = dict(identical_labels=same, M=int(got["M"]), M_ref=int(z["M"]), sizes=got["sizes"].tolist(),
                         sizes_ref=z["sizes"].tolist(), n_outer=int(got["n_outer"]), n_outer_ref=int(z["n_outer"]),
                         first_mismatching_iteration=first_bad,
                         elbo_rel_err=(np.abs(got["elbo"][:el] - z["elbo"][:el]) / np.abs(z["elbo"][:el])).tolist(),
                         kernels=got["kernels"].tolist(), kernels_ref=z["kernels"].tolist(),
                         device_s=round(wall, 2), cpu_reference_s=float(z["cpu_fit_seconds"]), gpu_launches=launches,
                         total_s=round(time.time() - t0, 2), seam_calls_seconds=stages)
        print(name, json.dumps(res[name]), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"fit_{name}.json"), "w") as f:
            json.dump(res[name], f, indent=1)


if __name__ == "__main__":
    main()
