#!/usr/bin/env python3
"""cProfile of one whole-fit scenario on the device path (host side: where the Python / launch / sync time of a fit goes).
usage: python tools/profile_fit.py rec100_online [n_rows]   -> gpurun_out/profile_<name>.txt"""
import cProfile, io, os, pstats, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
name = sys.argv[1] if len(sys.argv) > 1 else "rec100_online"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 45
from test_reference_fit_gpu import run_on_device
pr = cProfile.Profile()
pr.enable()
got, wall, launches, twins = run_on_device(name)
pr.disable()
out = io.StringIO()
out.write(f"{name}: device wall {wall:.1f} s, {launches} launches\n")
for key in ("cumulative", "tottime"):
    ps = pstats.Stats(pr, stream=out).strip_dirs().sort_stats(key)
    ps.print_stats(rows)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", f"profile_{name}.txt"), "w").write(out.getvalue())
print(out.getvalue()[:200])
