"""The reference's OWN drivers (GPI_HDP.include_batch, GPI_HDP.py:805; GPI_HDP.include_sample, :1906) running end to end
on the device path: the unmodified reference package (mounted in the dev container, or the copy tools/make_ref.sh stages
under oracle/_ref/ for the GPU box) with `hdpgpc_b200.integration.enable()` -- every seam call of SURVEY.md section 8b goes
to the CUDA library, the VI control flow stays the reference's.

North star: "identical cluster assignments and cluster counts on bundled MIT-BIH records".  Each scenario of
tests/fit_scenarios.py (the entry scripts' hyper-parameters) was run once on the CPU by the unmodified reference
(tests/golden/generate_fit_golden.py -> tests/golden/fit_*.npz); here the same driver runs on the GPU and must arrive at the
SAME label vector after EVERY outer iteration, the same cluster count and the same ELBO trace (1e-6 relative: the ELBO sums
~1e5 scores that each agree to ~1e-10).  Wall times go to gpurun_out/fit_times.json (bench.py reports them)."""
import contextlib
import io
import json
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _reference():
    from oracle import refshim
    root = refshim.find_reference()
    if root is None:
        pytest.skip("reference package neither mounted nor staged (tools/make_ref.sh)")
    return refshim, root


def run_on_device(name, quiet=True, profile=False):
    """(summary, wall seconds, kernel launches) of scenario `name` with the seam patched onto the device."""
    import hdpgpc_b200 as hb
    import hdpgpc_b200.integration as hgi
    from fit_scenarios import SCENARIOS, run_fit, summarize
    refshim, root = _reference()
    hdp = refshim.install(root)
    import hdpgpc.GPI as gpi
    import hdpgpc.GPI_model as gm
    hgi.seam_times.clear()
    if os.environ.get("HGP_TUNE_MALLOC"):
        hgi.tune_host_allocator()
    if os.environ.get("HGP_FIT_TORCH_THREADS"):
        import torch
        torch.set_num_threads(int(os.environ["HGP_FIT_TORCH_THREADS"]))
    if os.environ.get("HGP_FIT_BLAS_THREADS"):
        import threadpoolctl
        threadpoolctl.threadpool_limits(int(os.environ["HGP_FIT_BLAS_THREADS"]), user_api="blas")
    hgi.enable(gm.GPI_model, hdp.GPI_HDP, gpi.IterativeGaussianProcess, profile=profile)
    launches0 = hb.ops.launch_count()
    try:
        with contextlib.redirect_stdout(io.StringIO() if quiet else sys.stdout):
            sw, wall = run_fit(hdp, SCENARIOS[name], os.path.join(root, "data", "mitbih"))
            out = summarize(sw)
            twins = sum(hgi.twin_of(gp) is not None for lead in sw.gpmodels for gp in lead)
    finally:
        hgi.disable()
    return out, wall, hb.ops.launch_count() - launches0, twins


def _record(name, wall, cpu_s, launches):
    path = os.path.join(ROOT, "gpurun_out", "fit_times.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    rec = {}
    if os.path.exists(path):
        with open(path) as f:
            rec = json.load(f)
    rec[name] = dict(device_s=round(wall, 3), cpu_reference_s=round(cpu_s, 3), gpu_launches=int(launches))
    with open(path, "w") as f:
        json.dump(rec, f, indent=1)


@pytest.mark.parametrize("name", ["rec100_limit30", "rec102_warp", "rec100_offline", "rec100_online"])
def test_reference_driver_on_device_reproduces_cpu_fit(name):
    fixture = os.path.join(GOLDEN, f"fit_{name}.npz")
    if not os.path.exists(fixture):
        pytest.skip(f"{fixture} not generated")
    z = np.load(fixture)
    got, wall, launches, twins = run_on_device(name)
    _record(name, wall, float(z["cpu_fit_seconds"]), launches)
    assert launches > 0 and twins > 0, "the fit did not run on the device library"
    assert int(got["M"]) == int(z["M"]), (int(got["M"]), int(z["M"]))
    assert np.array_equal(got["sizes"], z["sizes"]), (got["sizes"], z["sizes"])
    assert np.array_equal(got["labels"], z["labels"])
    assert int(got["n_outer"]) == int(z["n_outer"])
    if "labels_per_iteration" in z.files:
        assert np.array_equal(got["labels_per_iteration"], z["labels_per_iteration"])
    assert np.array_equal(got["members"], z["members"])
    np.testing.assert_allclose(got["elbo"], z["elbo"], rtol=1e-6)
    # the kernels the device hyper-fit returned (outputscale, noise) vs the CPU restatement's
    np.testing.assert_allclose(got["kernels"], z["kernels"], rtol=1e-5)
