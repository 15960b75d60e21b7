"""The seam patch of hdpgpc_b200.integration against the UNMODIFIED reference package (dev container only: the
reference is not available on the GPU box, where these tests skip).  No GPU here, so what can be checked is the wiring:
the patch replaces exactly the seam methods, the patched calls refuse to run without a device (no CPU fallback), and
`disable()` restores the reference's own methods."""
import contextlib
import io
import os

import numpy as np
import pytest
import torch

REF = "/root/reference/hdpgpc"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference package not mounted")


@pytest.fixture(scope="module")
def ref():
    from oracle import refshim          # test infrastructure: import shims for the reference's missing dependencies
    hdp_mod = refshim.install()
    import hdpgpc.GPI_model as gm
    return hdp_mod, gm


def _tiny_model(hdp_mod):
    rng = np.random.default_rng(0)
    T, N = 12, 6
    x = np.atleast_2d(np.arange(T, dtype=np.float64)).T
    data = (30.0 * np.sin(np.arange(T) / 2.0)[None, :, None] + rng.normal(size=(N, T, 1))).astype(np.float64)
    with contextlib.redirect_stdout(io.StringIO()):
        sw = hdp_mod.GPI_HDP(x, x_basis_warp=x[::2], n_outputs=1, ini_lengthscale=3.0, bound_lengthscale=(1.0, 20.0),
                             ini_gamma=1.0, ini_sigma=1.0, ini_outputscale=300.0, noise_warp=0.1, bound_sigma=(0.5, 5.0),
                             bound_gamma=(0.5, 5.0), bound_noise_warp=(0.01, 0.02), verbose=False, hmm_switch=True,
                             bayesian_params=True, inducing_points=False, n_explore_steps=1, free_deg_MNIV=5)
        gp = sw.create_gp_default()
        resp = torch.ones(N, dtype=torch.float64)
        xt = torch.from_numpy(np.array([x] * N))
        gp.full_pass_weighted(xt, torch.from_numpy(data), resp)
    return sw, gp, xt, torch.from_numpy(data)


def test_patch_applies_fails_loudly_and_restores(ref):
    import hdpgpc_b200 as hb
    import hdpgpc_b200.integration as hgi
    hdp_mod, gm = ref
    sw, gp, xt, yt = _tiny_model(hdp_mod)
    with contextlib.redirect_stdout(io.StringIO()):
        q_ref = gp.compute_sq_err_all(xt, yt)
    originals = {n: getattr(gm.GPI_model, n) for n in hgi._PATCHES["GPI_model"]}
    originals_h = {n: getattr(hdp_mod.GPI_HDP, n) for n in hgi._PATCHES["GPI_HDP"]}
    patched = hgi.enable(gm.GPI_model, hdp_mod.GPI_HDP)
    try:
        assert set(patched) == {f"{cls}.{n}" for cls, table in hgi._PATCHES.items() for n in table}
        for n, f in originals.items():
            assert getattr(gm.GPI_model, n) is not f
        # everything outside the seam is untouched: the VI control flow stays the reference's
        # (include_batch only gains a keyword shim: tests/test_offline.py calls it with `with_warp=`)
        for n in ("include_sample", "estimate_q_all", "variational_local_terms_batch", "compute_q_elbo"):
            assert getattr(hdp_mod.GPI_HDP, n).__module__ == hdp_mod.GPI_HDP.__module__
        if not torch.cuda.is_available():
            with pytest.raises(hb.HgpError):          # no device -> loud failure, never a silent CPU path
                gp.compute_sq_err_all(xt, yt)
            with pytest.raises(hb.HgpError):
                sw.cluster_new_batch(xt.numpy(), yt.numpy())
    finally:
        hgi.disable()
    for n, f in originals.items():
        assert getattr(gm.GPI_model, n) is f
    for n, f in originals_h.items():
        assert getattr(hdp_mod.GPI_HDP, n) is f
    with contextlib.redirect_stdout(io.StringIO()):
        assert torch.equal(gp.compute_sq_err_all(xt, yt), q_ref)


def test_kernel_triple_and_cache_key(ref):
    import hdpgpc_b200.integration as hgi
    hdp_mod, gm = ref
    sw, gp, xt, yt = _tiny_model(hdp_mod)
    c, ell, noise = hgi._kernel_triple(gp)
    kp = gp.gp.kernel.get_params()
    assert (c, ell, noise) == (kp["k1__k1__constant_value"], kp["k1__k2__length_scale"], kp["k2__noise_level"])
    assert ell == 1.2                                   # fit_torch overrides the fitted lengthscale (GPI.py:711)
