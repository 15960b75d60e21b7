"""Wiring of hdpgpc_b200.integration against the UNMODIFIED reference on the CPU (dev container only).

The device twins are replaced by the stand-ins of tests/fake_twin.py (same method surface, arithmetic = the reference's own
un-patched code on a private copy of each model), so a whole fit driven by the reference's `include_batch` /
`include_sample` through the patched seam must reproduce the un-patched run BIT FOR BIT: same labels, same cluster count,
same ELBO trace, same final states.  That pins everything the patch itself decides -- list views, host bookkeeping,
copy-on-write copies, re-initialisation, which call goes to which twin, the HMM and warp call protocol -- independently
of the device arithmetic, which the -m gpu tests check."""
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
    from oracle import refshim
    hdp_mod = refshim.install()
    import hdpgpc.GPI as gpi
    import hdpgpc.GPI_model as gm
    return hdp_mod, gm, gpi


def _data(rec, n, stride, leads, start=0):
    d = np.load(os.path.join(REF, "data", "mitbih", rec + ".npy"))[start:start + n][:, ::stride, leads]
    return np.ascontiguousarray(d)


def _model(hdp_mod, data, free_deg=5, n_explore_steps=2, estimation_limit=None):
    from hdpgpc.get_data import compute_estimators_LDS
    N, T, L = data.shape
    std, std_dif, bound_sigma, bound_gamma = compute_estimators_LDS(data)
    x_basis = np.atleast_2d(np.arange(0, T, 1, dtype=np.float64)).T
    noise_warp = std * 0.1
    sw = hdp_mod.GPI_HDP(x_basis, x_basis_warp=x_basis[::2], n_outputs=L, kernels=None, model_type='dynamic',
                         ini_lengthscale=3.0, bound_lengthscale=(1.0, 20.0), ini_gamma=std_dif, ini_sigma=std,
                         ini_outputscale=300.0, noise_warp=noise_warp, bound_sigma=bound_sigma, bound_gamma=bound_gamma,
                         bound_noise_warp=(noise_warp * 0.1, noise_warp * 0.2), warp_updating=False,
                         method_compute_warp='greedy', verbose=False, hmm_switch=True, max_models=100, mode_warp='rough',
                         bayesian_params=True, inducing_points=False, estimation_limit=estimation_limit,
                         reestimate_initial_params=True, n_explore_steps=n_explore_steps, free_deg_MNIV=free_deg)
    return sw, x_basis


def _summary(sw):
    out = dict(M=sw.M, labels=[r.numpy().copy() for r in sw.resp_assigned],
               elbo=[float(e) for e in sw.train_elbo])
    states = []
    for lead in sw.gpmodels:
        for gp in lead:
            states.append((list(gp.indexes), gp.N, len(gp.f_star), len(gp.A),
                           gp.f_star_sm[-1].clone(), gp.cov_f_sm[-1].clone(), gp.Sigma[-1].clone(), gp.A[-1].clone()))
    out["states"] = states
    return out


def _same(a, b):
    assert a["M"] == b["M"]
    assert len(a["labels"]) == len(b["labels"])
    for x, y in zip(a["labels"], b["labels"]):
        assert np.array_equal(x, y)
    assert a["elbo"] == b["elbo"]
    assert len(a["states"]) == len(b["states"])
    for sa, sb in zip(a["states"], b["states"]):
        assert sa[:4] == sb[:4]
        for ta, tb in zip(sa[4:], sb[4:]):
            assert torch.equal(ta, tb)


@contextlib.contextmanager
def patched(ref):
    import hdpgpc_b200.integration as hgi
    from oracle.hyperfit import fit_exact_gp
    from fake_twin import make_fakes
    hdp_mod, gm, gpi = ref
    fake_model, fake_hdp, fake_ops = make_fakes(gm.GPI_model, hdp_mod.GPI_HDP, gpi.IterativeGaussianProcess, fit_exact_gp)
    keep = (hgi._DevModel, hgi._hdp, hgi._ops, hgi._DEVICE)
    hgi._DevModel, hgi._hdp, hgi._ops, hgi._DEVICE = fake_model, fake_hdp, fake_ops, "cpu"
    hgi.enable(gm.GPI_model, hdp_mod.GPI_HDP, gpi.IterativeGaussianProcess)
    try:
        yield hgi
    finally:
        hgi.disable()
        hgi._DevModel, hgi._hdp, hgi._ops, hgi._DEVICE = keep


def _offline(hdp_mod, data, spelling):
    sw, x_basis = _model(hdp_mod, data)
    x_trains = np.array([x_basis] * data.shape[0])
    with contextlib.redirect_stdout(io.StringIO()):
        if spelling == "with_warp":
            sw.include_batch(x_trains, data, with_warp=False)
        else:
            sw.include_batch(x_trains, data, warp=False)
    return sw


@pytest.mark.parametrize("rec,n,stride,leads", [("100", 40, 3, [0]), ("102", 48, 3, [0, 1])])
def test_offline_fit_through_the_patch_equals_reference(ref, rec, n, stride, leads):
    hdp_mod = ref[0]
    data = _data(rec, n, stride, leads)
    plain = _summary(_offline(hdp_mod, data, "warp"))
    with patched(ref) as hgi:
        sw = _offline(hdp_mod, data, "with_warp")        # the spelling tests/test_offline.py:79 uses
        assert all(hgi.twin_of(gp) is not None for lead in sw.gpmodels for gp in lead if gp.N > 0)
        got = _summary(sw)
        # back to plain reference objects (pickling, keep_last_all): same states, no twin left
        for lead in sw.gpmodels:
            for gp in lead:
                hgi.to_host(gp)
                assert hgi.twin_of(gp) is None and isinstance(gp.f_star, list)
        _same(got, _summary(sw))
    _same(plain, got)


def test_online_fit_through_the_patch_equals_reference(ref):
    hdp_mod = ref[0]
    data = _data("100", 14, 3, [0])

    def run():
        sw, x_basis = _model(hdp_mod, data, free_deg=20)
        with contextlib.redirect_stdout(io.StringIO()):
            for i in range(data.shape[0]):
                sw.include_sample(x_basis, data[i], with_warp=False)
        return sw

    plain = _summary(run())
    with patched(ref):
        got = _summary(run())
    _same(plain, got)
