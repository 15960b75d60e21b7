"""CPU stand-ins for the device classes `hdpgpc_b200.integration` drives -- TEST INFRASTRUCTURE ONLY.

`hdpgpc_b200.integration` re-points the reference's seam at device twins (`hdpgpc_b200.GPI_model`, `hdpgpc_b200.hdp`).
There is no GPU in the dev container, so the WIRING of that patch (list views, host bookkeeping, copy semantics, which
call goes where, in which order) is checked here against stand-ins with the same method surface whose arithmetic is the
reference's own, un-patched code running on a private copy of the model.  With them a whole `include_batch` /
`include_sample` run of the patched reference must reproduce the un-patched run bit for bit
(tests/test_integration_wiring.py).  The arithmetic of the real twins is checked on the GPU (tests/test_gpu_parity.py,
tests/test_reference_fit_gpu.py)."""
import copy
import types

import numpy as np
import torch

F64 = torch.float64

MODEL_METHODS = ("compute_sq_err_all", "compute_q_lat_all", "log_sq_error", "return_LDS_param_likelihood",
                 "full_pass_weighted", "include_weighted_sample", "backwards_pair", "bayesian_new_params", "backwards",
                 "reinit_GP", "reinit_LDS", "posterior_weighted")
HDP_METHODS = ("compute_snr", "compute_snr_ini", "estimate_new", "gpmodel_deepcopy", "cluster_new_batch", "include_batch",
               "forward", "backward", "coupled_state_coef", "warp_batch_by_resp_amtgp_cached")


def make_fakes(gpi_model_cls, gpi_hdp_cls, igp_cls, fit_exact_gp):
    """Build the stand-ins from the UN-PATCHED reference classes (call before integration.enable())."""
    Pristine = type("PristineGPIModel", (gpi_model_cls,), {n: getattr(gpi_model_cls, n) for n in MODEL_METHODS})
    PristineIGP = type("PristineIGP", (igp_cls,), {"fit_torch": igp_cls.fit_torch})
    PristineHDP = type("PristineHDP", (gpi_hdp_cls,), {n: getattr(gpi_hdp_cls, n) for n in HDP_METHODS})
    LISTS = ("f_star", "f_star_sm", "cov_f", "cov_f_sm", "A", "Gamma", "C", "Sigma", "y_var", "var", "x_train",
             "y_train", "likelihood", "indexes")

    def private_copy(gp):
        inner = copy.copy(gp)
        inner.__dict__.pop("_hgp", None)
        inner.__class__ = Pristine
        for n in LISTS:
            setattr(inner, n, list(getattr(gp, n)))
        inner.gp = copy.copy(gp.gp)
        inner.gp.__class__ = PristineIGP
        inner.gp.kernel = gp.gp.kernel.clone_with_theta(gp.gp.kernel.theta)
        return inner

    class FakeModel:
        def __init__(self, inner):
            self.inner = inner

        @classmethod
        def from_reference(cls, gp, device="cuda"):
            if len(torch.nonzero(gp.Gamma[-1])) < 1:
                raise RuntimeError("static model")
            return cls(private_copy(gp))

        # ---- struct-of-arrays attributes the list views read ----
        def _stack(self, name, col):
            lst = getattr(self.inner, name)
            if not len(lst):
                return torch.zeros((0,))
            t = torch.stack([torch.as_tensor(np.asarray(v), dtype=F64) for v in lst])
            return t[:, :, 0] if col else t

        f_star = property(lambda s: s._stack("f_star", True))
        f_star_sm = property(lambda s: s._stack("f_star_sm", True))
        cov_f = property(lambda s: s._stack("cov_f", False))
        cov_f_sm = property(lambda s: s._stack("cov_f_sm", False))
        A = property(lambda s: s._stack("A", False))
        Gamma = property(lambda s: s._stack("Gamma", False))
        C = property(lambda s: s._stack("C", False))
        Sigma = property(lambda s: s._stack("Sigma", False))
        indexes = property(lambda s: list(s.inner.indexes))
        N = property(lambda s: s.inner.N)
        fitted = property(lambda s: s.inner.fitted)

        @property
        def estimation_limit(self):
            return self.inner.estimation_limit

        @estimation_limit.setter
        def estimation_limit(self, v):
            self.inner.estimation_limit = v

        @property
        def annealing(self):
            return self.inner.annealing

        @annealing.setter
        def annealing(self, v):
            self.inner.annealing = v

        def _mniw(self, p):
            return dict(m_mean=p.m_mean, m_r_cov=p.m_r_cov, scale=p.scale, n0=torch.tensor([float(p.n0)]))

        internal = property(lambda s: s._mniw(s.inner.internal_params))
        observation = property(lambda s: s._mniw(s.inner.observation_params))

        def invalidate_caches(self):
            pass

        def to_reference_lists(self):
            d = {n: list(getattr(self.inner, n)) for n in ("f_star", "f_star_sm", "cov_f", "cov_f_sm", "A", "Gamma", "C",
                                                           "Sigma")}
            d.update(indexes=list(self.inner.indexes), N=self.inner.N)
            return d

        def clone(self):
            return FakeModel(private_copy(self.inner))

        # ---- the seam ----
        def include_weighted_sample(self, index, x_train, x_warped, y, h, snr=None):
            return self.inner.include_weighted_sample(index, x_train, x_warped, y, h, snr=snr)

        def backwards_pair(self, h, snr=None):
            return self.inner.backwards_pair(h, snr=snr)

        def bayesian_new_params(self, h, **kw):
            return self.inner.bayesian_new_params(h, **kw)

        def full_pass_weighted(self, x_trains, y_trains, resp, q=None, q_lat=None, snr=None):
            return self.inner.full_pass_weighted(x_trains, y_trains, resp, q=q, q_lat=q_lat, snr=snr)

        def reinit_GP(self, save_last=False, save_index=False):
            return self.inner.reinit_GP(save_last=save_last, save_index=save_index)

        def reinit_LDS(self, save_last=False, **kw):
            return self.inner.reinit_LDS(save_last=save_last, **kw)

        def compute_sq_err_all(self, x_trains, y_trains, no_first=False):
            return self.inner.compute_sq_err_all(x_trains, y_trains, no_first=no_first)

        def compute_q_lat_all(self, x_trains, h_ini=1.0):
            return self.inner.compute_q_lat_all(x_trains, h_ini=h_ini)

        def log_sq_error(self, x_train, y, **kw):
            return self.inner.log_sq_error(x_train, y, **kw)

        def return_LDS_param_likelihood(self, first=False):
            return self.inner.return_LDS_param_likelihood(first=first)

        def posterior_weighted(self, x_train, y, h, t=None):
            f, cov = self.inner.posterior_weighted(x_train, y, h, t=t)
            return f[:, 0], cov

        def estimate_new(self, x_train, y, h=1.0):
            return PristineHDP.estimate_new(None, 0, self.inner, x_train, y, h=h)

    class FakeHDP:
        """Stand-in for hdpgpc_b200.hdp.GPI_HDP as integration.py uses it."""

        def __init__(self, gpmodels, transTheta, startTheta, snr_norm=None, use_snr=True, device="cuda"):
            self.gpmodels = gpmodels
            h = object.__new__(PristineHDP)
            h.transTheta, h.startTheta = transTheta, startTheta
            h.cuda, h.device, h.verbose, h.use_snr = False, "cpu", False, use_snr
            h.inducing_points = []                # falsy: the f_star_sm branch of compute_snr (same values, see GPI.py:514)
            h.snr_norm = snr_norm
            h.trans_A, h.fmsg, h.margPrObs, h.M = 0, None, None, 0     # read but unused (GPI_HDP.py:3566, :3580)
            self.h = h

        def compute_snr(self, y_trains, gp):
            return self.h.compute_snr(y_trains, gp.inner)

        def compute_snr_ini(self, y_trains):
            self.h.compute_snr_ini(y_trains)
            return self.h.snr_norm

        def _smooth(self, pi, q):
            alpha, marg = self.h.forward(pi, 0, q)
            beta = self.h.backward(0, q, marg)
            pair = self.h.coupled_state_coef(alpha, beta, 0, q, marg)
            zpair = torch.argmax(pair.reshape(pair.shape[0], -1), dim=1)
            return types.SimpleNamespace(alpha=alpha, marg=marg, beta=beta, zpair=zpair)

        def cluster_new_batch(self, x_trains, y_trains):
            self.h.gpmodels = [[g.inner for g in lead] for lead in self.gpmodels]
            self.h.n_outputs, self.h.M = len(self.gpmodels), len(self.gpmodels[0])
            return self.h.cluster_new_batch(x_trains, y_trains)

    def hyperfit_batched(x, Y, noise_bounds, **kw):
        s, ell, noise, n_it = fit_exact_gp(x.cpu(), Y[0].cpu(), noise_bounds)
        return torch.tensor([[s, ell, noise, 0.0, 0.0, float(n_it), 0.0, 0.0]], dtype=F64)

    fake_ops = types.SimpleNamespace(hyperfit_batched=hyperfit_batched,
                                     _lib=types.SimpleNamespace(require_cuda=lambda: None))
    fake_hdp = types.SimpleNamespace(GPI_HDP=FakeHDP)
    return FakeModel, fake_hdp, fake_ops
