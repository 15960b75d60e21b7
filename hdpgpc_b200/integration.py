"""Re-point the reference's E-step seam at the device library.

    import hdpgpc.GPI_HDP, hdpgpc.GPI_model            # the unmodified reference package
    import hdpgpc_b200.integration as hgi
    hgi.enable()                                        # patches the classes in place; hgi.disable() restores them

This is the patch of INTEGRATION.md section 3 as code: the reference keeps its VI control flow (`include_batch`,
births, accept / reject -- SURVEY.md section 2 #5, out of scope here) and its Python-list state; the seam methods
listed in SURVEY.md section 8b are replaced by wrappers that mirror the calling object on the device
(`GPI_model.from_reference`, cached while the object's histories do not change) and return CPU float64 tensors in the
reference's shapes.  There is NO CPU fallback: with the patch enabled and no CUDA device / library the calls raise
`HgpError`.  Patched:

  GPI_model.compute_sq_err_all (GPI_model.py:488-547)      GPI_model.compute_q_lat_all (:549-559)
  GPI_model.log_sq_error (:250-286, params=None)           GPI_model.return_LDS_param_likelihood (:459-486)
  GPI_HDP.compute_snr (GPI_HDP.py:732-748)                 GPI_HDP.cluster_new_batch (learning=False, :2975-3001)
  GPI_HDP.forward / backward / coupled_state_coef (:3546-3699) -- as one cached smoothing per (pi, q) pair

The reference package is not importable on the GPU test box, so this module is exercised by the CPU suite only
(tests/test_integration_patch.py: the patch applies, restores, and fails loudly without a device).
"""
import numpy as np
import torch

from . import hdp as _hdp
from ._lib import HgpError
from .model import GPI_model as _DevModel

_saved = {}


def _kernel_triple(gp):
    """(constant, length scale, white noise) of the reference model's sklearn kernel C * RBF + White."""
    try:
        kp = gp.gp.kernel.get_params()
        return (float(kp["k1__k1__constant_value"]), float(kp["k1__k2__length_scale"]), float(kp["k2__noise_level"]))
    except Exception:
        return None


def device_model(gp, device="cuda"):
    """Device mirror of a reference GPI_model, cached on the object until its histories change (the reference never
    mutates a stored tensor in place -- every update rebinds or appends, GPI.py:298-299 -- so the lengths of the lists
    and the identity of their last elements identify the state)."""
    key = (len(gp.f_star), len(gp.A), id(gp.f_star_sm[-1]), id(gp.cov_f_sm[-1]), id(gp.Sigma[-1]), tuple(gp.indexes[-2:]))
    cached = getattr(gp, "_hgp_dev", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    dev = _DevModel(gp.x_basis, gp.f_star, gp.f_star_sm, gp.C, gp.Sigma, gp.indexes,
                    estimation_limit=getattr(gp, "estimation_limit", None), A=gp.A, Gamma=gp.Gamma,
                    cov_f_sm=gp.cov_f_sm, cov_f=gp.cov_f, kernel=_kernel_triple(gp), device=device)
    if getattr(gp, "A_def", None) is not None:
        up = lambda t: torch.as_tensor(np.asarray(t.detach().cpu()), dtype=torch.float64).to(dev.device)
        dev.defaults = dict(A=up(gp.A_def), Gamma=up(gp.Gamma_def), C=up(gp.C_def), Sigma=up(gp.Sigma_def))
    gp._hgp_dev = (key, dev)
    return dev


def _cpu(t):
    return t.detach().to("cpu", torch.float64)


# ---- GPI_model seam ------------------------------------------------------------------------------------------------
def _compute_sq_err_all(self, x_trains, y_trains, no_first=False):
    return _cpu(device_model(self).compute_sq_err_all(x_trains, y_trains, no_first=no_first))


def _compute_q_lat_all(self, x_trains, h_ini=1.0):
    return _cpu(device_model(self).compute_q_lat_all(x_trains, h_ini=h_ini))


def _log_sq_error(self, x_train, y, mean=None, cov=None, C=None, Sigma=None, i=None, proj=False, first=False):
    if mean is not None or proj:
        return _saved["GPI_model.log_sq_error"](self, x_train, y, mean=mean, cov=cov, C=C, Sigma=Sigma, i=i, proj=proj,
                                                first=first)      # explicit-parameter form: estimate_new path, host
    return _cpu(device_model(self).log_sq_error(x_train, y, i=i, first=first))


def _return_LDS_param_likelihood(self, first=False):
    if first:
        return _saved["GPI_model.return_LDS_param_likelihood"](self, first=True)
    return _cpu(device_model(self).return_LDS_param_likelihood())


# ---- GPI_HDP seam --------------------------------------------------------------------------------------------------
def _mirror(sw):
    models = [[device_model(gp) for gp in lead] for lead in sw.gpmodels]
    return _hdp.GPI_HDP(models, sw.transTheta, sw.startTheta, snr_norm=getattr(sw, "snr_norm", None),
                        use_snr=getattr(sw, "use_snr", True))


def _compute_snr(self, y_trains, gp):
    if not getattr(self, "use_snr", True):
        return torch.ones(y_trains.shape[0])
    m = _hdp.GPI_HDP([[device_model(gp)]], self.transTheta, self.startTheta)
    return _cpu(m.compute_snr(y_trains, m.gpmodels[0][0]))


def _cluster_new_batch(self, x_trains, y_trains, learning=False, it_limit=None, warp=False):
    if learning or warp:
        return _saved["GPI_HDP.cluster_new_batch"](self, x_trains, y_trains, learning=learning, it_limit=it_limit, warp=warp)
    return _mirror(self).cluster_new_batch(x_trains, y_trains).cpu()


def _smoothing(self, pi, q):
    """One device smoothing per (pi, q): forward, backward and coupled_state_coef are three views of it."""
    key = (id(q), id(pi))
    c = getattr(self, "_hgp_smooth", None)
    if c is None or c[0] != key:
        m = _hdp.GPI_HDP([[]], self.transTheta, self.startTheta)
        hm = m._smooth(pi, q)
        ops_ = m._operands(pi, q.shape[1])
        c = (key, hm, ops_, pi)
        self._hgp_smooth = c
    return c


def _forward(self, pi, trans_A, q):
    _, hm, _, _ = _smoothing(self, pi, q)
    return _cpu(hm.alpha), _cpu(hm.marg)


def _backward(self, trans_A, q, margprob):
    c = getattr(self, "_hgp_smooth", None)
    if c is None or c[0][0] != id(q):
        raise HgpError("backward() without the matching forward() call: the device path smooths both directions at once")
    return _cpu(c[1].beta)


def _coupled_state_coef(self, alpha, beta, trans_A, q, margprobs):
    """log respPair (N, K, K): -inf everywhere except the arg-max pair of every beat -- all `_safe_exp` (:338-350)
    looks at; row 0 keeps the reference's all -inf row."""
    c = getattr(self, "_hgp_smooth", None)
    if c is None or c[0][0] != id(q):
        raise HgpError("coupled_state_coef() without the matching forward() call")
    zp = c[1].zpair.cpu().long()
    N, K = q.shape
    out = torch.full((N, K * K), -float("inf"), dtype=torch.float64)
    out[torch.arange(1, N), zp[1:]] = 0.0
    return out.reshape(N, K, K)


_PATCHES = {
    "GPI_model": {"compute_sq_err_all": _compute_sq_err_all, "compute_q_lat_all": _compute_q_lat_all,
                  "log_sq_error": _log_sq_error, "return_LDS_param_likelihood": _return_LDS_param_likelihood},
    "GPI_HDP": {"compute_snr": _compute_snr, "cluster_new_batch": _cluster_new_batch, "forward": _forward,
                "backward": _backward, "coupled_state_coef": _coupled_state_coef},
}


def enable(gpi_model_cls=None, gpi_hdp_cls=None):
    """Patch the reference classes (found in `hdpgpc.GPI_model` / `hdpgpc.GPI_HDP` unless given)."""
    if gpi_model_cls is None or gpi_hdp_cls is None:
        import importlib
        gpi_model_cls = gpi_model_cls or importlib.import_module("hdpgpc.GPI_model").GPI_model
        gpi_hdp_cls = gpi_hdp_cls or importlib.import_module("hdpgpc.GPI_HDP").GPI_HDP
    for cname, cls in (("GPI_model", gpi_model_cls), ("GPI_HDP", gpi_hdp_cls)):
        for name, fn in _PATCHES[cname].items():
            key = f"{cname}.{name}"
            if key not in _saved:
                _saved[key] = getattr(cls, name)
                _saved[key + "/cls"] = cls
            setattr(cls, name, fn)
    return sorted(k for k in _saved if not k.endswith("/cls"))


def disable():
    """Restore the reference's own methods."""
    for key in [k for k in _saved if not k.endswith("/cls")]:
        cls = _saved.pop(key + "/cls")
        setattr(cls, key.split(".", 1)[1], _saved.pop(key))
