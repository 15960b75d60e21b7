"""Run the UNMODIFIED reference package on the device library.

    import hdpgpc.GPI_HDP, hdpgpc.GPI_model            # the reference package
    import hdpgpc_b200.integration as hgi
    hgi.enable()                                        # patches the classes in place; hgi.disable() restores them
    sw_gp.include_batch(x_trains, data, with_warp=False)   # hdpgpc/tests/test_offline.py:79, now on the GPU

This is the patch of INTEGRATION.md section 3 as code.  The reference keeps its VI control flow (`include_batch`,
`include_sample`, births, accept / reject -- SURVEY.md section 2 #5, out of scope here); every seam method of SURVEY.md
section 8b is re-pointed at a DEVICE TWIN of the calling model:

* a reference `GPI_model` gets a twin (`hdpgpc_b200.GPI_model`, struct-of-arrays on the GPU) the first time a seam
  method needs one; from then on the twin is the state and the object's Python lists (`f_star`, `cov_f`, `A`, ... --
  GPI_model.py:35-47) are read-only VIEWS of it (`_DevList`: an element is downloaded when somebody indexes it), so
  plots, `print_results` and un-patched read-only helpers keep working, nothing is uploaded per call, and
  `gpmodel_deepcopy` is O(1) (copy-on-write device histories);
* `to_host(gp)` turns the object back into a plain reference model (lists of CPU tensors, host MNIW objects), e.g.
  before pickling or for the rare paths that edit the lists directly (`reinit_*(save_last=True)`, `keep_last_all`).

Patched (reference line numbers):

  GPI_model.full_pass_weighted (GPI_model.py:377-406)        chain replay + scores, one launch per chain
  GPI_model.include_weighted_sample (:353-375), backwards_pair (:705-724), bayesian_new_params (:966-1115), backwards
  GPI_model.reinit_GP (:408-434), reinit_LDS (:437-456)
  GPI_model.compute_sq_err_all (:488-547), compute_q_lat_all (:549-559), log_sq_error (:250-286),
            return_LDS_param_likelihood (:459-486), posterior_weighted (:561-582)
  IterativeGaussianProcess.fit_torch (GPI.py:610-770)        one-beat hyper-fit (ExactGP branch)
  GPI_HDP.compute_snr (GPI_HDP.py:732-748), compute_snr_ini (:715-730), estimate_new (:2830-2842),
          gpmodel_deepcopy (:4037-4064), cluster_new_batch(learning=False) (:2975-3001),
          forward / backward / coupled_state_coef (:3546-3699), warp_batch_by_resp_amtgp_cached (:3412-3517),
          include_batch (:805; accepts the `with_warp=` spelling of tests/test_offline.py:79 as well)

There is NO CPU fallback: with the patch enabled and no CUDA device / library every seam call raises `HgpError`.
"""
import numpy as np
import torch

from . import hdp as _hdp
from . import ops as _ops
from . import warp as _warp
from ._lib import HgpError
from .model import GPI_model as _DevModel

F64 = torch.float64
_DEVICE = "cuda"
_saved = {}

_COLS = ("f_star", "f_star_sm")
_MATS = ("cov_f", "cov_f_sm", "A", "Gamma", "C", "Sigma")


# ---- list views ----------------------------------------------------------------------------------------------------
class _DevList:
    """Read-only view of one device history in the reference's list form: element i is a CPU float64 tensor, (T, 1) for
    means and (T, T) for matrices (GPI_model.py:35-47).  The view follows its owner's twin, so it never goes stale."""

    def __init__(self, owner, name):
        self._owner, self._name, self._col = owner, name, name in _COLS

    def _t(self):
        tw = self._owner.__dict__.get("_hgp")
        if tw is None:
            raise HgpError("list view without a device twin (use hdpgpc_b200.integration.to_host)")
        return getattr(tw, self._name)

    def __len__(self):
        t = self._t()
        return 0 if t is None else int(t.shape[0])

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        t = self._t()
        n = int(t.shape[0])
        i = int(i)
        i = i + n if i < 0 else i
        if not 0 <= i < n:
            raise IndexError("list index out of range")
        v = t[i].detach().to("cpu", F64)
        return v.reshape(-1, 1) if self._col else v

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def copy(self):
        return list(self)

    def __setitem__(self, i, v):
        raise HgpError(f"{self._name}[...] = ...: the device twin owns this history; call integration.to_host(model) first")

    def append(self, v):
        raise HgpError(f"{self._name}.append: the device twin owns this history; call integration.to_host(model) first")


def twin_of(gp):
    return gp.__dict__.get("_hgp")


def _kernel_triple(gp):
    """(constant, length scale, white noise) of the reference model's sklearn kernel C * RBF + White."""
    try:
        kp = gp.gp.kernel.get_params()
        return (float(kp["k1__k1__constant_value"]), float(kp["k1__k2__length_scale"]), float(kp["k2__noise_level"]))
    except Exception:
        return None


def _attach(gp, tw):
    gp._hgp = tw
    for n in _COLS + _MATS:
        setattr(gp, n, _DevList(gp, n))
    return tw


def device_model(gp, device=None):
    """The device twin of a reference GPI_model; built once from the object's (short) host lists, then authoritative."""
    tw = gp.__dict__.get("_hgp")
    if tw is None:
        tw = _attach(gp, _DevModel.from_reference(gp, device=device or _DEVICE))
    return tw


def to_host(gp):
    """Back to a plain reference model: lists of CPU tensors, host MNIW objects; the twin is dropped."""
    tw = gp.__dict__.pop("_hgp", None)
    if tw is None:
        return gp
    lists = tw.to_reference_lists()
    for n in _COLS + _MATS:
        setattr(gp, n, lists[n])
    gp.indexes, gp.N = list(tw.indexes), tw.N
    mniw = type(gp.internal_params) if getattr(gp, "internal_params", None) is not None else None
    if mniw is not None and hasattr(tw, "internal"):
        for attr, st in (("internal_params", tw.internal), ("observation_params", tw.observation)):
            setattr(gp, attr, mniw(st["m_mean"].cpu().clone(), st["m_r_cov"].cpu().clone(), float(st["n0"][0]),
                                   st["scale"].cpu().clone()))
    return gp


def _cpu(t):
    return t.detach().to("cpu", F64)


def _is_static(gp):
    return len(gp.Gamma) and len(torch.nonzero(torch.as_tensor(np.asarray(gp.Gamma[-1])))) < 1


# ---- IterativeGaussianProcess.fit_torch -----------------------------------------------------------------------------
def _fit_torch(self, x, y, alpha_ini, gamma_ini, reduced_points=False, verbose=False):
    """GPI.py:610-770, ExactGPModel branch: the Adam optimisation of the marginal likelihood runs on the device
    (hgp_hyperfit_batched); the side effects on `self` are the reference's (:704-769)."""
    if self.fitted:
        return self.fitted
    x_ = x.detach().T[0]
    y_ = y.detach().T[0]
    x_basis = self.x_basis.T[0].detach().clone()
    if reduced_points or not torch.equal(x_basis, x_):
        raise HgpError("fit_torch: only the ExactGPModel branch (x_train == x_basis, no inducing points) is built")
    _ops._lib.require_cuda()
    out = _ops.hyperfit_batched(x_.to(_DEVICE, F64), y_.to(_DEVICE, F64).reshape(1, -1),
                                self.kernel.k2.noise_level_bounds)[0].cpu().numpy()
    if int(out[6]):
        raise HgpError("hyper-fit: kernel matrix lost positive-definiteness")
    self.hyperfit_iterations = int(out[5])
    if hasattr(self.kernel.k1, "k1"):
        self.kernel.k1.k1.theta = np.log(np.array([float(out[0])]))
    if hasattr(self.kernel.k1, "k2"):
        self.kernel.k1.k2.theta = np.log(np.array([1.2]))
    else:
        self.kernel.k1.theta = np.log(np.array([float(out[1])]))
    self.kernel.k2.theta = np.log(np.array([float(out[2])]))
    xb = self.cond_to_numpy(self.x_basis)
    self.K_X_X = self.cond_to_torch(self.kernel(xb, xb))
    self.K_inv = self.inv_r("kernelMat", self.K_X_X)
    self.fitted = True
    ident = torch.eye(self.x_basis.shape[0])
    alph_ = torch.mul(self.cond_to_torch(self.kernel.k2.noise_level), ident)
    gam_ = torch.mul(self.cond_to_torch(gamma_ini), ident)
    self.assign_alpha_ini(alph_, gam_)
    return self.fitted


# ---- GPI_model seam ------------------------------------------------------------------------------------------------
def _fit_first(self, x_train, y):
    """GPI_model.include_weighted_sample :361-365: an unfitted, empty model fits its kernel on the first beat.  The
    bookkeeping of fit_kernel_params (:207-241) stays the reference's own code on host lists (one element each); the
    optimisation inside it is `_fit_torch`."""
    if twin_of(self) is not None:
        to_host(self)
    valid = bool(torch.allclose(torch.from_numpy(self.gp.kernel.theta), torch.from_numpy(self.ini_kernel_theta)))
    new_x_basis, _ = self.fit_kernel_params(x_train, y, self.Sigma[-1], self.Gamma[-1], valid=valid)
    return new_x_basis


def _include_weighted_sample(self, index, x_train, x_warped, y, h, snr=None):
    if snr is not None:
        raise HgpError("include_weighted_sample(snr=...) is not built")
    y = self.cond_to_cuda(self.cond_to_torch(y))
    x_train = self.cond_to_cuda(self.cond_to_torch(x_train))
    new_x_basis = self.x_basis
    if h != 1.0:
        return new_x_basis                 # include_sample(posterior=False): nothing is stored (:343-351)
    if self.N == 0 and not self.fitted:
        new_x_basis = _fit_first(self, x_train, y)
    tw = device_model(self)
    tw.include_weighted_sample(index, x_train, x_warped, y, 1.0)
    self.N = self.N + 1
    self.indexes.append(index)
    self.x_train.append(x_train)
    self.y_train.append(self.cond_to_torch(y))
    return new_x_basis


def _backwards_pair(self, h, snr=None):
    if snr is not None:
        raise HgpError("backwards_pair(snr=...) is not built")
    if len(self.indexes) > 1 and h == 1.0:
        device_model(self).backwards_pair(h)


def _bayesian_new_params(self, h, model_type="dynamic", full_data=False, q=None, force=False, snr=1.0):
    if h != 1.0:
        return                             # the reference does nothing for a beat the cluster did not take (:972)
    device_model(self).bayesian_new_params(h, model_type=model_type, full_data=full_data, q=q, force=force, snr=snr)


def _backwards(self, h=1.0):
    if twin_of(self) is None and len(self.f_star) <= 2:
        return _saved["GPI_model.backwards"](self, h)
    raise HgpError("backwards() on its own is not built: the full RTS pass runs inside full_pass_weighted")


def _full_pass_weighted(self, x_trains, y_trains, resp, q=None, q_lat=None, snr=None):
    resp_t = torch.as_tensor(np.asarray(resp)) if not isinstance(resp, torch.Tensor) else resp
    active = torch.nonzero(resp_t > 0.99, as_tuple=False).squeeze(1)
    if active.numel() == 0:
        return q, q_lat
    if _is_static(self):
        raise HgpError("full_pass_weighted: static models (Gamma = 0) are outside the built path; no CPU fallback")
    if self.N != 0:
        raise HgpError("full_pass_weighted on a model that already has members is not built (the reference's drivers "
                       "re-initialise first: GPI_HDP.py:2890-2901)")
    first = int(active[0])
    if not self.fitted:
        _fit_first(self, self.cond_to_torch(x_trains[first]), self.cond_to_torch(y_trains[first]))
    tw = device_model(self)
    q_, q_lat_ = tw.full_pass_weighted(x_trains, y_trains, resp_t)
    idx = [int(i) for i in active.tolist()]
    self.N = len(idx)
    self.indexes = idx
    self.x_train = [x_trains[i] for i in idx]
    self.y_train = [self.cond_to_torch(y_trains[i]) for i in idx]
    return _cpu(q_), _cpu(q_lat_)


def _reset_host_bookkeeping(self):
    self.y_var = self.y_var[:1]
    self.var = self.var[:1]
    self.indexes = []
    self.y_train = []
    self.x_train = []
    self.likelihood = []
    self.N = 0


def _reinit_GP(self, save_last=False, save_index=False):
    tw = twin_of(self)
    if tw is None or save_last or not getattr(tw, "fitted", True):
        to_host(self)
        return _saved["GPI_model.reinit_GP"](self, save_last=save_last, save_index=save_index)
    tw.reinit_GP()
    _reset_host_bookkeeping(self)


def _reinit_LDS(self, save_last=False, save_last_diag=False, return_likelihood=False):
    tw = twin_of(self)
    if tw is None or save_last or return_likelihood:
        to_host(self)
        return _saved["GPI_model.reinit_LDS"](self, save_last=save_last, save_last_diag=save_last_diag,
                                              return_likelihood=return_likelihood)
    tw.reinit_LDS()


def _compute_sq_err_all(self, x_trains, y_trains, no_first=False):
    if len(self.indexes) == 0:
        _ops._lib.require_cuda()
        return torch.zeros(x_trains.shape[0], dtype=F64)                   # :494-495
    return _cpu(device_model(self).compute_sq_err_all(x_trains, y_trains, no_first=no_first))


def _compute_q_lat_all(self, x_trains, h_ini=1.0):
    if self.N == 0:
        _ops._lib.require_cuda()
        return torch.zeros(x_trains.shape[0], dtype=F64)                   # :553-554
    return _cpu(device_model(self).compute_q_lat_all(x_trains, h_ini=h_ini))


def _log_sq_error(self, x_train, y, mean=None, cov=None, C=None, Sigma=None, i=None, proj=False, first=False):
    if x_train is None:
        x_train = self.x_basis
    if i is None and mean is None:
        raise HgpError("log_sq_error(i=None) without explicit parameters (step_forward_last) is not built")
    return _cpu(device_model(self).log_sq_error(x_train, y, mean=mean, cov=cov, C=C, Sigma=Sigma, i=i, proj=proj,
                                                first=first))


def _return_LDS_param_likelihood(self, first=False):
    tw = device_model(self)
    v = tw.return_LDS_param_likelihood(first=first)
    if first:
        return _cpu(v)
    c = tw.__dict__.get("_lds_lik_host")            # host copy of the twin's cached value: one download per parameter update
    if c is None or c[0] is not v:
        c = (v, _cpu(v))
        tw._lds_lik_host = c
    return c[1]


def _posterior_weighted(self, x_train, y, h, t=None):
    f, cov = device_model(self).posterior_weighted(x_train, y, h, t=t)
    return _cpu(f).reshape(-1, 1), _cpu(cov)


# ---- GPI_HDP seam --------------------------------------------------------------------------------------------------
def _mirror(sw):
    models = [[device_model(gp) for gp in lead] for lead in sw.gpmodels]
    return _hdp.GPI_HDP(models, sw.transTheta, sw.startTheta, snr_norm=getattr(sw, "snr_norm", None),
                        use_snr=getattr(sw, "use_snr", True))


def _compute_snr(self, y_trains, gp):
    if not getattr(self, "use_snr", True):
        return torch.ones(y_trains.shape[0])
    # `self.inducing_points` is a per-model list, so the reference always goes through resample_latent_mean (:740-741),
    # which is the identity on the basis grid (GPI.py:514-516); other grids are not built
    xt = self.x_train[-1] if len(self.x_train) else None
    if xt is not None and not torch.equal(torch.as_tensor(np.asarray(xt), dtype=F64).reshape(-1),
                                          torch.as_tensor(np.asarray(gp.x_basis), dtype=F64).reshape(-1)):
        raise HgpError("compute_snr on a grid other than x_basis (resample_latent_mean) is not built")
    m = _hdp.GPI_HDP([[device_model(gp)]], self.transTheta, self.startTheta)
    return _cpu(m.compute_snr(y_trains, m.gpmodels[0][0]))


def _compute_snr_ini(self, y_trains):
    if not getattr(self, "use_snr", True):
        self.snr_norm = torch.ones(y_trains.shape[0], y_trains.shape[2])
        return
    m = _hdp.GPI_HDP([[]], self.transTheta, self.startTheta)
    self.snr_norm = _cpu(m.compute_snr_ini(np.ascontiguousarray(np.asarray(y_trains, dtype=np.float64))))


def _estimate_new(self, t, gpmodel, x_train, y, h=1.0):
    return _cpu(device_model(gpmodel).estimate_new(x_train, y, h)).reshape(())


def _gpmodel_deepcopy(self, gpmodel):
    """GPI_HDP.gpmodel_deepcopy (:4037-4064): a new model object that shares the immutable histories.  Quirks kept: the
    copy is built with the constructor's defaults, so it forgets `estimation_limit` (-> inf), `annealing` (-> True)
    and `bayesian` (-> False)."""
    tw = twin_of(gpmodel)
    if tw is None:
        return _saved["GPI_HDP.gpmodel_deepcopy"](self, gpmodel)
    cls = type(gpmodel)
    gp_ = cls(gpmodel.gp.kernel.clone_with_theta(gpmodel.gp.kernel.theta), gpmodel.x_basis.clone(), verbose=self.verbose)
    gp_.y_train = list(gpmodel.y_train)
    gp_.x_train = list(gpmodel.x_train)
    gp_.y_var = list(gpmodel.y_var)
    gp_.var = list(gpmodel.var)
    tw2 = tw.clone()
    if tw2.estimation_limit != np.inf or not getattr(tw2, "annealing", True):
        tw2.estimation_limit = np.inf
        tw2.annealing = True
        tw2.invalidate_caches()
    _attach(gp_, tw2)
    gp_.gp.assign_alpha_ini(gp_.Sigma[0], gp_.Gamma[0])
    gp_.likelihood = list(gpmodel.likelihood)
    gp_.N = gpmodel.N
    gp_.indexes = list(gpmodel.indexes)
    gp_.fitted = gpmodel.fitted
    gp_.ini_cov_def = gpmodel.ini_cov_def
    gp_.A_def, gp_.Gamma_def = gpmodel.A_def, gpmodel.Gamma_def
    gp_.C_def, gp_.Sigma_def = gpmodel.C_def, gpmodel.Sigma_def
    gp_.internal_params = gpmodel.internal_params
    gp_.observation_params = gpmodel.observation_params
    gp_.ini_kernel_theta = gpmodel.ini_kernel_theta
    gp_.free_deg_MNIV = gpmodel.free_deg_MNIV
    return gp_


def _cluster_new_batch(self, x_trains, y_trains, learning=False, it_limit=None, warp=False):
    if learning or warp:
        return _saved["GPI_HDP.cluster_new_batch"](self, x_trains, y_trains, learning=learning, it_limit=it_limit, warp=warp)
    return _mirror(self).cluster_new_batch(x_trains, y_trains).cpu()


def _include_batch(self, x_trains, y_trains, it_limit=None, warp=False, with_warp=None):
    """tests/test_offline.py:79 passes `with_warp=`; the signature says `warp=` (GPI_HDP.py:805).  Both are accepted."""
    if with_warp is not None:
        warp = with_warp
    return _saved["GPI_HDP.include_batch"](self, x_trains, y_trains, it_limit=it_limit, warp=warp)


def _smoothing(self, pi, q):
    """One device smoothing per (pi, q): forward, backward and coupled_state_coef are three views of it.  The cache
    entry holds the keyed tensors themselves (compared with `is`), so a recycled id() can never match."""
    c = getattr(self, "_hgp_smooth", None)
    if c is None or c[0] is not q or c[1] is not pi:
        m = _hdp.GPI_HDP([[]], self.transTheta, self.startTheta)
        c = (q, pi, m._smooth(pi, q))
        self._hgp_smooth = c
    return c[2]


def _forward(self, pi, trans_A, q):
    hm = _smoothing(self, pi, q)
    return _cpu(hm.alpha), _cpu(hm.marg)


def _backward(self, trans_A, q, margprob):
    c = getattr(self, "_hgp_smooth", None)
    if c is None or c[0] is not q:
        raise HgpError("backward() without the matching forward() call: the device path smooths both directions at once")
    return _cpu(c[2].beta)


def _coupled_state_coef(self, alpha, beta, trans_A, q, margprobs):
    """log respPair (N, K, K): -inf everywhere except the arg-max pair of every beat -- all `_safe_exp` (:338-350)
    looks at; row 0 keeps the reference's all -inf row."""
    c = getattr(self, "_hgp_smooth", None)
    if c is None or c[0] is not q:
        raise HgpError("coupled_state_coef() without the matching forward() call")
    zp = c[2].zpair.cpu().long()
    N, K = q.shape
    out = torch.full((N, K * K), -float("inf"), dtype=F64)
    out[torch.arange(1, N), zp[1:]] = 0.0
    return out.reshape(N, K, K)


def _dev_warper(ws):
    """Device mirror of a reference Warping_system (amtgp_warping_system.py:266-330), cached on the object."""
    d = ws.__dict__.get("_hgp")
    if d is None:
        d = _warp.Warping_system(ws.x_basis, ws.noise_warp_default, ws.noise_bounds, recursive=ws.recursive,
                                 bayesian=ws.bayesian, mode=ws.mode, n_ctrl=ws.n_ctrl, lr=ws.lr,
                                 lambda_smooth=ws.lambda_smooth_base, lambda_amp=ws.lambda_amp_base)
        d.warp_gp.noise_warp = float(ws.warp_gp.noise_warp)
        d.warp_gp.noise_bounds = tuple(float(b) for b in ws.warp_gp.noise_bounds)
        ws._hgp = d
    return d


def _warp_batch_by_resp_amtgp_cached(self, x_trains, y_trains, resp_temp, f_ind_old=None, train_iter=50, batch_size=128):
    """GPI_HDP.warp_batch_by_resp_amtgp_cached (:3412-3517): same cache (`_warp_cache_full`, keyed by lead and
    representative beat), the chunk loop (:3476-3504) replaced by one launch for all beats x all uncached
    representatives of a lead (hdpgpc_b200.warp.warp_batch_by_resp)."""
    if not self.warp:
        return _saved["GPI_HDP.warp_batch_by_resp_amtgp_cached"](self, x_trains, y_trains, resp_temp, f_ind_old=f_ind_old,
                                                                  train_iter=train_iter, batch_size=batch_size)
    _ops._lib.require_cuda()
    x_trains = self.cond_to_torch(x_trains)
    y_trains = self.cond_to_torch(y_trains)
    resp_temp = self.cond_to_torch(resp_temp)
    if f_ind_old is None:
        f_ind_old = self.f_ind_old
    N, T, D_out = y_trains.shape
    M = resp_temp.shape[1]
    assert D_out == self.n_outputs
    if not hasattr(self, "_warp_cache_full"):
        self._warp_cache_full = {}
    y_trains_w = torch.empty((N, T, D_out, M), dtype=F64)
    x_w = torch.empty((N, T, self.n_outputs, M), dtype=F64)
    liks_full = torch.zeros((N, M, self.n_outputs), dtype=F64)
    theta = self.kernel_def.get_params()["k1__k2__length_scale"]
    noise_scalar = float(np.sqrt(self.ini_sigma_def))
    for ld in range(self.n_outputs):
        refs = [int(f_ind_old[m].item()) for m in range(M)]
        todo = []
        for m, ref in enumerate(refs):
            if (ld, ref) not in self._warp_cache_full and ref not in [r for _, r in todo]:
                todo.append((m, ref))
        if todo:
            Yd = y_trains[:, :, ld].to(_DEVICE, F64).contiguous()
            x0 = x_trains[todo[0][1]]
            if any(not torch.equal(x_trains[r], x0) for _, r in todo):
                raise HgpError("warp driver: representatives on different grids are not built")
            warpers = [_dev_warper(self.wp_sys[ld][min(m, len(self.wp_sys[ld]) - 1)]) for m, _ in todo]
            base = _dev_warper(self.wp_sys[ld][-1])
            noise_vec = noise_scalar * torch.ones(T, dtype=F64)
            yw, xw, lk = _warp.warp_batch_by_resp(x0, Yd, [r for _, r in todo], warpers, base, theta, noise_vec,
                                                  train_iter=train_iter, batch_size=batch_size)
            yw, xw, lk = _cpu(yw), _cpu(xw), _cpu(lk)
            for k, (_, ref) in enumerate(todo):
                self._warp_cache_full[(ld, ref)] = (xw[:, :, k].contiguous(), yw[:, :, k].contiguous(),
                                                    lk[:, k].contiguous())
        for m, ref in enumerate(refs):
            xw_all_2d, yw_all_2d, lik_all = self._warp_cache_full[(ld, ref)]
            liks_full[:, m, ld] = lik_all
            y_trains_w[:, :, ld, m] = yw_all_2d
            x_w[:, :, ld, m] = xw_all_2d
    return y_trains_w, x_w, liks_full


_PATCHES = {
    "GPI_model": {"compute_sq_err_all": _compute_sq_err_all, "compute_q_lat_all": _compute_q_lat_all,
                  "log_sq_error": _log_sq_error, "return_LDS_param_likelihood": _return_LDS_param_likelihood,
                  "full_pass_weighted": _full_pass_weighted, "include_weighted_sample": _include_weighted_sample,
                  "backwards_pair": _backwards_pair, "bayesian_new_params": _bayesian_new_params, "backwards": _backwards,
                  "reinit_GP": _reinit_GP, "reinit_LDS": _reinit_LDS, "posterior_weighted": _posterior_weighted},
    "GPI_HDP": {"compute_snr": _compute_snr, "compute_snr_ini": _compute_snr_ini, "estimate_new": _estimate_new,
                "gpmodel_deepcopy": _gpmodel_deepcopy, "cluster_new_batch": _cluster_new_batch,
                "include_batch": _include_batch, "forward": _forward, "backward": _backward,
                "coupled_state_coef": _coupled_state_coef,
                "warp_batch_by_resp_amtgp_cached": _warp_batch_by_resp_amtgp_cached},
    "IterativeGaussianProcess": {"fit_torch": _fit_torch},
}


seam_times = {}        # "Class.method" -> [calls, seconds] (wall clock around the patched call, nested calls included)


def _timed(key, fn):
    import functools
    import time

    @functools.wraps(fn)
    def wrapper(*a, **k):
        t0 = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            rec = seam_times.setdefault(key, [0, 0.0])
            rec[0] += 1
            rec[1] += time.perf_counter() - t0
    return wrapper


def tune_host_allocator():
    """Keep freed host memory mapped (glibc mallopt: no trimming, no mmap per large block).  With the numerics on the
    device the reference's drivers spend their host time on short-lived megabyte tensors (the (N, K, K) respPair of every
    `variational_local_terms`, the (N, M) score planes): glibc hands such blocks back to the kernel on free and maps
    fresh pages for the next one, and a page fault inside a microVM costs microseconds -- `torch.full((500, 441))` was
    measured at 2.2 ms inside an online fit against 8 us in isolation.  Process-wide and optional; returns True if applied."""
    import ctypes
    try:
        libc = ctypes.CDLL("libc.so.6")
        M_TRIM_THRESHOLD, M_TOP_PAD, M_MMAP_THRESHOLD = -1, -2, -3
        ok = libc.mallopt(M_MMAP_THRESHOLD, 32 << 20) and libc.mallopt(M_TRIM_THRESHOLD, 1 << 30) and \
            libc.mallopt(M_TOP_PAD, 64 << 20)
        return bool(ok)
    except Exception:
        return False


_host_threads_saved = {}


def _limit_host_threads(n, blas=1):
    """With the numerics on the device, what is left on the host are the reference's own small tensor operations (LogLik
    over an (N, K, K) respPair, the L-BFGS of the HDP weights in scipy).  Sixteen-thread OpenMP and BLAS pools waking up
    for 200 k-element tensors -- and spinning against each other between calls -- cost milliseconds per operation:
    record 100 online, 500 beats: `torch.max` 19.5 s of a 93 s fit with the defaults, 0.23 s with the pools limited
    (fit 93 -> 50 s).  Limits torch's intra-op pool to `n` threads and the BLAS pools to one; `disable()` restores both."""
    import os
    if os.environ.get("HGP_HOST_THREADS"):              # A/B override: "0" = leave alone, "4" / "4,1" = torch[,blas]
        parts = os.environ["HGP_HOST_THREADS"].split(",")
        n = int(parts[0]) or None
        blas = int(parts[1]) if len(parts) > 1 else blas
    if n is None or _host_threads_saved:
        return
    _host_threads_saved["torch"] = torch.get_num_threads()
    torch.set_num_threads(max(1, min(int(n), _host_threads_saved["torch"])))
    if blas:
        try:
            import threadpoolctl
            _host_threads_saved["blas"] = threadpoolctl.threadpool_limits(int(blas), user_api="blas")
        except Exception:
            pass


def _restore_host_threads():
    if "torch" in _host_threads_saved:
        torch.set_num_threads(_host_threads_saved["torch"])
    lim = _host_threads_saved.get("blas")
    if lim is not None:
        try:
            lim.restore_original_limits()
        except Exception:
            pass
    _host_threads_saved.clear()


def enable(gpi_model_cls=None, gpi_hdp_cls=None, igp_cls=None, profile=False, host_threads=4):
    """Patch the reference classes (found in `hdpgpc.GPI_model` / `hdpgpc.GPI_HDP` / `hdpgpc.GPI` unless given).
    profile=True accumulates calls and wall seconds per seam method in `seam_times`.  host_threads: size of torch's
    intra-op CPU pool while the patch is active (BLAS pools go to one thread; None leaves both alone) -- see
    `_limit_host_threads`."""
    import importlib
    _limit_host_threads(host_threads)
    gpi_model_cls = gpi_model_cls or importlib.import_module("hdpgpc.GPI_model").GPI_model
    gpi_hdp_cls = gpi_hdp_cls or importlib.import_module("hdpgpc.GPI_HDP").GPI_HDP
    igp_cls = igp_cls or importlib.import_module("hdpgpc.GPI").IterativeGaussianProcess
    for cname, cls in (("GPI_model", gpi_model_cls), ("GPI_HDP", gpi_hdp_cls), ("IterativeGaussianProcess", igp_cls)):
        for name, fn in _PATCHES[cname].items():
            key = f"{cname}.{name}"
            if key not in _saved:
                _saved[key] = getattr(cls, name)
                _saved[key + "/cls"] = cls
            setattr(cls, name, _timed(key, fn) if profile else fn)
    return sorted(k for k in _saved if not k.endswith("/cls"))


def disable():
    """Restore the reference's own methods (models that carry a twin keep it; use to_host to convert them)."""
    for key in [k for k in _saved if not k.endswith("/cls")]:
        cls = _saved.pop(key + "/cls")
        setattr(cls, key.split(".", 1)[1], _saved.pop(key))
    _restore_host_threads()
