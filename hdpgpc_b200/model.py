"""Device-resident mirror of the reference's per-(cluster, lead) model `GPI_model`
(reference hdpgpc/GPI_model.py:16) for the E-step seam: same method names, argument meaning and
return types (torch float64 tensors), state held as struct-of-arrays on the GPU instead of Python
lists of T x T tensors.  Index logic (which historical state scores which beat) is host-side
integer work; all floating-point arithmetic runs in the CUDA library.
"""
from bisect import bisect_right

import numpy as np
import torch

from . import ops
from ._lib import HgpError

F64 = torch.float64


class LinAlgError(HgpError):
    """Raised where the reference would raise torch.linalg.LinAlgError (non-SPD covariance)."""


def _stack(lst, device):
    if isinstance(lst, torch.Tensor):
        return lst.to(device=device, dtype=F64).contiguous()
    if isinstance(lst, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(lst, dtype=np.float64)).to(device)
    arrs = [v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v) for v in lst]
    return torch.from_numpy(np.ascontiguousarray(np.stack(arrs), dtype=np.float64)).to(device)


def _vecs(lst, device):
    t = _stack(lst, device)
    if t.dim() == 3:  # reference carries (T, 1) column vectors
        t = t[:, :, 0].contiguous()
    return t


def state_index_map(indexes, n_samps, no_first=False):
    """i_vals / first_mask of GPI_model.compute_sq_err_all (GPI_model.py:497-513)."""
    idx = np.asarray(indexes, dtype=np.int64)
    pos = np.full(n_samps, -1, dtype=np.int64)
    pos[idx] = np.arange(idx.size)
    exact = pos >= 0
    closest = np.maximum(np.searchsorted(idx, np.arange(n_samps), side="right") - 1, 0)
    i_vals = np.where(exact, pos + 1, np.maximum(closest, 1))
    first = exact & (i_vals == 1) & (not no_first)
    return i_vals, first


def snr_state_index(indexes, n_states, n_samps):
    """j of GPI_HDP.compute_snr (GPI_HDP.py:739): clip(find_closest_lower(t), 1, len(f_star_sm) - 1)."""
    idx = np.asarray(indexes, dtype=np.int64)
    p = np.searchsorted(idx, np.arange(n_samps), side="right")
    j = np.where(p > 0, p - 1, 0)
    return np.minimum(np.maximum(j, 1), n_states - 1)


class _Appended:
    """history + one extra element, indexable like the list the reference builds with .copy() + .append()."""

    def __init__(self, hist, extra):
        self.hist, self.extra = hist, extra

    def __len__(self):
        return self.hist.shape[0] + 1

    def __getitem__(self, i):
        n = len(self)
        i = i + n if i < 0 else i
        return self.extra if i == n - 1 else self.hist[i]


class GPI_model:
    def __init__(self, x_basis, f_star, f_star_sm, C, Sigma, indexes, estimation_limit=None,
                 A=None, Gamma=None, cov_f_sm=None, cov_f=None, kernel=None, device="cuda"):
        ops._lib.require_cuda()          # fail loudly: there is no CPU path
        self.device = torch.device(device)
        self.x_basis = np.asarray(x_basis, dtype=np.float64).reshape(-1)
        self.T = self.x_basis.shape[0]
        self.f_star = _vecs(f_star, self.device)          # [nF, T] filtered means (observe uses these)
        self.f_star_sm = _vecs(f_star_sm, self.device)    # [nF, T] smoothed means (SNR, q_lat)
        self.C = _stack(C, self.device)                   # [nC, T, T]
        self.Sigma = _stack(Sigma, self.device)           # [nC, T, T]
        self.A = None if A is None else _stack(A, self.device)
        self.Gamma = None if Gamma is None else _stack(Gamma, self.device)
        self.cov_f_sm = None if cov_f_sm is None else _stack(cov_f_sm, self.device)
        self.cov_f = None if cov_f is None else _stack(cov_f, self.device)
        self.kernel = None if kernel is None else tuple(float(v) for v in kernel)   # (const, length, noise)
        self.indexes = [int(i) for i in indexes]
        self.N = len(self.indexes)
        self.estimation_limit = np.inf if estimation_limit is None else estimation_limit
        self._tables = None

    # ---- construction helpers ----
    @classmethod
    def from_reference(cls, gp, device="cuda"):
        """Device twin of a reference GPI_model (or any object with the same attributes): the histories, the prior
        defaults (A_def ...), the MNIW posteriors (internal_params / observation_params, GPI_model.py:171-175), the fitted
        kernel and the flags the chain needs.  Static models (Gamma == 0, inv_wishart observation prior) are refused."""
        if len(gp.Gamma) and not bool(torch.any(torch.as_tensor(np.asarray(gp.Gamma[-1])) != 0)):
            raise HgpError("static models (Gamma = 0) are outside the built path; no CPU fallback")
        kernel = getattr(gp, "kernel", None)
        noise_bounds = None
        if hasattr(gp, "gp") and hasattr(gp.gp, "kernel"):
            kp = gp.gp.kernel.get_params()
            kernel = (float(kp["k1__k1__constant_value"]), float(kp["k1__k2__length_scale"]), float(kp["k2__noise_level"]))
            noise_bounds = tuple(float(v) for v in gp.gp.kernel.k2.noise_level_bounds)
        self = cls(gp.x_basis, gp.f_star, gp.f_star_sm, gp.C, gp.Sigma, gp.indexes,
                   estimation_limit=getattr(gp, "estimation_limit", None), A=gp.A, Gamma=gp.Gamma,
                   cov_f_sm=gp.cov_f_sm, cov_f=gp.cov_f, kernel=kernel, device=device)
        up = lambda t: _stack([t], self.device)[0]
        if getattr(gp, "A_def", None) is not None:
            self.defaults = dict(A=up(gp.A_def), Gamma=up(gp.Gamma_def), C=up(gp.C_def), Sigma=up(gp.Sigma_def))
        for name, attr in (("internal", "internal_params"), ("observation", "observation_params")):
            p = getattr(gp, attr, None)
            if p is None:
                continue
            if not hasattr(p, "m_r_cov"):
                raise HgpError("inverse-Wishart (static) observation priors are outside the built path")
            setattr(self, name, dict(m_mean=up(p.m_mean).clone(), m_r_cov=up(p.m_r_cov).clone(), scale=up(p.scale).clone(),
                                     n0=torch.tensor([float(p.n0)], dtype=F64, device=self.device)))
        self.fitted = bool(getattr(gp, "fitted", True))
        self.annealing = bool(getattr(gp, "annealing", True))
        self.free_deg = float(getattr(gp, "free_deg_MNIV", 5.0))
        if noise_bounds is not None:
            self.noise_bounds = noise_bounds
        if self.N == 0 and hasattr(gp, "gp") and self.cov_f_sm is not None:
            # IterativeGaussianProcess.posterior (GPI.py:136) treats the first member specially when the prior covariance
            # IS the kernel matrix (torch.equal on the host's own matrices); the decision travels as a flag
            xb = np.asarray(torch.as_tensor(np.asarray(gp.x_basis)).detach().cpu(), dtype=np.float64).reshape(-1, 1)
            K = torch.from_numpy(np.asarray(gp.gp.kernel(xb, xb), dtype=np.float64))
            self.ini_cov_is_prior = bool(torch.equal(torch.as_tensor(np.asarray(gp.cov_f_sm[-1]), dtype=F64), K))
        return self

    @classmethod
    def from_dump(cls, z, prefix, device="cuda"):
        """From a tests/golden fixture written by generate_golden.dump_gp(full=True)."""
        g = lambda k: z[prefix + k]
        self = cls(g("x_basis"), g("f_star"), g("f_star_sm"), g("C"), g("Sigma"), g("indexes"),
                   estimation_limit=float(g("estimation_limit")), A=g("A"), Gamma=g("Gamma"),
                   cov_f_sm=g("cov_f_sm"), cov_f=g("cov_f"), kernel=g("kernel"), device=device)
        if prefix + "A_def" in z:
            self.defaults = {k: _stack(g(k + "_def")[None], self.device)[0] for k in ("A", "Gamma", "C", "Sigma")}
        if prefix + "int_m_mean" in z:      # MNIW posteriors: needed to continue the chain online
            for name, pre in (("internal", "int_"), ("observation", "obs_")):
                st = {k: _stack(g(pre + k)[None], self.device)[0].clone() for k in ("m_mean", "m_r_cov", "scale")}
                st["n0"] = torch.tensor([float(g(pre + "n0"))], dtype=F64, device=self.device)
                setattr(self, name, st)
            self.fitted = bool(g("fitted")) if prefix + "fitted" in z else True
            self.annealing = bool(z[prefix + "annealing"]) if prefix + "annealing" in z else True
            self.free_deg = float(z[prefix + "free_deg"]) if prefix + "free_deg" in z else 5.0
        return self

    # ---- trial copies and resets of the online driver ----
    def clone(self):
        """GPI_HDP.gpmodel_deepcopy (GPI_HDP.py:4037-4064): an independent copy of the device-resident state.  Like the
        reference (which copies the lists and shares the immutable tensors) the copy is O(1) in the history length: the
        histories are SHARED and become copy-on-write for both parties -- the only in-place writers are the online
        steps, which go through `_reserve`; dropping `_store` on both sides makes the next such step of either model move
        its histories to storage of its own first.  The small in-place state (MNIW posteriors, q_lat values) is copied;
        the score tables travel with the copy: they are functions of the state only."""
        new = object.__new__(type(self))
        shared = set(self._HIST + self._PAR)
        for k, v in self.__dict__.items():
            if k in ("_store", "_online_work"):
                continue
            if k in shared:
                new.__dict__[k] = v
            elif isinstance(v, torch.Tensor):
                new.__dict__[k] = v.clone()
            elif isinstance(v, dict) and k != "_tables":
                new.__dict__[k] = {kk: (vv.clone() if isinstance(vv, torch.Tensor) else vv) for kk, vv in v.items()}
            elif isinstance(v, list):
                new.__dict__[k] = list(v)
            else:
                new.__dict__[k] = v
        self._store = {}
        return new

    __deepcopy__ = lambda self, memo: self.clone()

    def reinit_GP(self, save_last=False, save_index=False):
        """GPI_model.reinit_GP (GPI_model.py:408-434), save_last=False: the initial state again, prior covariance of the
        fitted kernel (ini_cov_def), no members."""
        if save_last:
            raise HgpError("reinit_GP(save_last=True) is not built")
        if self.kernel is None or not getattr(self, "fitted", True):
            raise HgpError("reinit_GP needs the fitted kernel of the model (prior covariance)")
        xb = torch.from_numpy(self.x_basis).to(self.device)
        K = ops.rbf_kernel_matrix(xb, xb, self.kernel[0], self.kernel[1])
        self.f_star = self.f_star[:1].clone()
        self.f_star_sm = self.f_star.clone()
        self.cov_f, self.cov_f_sm = K[None].clone(), K[None].clone()
        self.ini_cov_is_prior = True
        self.indexes, self.N = [], 0
        self._store = {}
        self.invalidate_caches()

    def reinit_LDS(self, save_last=False, save_last_diag=False, return_likelihood=False):
        """GPI_model.reinit_LDS (:437-456), save_last=False: default parameters, fresh MNIW priors."""
        if save_last or return_likelihood:
            raise HgpError("reinit_LDS(save_last=True / return_likelihood=True) is not built")
        d = self._prior_defaults()
        for k in self._PAR:
            setattr(self, k, d[k][None].clone())
        for k in ("internal", "observation"):
            self.__dict__.pop(k, None)
        self._store = {}
        self.invalidate_caches()

    # ---- index rules (host integer work) ----
    def find_closest_lower(self, t):
        """GPI_model.find_closest_lower (GPI_model.py:584-593)."""
        idx = bisect_right(self.indexes, t)
        return idx - 1 if idx else 0

    def param_index(self, t):
        """Which (C, Sigma) a state index uses: GPI_model.observe (:626-662) + get_params (:664-669)."""
        nC = self.C.shape[0]
        nF = self.f_star.shape[0]
        if self.N == 0:
            return 0, 0
        if t < 0:
            t = nF + t
            return (t if t < nC else nC - 1), t   # python negative indexing of both lists
        if self.N <= t:
            return nC - 1, nF - 1
        if self.estimation_limit <= t:
            return nC - 1, t
        return (t if t < nC else nC - 1), t

    def tables(self):
        """Emission-mean table and whitening factors for every state (cached; states are immutable
        between chain updates).  Factor F-1 is the `first`-jitter variant of state 1 (GPI_model.py:527-529)."""
        if self._tables is not None:
            return self._tables
        nF = self.f_star.shape[0]
        cidx = np.zeros(nF, dtype=np.int32)
        fidx = np.zeros(nF, dtype=np.int32)
        for t in range(nF):
            c, f = self.param_index(t)
            cidx[t], fidx[t] = c, f
        dev = self.device
        mu = ops.emission_means(self.C, self.f_star, torch.from_numpy(cidx).to(dev), torch.from_numpy(fidx).to(dev))
        uniq, inv = np.unique(cidx, return_inverse=True)
        n_u = len(uniq)
        sig = self.Sigma.index_select(0, torch.from_numpy(uniq.astype(np.int64)).to(dev))
        add = torch.zeros(n_u + 1, dtype=F64, device=dev)
        first_state = min(1, nF - 1)
        sig = torch.cat([sig, self.Sigma[int(cidx[first_state])].unsqueeze(0)], dim=0)
        add[n_u] = 1e-2 * torch.mean(torch.diagonal(self.Sigma[0]))
        Lf, info = ops.chol_batched(sig, add_diag=add)
        W = ops.tri_inverse_batched(Lf)
        bad = torch.nonzero(info).flatten()
        if bad.numel():
            raise LinAlgError(f"linalg.cholesky: factor {int(bad[0])} is not positive-definite "
                              f"(leading minor {int(info[bad[0]])})")
        self._tables = dict(mu=mu, W=W, factor_of_state=inv.astype(np.int32), first_factor=n_u, cidx=cidx)
        return self._tables

    # ---- the seam methods ----
    def _off_grid(self, x_trains):
        """None when every grid equals x_basis (pred_dist short-circuit, GPI.py:467-468); otherwise the grids as a
        float64 array [n, nx] (n = 1 for a single / shared grid)."""
        if x_trains is None:
            return None
        x = x_trains.detach().cpu().numpy() if isinstance(x_trains, torch.Tensor) else np.asarray(x_trains)
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 3:
            x = x[:, :, 0]
        elif x.ndim == 2 and x.shape[1] == 1:
            x = x[:, 0][None, :]
        elif x.ndim == 1:
            x = x[None, :]
        if x.shape[1] == self.T and np.all(x == self.x_basis[None, :]):
            return None
        return np.ascontiguousarray(x)

    def _state_params(self):
        """(C index, f_star index) of every state, as GPI_model.observe (:626-662) picks them."""
        tb = self.tables()
        return tb["cidx"], tb["mu"]

    def observe(self, x_post, t, params=None, proj=False):
        """GPI_model.observe (GPI_model.py:626-662): emission distribution of state t resampled on x_post.
        Returns (mean (nx,), cov (nx, nx)) CUDA tensors."""
        if proj or params is not None:
            raise HgpError("observe(proj=True / params=...) is not built")
        c_i, f_i = self.param_index(int(t))
        grid = self._off_grid(x_post)
        tb = self.tables()
        if grid is None:
            return tb["mu"][f_i].clone(), self.Sigma[c_i].clone()
        if self.kernel is None:
            raise HgpError("observe on a grid other than x_basis needs the fitted kernel (const, length, noise)")
        dev = self.device
        one = lambda v: torch.tensor([v], dtype=torch.int32, device=dev)
        f, cov, info = ops.pred_dist_inducing(torch.from_numpy(self.x_basis).to(dev), torch.from_numpy(grid[0]).to(dev),
                                              tb["mu"], one(f_i), self.Sigma, one(c_i), self.kernel)
        if int(info[0]):
            raise LinAlgError("linalg.cholesky: K(x_basis, x_basis) + jitter is not positive-definite")
        return f[0], cov[0]

    def _score_off_grid(self, grids, Y, t_idx, first, chunk=256):
        """Score beats on grids that differ from x_basis: per item pred_dist (kernel matrices, Cholesky, projected
        covariance), then the Gaussian score under that item's own covariance (GPI_model.py:535-547 / :516-533 with a
        shared off-basis grid).  grids [1 | n, nx]; Y [n, nx]; t_idx, first: per-beat state index / first flag."""
        if self.kernel is None:
            raise HgpError("scoring on a grid other than x_basis needs the fitted kernel (const, length, noise)")
        dev = self.device
        tb = self.tables()
        n = Y.shape[0]
        pidx = np.array([self.param_index(int(t)) for t in t_idx], dtype=np.int64).reshape(n, 2)
        first = np.asarray(first, dtype=bool)
        shared = grids.shape[0] == 1
        out = torch.empty(n, dtype=F64, device=dev)
        xb = torch.from_numpy(self.x_basis).to(dev)
        jit1 = 1e-2 * torch.mean(torch.diagonal(self.Sigma[0]))
        i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)
        if shared:
            # one item per distinct (state, first) group, every beat of the group scored under it
            key = pidx[:, 0] * (2 * (pidx[:, 1].max() + 1)) + pidx[:, 1] * 2 + first
            uniq, rep, inv = np.unique(key, return_index=True, return_inverse=True)
            items = [(rep, inv)]
            xg = torch.from_numpy(grids[0]).to(dev)
        else:
            items = [(np.arange(s, min(s + chunk, n)), None) for s in range(0, n, chunk)]
        for rep, inv in items:
            xp = xg if shared else torch.from_numpy(grids[rep]).to(dev)
            f, cov, info = ops.pred_dist_inducing(xb, xp, tb["mu"], i32(pidx[rep, 1]), self.Sigma, i32(pidx[rep, 0]),
                                                  self.kernel)
            if bool(torch.any(info != 0)):
                raise LinAlgError("linalg.cholesky: K(x_basis, x_basis) + jitter is not positive-definite")
            add = torch.from_numpy(first[rep].astype(np.float64)).to(dev) * jit1
            Lf, info = ops.chol_batched(cov, add_diag=add)
            if bool(torch.any(info != 0)):
                raise LinAlgError("linalg.cholesky: projected emission covariance is not positive-definite")
            W = ops.tri_inverse_batched(Lf)
            k = len(rep)
            if shared:
                q = ops.score_pairs(Y, f, W, i32(inv.reshape(n, 1)), i32(np.arange(k)))
                out[:] = q[:, 0]
            else:
                q = ops.score_pairs(Y[rep[0]:rep[-1] + 1].contiguous(), f, W, i32(np.arange(k).reshape(k, 1)),
                                    i32(np.arange(k)))
                out[rep[0]:rep[-1] + 1] = q[:, 0]
        return out

    def _beats(self, y_trains):
        y = y_trains
        if not isinstance(y, torch.Tensor):
            y = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64))
        y = y.to(device=self.device, dtype=F64)
        if y.dim() == 3:
            y = y[:, :, 0]
        return y.contiguous()

    def compute_sq_err_all(self, x_trains, y_trains, no_first=False):
        """GPI_model.compute_sq_err_all (GPI_model.py:488-547): score of every beat under the
        time-indexed state of this cluster.  Returns (N,) float64 CUDA tensor."""
        Y = self._beats(y_trains)
        N = Y.shape[0]
        if self.N == 0:
            return torch.zeros(N, dtype=F64, device=self.device)
        grids = self._off_grid(x_trains)
        i_vals, first = state_index_map(self.indexes, N, no_first)
        if grids is not None:
            if grids.shape[0] not in (1, N):
                raise HgpError("compute_sq_err_all: one grid per beat expected")
            if grids.shape[0] == N and np.all(grids == grids[0:1]):
                grids = grids[0:1]          # the reference's shared-grid test (GPI_model.py:516)
            return self._score_off_grid(grids, Y, i_vals, first)
        tb = self.tables()
        # `first` beats use a duplicate of state 1 whose factor carries the extra jitter
        fos = np.concatenate([tb["factor_of_state"], [tb["first_factor"]]]).astype(np.int32)
        mu = torch.cat([tb["mu"], tb["mu"][1:2] if tb["mu"].shape[0] > 1 else tb["mu"][0:1]], dim=0)
        s_of = np.where(first, len(fos) - 1, i_vals).astype(np.int32).reshape(N, 1)
        dev = self.device
        q = ops.score_pairs(Y, mu, tb["W"], torch.from_numpy(s_of).to(dev), torch.from_numpy(fos).to(dev))
        return q[:, 0]

    def log_sq_error(self, x_train, y, mean=None, cov=None, C=None, Sigma=None, i=None, proj=False, first=False):
        """GPI_model.log_sq_error (GPI_model.py:250-286) for params=None; i=None / -1 = last state."""
        if proj:
            raise HgpError("log_sq_error(proj=True) is not built")
        grids = self._off_grid(x_train)
        if grids is not None:
            if mean is not None:
                raise HgpError("log_sq_error with explicit parameters on a grid other than x_basis is not built")
            Yg = self._beats(torch.as_tensor(np.asarray(y.detach().cpu() if isinstance(y, torch.Tensor) else y,
                                                        dtype=np.float64).reshape(1, -1)))
            nF = self.f_star.shape[0]
            t = -1 if i is None else i
            t = t if t >= 0 else nF + t
            return self._score_off_grid(grids[0:1], Yg, [t], [bool(first)])[0]
        if mean is not None:
            # explicit parameters (estimate_new path, :261-264 -> observe :657-660): mean' = C @ mean, cov = Sigma
            dev = self.device
            Yb = self._beats(torch.as_tensor(np.asarray(y.detach().cpu() if isinstance(y, torch.Tensor) else y,
                                                        dtype=np.float64).reshape(1, -1)))
            Cm = _stack(C, dev).reshape(1, self.T, self.T)
            mv = _vecs(mean, dev).reshape(1, self.T)
            zero = torch.zeros(1, dtype=torch.int32, device=dev)
            mu = ops.emission_means(Cm, mv, zero, zero)
            add = None
            if first:
                add = (1e-2 * torch.mean(torch.diagonal(self.Sigma[0]))).reshape(1)
            Lf, info = ops.chol_batched(_stack(Sigma, dev).reshape(1, self.T, self.T), add_diag=add)
            if int(info[0]):
                raise LinAlgError("linalg.cholesky: explicit Sigma is not positive-definite")
            W = ops.tri_inverse_batched(Lf)
            q = ops.score_pairs(Yb, mu, W, torch.zeros((1, 1), dtype=torch.int32, device=dev), zero)
            return q[0, 0]
        Y = self._beats(torch.as_tensor(np.asarray(y.detach().cpu() if isinstance(y, torch.Tensor) else y,
                                                   dtype=np.float64).reshape(1, -1)))
        tb = self.tables()
        nF = self.f_star.shape[0]
        if i is None:
            i = -1
        t = i if i >= 0 else nF + i
        t = min(t, nF - 1) if self.N > 0 else 0
        fos = np.concatenate([tb["factor_of_state"], [tb["first_factor"]]]).astype(np.int32)
        mu = tb["mu"]
        if first:
            # jitter is defined on Sigma[0]; the factor table carries it for state 1 only
            if t != min(1, nF - 1):
                raise HgpError("first=True is only defined for the first member state")
            mu = torch.cat([mu, mu[t:t + 1]], dim=0)
            s = len(fos) - 1
        else:
            s = t
        dev = self.device
        q = ops.score_pairs(Y, mu, tb["W"], torch.tensor([[s]], dtype=torch.int32, device=dev),
                            torch.from_numpy(fos).to(dev))
        return q[0, 0]

    def _qlat_dirty(self, first_member):
        """Members >= first_member score against inputs that have changed since the last compute_q_lat_all."""
        self._qlat_stable = max(0, min(getattr(self, "_qlat_stable", 0), int(first_member)))

    def invalidate_caches(self):
        """Call after editing the histories in place from outside the model's own methods."""
        self._tables = None
        self._qlat_stable = 0
        self._lds_lik = None

    def compute_q_lat_all(self, x_trains, h_ini=1.0):
        """GPI_model.compute_q_lat_all (GPI_model.py:549-559) -> log_lat_error (:288-323): per member j the
        latent transition score of its smoothed state under (A, Gamma) of step j+1; zero elsewhere.

        Incremental on the online path: member j reads the smoothed states j, j+1 and the parameter
        set min(j+1, last), so after include_weighted_sample / backwards_pair / bayesian_new_params only member 0 (it
        reads the LAST parameter set) and the trailing members see changed inputs.  The methods that rewrite histories
        mark the first stale member (`_qlat_dirty`); the others keep their stored value, which is what the reference's
        O(t) recomputation (GPI_HDP.py:1972) would return for them."""
        n = x_trains.shape[0] if hasattr(x_trains, "shape") else len(x_trains)
        out = torch.zeros(n, dtype=F64, device=self.device)
        if self.N == 0 or self.Gamma is None or not bool(torch.any(self.Gamma[-1] != 0)):
            return out
        nG = self.Gamma.shape[0]
        J = self.N
        dev = self.device
        vals = getattr(self, "_qlat_vals", None)
        stable = min(getattr(self, "_qlat_stable", 0), J) if vals is not None else 0
        if vals is None or vals.numel() < J:
            grown = torch.zeros(max(J, 2 * (vals.numel() if vals is not None else 0), 8), dtype=F64, device=dev)
            if stable:
                grown[:stable] = vals[:stable]
            vals = self._qlat_vals = grown
        todo = np.concatenate([[0], np.arange(max(1, stable), J)]).astype(np.int64)
        p_idx = np.where(todo == 0, 1, todo).astype(np.int32)
        a_idx = np.where((todo == 0) | (todo + 1 >= nG), nG - 1, todo + 1).astype(np.int32)
        fc_idx = (todo + 1).astype(np.int32)
        scale = np.where(todo == 0, float(h_ini), 1.0)
        t = lambda a, dt=torch.int32: torch.from_numpy(a).to(dev).to(dt)
        new, info = ops.qlat_batched(self.A, self.Gamma, self.cov_f_sm, self.f_star_sm, t(a_idx), t(a_idx), t(p_idx),
                                     t(p_idx), t(fc_idx), gamma_scale=t(scale, F64))
        bad = torch.nonzero(info).flatten()
        if bad.numel():
            self._qlat_stable = 0
            raise LinAlgError(f"linalg.cholesky: Gamma of member {int(todo[int(bad[0])])} is not positive-definite")
        vals[t(todo, torch.long)] = new
        self._qlat_stable = J
        out[torch.as_tensor(self.indexes, device=dev, dtype=torch.long)] = vals[:J]
        return out

    # ---- state export ------------------------------------------------------------------------------
    def to_reference_lists(self):
        """The device-resident state in the reference's own form: Python lists of CPU float64 tensors, (T, 1) column
        vectors for means and (T, T) matrices (GPI_model.py:35-47), plus `indexes` and `N` -- what `save_swgp`
        (GPI_HDP.py:3946-3950), `gpmodel_deepcopy` (:4037-4064) and the plotting helpers (util_plots.py:269-299) read."""
        col = lambda t: [t[i].detach().cpu().reshape(-1, 1).clone() for i in range(t.shape[0])]
        mat = lambda t: [] if t is None else [t[i].detach().cpu().clone() for i in range(t.shape[0])]
        return dict(f_star=col(self.f_star), f_star_sm=col(self.f_star_sm), cov_f=mat(self.cov_f),
                    cov_f_sm=mat(self.cov_f_sm), A=mat(self.A), Gamma=mat(self.Gamma), C=mat(self.C), Sigma=mat(self.Sigma),
                    indexes=list(self.indexes), N=self.N, x_basis=torch.from_numpy(self.x_basis.copy()).reshape(-1, 1),
                    kernel=self.kernel)

    def adopt_into(self, ref_gp):
        """Write the exported lists into a reference GPI_model object (attribute names are the reference's)."""
        for k, v in self.to_reference_lists().items():
            if k not in ("kernel", "x_basis"):
                setattr(ref_gp, k, v)
        return ref_gp

    # ---- ELBO term of the LDS parameters -----------------------------------------------------------
    def _prior_defaults(self):
        """(A_def, Gamma_def, C_def, Sigma_def): the prior the chain started from (GPI_model.py:466)."""
        d = getattr(self, "defaults", None)
        if d is None:
            d = dict(A=self.A[0], Gamma=self.Gamma[0], C=self.C[0], Sigma=self.Sigma[0])
        return d

    def return_LDS_param_likelihood(self, first=False):
        """GPI_model.return_LDS_param_likelihood (GPI_model.py:459-486), first=False.  The value depends on the last
        parameter set and the prior defaults only, so it is kept until a parameter update / re-initialisation (the
        drivers ask for it inside every compute_q_elbo, GPI_HDP.py:1838-1864, far more often than parameters change)."""
        if first:
            return lds_param_likelihood_batch([self], first=first)[0]
        if getattr(self, "_lds_lik", None) is None:
            self._lds_lik = lds_param_likelihood_batch([self])[0]
        return self._lds_lik

    # ---- online extras ------------------------------------------------------------------------------
    def posterior_weighted(self, x_train, y, h, t=None):
        """GPI_model.posterior_weighted (GPI_model.py:561-582), t=None: one Kalman update
        (IterativeGaussianProcess.posterior, GPI.py:72-151) from the last filtered state with Gamma/h, Sigma/h.
        Nothing is stored.  Returns (f (T,), cov (T,T))."""
        if t is not None:
            raise HgpError("posterior_weighted(t=...) is not built")
        if self._off_grid(x_train) is not None:
            raise HgpError("posterior_weighted on a grid other than x_basis is not built; no CPU fallback")
        if not h > 0.0:
            return self.f_star[-1].clone(), self.cov_f[-1].clone()
        T, dev = self.T, self.device
        Y = self._beats(torch.as_tensor(np.asarray(y.detach().cpu() if isinstance(y, torch.Tensor) else y,
                                                   dtype=np.float64).reshape(1, -1)))
        z = lambda *shape: torch.zeros(shape, dtype=F64, device=dev)
        hist = {k: z(2, T, T) for k in ("cov_f", "cov_f_sm", "A", "Gamma", "C", "Sigma")}
        f_star, f_star_sm = z(2, T), z(2, T)
        f_star[0] = f_star_sm[0] = self.f_star[-1]
        hist["cov_f"][0] = hist["cov_f_sm"][0] = self.cov_f[-1]
        hist["A"][0], hist["C"][0] = self.A[-1], self.C[-1]
        hist["Gamma"][0], hist["Sigma"][0] = self.Gamma[-1] / h, self.Sigma[-1] / h
        prior = self.N == 0 and self.kernel is not None
        r_first = 0.0
        if prior:
            c, ell, noise = self.kernel
            r_first = ((c + noise) - c) / h
        lib = ops._lib.load()
        scratch = lambda: z(T, T)
        desc = dict(n_members=1, first_is_prior=int(prior), annealing=0, estimation_limit=0, r_first=r_first,
                    member_beats=torch.zeros(1, dtype=torch.int32, device=dev), Y=Y, f_star=f_star, f_star_sm=f_star_sm,
                    **hist, int_m_mean=scratch(), int_m_r_cov=scratch(), int_scale=scratch(),
                    int_n0=torch.tensor([5.0], dtype=F64, device=dev), obs_m_mean=scratch(), obs_m_r_cov=scratch(),
                    obs_scale=scratch(), obs_n0=torch.tensor([5.0], dtype=F64, device=dev),
                    work=z(int(lib.hgp_chain_work_doubles(T))), piv=torch.zeros(T, dtype=torch.int32, device=dev),
                    status=torch.zeros(2, dtype=torch.int32, device=dev))
        ops.chain_run([desc], T)
        return f_star[1], hist["cov_f"][1]

    def smoother_weighted(self, x_train, y, h):
        """GPI_model.smoother_weighted (:726-738): histories with the would-be posterior appended.  Returns views
        that support [-1] / len() like the reference's lists without copying the histories."""
        f, cov = self.posterior_weighted(x_train, y, h)
        return (_Appended(self.f_star, f), _Appended(self.cov_f, cov), _Appended(self.C, self.C[-1]),
                _Appended(self.Sigma, self.Sigma[-1]))

    def estimate_new(self, x_train, y, h=1.0):
        """GPI_HDP.estimate_new (GPI_HDP.py:2830-2842): score of y as if it had already been absorbed."""
        mean_, cov_, C_, Sigma_ = self.smoother_weighted(x_train, y, h)
        return self.log_sq_error(x_train, y, mean=mean_[-1], cov=cov_[-1], C=C_[-1], Sigma=Sigma_[-1], i=-1,
                                 first=(len(self.indexes) == 1))

    # ---- online assimilation: the reference's three seam calls, one by one -----------------------------------------
    _HIST = ("f_star", "f_star_sm", "cov_f", "cov_f_sm")
    _PAR = ("A", "Gamma", "C", "Sigma")

    def _reserve(self, n_states, n_params):
        """Make room for n_states states / n_params parameter sets: the histories are views of storage tensors that
        grow geometrically, so an online append does not copy the chain (reference: Python list append)."""
        store = getattr(self, "_store", None)
        if store is None:
            store = self._store = {}
        for names, need in ((self._HIST, n_states), (self._PAR, n_params)):
            for k in names:
                cur = getattr(self, k)
                st = store.get(k)
                shares = st is not None and st.data_ptr() == cur.data_ptr() and st.shape[0] >= cur.shape[0]
                if not shares or st.shape[0] < need:
                    cap = max(need, 2 * cur.shape[0], 8)
                    st = torch.zeros((cap,) + tuple(cur.shape[1:]), dtype=F64, device=self.device)
                    st[:cur.shape[0]] = cur
                    store[k] = st
                    setattr(self, k, st[:cur.shape[0]])

    def _online_desc(self, y_row, phases, start_members):
        T, dev = self.T, self.device
        if not hasattr(self, "internal"):
            raise HgpError("online assimilation needs the MNIW state of the chain (a model built by full_pass_weighted, "
                           "GPI_model.fresh / unfitted, or from_dump of a full dump)")
        lib = ops._lib.load()
        st = self._store
        c, ell, noise = self.kernel
        if getattr(self, "_online_work", None) is None:
            self._online_work = (torch.zeros(int(lib.hgp_chain_work_doubles(T)), dtype=F64, device=dev),
                                 torch.zeros(T, dtype=torch.int32, device=dev))
        work, piv = self._online_work
        return dict(n_members=1, first_is_prior=int(start_members == 0 and getattr(self, "ini_cov_is_prior", False)),
                    annealing=int(getattr(self, "annealing", True)),
                    estimation_limit=0 if np.isinf(self.estimation_limit) else int(self.estimation_limit),
                    r_first=(c + noise) - c, member_beats=torch.zeros(1, dtype=torch.int32, device=dev), Y=y_row,
                    **{k: st[k] for k in self._HIST + self._PAR},
                    **{"int_" + k: v for k, v in self.internal.items()}, **{"obs_" + k: v for k, v in self.observation.items()},
                    work=work, piv=piv, status=torch.zeros(2, dtype=torch.int32, device=dev),
                    start_members=start_members, start_params=self.A.shape[0] - 1, phases=phases)

    def _init_mniw(self):
        if not hasattr(self, "internal"):
            eye = torch.eye(self.T, dtype=F64, device=self.device)
            n0 = lambda: torch.tensor([getattr(self, "free_deg", 5.0)], dtype=F64, device=self.device)
            self.internal = dict(m_mean=self.A[0].clone(), m_r_cov=eye.clone(), scale=self.Gamma[0].clone(), n0=n0())
            self.observation = dict(m_mean=self.C[0].clone(), m_r_cov=eye.clone(), scale=self.Sigma[0].clone(), n0=n0())

    def include_weighted_sample(self, index, x_train, x_warped, y, h, snr=None):
        """GPI_model.include_weighted_sample (GPI_model.py:353-375) -> include_sample (:325-351) ->
        IterativeGaussianProcess.posterior (GPI.py:72-151): one Kalman update in Joseph form from the last SMOOTHED state
        with the last parameter set; the new state is appended to the filtered and smoothed histories.  The first beat
        of an unfitted model runs the hyper-fit first (:361-365).  Returns x_basis like the reference."""
        if snr is not None:
            raise HgpError("include_weighted_sample(snr=...) is not built")
        if self._off_grid(x_train) is not None:
            raise HgpError("include_weighted_sample on a grid other than x_basis is not built; no CPU fallback")
        if h != 1.0:
            return self.x_basis          # include_sample(posterior=False): nothing is stored (:343-351)
        Y = self._beats(torch.as_tensor(np.asarray(y.detach().cpu() if isinstance(y, torch.Tensor) else y,
                                                   dtype=np.float64).reshape(1, -1)))
        if self.N == 0 and not getattr(self, "fitted", True):
            self.fit_kernel_params(None, Y[0])
        self._init_mniw()
        nF = self.f_star.shape[0]
        self._reserve(nF + 1, self.A.shape[0] + 1)
        desc = self._online_desc(Y, 1, self.N)
        ops.chain_run([desc], self.T)
        for k in self._HIST:
            setattr(self, k, self._store[k][:nF + 1])
        self.indexes.append(int(index))
        self.N += 1
        self._last_y = Y
        self._tables = None
        return self.x_basis

    def backwards_pair(self, h, snr=None):
        """GPI_model.backwards_pair (:705-724) -> backward_notrange (GPI.py:272-300): RTS smoothing of the last two
        states; rewrites f_star_sm[-2:], cov_f_sm[-2:]."""
        if snr is not None:
            raise HgpError("backwards_pair(snr=...) is not built")
        if len(self.indexes) > 1 and h == 1.0:
            self._reserve(self.f_star.shape[0], self.A.shape[0])     # own storage first (histories may be shared)
            self._init_mniw()
            ops.chain_run([self._online_desc(self._last_y, 2, self.N - 1)], self.T)
            self._tables = None
            self._qlat_dirty(self.N - 2)       # smoothed states N-1 and N were rewritten

    def bayesian_new_params(self, h, model_type="dynamic", full_data=False, q=None, force=False, snr=1.0):
        """GPI_model.bayesian_new_params (:966-1115), 1-step dynamic form: MNIW update of (A, Gamma) from the last two
        smoothed means and of (C, Sigma) from (y, last smoothed mean) when 1 < N < estimation_limit, then a new
        parameter set (with the annealing terms) is appended while N < estimation_limit."""
        if h != 1.0:
            return                       # the reference does nothing for a beat the cluster did not take (:972)
        if full_data or force or snr != 1.0 or model_type != "dynamic":
            raise HgpError("bayesian_new_params: only the 1-step dynamic update is built")
        nC = self.A.shape[0]
        self._reserve(self.f_star.shape[0], nC + 1)
        self._init_mniw()
        desc = self._online_desc(self._last_y, 4, self.N - 1)
        ops.chain_run([desc], self.T)
        fail, n_par = (int(v) for v in desc["status"])
        if fail:
            # the reference keeps the previous parameters and prints (GPI_model.py:1068-1071); the kernel already did
            print("Alg error matrix ill conditioned.")
        for k in self._PAR:
            setattr(self, k, self._store[k][:n_par])
        self._tables = None
        self._lds_lik = None
        self._qlat_dirty(nC - 2)               # members that read "the last parameter set" now read another one

    # ---- chain replay -------------------------------------------------------------------------------
    @classmethod
    def fresh(cls, x_basis, kernel, ini_sigma, ini_gamma, free_deg=5, estimation_limit=None, annealing=True,
              device="cuda"):
        """A fitted, empty dynamic model: GPI_HDP.create_gp_default (GPI_HDP.py:496-535) + GPI_model.GPR_dynamic /
        initial_conditions (GPI_model.py:115-205) + fit_kernel_params' side effects (:207-241) for the given
        fitted kernel (const, length, noise) of ConstantKernel*RBF + WhiteKernel."""
        x = np.asarray(x_basis, dtype=np.float64).reshape(-1)
        T = x.shape[0]
        I = np.eye(T)
        self = cls(x, [np.zeros(T)], [np.zeros(T)], [I], [ini_sigma * I], [], estimation_limit=estimation_limit,
                   A=[I], Gamma=[ini_gamma * I], cov_f_sm=[I], device=device)
        self.free_deg = float(free_deg)
        self.annealing = annealing
        self._set_kernel(kernel)
        return self

    @classmethod
    def unfitted(cls, x_basis, noise_bounds, ini_sigma, ini_gamma, free_deg=5, estimation_limit=None, annealing=True,
                 device="cuda"):
        """A default model whose kernel is still to be fitted: full_pass_weighted runs the one-beat hyper-fit on its
        first member (GPI_model.include_weighted_sample :353-375 -> fit_kernel_params :207-241 -> fit_torch)."""
        self = cls.fresh(x_basis, (1.0, 1.2, float(noise_bounds[0])), ini_sigma, ini_gamma, free_deg, estimation_limit,
                         annealing, device)
        self.fitted = False
        self.noise_bounds = (float(noise_bounds[0]), float(noise_bounds[1]))
        return self

    def _set_kernel(self, kernel):
        """fit_kernel_params' side effects (:228-231): prior covariance = kernel(x_basis, x_basis) without the white part."""
        c, ell, noise = (float(v) for v in kernel)
        self.kernel = (c, ell, noise)
        xb = torch.from_numpy(self.x_basis).to(self.device)
        K = ops.rbf_kernel_matrix(xb, xb, c, ell)
        self.cov_f = K[None].clone()
        self.cov_f_sm = K[None].clone()
        self.ini_cov_is_prior = True
        self.fitted = True
        self.invalidate_caches()

    def fit_kernel_params(self, x_train, y):
        """GPI_model.fit_kernel_params (:207-241): hyper-fit on one beat (device, see csrc/hgp_hyperfit.cu); keeps the
        fitted outputscale and noise, sets the RBF lengthscale to 1.2 (GPI.py:711)."""
        if self._off_grid(x_train) is not None:
            raise HgpError("fit_kernel_params on a grid other than x_basis (ProjectedGPModel branch) is not built")
        Y = self._beats(torch.as_tensor(np.asarray(y.detach().cpu() if isinstance(y, torch.Tensor) else y,
                                                   dtype=np.float64).reshape(1, -1)))
        out = ops.hyperfit_batched(torch.from_numpy(self.x_basis).to(self.device), Y, self.noise_bounds)[0].cpu().numpy()
        if int(out[6]):
            raise LinAlgError("hyper-fit: kernel matrix lost positive-definiteness")
        self.hyperfit_iterations = int(out[5])
        self._set_kernel((out[0], 1.2, out[2]))
        return self.kernel

    def _chain_prepare(self, Y, resp):
        """Allocate the histories of a fresh chain and fill its hgp_chain_desc (None if no member)."""
        if self.N != 0 or not hasattr(self, "fitted"):
            raise HgpError("device full_pass_weighted replays a chain from an empty model (GPI_model.fresh / unfitted)")
        r = resp if isinstance(resp, torch.Tensor) else torch.from_numpy(np.asarray(resp, dtype=np.float64))
        active = torch.nonzero(r.to(self.device) > 0.99).flatten()
        n = int(active.numel())
        if n == 0:
            return None
        if not self.fitted:
            self.fit_kernel_params(None, Y[int(active[0])])
        T, dev = self.T, self.device
        z = lambda *shape: torch.zeros(shape, dtype=F64, device=dev)
        hist = {k: z(n + 1, T, T) for k in ("cov_f", "cov_f_sm", "A", "Gamma", "C", "Sigma")}
        f_star, f_star_sm = z(n + 1, T), z(n + 1, T)
        f_star[0], f_star_sm[0] = self.f_star[0], self.f_star_sm[0]
        hist["cov_f"][0], hist["cov_f_sm"][0] = self.cov_f[0], self.cov_f_sm[0]
        hist["A"][0], hist["Gamma"][0], hist["C"][0], hist["Sigma"][0] = self.A[0], self.Gamma[0], self.C[0], self.Sigma[0]
        eye = torch.eye(T, dtype=F64, device=dev)
        c, ell, noise = self.kernel
        lib = ops._lib.load()
        return dict(n_members=n, first_is_prior=1, annealing=int(self.annealing),
                    estimation_limit=0 if np.isinf(self.estimation_limit) else int(self.estimation_limit),
                    r_first=(c + noise) - c, member_beats=active.to(torch.int32).contiguous(), Y=Y,
                    f_star=f_star, f_star_sm=f_star_sm, **hist,
                    int_m_mean=self.A[0].clone(), int_m_r_cov=eye.clone(), int_scale=self.Gamma[0].clone(),
                    int_n0=torch.tensor([self.free_deg], dtype=F64, device=dev),
                    obs_m_mean=self.C[0].clone(), obs_m_r_cov=eye.clone(), obs_scale=self.Sigma[0].clone(),
                    obs_n0=torch.tensor([self.free_deg], dtype=F64, device=dev),
                    work=z(int(lib.hgp_chain_work_doubles(T))), piv=torch.zeros(T, dtype=torch.int32, device=dev),
                    status=torch.zeros(2, dtype=torch.int32, device=dev),
                    rts_cache=torch.empty(max(1, int(lib.hgp_chain_rts_cache_doubles(T, n + 1))), dtype=F64, device=dev))

    def _chain_finish(self, desc):
        fail, n_par = (int(v) for v in desc["status"])
        if fail:
            # torch.linalg.LinAlgError inside bayesian_new_params is caught by the reference, which prints and keeps the
            # previous MNIW posteriors for that member (GPI_model.py:1068-1071); the kernel has done exactly that
            print("Alg error matrix ill conditioned.")
        self.mniw_first_failed_member = fail - 1 if fail else None
        self.f_star, self.f_star_sm = desc["f_star"], desc["f_star_sm"]
        self.cov_f, self.cov_f_sm = desc["cov_f"], desc["cov_f_sm"]
        self.A, self.Gamma, self.C, self.Sigma = (desc[k][:n_par] for k in ("A", "Gamma", "C", "Sigma"))
        self.internal = {k[4:]: desc[k] for k in ("int_m_mean", "int_m_r_cov", "int_scale", "int_n0")}
        self.observation = {k[4:]: desc[k] for k in ("obs_m_mean", "obs_m_r_cov", "obs_scale", "obs_n0")}
        self.indexes = [int(i) for i in desc["member_beats"].cpu()]
        self.N = desc["n_members"]
        self._last_y = desc["Y"][self.indexes[-1]].reshape(1, -1).clone()      # y_train[-1] of the reference (:1029)
        self._store = {}
        self.invalidate_caches()

    def full_pass_weighted(self, x_trains, y_trains, resp, q=None, q_lat=None, snr=None):
        """GPI_model.full_pass_weighted (GPI_model.py:377-406) for a fresh fitted dynamic model: assimilate the
        beats with resp > 0.99 in time order (Kalman + pair smoother + MNIW per member), full RTS pass, then
        (q, q_lat) over all beats.  One persistent CTA runs the whole chain on the device."""
        if self._off_grid(x_trains) is not None:
            raise HgpError("full_pass_weighted on a grid other than x_basis is not built; no CPU fallback")
        Y = self._beats(y_trains)
        desc = self._chain_prepare(Y, resp)
        if desc is None:
            return q, q_lat
        ops.chain_run([desc], self.T)
        self._chain_finish(desc)
        return self.compute_sq_err_all(x_trains, y_trains), self.compute_q_lat_all(Y)


def full_pass_weighted_batch(models, Y_planes, resp):
    """All (cluster, lead) chains of one sweep in ONE launch (one CTA per chain): models[ld][m] fresh fitted
    GPI_model objects, Y_planes [L, N, T], resp [N, M] one-hot.  The reference runs them one after the other
    (estimate_q_all, GPI_HDP.py:2879-2907).  Returns (q [N, M, L], q_lat [N, M, L])."""
    L, N, T = Y_planes.shape
    M = len(models[0])
    descs, owners = [], []
    for ld in range(L):
        for m in range(M):
            d = models[ld][m]._chain_prepare(Y_planes[ld], resp[:, m])
            if d is not None:
                descs.append(d)
                owners.append((ld, m))
    if descs:
        ops.chain_run(descs, T)
    dev = Y_planes.device
    q = torch.zeros((N, M, L), dtype=F64, device=dev)
    q_lat = torch.zeros((N, M, L), dtype=F64, device=dev)
    for d, (ld, m) in zip(descs, owners):
        gp = models[ld][m]
        gp._chain_finish(d)
        q[:, m, ld] = gp.compute_sq_err_all(None, Y_planes[ld])
        q_lat[:, m, ld] = gp.compute_q_lat_all(Y_planes[ld])
    return q, q_lat


def lds_param_likelihood_batch(models, first=False):
    """GPI_model.return_LDS_param_likelihood (GPI_model.py:459-486) for a list of models in one batched call:
    MNIW log-likelihood of (A[-1], Gamma[-1]) under MNIW(A_def, I, Gamma_def) (skipped when Gamma_def is all zero) plus
    that of (C[-1], Sigma[-1]) under MNIW(C_def, I, Sigma_def), times 100 / T.  Returns a float64 CUDA tensor [len]."""
    if first:
        raise HgpError("return_LDS_param_likelihood(first=True) is not built (unused by the reference's drivers)")
    if not models:
        return torch.zeros(0, dtype=F64)
    dev, T = models[0].device, models[0].T
    mats, owner, use = [], [], []
    eye = torch.eye(T, dtype=F64, device=dev)
    for k, gp in enumerate(models):
        d = gp._prior_defaults()
        pairs = [(gp.C[-1], gp.Sigma[-1], d["C"], d["Sigma"])]
        if bool(torch.any(d["Gamma"] != 0)):
            pairs.append((gp.A[-1], gp.Gamma[-1], d["A"], d["Gamma"]))
        for M_, S_, pm, ps in pairs:
            mats.append((M_, S_, pm, ps))
            owner.append(k)
    J = len(mats)
    stack = lambda i: torch.stack([m[i] for m in mats], dim=0)
    ar = torch.arange(J, dtype=torch.int32, device=dev)
    zero = torch.zeros(J, dtype=torch.int32, device=dev)
    vals, info = ops.mniw_loglik_batched(stack(0), ar, stack(1), ar, stack(2), ar, eye[None], zero, stack(3), ar)
    bad = torch.nonzero(info).flatten()
    if bad.numel():
        raise LinAlgError(f"linalg.cholesky: LDS noise covariance of model {owner[int(bad[0])]} is not positive-definite")
    out = torch.zeros(len(models), dtype=F64, device=dev)
    out.index_add_(0, torch.as_tensor(owner, device=dev, dtype=torch.long), vals)
    return out / T * 100
