"""Device mirror of the reference's alignment classes (hdpgpc/amtgp_warping_system.py): `WarpPriorAMTGP` (:106-264)
and `Warping_system` (:266-736), same method names / argument meaning / return shapes, arithmetic in the CUDA library
(one warp per fit, closed-form gradient, see csrc/hgp_warp.cu).  D = 1 (the reference always warps one lead at a
time: GPI_HDP.py:3480).  No CPU fallback."""
import numpy as np
import torch

from . import ops
from ._lib import HgpError

F64 = torch.float64


def _dev_tensor(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=F64)
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to(device)


def _as_1d(x, device):
    return _dev_tensor(x, device).reshape(-1).contiguous()


class WarpPriorAMTGP:
    """GP prior on warps: full log density (quad + logdet + const), Cholesky cached per (grid, theta, noise)."""

    def __init__(self, noise_warp, bound_noise_warp=(1e-8, 1e2), jitter=1e-6, default_rho=1.0, default_omega=1.0,
                 normalize_x=True, device="cuda"):
        self.noise_warp = float(noise_warp)
        self.noise_bounds = tuple(float(b) for b in bound_noise_warp)
        self.jitter = float(jitter)
        self.default_rho, self.default_omega = float(default_rho), float(default_omega)
        self.normalize_x = bool(normalize_x)
        self.theta = None
        self.device = torch.device(device)
        self._cache_key = None
        self._cache = None

    def _parse_theta(self):
        """(:140-153): only tuple/list/dict thetas override the defaults (a float theta is ignored)."""
        rho, omega = self.default_rho, self.default_omega
        th = self.theta
        try:
            if isinstance(th, (tuple, list)) and len(th) >= 2:
                rho, omega = float(th[0]), float(th[1])
            elif isinstance(th, dict):
                rho, omega = float(th.get("rho", rho)), float(th.get("omega", omega))
        except Exception:
            pass
        return max(rho, 1e-12), max(omega, 1e-12)

    def _clamped_noise(self):
        lo, hi = self.noise_bounds
        return min(max(self.noise_warp, lo), hi)

    def _ensure_cache(self, x):
        rho, omega = self._parse_theta()
        n2 = self._clamped_noise()
        key = (int(x.numel()), rho, omega, n2, float(x[0]), float(x[-1]), self.normalize_x)
        if key != self._cache_key:
            self._cache = ops.warp_prior_factor(x, rho, omega, n2 + self.jitter, self.normalize_x)
            self._cache_key = key
        return self._cache

    def log_sq_error_batch(self, x_model, x_warp_batch):
        x = _as_1d(x_model, self.device)
        W = x_warp_batch
        if isinstance(W, list):
            W = torch.stack([_dev_tensor(w, self.device).reshape(-1) for w in W], dim=0)
        W = _dev_tensor(W, self.device)
        if W.dim() == 3 and W.shape[-1] == 1:
            W = W[..., 0]
        if W.shape[0] == x.numel() and W.shape[1] != x.numel():
            W = W.transpose(0, 1)
        if W.dim() != 2 or W.shape[1] != x.numel():
            raise HgpError(f"Expected (B,T), got {tuple(W.shape)}")
        fac, logdet = self._ensure_cache(x)
        return ops.warp_prior_score(fac, logdet, W.contiguous())

    def log_sq_error(self, x_model, x_warp):
        return self.log_sq_error_batch(x_model, _dev_tensor(x_warp, self.device).reshape(1, -1))[0]


class Warping_system:
    def __init__(self, x_basis_warp, noise_warp=1e-2, bound_noise_warp=(1e-6, 1e2), recursive=True, cuda=True,
                 bayesian=True, mode="balanced", n_ctrl=8, lr=5e-2, lambda_smooth=200.0, lambda_amp=1e-3, device="cuda"):
        self.device = torch.device(device)
        self.x_basis = _as_1d(x_basis_warp, self.device)
        self.T = self.x_basis.numel()
        self.noise_warp_default = float(noise_warp)
        self.noise_bounds = tuple(float(b) for b in bound_noise_warp)
        self.recursive, self.bayesian, self.mode = bool(recursive), bool(bayesian), str(mode)
        self.n_ctrl = int(max(4, min(n_ctrl, self.T)))
        self.lr = float(lr)
        self.lambda_smooth_base, self.lambda_amp_base = float(lambda_smooth), float(lambda_amp)
        self._u_ctrl_prev = None
        self.warp_gp = WarpPriorAMTGP(noise_warp, bound_noise_warp, device=device)

    def _theta_to_lambdas(self, theta):
        """(:376-405)."""
        lam_s, lam_a = self.lambda_smooth_base, self.lambda_amp_base
        try:
            if isinstance(theta, (tuple, list)) and len(theta) >= 2:
                rho, omg = float(theta[0]), float(theta[1])
            elif isinstance(theta, dict):
                rho, omg = float(theta.get("rho", 1.0)), float(theta.get("omega", 1.0))
            else:
                return lam_s, lam_a
            return lam_s / (rho * rho + 1e-12), lam_a / (omg * omg + 1e-12)
        except Exception:
            return lam_s, lam_a

    def _regrid(self, x_model):
        """(:597-607): a call on another grid length re-bases the object and resets its prior."""
        T = x_model.numel()
        if T != self.T:
            self.x_basis, self.T = x_model, T
            self.n_ctrl = int(max(4, min(self.n_ctrl, T)))
            self.warp_gp = WarpPriorAMTGP(self.noise_warp_default, self.noise_bounds, device=self.device)

    def _noise_scalar(self, noise):
        """(:612-619)."""
        if noise is None:
            return self.noise_warp_default
        nz = _dev_tensor(noise, self.device)
        n = float(nz.mean()) if nz.numel() > 1 else float(nz.reshape(()))
        lo, hi = self.noise_bounds
        return min(max(n, lo), hi)

    def fit(self, x_model, Y, y_model, theta=None, noise=None, train_iter=50, grad_scale=None, u0=None, want_trace=False):
        """All (beat, template) fits in one launch.  Y [N, T]; y_model [R, T].  Returns x_warp, y_warp [R, N, T], u, trace."""
        x = _as_1d(x_model, self.device)
        self._regrid(x)
        self.warp_gp.theta = theta
        lam_s, lam_a = self._theta_to_lambdas(theta)
        return ops.warp_fit_batched(x, Y, y_model, self._noise_scalar(noise), lam_s, lam_a, self.n_ctrl, self.lr,
                                    train_iter, u0=u0, grad_scale=grad_scale, want_u=True, want_trace=want_trace)

    def compute_warp_batch(self, x_model, y_target_batch, y_model, theta=None, noise=None, weights=None, visualize=False,
                           verbose=False, train_iter=50):
        """Warping_system.compute_warp_batch (:548-736).  Returns (x_warp (B,T,1), y_warp (B,T,1), lik (B,), trace)."""
        x = _as_1d(x_model, self.device)
        T = x.numel()
        Yt = _dev_tensor(y_target_batch, self.device)
        if Yt.dim() == 1:
            Yt = Yt[None, :, None]
        elif Yt.dim() == 2:
            Yt = Yt[:, :, None]
        if Yt.shape[1] != T:
            raise HgpError(f"y_target_batch length mismatch: got {Yt.shape[1]} expected {T}")
        if Yt.shape[2] != 1:
            raise HgpError("compute_warp_batch: only D = 1 is built (the reference warps one lead at a time)")
        B = Yt.shape[0]
        Ym = _dev_tensor(y_model, self.device)
        if Ym.dim() == 3:
            raise HgpError("compute_warp_batch: per-beat templates are not built")
        Ym = Ym.reshape(1, T)
        if weights is None:
            wgt = torch.ones(B, dtype=F64, device=self.device)
        else:
            wgt = torch.clamp(_dev_tensor(weights, self.device).reshape(-1), min=0.0)
        scale = wgt / (torch.sum(wgt) + 1e-12)
        u0 = None
        if self.recursive and self._u_ctrl_prev is not None and self._u_ctrl_prev.numel() == self.n_ctrl:
            u0 = self._u_ctrl_prev.reshape(1, -1)
        xw, yw, u, tr = self.fit(x, Yt[:, :, 0].contiguous(), Ym, theta=theta, noise=noise, train_iter=train_iter,
                                 grad_scale=scale, u0=u0, want_trace=True)
        lik = self.warp_gp.log_sq_error_batch(x, xw[0])
        if self.recursive:
            self._u_ctrl_prev = torch.mean(u[0], dim=0)
        trace = {"loss": [float(v) for v in (tr[:, 0, :] @ scale).cpu()]}
        return xw[0][:, :, None], yw[0][:, :, None], lik, trace


def warp_batch_by_resp(x_model, Y, ref_index, warpers, base_warper, theta, noise, train_iter=50, batch_size=128):
    """One lead of GPI_HDP.warp_batch_by_resp_amtgp_cached (GPI_HDP.py:3412-3517) without the cache bookkeeping:
    every beat against every cluster's representative beat Y[ref_index[m]], all N x M fits in ONE launch.  The
    reference fits chunks of `batch_size` beats and optimises each chunk's mean loss, so beat n carries the gradient
    scale 1 / (size of its chunk + 1e-12).  Returns y_w, x_w [N, T, M] and liks [N, M] (fitted warper's prior score +
    the base warper's, :3494-3495).  warpers[m] must share grid / hyper-parameters apart from their priors
    (recursive_warp=False, the only mode the reference's drivers use)."""
    dev = Y.device
    N, T = Y.shape
    M = len(ref_index)
    if any(w.recursive for w in warpers[:M]):
        raise HgpError("warp_batch_by_resp: recursive warm start couples the chunks; use compute_warp_batch per chunk")
    sizes = np.minimum(batch_size, N - (np.arange(N) // batch_size) * batch_size).astype(np.float64)
    scale = torch.from_numpy(1.0 / (sizes + 1e-12)).to(dev)
    Ym = Y[torch.as_tensor(np.asarray(ref_index, dtype=np.int64), device=dev)]
    w0 = warpers[0]
    xw, yw, _, _ = w0.fit(x_model, Y, Ym, theta=theta, noise=noise, train_iter=train_iter, grad_scale=scale)
    x = _as_1d(x_model, dev)
    liks = torch.empty((N, M), dtype=F64, device=dev)
    for m in range(M):
        w = warpers[min(m, len(warpers) - 1)]
        w._regrid(x)
        w.warp_gp.theta = theta
        liks[:, m] = w.warp_gp.log_sq_error_batch(x, xw[m]) + base_warper.warp_gp.log_sq_error_batch(x, xw[m])
    return yw.permute(1, 2, 0), xw.permute(1, 2, 0), liks
