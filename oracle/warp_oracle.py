"""CPU restatement (numpy) of the reference's batched alignment ("warp") path.

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg as the
checker; the product package never imports it.

Restates, with closed-form gradients instead of autograd:
  Warping_system.compute_warp_batch        reference hdpgpc/amtgp_warping_system.py:548-736
  WarpPriorAMTGP.log_sq_error_batch        :223-264   (_rbf_cov :160-173, _ensure_cache :175-194)
  GPI_HDP.warp_batch_by_resp_amtgp_cached  hdpgpc/GPI_HDP.py:3412-3517 (chunks of 128 beats, fit + prior + base prior)
Pinned against the unmodified reference by tests/golden/warp_rec102_T90.npz (tests/test_oracle_vs_golden.py).
"""
import math

import numpy as np

ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-8     # torch.optim.Adam defaults (amtgp_warping_system.py:639)


def ctrl_interp_table(n_ctrl, T):
    """F.interpolate(mode="linear", align_corners=True) from n_ctrl to T points (:679): source index and weight."""
    scale = (n_ctrl - 1) / (T - 1) if T > 1 else 0.0
    src = scale * np.arange(T, dtype=np.float64)
    i0 = np.minimum(np.floor(src).astype(np.int64), n_ctrl - 1)
    i1 = np.minimum(i0 + 1, n_ctrl - 1)
    lam = src - i0
    return i0, i1, lam


def softplus(x):
    return np.where(x > 20.0, x, np.log1p(np.exp(np.minimum(x, 20.0))))


def monotone_grid(u, x, tab):
    """build_monotone_grid (:672-690).  u [B, n_ctrl] -> (g, xw, uT, s) each [B, T]."""
    i0, i1, lam = tab
    uT = (1.0 - lam)[None, :] * u[:, i0] + lam[None, :] * u[:, i1]
    inc = softplus(uT) + 1e-6
    s = np.cumsum(inc, axis=1)
    den = s[:, -1:] - s[:, :1] + 1e-12
    r = (s - s[:, :1]) / den
    g = x[0] + (x[-1] - x[0]) * r
    return g, g - x[None, :], uT, r, den


def interp_bins(x, g):
    """lin_interp_batch (:641-666): clamp, searchsorted(right=False), clamp to [1, T-1]."""
    gq = np.clip(g, x[0], x[-1])
    hi = np.clip(np.searchsorted(x, gq.reshape(-1), side="left").reshape(g.shape), 1, x.size - 1)
    lo = hi - 1
    w = (gq - x[lo]) / (x[hi] - x[lo] + 1e-12)
    return gq, lo, hi, w


def loss_and_grad(u, x, Y, Ym, tab, noise, lam_s, lam_a, scale):
    """Per-beat loss terms (:697-705) and d(sum_b scale_b * loss_b)/du.  Y, Ym [B, T]."""
    B, T = Y.shape
    i0, i1, lam = tab
    g, xw, uT, r, den = monotone_grid(u, x, tab)
    gq, lo, hi, w = interp_bins(x, g)
    ylo = np.take_along_axis(Y, lo, axis=1)
    yhi = np.take_along_axis(Y, hi, axis=1)
    Yw = (1.0 - w) * ylo + w * yhi
    resid = Yw - Ym
    sse = np.sum(resid * resid, axis=1)
    data = 0.5 * sse / (noise + 1e-12)
    d2 = xw[:, :-2] - 2.0 * xw[:, 1:-1] + xw[:, 2:]
    sp = np.sum(d2 * d2, axis=1)
    ap = np.sum(xw * xw, axis=1)
    loss = data + lam_s * sp + lam_a * ap
    # ---- backward
    sc = scale[:, None]
    inside = (g >= x[0]) & (g <= x[-1])
    dg = sc * resid / (noise + 1e-12) * (yhi - ylo) / (x[hi] - x[lo] + 1e-12) * inside
    dxw = sc * 2.0 * lam_a * xw
    dd2 = sc * 2.0 * lam_s * d2
    dxw[:, :-2] += dd2
    dxw[:, 1:-1] -= 2.0 * dd2
    dxw[:, 2:] += dd2
    dr = (x[-1] - x[0]) * (dg + dxw)
    ds = dr / den
    dden = -np.sum(dr * r, axis=1) / den[:, 0]
    ds[:, 0] -= np.sum(dr, axis=1) / den[:, 0] + dden
    ds[:, -1] += dden
    dinc = np.cumsum(ds[:, ::-1], axis=1)[:, ::-1]
    duT = dinc / (1.0 + np.exp(-uT))
    du = np.zeros_like(u)
    for t in range(T):
        du[:, i0[t]] += (1.0 - lam[t]) * duT[:, t]
        du[:, i1[t]] += lam[t] * duT[:, t]
    return dict(loss=loss, data=data, smooth=lam_s * sp, amp=lam_a * ap), du, xw, Yw


def theta_to_lambdas(theta, lam_s=200.0, lam_a=1e-3):
    """Warping_system._theta_to_lambdas (:376-405): only tuple/list/dict thetas change the base values."""
    if isinstance(theta, (tuple, list)) and len(theta) >= 2:
        return lam_s / (float(theta[0]) ** 2 + 1e-12), lam_a / (float(theta[1]) ** 2 + 1e-12)
    if isinstance(theta, dict):
        return lam_s / (float(theta.get("rho", 1.0)) ** 2 + 1e-12), lam_a / (float(theta.get("omega", 1.0)) ** 2 + 1e-12)
    return lam_s, lam_a


def fit_warp_batch(x, Y, Ym, noise, lam_s=200.0, lam_a=1e-3, n_ctrl=8, lr=5e-2, train_iter=50, u0=None, weights=None):
    """Warping_system.compute_warp_batch (:548-736), D = 1.  Returns dict(xw, yw, u, trace)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    Y = np.asarray(Y, dtype=np.float64)
    B, T = Y.shape
    Ym = np.broadcast_to(np.asarray(Ym, dtype=np.float64).reshape(-1, T), (B, T))
    n_ctrl = int(max(4, min(n_ctrl, T)))
    tab = ctrl_interp_table(n_ctrl, T)
    wgt = np.ones(B) if weights is None else np.maximum(np.asarray(weights, dtype=np.float64), 0.0)
    scale = wgt / (np.sum(wgt) + 1e-12)
    u = np.zeros((B, n_ctrl)) if u0 is None else np.repeat(np.asarray(u0, dtype=np.float64)[None, :], B, axis=0)
    m = np.zeros_like(u)
    v = np.zeros_like(u)
    trace = {"loss": [], "data": [], "smooth": [], "amp": []}
    for it in range(1, int(train_iter) + 1):
        terms, du, _, _ = loss_and_grad(u, x, Y, Ym, tab, noise, lam_s, lam_a, scale)
        for k in trace:
            trace[k].append(float(np.sum(scale * terms[k])))
        m = ADAM_B1 * m + (1.0 - ADAM_B1) * du
        v = ADAM_B2 * v + (1.0 - ADAM_B2) * du * du
        bc1 = 1.0 - ADAM_B1 ** it
        bc2 = 1.0 - ADAM_B2 ** it
        u = u - (lr / bc1) * m / (np.sqrt(v) / math.sqrt(bc2) + ADAM_EPS)
    g, xw, _, _, _ = monotone_grid(u, x, tab)
    _, lo, hi, w = interp_bins(x, g)
    Yw = (1.0 - w) * np.take_along_axis(Y, lo, axis=1) + w * np.take_along_axis(Y, hi, axis=1)
    return dict(xw=xw, yw=Yw, u=u, trace=trace)


def warp_prior_cov(x, rho=1.0, omega=1.0, noise2=0.0, jitter=1e-6, normalize_x=True):
    """WarpPriorAMTGP._rbf_cov (:160-173)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    if normalize_x:
        xr = x - x[0]
        x = xr / (abs(xr[-1] - xr[0]) + 1e-12)
    dx = x[:, None] - x[None, :]
    return (omega * omega) * np.exp(-0.5 * (dx * dx) / (rho * rho)) + (noise2 + jitter) * np.eye(x.size)


def warp_prior_score(x, W, noise_warp, noise_bounds, theta=None, jitter=1e-6):
    """WarpPriorAMTGP.log_sq_error_batch (:223-264): full GP log density of every warp row of W [B, T]."""
    rho, omega = 1.0, 1.0
    if isinstance(theta, (tuple, list)) and len(theta) >= 2:
        rho, omega = float(theta[0]), float(theta[1])
    elif isinstance(theta, dict):
        rho, omega = float(theta.get("rho", rho)), float(theta.get("omega", omega))
    rho, omega = max(rho, 1e-12), max(omega, 1e-12)
    n2 = min(max(float(noise_warp), noise_bounds[0]), noise_bounds[1])
    K = warp_prior_cov(x, rho, omega, n2, jitter)
    L = np.linalg.cholesky(K)
    logdet = 2.0 * np.sum(np.log(np.diag(L)))
    Z = np.linalg.solve(L, np.asarray(W, dtype=np.float64).T)      # |L^-1 w|^2 = w^T K^-1 w
    quad = np.sum(Z * Z, axis=0)
    T = K.shape[0]
    return -0.5 * (quad + logdet + T * math.log(2.0 * math.pi))


def warp_all_beats(x, Y, y_model, noise, noise_warp, noise_bounds, base_noise_warp=None, base_noise_bounds=None,
                   theta=None, train_iter=50, batch_size=128, n_ctrl=8, lr=5e-2):
    """One (lead, representative beat) column of GPI_HDP.warp_batch_by_resp_amtgp_cached (GPI_HDP.py:3466-3506):
    chunks of `batch_size` beats fitted independently (recursive_warp=False), lik = fitted warper's prior score +
    the base warper's prior score (:3494-3495)."""
    N, T = Y.shape
    lam_s, lam_a = theta_to_lambdas(theta)
    n = float(np.clip(np.mean(noise), noise_bounds[0], noise_bounds[1]))
    xw = np.zeros((N, T)); yw = np.zeros((N, T)); lik = np.zeros(N)
    for s in range(0, N, batch_size):
        sl = slice(s, min(s + batch_size, N))
        r = fit_warp_batch(x, Y[sl], y_model, n, lam_s, lam_a, n_ctrl, lr, train_iter)
        xw[sl], yw[sl] = r["xw"], r["yw"]
        lik[sl] = warp_prior_score(x, r["xw"], noise_warp, noise_bounds, theta)
        bw = noise_warp if base_noise_warp is None else base_noise_warp
        bb = noise_bounds if base_noise_bounds is None else base_noise_bounds
        lik[sl] += warp_prior_score(x, r["xw"], bw, bb, theta)
    return xw, yw, lik
