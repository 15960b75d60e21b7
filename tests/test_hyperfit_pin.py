"""An EXTERNAL pin for the one-beat hyper-parameter fit.

The reference runs this step with gpytorch 1.13 (GPI.py:610-770, GPI_models_pytorch.py:10-22), which cannot be installed in
this image, so neither the oracle's restatement (oracle/hyperfit.py, torch autograd) nor the device kernel
(csrc/hgp_hyperfit.cu, closed-form gradient) can be compared with the real library.  What can be done without it is to
check both against something that shares NO code with either: the objective written down here from gpytorch's documented
parameterisation, in plain numpy, with its gradient taken by central finite differences and Adam stepped by hand.

  ConstantMean            constant = raw_constant                                  (initial 0)
  ScaleKernel(RBFKernel)  outputscale = softplus(raw_outputscale), lengthscale = softplus(raw_lengthscale)   (initial raw 0)
  GaussianLikelihood(noise_constraint=Interval(lo, hi))   noise = lo + (hi - lo) sigmoid(raw_noise)          (initial raw 0)
  ExactMarginalLogLikelihood / T:  loss = -[ -1/2 r^T K^-1 r - 1/2 log det K - T/2 log 2 pi ] / T,  K = s exp(-d^2 / 2 l^2) + noise I
  torch.optim.Adam(lr = 0.1, betas = (0.9, 0.999), eps = 1e-8)

Pinned: the loss at the initial point, the first Adam step (decided by the SIGN of every partial derivative), and a
25-step trajectory -- for the oracle on the CPU and for the device kernel on the GPU."""
import numpy as np
import pytest

LO, HI = 0.5, 30.0
X = np.arange(8, dtype=np.float64)
Y = np.array([3.0, 7.5, 9.0, 4.0, -2.5, -6.0, -1.0, 2.0])


def _transform(raw):
    c, rs, rl, rn = raw
    softplus = lambda v: np.log1p(np.exp(v))
    return c, softplus(rs), softplus(rl), LO + (HI - LO) / (1.0 + np.exp(-rn))


def _loss(raw, x=X, y=Y):
    c, s, ell, noise = _transform(raw)
    d = x[:, None] - x[None, :]
    K = s * np.exp(-0.5 * d * d / (ell * ell)) + noise * np.eye(x.size)
    r = y - c
    sign, logdet = np.linalg.slogdet(K)
    mll = -0.5 * r @ np.linalg.solve(K, r) - 0.5 * logdet - 0.5 * x.size * np.log(2.0 * np.pi)
    return -mll / x.size


def _grad_fd(raw, h=1e-6):
    g = np.zeros(4)
    for p in range(4):
        e = np.zeros(4)
        e[p] = h
        g[p] = (_loss(raw + e) - _loss(raw - e)) / (2.0 * h)
    return g


def _adam(n_steps, lr=0.1):
    raw, m, v = np.zeros(4), np.zeros(4), np.zeros(4)
    losses = []
    for t in range(1, n_steps + 1):
        losses.append(_loss(raw))
        g = _grad_fd(raw)
        m = 0.9 * m + 0.1 * g
        v = 0.999 * v + 0.001 * g * g
        raw = raw - lr * (m / (1 - 0.9 ** t)) / (np.sqrt(v / (1 - 0.999 ** t)) + 1e-8)
    return raw, losses


def test_hand_written_objective_is_self_consistent():
    # value at the initial point by hand: s = l = ln 2, noise = (lo + hi) / 2
    c, s, ell, noise = _transform(np.zeros(4))
    assert c == 0.0 and abs(s - np.log(2.0)) < 1e-15 and abs(ell - np.log(2.0)) < 1e-15 and noise == 0.5 * (LO + HI)
    # the finite-difference gradient against the textbook derivative of the marginal likelihood,
    # d mll / d theta = 1/2 tr((a a^T - K^-1) dK/dtheta), a = K^-1 r, for the noise and the constant mean
    d = X[:, None] - X[None, :]
    K = s * np.exp(-0.5 * d * d / (ell * ell)) + noise * np.eye(X.size)
    Ki = np.linalg.inv(K)
    a = Ki @ (Y - c)
    dnoise_draw = (HI - LO) * 0.25                                   # sigmoid'(0) = 1/4
    g_noise = -0.5 * (a @ a - np.trace(Ki)) * dnoise_draw / X.size
    g_const = -np.sum(a) / X.size
    g = _grad_fd(np.zeros(4))
    assert abs(g[3] - g_noise) < 1e-7 * abs(g_noise) and abs(g[0] - g_const) < 1e-7 * abs(g_const)


def test_oracle_follows_the_hand_written_objective():
    import torch
    from oracle import hyperfit
    raw1, losses = _adam(1)
    s, ell, noise, n_it = hyperfit.fit_exact_gp(torch.from_numpy(X), torch.from_numpy(Y), (LO, HI), training_iter=1)
    c1, s1, l1, n1 = _transform(raw1)
    assert n_it == 1
    assert abs(s - s1) < 1e-7 * s1 and abs(ell - l1) < 1e-7 * l1 and abs(noise - n1) < 1e-7 * n1
    raw25, _ = _adam(25)
    s, ell, noise, n_it = hyperfit.fit_exact_gp(torch.from_numpy(X), torch.from_numpy(Y), (LO, HI), training_iter=25)
    c25, s25, l25, n25 = _transform(raw25)
    assert abs(s - s25) < 1e-5 * s25 and abs(ell - l25) < 1e-5 * l25 and abs(noise - n25) < 1e-5 * n25


@pytest.mark.gpu
def test_device_fit_follows_the_hand_written_objective():
    import torch
    from hdpgpc_b200 import ops
    x, y = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda().reshape(1, -1)
    raw1, losses = _adam(26)
    out = ops.hyperfit_batched(x, y, (LO, HI), max_iter=1, min_iter=1000).cpu().numpy()[0]
    c1, s1, l1, n1 = _transform(_adam(1)[0])
    assert int(out[5]) == 1 and int(out[6]) == 0
    assert abs(out[4] - losses[0]) < 1e-12 * abs(losses[0])                  # the loss at the initial point
    assert abs(out[0] - s1) < 1e-7 * s1 and abs(out[1] - l1) < 1e-7 * l1 and abs(out[2] - n1) < 1e-7 * n1
    assert abs(out[3] - c1) < 1e-7 * max(abs(c1), 1e-3)
    out = ops.hyperfit_batched(x, y, (LO, HI), max_iter=26, min_iter=1000).cpu().numpy()[0]
    c25, s25, l25, n25 = _transform(_adam(26)[0])
    assert abs(out[4] - losses[25]) < 1e-6 * abs(losses[25])                 # loss evaluated at the 26th iterate
    assert abs(out[0] - s25) < 1e-5 * s25 and abs(out[1] - l25) < 1e-5 * l25 and abs(out[2] - n25) < 1e-5 * n25
