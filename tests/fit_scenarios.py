"""Whole-fit scenarios shared by tests/golden/generate_fit_golden.py (reference on the CPU, dev container) and
tests/test_reference_fit_gpu.py (the same reference driver on the device path, GPU box): the hyper-parameters of the
reference's entry scripts hdpgpc/tests/test_offline.py:37-75 and hdpgpc/tests/test_online.py:36-75, lead 0."""
import os
import time

import numpy as np

SCENARIOS = {
    # BASELINE.json configs[0]: python hdpgpc/tests/test_offline.py 100
    "rec100_offline": dict(rec="100", n=None, mode="offline", warp=False, free_deg=5, n_explore_steps=5,
                           estimation_limit=None),
    # BASELINE.json configs[2]: test_offline.py 102 with alignment enabled
    "rec102_warp": dict(rec="102", n=320, mode="offline", warp=True, free_deg=5, n_explore_steps=5,
                        estimation_limit=None),
    # BASELINE.json configs[1]: python hdpgpc/tests/test_online.py 100 (first 500 beats)
    "rec100_online": dict(rec="100", n=500, mode="online", warp=False, free_deg=20, n_explore_steps=10,
                          estimation_limit=None),
    # the benchmarked regime: finite estimation_limit (tests/test_online_warp.py:75 uses 100, test_step.ipynb 30)
    "rec100_limit30": dict(rec="100", n=160, mode="offline", warp=False, free_deg=5, n_explore_steps=5,
                           estimation_limit=30),
}


def load_beats(sc, data_dir):
    data = np.load(os.path.join(data_dir, sc["rec"] + ".npy"))
    labels = np.load(os.path.join(data_dir, sc["rec"] + "_labels.npy"))
    data = data[:, :, [0]]                                     # lead = 0 (tests/test_offline.py:35-36)
    if sc["n"] is not None:
        data, labels = data[:sc["n"]], labels[:sc["n"]]
    return np.ascontiguousarray(data), labels


def build(hdp, sc, data):
    """The constructor call of the entry scripts (test_offline.py:68-75 / test_online.py:68-75)."""
    from hdpgpc.get_data import compute_estimators_LDS
    num_samples, T, L = data.shape
    if sc["mode"] == "online":
        std, std_dif, bound_sigma, bound_gamma = compute_estimators_LDS(data, 30)
    else:
        std, std_dif, bound_sigma, bound_gamma = compute_estimators_LDS(data)
    sigma, gamma = std * 1.0, std_dif * 1.0
    noise_warp = std * 0.1
    bound_noise_warp = (noise_warp * 0.1, noise_warp * 0.2)
    x_basis = np.atleast_2d(np.arange(0, T, 1, dtype=np.float64)).T
    x_basis_warp = np.atleast_2d(np.arange(0, T, 2, dtype=np.float64)).T
    common = dict(x_basis_warp=x_basis_warp, n_outputs=L, kernels=None, model_type='dynamic', ini_lengthscale=3.0,
                  bound_lengthscale=(1.0, 20.0), ini_gamma=gamma, ini_sigma=sigma, ini_outputscale=300.0,
                  noise_warp=noise_warp, bound_sigma=bound_sigma, bound_gamma=bound_gamma,
                  bound_noise_warp=bound_noise_warp, method_compute_warp='greedy', verbose=False, hmm_switch=True,
                  max_models=100, mode_warp='rough', bayesian_params=True, inducing_points=False,
                  estimation_limit=sc["estimation_limit"], free_deg_MNIV=sc["free_deg"])
    if sc["mode"] == "online":
        sw = hdp.GPI_HDP(x_basis, warp_updating=sc["warp"], **common)
    else:
        sw = hdp.GPI_HDP(x_basis, warp_updating=False, reestimate_initial_params=True,
                         n_explore_steps=sc["n_explore_steps"], **common)
    return sw, x_basis


def run_fit(hdp, sc, data_dir):
    data, _ = load_beats(sc, data_dir)
    sw, x_basis = build(hdp, sc, data)
    t0 = time.time()
    if sc["mode"] == "online":
        for i in range(data.shape[0]):
            sw.include_sample(x_basis, data[i], with_warp=sc["warp"])
    else:
        x_trains = np.array([x_basis] * data.shape[0])
        sw.include_batch(x_trains, data, warp=sc["warp"])
    return sw, time.time() - t0


def summarize(sw):
    import torch
    last = sw.resp_assigned[-1].detach().cpu().numpy().astype(np.int64)
    out = dict(M=np.int64(sw.M), labels=last, n_outer=np.int64(len(sw.resp_assigned)),
               sizes=np.bincount(last, minlength=int(sw.M)).astype(np.int64),
               elbo=np.array([float(torch.as_tensor(e).reshape(-1)[0]) for e in sw.train_elbo], dtype=np.float64))
    if len(sw.resp_assigned) and all(r.shape == sw.resp_assigned[-1].shape for r in sw.resp_assigned):
        out["labels_per_iteration"] = np.stack([r.detach().cpu().numpy().astype(np.int64) for r in sw.resp_assigned])
    kern = []
    for gp in sw.gpmodels[0]:
        kp = gp.gp.kernel.get_params()
        kern.append([kp["k1__k1__constant_value"], kp["k1__k2__length_scale"], kp["k2__noise_level"], float(gp.N)])
    out["kernels"] = np.array(kern, dtype=np.float64).reshape(-1, 4)
    out["members"] = np.array([gp.N for gp in sw.gpmodels[0]], dtype=np.int64)
    return out
