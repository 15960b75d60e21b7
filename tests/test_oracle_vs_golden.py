"""Pins the CPU oracle (oracle/hdpgpc_oracle.py) to the reference: every fixture under tests/golden/
was produced by the unmodified reference (tests/golden/generate_golden.py).  CPU only."""
import numpy as np
import pytest

from oracle import hdpgpc_oracle as O

OFFLINE_FULL = ["offline_rec100_T30_L1", "offline_rec102_T30_L2", "offline_rec100_T30_L1_lim30"]   # lim30: estimation_limit = 30
OFFLINE_ALL = OFFLINE_FULL + ["offline_rec100_T90_L1"]


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0


def test_hmm_synthetic_cases(golden):
    z = golden("hmm_synth")
    for c in range(int(z["n_cases"])):
        g = lambda k: z[f"c{c}_{k}"]
        r = O.estep_responsibilities(g("q"), g("snr"), g("transTheta"), g("startTheta"))
        assert rel(r["qbar"], g("qbar")) < 1e-13
        assert np.max(np.abs(r["alpha"] - g("alpha"))) < 1e-12
        assert rel(r["beta"], g("beta")) < 1e-11
        assert rel(r["marg"], g("margprob")) < 1e-11
        assert np.array_equal(r["z"], g("z"))
        assert np.array_equal(r["zpair"], g("zpair"))
        assert np.array_equal(r["transStateCount"], g("transStateCount"))
        assert np.array_equal(r["startStateCount"], g("startStateCount"))
        assert np.array_equal(r["Nm"], g("Nm"))
        assert abs(r["Q_em"] - float(g("Q_em"))) <= 1e-12 * abs(float(g("Q_em")))
        assert str(g("respPair_dtype")) == "torch.float32"   # reference quirk kept in the mirror


@pytest.mark.parametrize("name", OFFLINE_FULL)
def test_seam_functions_from_dumped_state(golden, name):
    z = golden(name)
    M, L = int(z["M"]), int(z["L"])
    Y, Ynew = z["data"], z["new"]
    N = Y.shape[0]
    gps = [[O.OracleGP.from_dump(z, f"gp_{ld}_{m}_") for m in range(M)] for ld in range(L)]
    q = np.zeros((N, M, L)); qnf = np.zeros_like(q); ql = np.zeros_like(q); snr = np.zeros_like(q)
    qn = np.zeros((Ynew.shape[0], M, L)); sn = np.zeros_like(qn)
    for ld in range(L):
        for m in range(M):
            gp = gps[ld][m]
            q[:, m, ld] = gp.compute_sq_err_all(None, Y[:, :, ld])
            qnf[:, m, ld] = gp.compute_sq_err_all(None, Y[:, :, ld], no_first=True)
            ql[:, m, ld] = gp.compute_q_lat_all(N)
            snr[:, m, ld] = O.compute_snr(Y[:, :, ld], gp)
            qn[:, m, ld] = [gp.log_sq_error(None, y, i=-1) for y in Ynew[:, :, ld]]
            sn[:, m, ld] = O.compute_snr(Ynew[:, :, ld], gp)
    assert rel(q, z["q_all"]) < 1e-12
    assert rel(qnf, z["q_all_nofirst"]) < 1e-12
    assert rel(ql, z["q_lat_all"]) < 1e-12
    assert rel(snr, z["snr_all"]) < 1e-12
    assert rel(qn, z["q_new"]) < 1e-12
    assert rel(sn, z["snr_new"]) < 1e-12
    for pre, qq, ss in (("train_", z["q_all"], z["snr_all"]), ("new_", z["q_new"], z["snr_new"])):
        r = O.estep_responsibilities(qq, ss, z["transTheta"], z["startTheta"])
        assert np.array_equal(r["z"], z[pre + "z"])
        assert np.array_equal(r["zpair"], z[pre + "zpair"])
        assert np.max(np.abs(r["alpha"] - z[pre + "alpha"])) < 1e-12
        assert rel(r["beta"], z[pre + "beta"]) < 1e-10
        assert np.array_equal(r["transStateCount"], z[pre + "transStateCount"])
        assert np.array_equal(r["Nm"], z[pre + "Nm"])
        assert rel(r["Q_em"], z[pre + "Q_em"]) < 1e-13
    r = O.estep_responsibilities(z["q_all"], None, z["transTheta"], z["startTheta"], snr_norm=z["snr_norm"])
    assert np.array_equal(r["z"], z["saved_z"])
    assert np.array_equal(r["z"] if False else O.estep_responsibilities(
        z["q_new"], z["snr_new"], z["transTheta"], z["startTheta"])["z"], z["new_cluster_new_batch"])
    lik = np.array([[gps[ld][m].return_LDS_param_likelihood() for m in range(M)] for ld in range(L)])
    assert rel(lik, z["lds_param_lik"]) < 1e-12
    Nm = z["train_Nm"]
    fl = [O.full_LDS_elbo(gps[ld], Nm, 5) for ld in range(L)]
    assert rel(fl, z["elbo_full_LDS"]) < 1e-12


@pytest.mark.parametrize("name", OFFLINE_ALL)
def test_chain_replay(golden, name):
    """Fresh model -> Kalman / pair smoother / MNIW per member -> full RTS pass -> scores
    (GPI_model.full_pass_weighted).  The first Kalman step solves against K + noise*I with
    cond ~1e7, so different LAPACK builds already differ at ~1e-10; the bar is the north-star 1e-8."""
    z = golden(name)
    Y = z["data"]
    full = f"chain_0_Sigma" in z.files
    for m in range(int(z["n_chain"])):
        pre = f"chain_{m}_"
        lim = float(z[pre + "estimation_limit"])
        gp = O.OracleGP(z["x_basis"], z["kernel_def"], float(z["ini_sigma_def"]), float(z["ini_gamma_def"]),
                        free_deg=int(z["free_deg_MNIV"]), estimation_limit=None if np.isinf(lim) else lim)
        q, ql = gp.full_pass_weighted(Y[:, :, 0], z[pre + "resp"], fitted_kernel=z[pre + "kernel"])
        assert gp.indexes == [int(i) for i in z[pre + "indexes"]]
        assert rel(q, z[pre + "q"]) < 1e-8
        assert rel(ql, z[pre + "q_lat"]) < 1e-8
        assert rel(O.compute_snr(Y[:, :, 0], gp), z[pre + "snr"]) < 1e-8
        scale = np.max(np.abs(z[pre + "f_star"]))
        assert np.max(np.abs(np.stack(gp.f_star) - z[pre + "f_star"][:, :, 0])) < 1e-8 * scale
        assert np.max(np.abs(np.stack(gp.f_star_sm) - z[pre + "f_star_sm"][:, :, 0])) < 1e-8 * scale
        for nm in ["cov_f", "cov_f_sm", "A", "Gamma", "C", "Sigma"]:
            mine = getattr(gp, nm)
            if full:
                ref = z[pre + nm]
                assert len(mine) == ref.shape[0]
                assert np.max(np.abs(np.stack(mine) - ref)) < 1e-8 * np.max(np.abs(ref))
            else:
                assert len(mine) == int(z[pre + nm + "_len"])
                ref = z[pre + nm + "_last"]
                assert np.max(np.abs(mine[-1] - ref)) < 1e-8 * np.max(np.abs(ref))
                chk = np.array([[np.trace(v), np.linalg.norm(v)] for v in mine])
                assert rel(chk, z[pre + nm + "_chk"][:, :2]) < 1e-8


def test_inducing_grid(golden):
    """x_train != x_basis: pred_dist kernel branch (GPI.py:470-501)."""
    z = golden("inducing_T30")
    gp = O.OracleGP.from_dump(z, "gp_")
    Y = z["data"][:, :, 0]
    x_off, x_sub = z["x_off"].reshape(-1), z["x_sub"].reshape(-1)
    for k, i in enumerate([1, 4, 9]):
        f, c = gp.observe(x_off, i)
        assert np.max(np.abs(f - z[f"obs_off_{i}_mean"][:, 0])) < 1e-9 * np.max(np.abs(f))
        assert np.max(np.abs(c - z[f"obs_off_{i}_cov"])) < 1e-12 * np.max(np.abs(c))
        s = [gp.log_sq_error(x_off, Y[n], i=i) for n in range(6)]
        assert rel(s, z["score_off"][k]) < 1e-11
    assert rel([gp.log_sq_error(x_off, Y[n], i=-1) for n in range(6)], z["score_off_last"]) < 1e-11
    assert rel([gp.log_sq_error(x_sub, Y[n, ::2], i=3) for n in range(6)], z["score_sub"]) < 1e-11


def test_table_form_matches_object_form(golden):
    """score_states / snr_states (the struct-of-arrays form the device consumes) == per-object methods."""
    z = golden("offline_rec100_T30_L1")
    M = int(z["M"])
    Y = z["data"][:, :, 0]
    N = Y.shape[0]
    for m in range(M):
        gp = O.OracleGP.from_dump(z, f"gp_0_{m}_")
        i_vals, first = gp.state_index_map(N)
        nF = len(gp.f_star)
        mu = np.stack([gp.observe_state(t)[0] for t in range(nF)] + [gp.observe_state(1)[0]])
        Sig = np.stack([gp.observe_state(t)[1] for t in range(nF)] + [gp.observe_state(1)[1]])
        add = np.zeros(nF + 1); add[-1] = gp.first_jitter()
        s_of = np.where(first, nF, i_vals).reshape(N, 1)
        q = O.score_states(Y, mu, Sig, s_of, np.arange(nF + 1), add)[:, 0]
        assert rel(q, z["q_all"][:, m, 0]) < 1e-12


def test_online_extras(golden):
    """estimate_new / posterior_weighted."""
    z = golden("online_T30")
    Y = z["data"][:, :, 0]
    for tag in ("many", "one"):
        gp = O.OracleGP.from_dump(z, tag + "_")
        for k, n in enumerate(range(22, 30)):
            m, P = gp.posterior_weighted(None, Y[n], 1.0)
            assert np.max(np.abs(m - z[tag + "_post_mean"][k][:, 0])) < 1e-9 * np.max(np.abs(m))
            assert np.max(np.abs(P - z[tag + "_post_cov"][k])) < 1e-9 * np.max(np.abs(P))
            assert abs(gp.estimate_new(None, Y[n]) - z[tag + "_estimate_new"][k]) < 1e-9 * abs(z[tag + "_estimate_new"][k])
        m, P = gp.posterior_weighted(None, Y[25], 0.5)
        assert np.max(np.abs(m - z[tag + "_post_mean_h05"][:, 0])) < 1e-9 * np.max(np.abs(m))
        assert np.max(np.abs(P - z[tag + "_post_cov_h05"])) < 1e-9 * np.max(np.abs(P))


def test_warp_oracle_vs_reference(golden):
    """Closed-form-gradient restatement of Warping_system.compute_warp_batch + WarpPriorAMTGP.log_sq_error_batch +
    the chunked driver (GPI_HDP.py:3412-3517) against the reference's autograd/Adam outputs (record 102, T=90)."""
    from oracle import warp_oracle as W
    z = golden("warp_rec102_T90")
    x = z["x_basis"].reshape(-1)
    Y = z["data"][:, :, 0]
    nb = z["noise_bounds"]
    n = float(np.clip(z["noise"], nb[0], nb[1]))
    for tag, theta in (("f", float(z["theta_float"])), ("t", tuple(z["fit_t_theta"]))):
        ls, la = W.theta_to_lambdas(theta)
        r = W.fit_warp_batch(x, Y[:40], Y[3], n, ls, la, int(z["n_ctrl"]), float(z["lr"]), 50)
        assert np.max(np.abs(r["xw"] - z[f"fit_{tag}_xw"])) < 1e-10 * np.max(np.abs(z[f"fit_{tag}_xw"]))
        assert np.max(np.abs(r["yw"] - z[f"fit_{tag}_yw"])) < 1e-10 * np.max(np.abs(z[f"fit_{tag}_yw"]))
        tr = np.array([r["trace"][k] for k in ("loss", "data", "smooth", "amp")])
        assert rel(tr[:2], z[f"fit_{tag}_trace"][:2]) < 1e-10
        lik = W.warp_prior_score(x, r["xw"], float(z["noise_warp"]), nb, theta)
        assert rel(lik, z[f"fit_{tag}_lik"]) < 1e-10
    for m, ref in enumerate(z["drv_f_ind"]):
        xw, yw, lik = W.warp_all_beats(x, Y, Y[ref], np.full(x.size, z["noise"]), float(z["drv_fit_noise_warp"][m]),
                                       z["drv_fit_noise_bounds"][m], float(z["drv_base_noise_warp"]),
                                       z["drv_base_noise_bounds"], theta=float(z["theta_float"]))
        assert np.max(np.abs(xw - z["drv_xw"][:, :, 0, m])) < 1e-10 * np.max(np.abs(z["drv_xw"][:, :, 0, m]))
        assert np.max(np.abs(yw - z["drv_yw"][:, :, 0, m])) < 1e-10 * np.max(np.abs(z["drv_yw"][:, :, 0, m]))
        assert rel(lik, z["drv_liks"][:, m, 0]) < 1e-10


def test_oracle_replays_offline_fit_trace(golden):
    """The seam trace of a whole reference fit (tests/golden/generate_golden.py: trace_scenario): the oracle's chain
    replay and HMM blocks against what the reference's driver saw (every 3rd chain, every HMM block)."""
    z = golden("trace_rec102_T30_L2")
    Y = z["data"]
    N, T, L = Y.shape
    for i in range(0, int(z["n_chains"]), 3):
        resp = np.unpackbits(z[f"c{i}_resp"])[:N].astype(float)
        lead, fitted_before, n_states, s0, g0 = z[f"c{i}_meta"]
        kern = tuple(z[f"c{i}_kernel"])
        og = O.OracleGP(z["x_basis"], kern, float(s0), float(g0), free_deg=int(z["free_deg_MNIV"]))
        q, ql = og.full_pass_weighted(Y[:, :, int(lead)], resp, fitted_kernel=kern)
        assert len(og.f_star) == int(n_states)
        assert rel(q, z[f"c{i}_q"]) < 2e-7
    for i in range(int(z["n_hmm"])):
        K = z[f"h{i}_q"].shape[1]
        pi, PiT, Pi, Pc = O.hmm_operands(z[f"h{i}_transTheta"], z[f"h{i}_pi"], K)
        qn = z[f"h{i}_q"]
        alpha, _ = O.hmm_forward(pi, PiT, qn)
        beta = O.hmm_backward(Pi, qn)
        assert np.array_equal(O.hard_resp(alpha, beta), z[f"h{i}_z"])
        assert np.array_equal(O.hard_resp_pair(alpha, beta, Pc, qn), z[f"h{i}_zpair"])


@pytest.mark.parametrize("name", ["online_trace_rec100_T30_L1", "online_trace_rec100_T90_L1"])
def test_oracle_replays_online_fit_trace(golden, name):
    """The seam trace of a whole ONLINE reference fit (generate_golden.py: online_trace_scenario; the loop of
    hdpgpc/tests/test_online.py on 30 beats of record 100): every call GPI_HDP.include_sample made on a GPI_model, in order,
    replayed on oracle models (trial copies and resets included), and every HMM block of variational_local_terms."""
    import copy
    from online_replay import replay
    z = golden(name)
    chk = lambda m: [float(np.trace(m)), float(np.linalg.norm(m))]

    class Oracle:
        def new(self, s0, g0):
            return O.OracleGP(z["x_basis"], (1.0, 1.2, 1.0), s0, g0, free_deg=float(z["free_deg_MNIV"]),
                              noise_bounds=tuple(z["noise_bounds"]))

        def copy(self, g):
            return copy.deepcopy(g)

        def reinit_GP(self, g):
            g.reinit_GP()

        def reinit_LDS(self, g):
            g.reinit_LDS()

        def inc(self, g, index, y, h, kernel):
            if h != 1.0:                          # include_sample(posterior=False): nothing is stored (GPI_model.py:374)
                return None, None
            if g.N == 0 and not g.fitted:
                g.fit_kernel_params(g.x_basis, y, fitted_kernel=kernel)
            g.include_sample(index, None, y)
            return g.f_star[-1], chk(g.cov_f[-1])

        def pair(self, g, h):
            if h == 1.0:
                g.backwards_pair()
            return np.stack(g.f_star_sm[-2:]) if len(g.indexes) > 1 else None

        def par(self, g, h):
            if h == 1.0:
                g.bayesian_new_params()
            return len(g.A), np.array([chk(g.A[-1]), chk(g.Gamma[-1]), chk(g.C[-1]), chk(g.Sigma[-1])])

        def lsq(self, g, y, i):
            return g.log_sq_error(None, y, i=i)

        def est(self, g, y, h):
            return g.estimate_new(None, y, h)

        def qlat(self, g, n, h_ini):
            return g.compute_q_lat_all(n, h_ini)

        def lds(self, g):
            return g.return_LDS_param_likelihood()

        def final(self, g):
            return np.stack(g.f_star_sm), g.Sigma[-1], g.indexes

    worst = replay(z, Oracle(), tol_score=1e-9, tol_state=1e-9)
    import json
    n_ev = lambda op: sum(1 for e in json.loads(str(z["events"])) if e["op"] == op)
    assert worst["lsq"][0] == n_ev("lsq") > 50 and worst["est"][0] == n_ev("est") > 40 and worst["lds"][0] == n_ev("lds")
    assert int(z["M"]) == int(z["M_after"][-1]) == {"online_trace_rec100_T30_L1": 2, "online_trace_rec100_T90_L1": 10}[name]
    for i in range(int(z["n_hmm"])):
        K = z[f"h{i}_q"].shape[1]
        pi, PiT, Pi, Pc = O.hmm_operands(z[f"h{i}_transTheta"], z[f"h{i}_pi"], K)
        qn = z[f"h{i}_q"]
        alpha, _ = O.hmm_forward(pi, PiT, qn)
        beta = O.hmm_backward(Pi, qn)
        with np.errstate(divide="ignore"):
            assert np.array_equal(O.hard_resp(alpha, beta), z[f"h{i}_z"]), i
            assert np.array_equal(O.hard_resp_pair(alpha, beta, Pc, qn), z[f"h{i}_zpair"]), i
        assert rel(alpha[-1], z[f"h{i}_alpha_last"]) < 1e-12


def test_oracle_replays_warp_fit_trace(golden):
    """BASELINE.json configs[2] (offline fit with alignment enabled, record 102): every call of the cached all-beats
    driver GPI_HDP.warp_batch_by_resp_amtgp_cached the reference made during a whole include_batch(warp=True) fit, the
    chains on the warped beats and the HMM blocks, against the oracle."""
    from oracle import warp_oracle as W
    z = golden("trace_warp_rec102_T30_L1")
    Y = z["data"][:, :, 0]
    x = z["x_basis"].reshape(-1)
    N, T = Y.shape
    assert not bool(z["warp_recursive"]) and int(z["n_warp"]) > 30
    fits = {}
    for i in range(int(z["n_warp"])):
        for m, ref in enumerate(z[f"w{i}_refs"]):
            wp = min(m, int(z[f"w{i}_n_wp"]) - 1)
            key = (int(ref), wp)
            if key not in fits:
                fits[key] = W.warp_all_beats(x, Y, Y[int(ref)], np.full(T, float(z["warp_noise"])),
                                             float(z["warp_fit_noise_warp"][wp]), z["warp_fit_noise_bounds"][wp],
                                             float(z["warp_base_noise_warp"]), z["warp_base_noise_bounds"],
                                             theta=float(z["warp_theta"]))
            xw, yw, lik = fits[key]
            assert np.max(np.abs(xw - z[f"w{i}_xw"][:, :, 0, m])) < 1e-9 * max(np.max(np.abs(z[f"w{i}_xw"][:, :, 0, m])), 1e-3), (i, m)
            assert np.max(np.abs(yw - z[f"w{i}_yw"][:, :, 0, m])) < 1e-9 * np.max(np.abs(z[f"w{i}_yw"][:, :, 0, m])), (i, m)
            assert rel(lik, z[f"w{i}_liks"][:, m, 0]) < 1e-9, (i, m)
    assert len(fits) >= 5
    for i in range(0, int(z["n_chains"]), 2):
        resp = np.unpackbits(z[f"c{i}_resp"])[:N].astype(float)
        lead, fitted_before, n_states, s0, g0 = z[f"c{i}_meta"]
        kern = tuple(z[f"c{i}_kernel"])
        og = O.OracleGP(z["x_basis"], kern, float(s0), float(g0), free_deg=int(z["free_deg_MNIV"]))
        q, ql = og.full_pass_weighted(z[f"c{i}_Y"], resp, fitted_kernel=kern)
        assert len(og.f_star) == int(n_states)
        assert rel(q, z[f"c{i}_q"]) < 2e-7
    for i in range(int(z["n_hmm"])):
        K = z[f"h{i}_q"].shape[1]
        pi, PiT, Pi, Pc = O.hmm_operands(z[f"h{i}_transTheta"], z[f"h{i}_pi"], K)
        qn = z[f"h{i}_q"]
        alpha, _ = O.hmm_forward(pi, PiT, qn)
        beta = O.hmm_backward(Pi, qn)
        with np.errstate(divide="ignore"):
            assert np.array_equal(O.hard_resp(alpha, beta), z[f"h{i}_z"])
            assert np.array_equal(O.hard_resp_pair(alpha, beta, Pc, qn), z[f"h{i}_zpair"])
