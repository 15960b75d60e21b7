"""Replay of an online seam trace (tests/golden/generate_golden.py:online_trace_scenario) on any implementation of the
GPI_model seam: the oracle (CPU test) or the device model (GPU test) behind a small adapter.  Every number the reference's
driver (GPI_HDP.include_sample, GPI_HDP.py:1906-2208) read back from a model is compared with what the replayed model
returns at the same point of the same call sequence."""
import json

import numpy as np


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)) if b.size else 0.0


def replay(z, ad, tol_score, tol_state):
    """ad: adapter with new / copy / reinit_GP / reinit_LDS / inc / pair / par / lsq / qlat / lds / est (see the two
    adapters in the tests).  Returns {op: (count, worst relative error)}."""
    Y = z["data"][:, :, 0]
    events = json.loads(str(z["events"]))
    models, worst = {}, {}

    def note(op, err, tol, k):
        c, w = worst.get(op, (0, 0.0))
        worst[op] = (c + 1, max(w, err))
        assert err < tol, (k, op, err)

    for k, e in enumerate(events):
        op = e["op"]
        g = models.get(e["gp"])
        if op == "new":
            models[e["gp"]] = ad.new(e["sigma0"], e["gamma0"])
        elif op == "copy":
            models[e["gp"]] = ad.copy(models[e["src"]])
        elif op == "reinit_GP":
            ad.reinit_GP(g)
        elif op == "reinit_LDS":
            ad.reinit_LDS(g)
        elif op == "inc":
            f, cov_chk = ad.inc(g, e["index"], Y[e["beat"]], e["h"], e.get("kernel"))
            if f"e{k}_f" in z:
                note("inc", _rel(f, z[f"e{k}_f"]), tol_state, k)
                note("inc_cov", _rel(cov_chk, z[f"e{k}_cov"]), tol_state, k)
            else:
                assert f is None, k
        elif op == "pair":
            f = ad.pair(g, e["h"])
            if f"e{k}_f" in z:
                note("pair", _rel(f, z[f"e{k}_f"]), tol_state, k)
        elif op == "par":
            n_par, chk = ad.par(g, e["h"])
            assert n_par == e["lenA"], (k, n_par, e["lenA"])
            note("par", _rel(chk, z[f"e{k}_chk"]), 10 * tol_state, k)
        elif op == "lsq":
            note("lsq", _rel(ad.lsq(g, Y[e["beat"]], e["i"]), z[f"e{k}_out"]), tol_score, k)
        elif op == "est":
            note("est", _rel(ad.est(g, Y[e["beat"]], e["h"]), z[f"e{k}_out"]), tol_score, k)
        elif op == "qlat":
            ref = z[f"e{k}_out"]
            got = np.asarray(ad.qlat(g, e["n"], e["h_ini"]), dtype=np.float64)
            assert np.array_equal(got != 0, ref != 0), k
            if np.any(ref != 0):
                note("qlat", float(np.max(np.abs(got - ref)[ref != 0] / np.abs(ref[ref != 0]))), tol_score, k)
        elif op == "lds":
            note("lds", _rel(ad.lds(g), z[f"e{k}_out"]), tol_score, k)
        else:
            raise AssertionError(op)
    # the models the driver ended with
    for ld, row in enumerate(z["final_models"]):
        for m, gid in enumerate(row):
            f_sm, Sig, idx = ad.final(models[int(gid)])
            assert list(idx) == [int(v) for v in z[f"final_{ld}_{m}_indexes"]]
            note("final_f_sm", _rel(f_sm, z[f"final_{ld}_{m}_f_sm"]), tol_state, -1)
            note("final_Sigma", _rel(Sig, z[f"final_{ld}_{m}_Sigma_last"]), 10 * tol_state, -1)
    return worst
