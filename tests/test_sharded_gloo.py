"""N>1 host logic on CPU: the beat-sharded HMM boundary exchange (hdpgpc_b200.hdp.sharded_hmm_exchange) and
the statistics all-reduce, world_size 2 and 3 over gloo.  The per-slice scan is the numpy oracle here (the
CUDA scan is tested on the GPU box); what is under test is the exchange protocol and its exactness."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import hdpgpc_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slice_smooth(e, pi, PiT, Pi, bin_, has_prev, has_next):
    """Sequential forward/backward over one slice with boundary messages (same recursion as the oracle)."""
    N, K = e.shape
    alpha = np.zeros((N, K)); beta = np.ones((N, K))
    prev = bin_[:K]
    for t in range(N):
        a = (pi * e[0]) if (t == 0 and not has_prev) else (PiT @ prev) * e[t]
        alpha[t] = a / np.sum(a)
        prev = alpha[t]
    u = bin_[K:]
    for t in range(N - 1, -1, -1):
        if t == N - 1 and not has_next:
            beta[t] = 1.0
        else:
            b = Pi @ u
            beta[t] = b / np.sum(b[:-1])
        u = beta[t] * e[t]
    class R: pass
    r = R()
    r.alpha, r.beta = alpha, beta
    r.boundary_out = torch.from_numpy(np.concatenate([alpha[-1], u]))
    return r


def _worker(rank, world, port, q_all, tt, st, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hdpgpc_b200 import hdp
    N, K = q_all.shape
    startPi, _ = hdp.expected_log_pi(tt, st, K)
    pi, PiT, Pi, Pc = hdp.hmm_operands(tt, startPi, K)
    e_all = O._safe_exp_rows(O.loglik_normalise(q_all))
    bounds = np.linspace(0, N, world + 1).astype(int)
    e = e_all[bounds[rank]:bounds[rank + 1]]
    smooth = lambda b, hp, hn, prev=None: _slice_smooth(e, pi, PiT, Pi, b.numpy(), hp, hn)
    hm, rounds = hdp.sharded_hmm_exchange(smooth, K, rank, world, None, torch.device("cpu"))
    # statistics all-reduce (counts are integers -> exact in any order)
    z = np.argmax(hm.alpha * hm.beta, axis=1)
    packed = torch.from_numpy(np.bincount(z, minlength=K).astype(np.float64))
    dist.all_reduce(packed)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), alpha=hm.alpha, beta=hm.beta, rounds=rounds, Nm=packed.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_hmm_exchange_is_exact(tmp_path, world):
    rng = np.random.default_rng(world)
    N, K = 900, 5
    tt = rng.gamma(1.0, 1.0, size=(K + 1, K + 1)) + np.eye(K + 1) * 20.0
    st = rng.gamma(1.0, 1.0, size=K + 1)
    lab = np.zeros(N, dtype=int)
    for t in range(1, N):
        lab[t] = lab[t - 1] if rng.uniform() < 0.9 else rng.integers(K)
    q = rng.normal(size=(N, K)) - 40.0
    q[np.arange(N), lab] += 2.0
    port = _free_port()
    mp.spawn(_worker, args=(world, port, q, tt, st, str(tmp_path)), nprocs=world, join=True)
    from hdpgpc_b200 import hdp
    startPi, _ = hdp.expected_log_pi(tt, st, K)
    pi, PiT, Pi, Pc = hdp.hmm_operands(tt, startPi, K)
    q_norm = O.loglik_normalise(q)
    alpha, _ = O.hmm_forward(pi, PiT, q_norm)
    beta = O.hmm_backward(Pi, q_norm)
    parts = [np.load(os.path.join(str(tmp_path), f"r{r}.npz")) for r in range(world)]
    a = np.concatenate([p["alpha"] for p in parts]); b = np.concatenate([p["beta"] for p in parts])
    assert np.array_equal(a, alpha) and np.array_equal(b, beta)      # bit-identical to the sequential scan
    assert all(2 <= int(p["rounds"]) <= world + 2 for p in parts)
    Nm = np.bincount(np.argmax(alpha * beta, axis=1), minlength=K)
    assert all(np.array_equal(p["Nm"], Nm) for p in parts)
