"""Thin Python wrappers over the C ABI (include/hdpgpc_b200.h).  Tensors are torch CUDA tensors
(device memory + streams are torch's; the arithmetic is the library's).  No CPU fallback."""
import ctypes
import os

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

F64 = torch.float64
I32 = torch.int32

_NVTX = bool(os.environ.get("HGP_NVTX"))


class nvtx:
    """NVTX range around a stage of the path (HGP_NVTX=1; shows up in `ncu --nvtx` / Nsight timelines): the stages of a
    sweep (table build, scoring per lead, responsibilities, statistics), chain replays, hyper-fits, warp fits."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push("hgp:" + self.name)
        return self

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def _lib_ready():
    _lib.require_cuda()
    return _lib.load()


def _dev(t):
    if not t.is_cuda:
        raise _lib.HgpError("expected a CUDA tensor")
    return t


def launch_count():
    return int(_lib.load().hgp_launch_count())


def tile_beats():
    """Beats per tile of the tile kernel (hgp_tile_beats)."""
    return int(_lib.load().hgp_tile_beats())


def pack_leads(Y_ntl):
    """[N, T, L] -> [L, N, T] (reference layout -> per-lead planes)."""
    lib = _lib_ready()
    Y = _dev(Y_ntl).contiguous()
    N, T, L = Y.shape
    if L == 1:
        return Y.reshape(1, N, T)
    out = torch.empty((L, N, T), dtype=F64, device=Y.device)
    check(lib.hgp_pack_leads(ptr(Y), N, T, L, ptr(out), stream_ptr()), "hgp_pack_leads")
    return out


def pack_leads_slice(Y_ntl_slice, Yp, n0):
    """[n, T, L] slice (device) -> rows [n0, n0 + n) of every plane of Yp[L, N, T]."""
    lib = _lib_ready()
    Ys = _dev(Y_ntl_slice).contiguous()
    n, T, L = Ys.shape
    Lp, N, Tp = Yp.shape
    if (Lp, Tp) != (L, T) or n0 + n > N or not Yp.is_contiguous():
        raise _lib.HgpError("pack_leads_slice: shape mismatch")
    dst = Yp.data_ptr() + n0 * T * 8
    check(lib.hgp_pack_leads_slice(ptr(Ys), n, T, L, ctypes.c_void_p(dst), N * T, stream_ptr()), "hgp_pack_leads_slice")


def chol_batched(Sigma, add_diag=None, jitter_scale=1e-8, want_logdet=False):
    """GPI_model._chol_spd for a stack [F, T, T].  Returns (L, info[, logdet])."""
    lib = _lib_ready()
    S = _dev(Sigma).contiguous()
    F, T, _ = S.shape
    L = torch.empty_like(S)
    info = torch.empty(F, dtype=I32, device=S.device)
    logdet = torch.empty(F, dtype=F64, device=S.device) if want_logdet else None
    if add_diag is not None:
        add_diag = _dev(add_diag).to(F64).contiguous()
    check(lib.hgp_chol_batched(ptr(S), F, T, ptr(add_diag), float(jitter_scale), ptr(L), ptr(logdet), ptr(info),
                               stream_ptr()), "hgp_chol_batched")
    return (L, info, logdet) if want_logdet else (L, info)


def cholinv_batched(Sigma, add_diag=None, jitter_scale=1e-8, want_logdet=False):
    """chol_batched + tri_inverse_batched in one launch (hgp_cholinv_batched).  Returns (L, W, info[, logdet])."""
    lib = _lib_ready()
    S = _dev(Sigma).contiguous()
    F, T, _ = S.shape
    L = torch.empty_like(S)
    W = torch.empty_like(S)
    info = torch.empty(F, dtype=I32, device=S.device)
    logdet = torch.empty(F, dtype=F64, device=S.device) if want_logdet else None
    if add_diag is not None:
        add_diag = _dev(add_diag).to(F64).contiguous()
    check(lib.hgp_cholinv_batched(ptr(S), F, T, ptr(add_diag), float(jitter_scale), ptr(L), ptr(W), ptr(logdet), ptr(info),
                                  stream_ptr()), "hgp_cholinv_batched")
    return (L, W, info, logdet) if want_logdet else (L, W, info)


def tri_inverse_batched(Lfac):
    lib = _lib_ready()
    Lf = _dev(Lfac).contiguous()
    F, T, _ = Lf.shape
    W = torch.empty_like(Lf)
    check(lib.hgp_tri_inverse_batched(ptr(Lf), F, T, ptr(W), stream_ptr()), "hgp_tri_inverse_batched")
    return W


def pack_factors(W):
    lib = _lib_ready()
    W = _dev(W).contiguous()
    F, T, _ = W.shape
    nd = lib.hgp_packed_factor_bytes(T) // 8
    out = torch.empty((F, nd), dtype=F64, device=W.device)
    check(lib.hgp_pack_factors(ptr(W), F, T, ptr(out), stream_ptr()), "hgp_pack_factors")
    return out


def emission_means(C, f, c_idx, f_idx):
    lib = _lib_ready()
    C = _dev(C).contiguous()
    f = _dev(f).contiguous()
    S = c_idx.numel()
    T = f.shape[1]
    mu = torch.empty((S, T), dtype=F64, device=f.device)
    check(lib.hgp_emission_means(ptr(C), ptr(f), ptr(c_idx), ptr(f_idx), S, T, ptr(mu), stream_ptr()),
          "hgp_emission_means")
    return mu


def tile_uniform_states(state_of):
    """Per (64-beat tile, cluster): the state shared by all beats of the tile, or -2 (see hgp_tile_uniform_states)."""
    lib = _lib_ready()
    so = _dev(state_of).contiguous()
    N, M = so.shape
    bt = int(lib.hgp_tile_beats())
    out = torch.empty(((N + bt - 1) // bt, M), dtype=I32, device=so.device)
    check(lib.hgp_tile_uniform_states(ptr(so), N, M, ptr(out), stream_ptr()), "hgp_tile_uniform_states")
    return out


def whiten_means(mu, W, factor_of_state=None):
    """nu[s] = W[factor_of_state[s]] mu[s]: the state means in the whitened coordinates the tile kernel works in."""
    lib = _lib_ready()
    mu = _dev(mu).contiguous()
    W = _dev(W).contiguous()
    S, T = mu.shape
    nu = torch.empty_like(mu)
    if factor_of_state is not None:
        factor_of_state = _dev(factor_of_state).to(I32).contiguous()
    check(lib.hgp_whiten_means(ptr(mu), ptr(W), ptr(factor_of_state), S, T, ptr(nu), stream_ptr()), "hgp_whiten_means")
    return nu


def whiten_plan(factor_of_state, max_factors_per_tile=4):
    """Work lists of hgp_whiten_means_tiles for a state table (they depend on factor_of_state only, so a table build on
    unchanged index maps re-uses them): the (tile, factor) pairs of the 64-state tiles that mix at most
    `max_factors_per_tile` factors, and the states of all other tiles.  Returns (items [n, 2] int32, state_list int32)."""
    lib = _lib_ready()
    fos = _dev(factor_of_state).to(torch.int64).contiguous()
    S = fos.numel()
    bt = int(lib.hgp_tile_beats())
    n_tiles = (S + bt - 1) // bt
    pad = torch.full((n_tiles * bt,), -1, dtype=torch.int64, device=fos.device)
    pad[:S] = fos
    pad = pad.view(n_tiles, bt)
    srt = torch.sort(pad, dim=1).values
    new = torch.ones_like(srt, dtype=torch.bool)
    new[:, 1:] = srt[:, 1:] != srt[:, :-1]
    new &= srt >= 0
    fast = new.sum(dim=1) <= max_factors_per_tile                       # [n_tiles]
    sel = new & fast[:, None]
    tiles = torch.arange(n_tiles, device=fos.device)[:, None].expand_as(srt)
    items = torch.stack([tiles[sel], srt[sel]], dim=1).to(I32).contiguous()
    slow_states = torch.nonzero(~fast[torch.arange(S, device=fos.device) // bt]).reshape(-1).to(I32).contiguous()
    return items, slow_states


def whiten_means_tiles(mu, W, Wpacked, factor_of_state, plan):
    """nu = whiten_means(mu, W, factor_of_state) computed on the tile kernel's pipeline (see hgp_whiten_means_tiles)."""
    lib = _lib_ready()
    mu = _dev(mu).contiguous()
    S, T = mu.shape
    items, slow = plan
    W = _dev(W).contiguous()
    factor_of_state = _dev(factor_of_state).to(I32).contiguous()
    nu = torch.empty_like(mu)
    check(lib.hgp_whiten_means_tiles(ptr(mu), ptr(W), ptr(Wpacked), ptr(factor_of_state), S, T,
                                     ptr(items) if items.numel() else None, items.shape[0],
                                     ptr(slow) if slow.numel() else None, slow.numel(), ptr(nu), stream_ptr()),
          "hgp_whiten_means_tiles")
    return nu


def score_tiles(Y, nu, Wpacked, state_of, factor_of_cluster, out=None, tile_state=None, mu_sm=None, snr_state_of=None,
                snr_out=None):
    """Tensor-core emission scores of one lead plane; nu = whiten_means(mu, W, factor_of_state)."""
    lib = _lib_ready()
    N, T = Y.shape
    M = state_of.shape[1]
    q = out if out is not None else torch.empty((N, M), dtype=F64, device=Y.device)
    if tile_state is None:
        tile_state = tile_uniform_states(state_of)
    check(lib.hgp_score_tiles(ptr(Y), N, T, ptr(nu), ptr(Wpacked), ptr(state_of), ptr(tile_state), ptr(factor_of_cluster),
                              M, ptr(q), ptr(mu_sm), ptr(snr_state_of), ptr(snr_out), stream_ptr()), "hgp_score_tiles")
    return q


def score_blocks(Y, nu, W, state_of, factor_of_cluster, out=None):
    """Tensor-core emission scores of one lead plane for beats longer than the tile kernel's 256 samples; plain factors
    W [F, T, T], nu = whiten_means(mu, W, factor_of_state)."""
    lib = _lib_ready()
    N, T = Y.shape
    M = state_of.shape[1]
    q = out if out is not None else torch.empty((N, M), dtype=F64, device=Y.device)
    check(lib.hgp_score_blocks(ptr(Y), N, T, ptr(nu), ptr(W), ptr(state_of), ptr(factor_of_cluster), M, ptr(q),
                               stream_ptr()), "hgp_score_blocks")
    return q


def score_pairs(Y, mu, W, state_of, factor_of_state, pair_n=None, pair_m=None, out=None):
    lib = _lib_ready()
    N, T = Y.shape
    M = state_of.shape[1]
    q = out if out is not None else torch.zeros((N, M), dtype=F64, device=Y.device)
    n_pairs = N * M if pair_n is None else pair_n.numel()
    check(lib.hgp_score_pairs(ptr(Y), N, T, ptr(mu), ptr(W), ptr(state_of), ptr(factor_of_state), M, ptr(pair_n),
                              ptr(pair_m), n_pairs, ptr(q), stream_ptr()), "hgp_score_pairs")
    return q


def group_plan(state_of, factor_of_state, pair_n=None, pair_m=None, max_pairs=None):
    """Work lists of hgp_score_groups: the (beat, cluster) pairs with a state, sorted by factor and cut into chunks of at
    most hgp_score_groups_max_pairs() pairs that share one.  Depends on the index maps only (re-usable across sweeps).
    Returns dict(pair_n, pair_m, chunk_start, n_chunks, invalid) -- invalid: flat indices into q of pairs without a state."""
    lib = _lib_ready()
    so = _dev(state_of).to(I32).contiguous()
    N, M = so.shape
    dev = so.device
    if pair_n is None:
        flat = torch.arange(N * M, device=dev)
    else:
        flat = _dev(pair_n).long() * M + _dev(pair_m).long()
    st = so.reshape(-1)[flat].long()
    ok = st >= 0
    invalid = flat[~ok]
    flat, st = flat[ok], st[ok]
    fac = st if factor_of_state is None else _dev(factor_of_state).long()[st]
    order = torch.argsort(fac, stable=True)
    flat, fac = flat[order], fac[order]
    n = flat.numel()
    mp = int(lib.hgp_score_groups_max_pairs())
    if n == 0:
        z = torch.zeros(1, dtype=I32, device=dev)
        return dict(pair_n=z, pair_m=z, chunk_start=z, n_chunks=0, invalid=invalid, max_pairs=mp)
    if max_pairs is None:
        # chunks of 16 unless most pairs sit in groups that fill more than two 8-pair tiles: an idle tile still costs
        # tensor-pipe time, a second chunk of the same factor costs a (mostly L2-served) second read of it
        sizes = torch.bincount(fac)
        in_large = sizes[sizes > 16].sum()
        max_pairs = mp if int(in_large) * 2 > n else 16
    mp = int(max_pairs)
    idx = torch.arange(n, device=dev)
    new_grp = torch.ones(n, dtype=torch.bool, device=dev)
    new_grp[1:] = fac[1:] != fac[:-1]
    grp_first = torch.cummax(torch.where(new_grp, idx, torch.zeros_like(idx)), dim=0).values      # start of each pair's group
    starts = torch.nonzero(((idx - grp_first) % mp) == 0).reshape(-1)
    chunk_start = torch.cat([starts, torch.tensor([n], device=dev)]).to(I32).contiguous()
    return dict(pair_n=(flat // M).to(I32).contiguous(), pair_m=(flat % M).to(I32).contiguous(), chunk_start=chunk_start,
                n_chunks=int(starts.numel()), invalid=invalid, max_pairs=mp)


def score_groups(Y, mu, W, state_of, factor_of_state, plan, out=None):
    """score_pairs for a whole plane whose pairs are grouped by factor (hgp_score_groups; plan = group_plan(...))."""
    lib = _lib_ready()
    N, T = Y.shape
    M = state_of.shape[1]
    q = out if out is not None else torch.zeros((N, M), dtype=F64, device=Y.device)
    check(lib.hgp_score_groups(ptr(Y), N, T, ptr(mu), ptr(W), ptr(state_of), ptr(factor_of_state), M, ptr(plan["pair_n"]),
                               ptr(plan["pair_m"]), ptr(plan["chunk_start"]), plan["n_chunks"], int(plan["max_pairs"]), ptr(q), stream_ptr()),
          "hgp_score_groups")
    if plan["invalid"].numel():
        q.view(-1)[plan["invalid"]] = 0.0            # "cluster has no members" (GPI_model.py:494-495)
    return q


def snr_states(Y, mu_sm, snr_state_of, out=None):
    lib = _lib_ready()
    N, T = Y.shape
    M = snr_state_of.shape[1]
    snr = out if out is not None else torch.empty((N, M), dtype=F64, device=Y.device)
    check(lib.hgp_snr_states(ptr(Y), N, T, ptr(mu_sm), ptr(snr_state_of), M, ptr(snr), stream_ptr()), "hgp_snr_states")
    return snr


def mean_beat(Y):
    """Mean over the beats of one lead plane Y [N, T] -> [T] (fixed summation order)."""
    lib = _lib_ready()
    Y = _dev(Y).contiguous()
    N, T = Y.shape
    out = torch.empty(T, dtype=F64, device=Y.device)
    work = torch.empty(int(lib.hgp_mean_beat_work_doubles(T)), dtype=F64, device=Y.device)
    check(lib.hgp_mean_beat(ptr(Y), N, T, ptr(out), ptr(work), stream_ptr()), "hgp_mean_beat")
    return out


def lead_weights(q_lnm, snr_lnm=None, lead_w=None):
    """q, snr: [L, N, M].  Returns (qbar [N,M], e [N,M], w [N,L], any_inf flag tensor)."""
    lib = _lib_ready()
    L, N, M = q_lnm.shape
    dev = q_lnm.device
    qbar = torch.empty((N, M), dtype=F64, device=dev)
    e = torch.empty((N, M), dtype=F64, device=dev)
    w = torch.empty((N, L), dtype=F64, device=dev)
    flags = torch.zeros(1, dtype=I32, device=dev)
    check(lib.hgp_lead_weights(ptr(q_lnm), ptr(snr_lnm), ptr(lead_w), N, M, L, ptr(qbar), ptr(e), ptr(w), ptr(flags),
                               stream_ptr()), "hgp_lead_weights")
    return qbar, e, w, flags


class HmmResult:
    __slots__ = ("alpha", "beta", "marg", "z", "zpair", "boundary_out", "rounds", "workspace")


def hmm_smooth(e, pi, PiT, Pi, Pc, boundary_in=None, has_prev=False, has_next=False, workspace=None, prev=None):
    """HMM smoothing + hard responsibilities of one beat slice.  prev: the HmmResult of an earlier call on the SAME e
    with another boundary_in -- its arrays are repaired in place (hgp_hmm_resmooth) instead of scanning the slice again."""
    lib = _lib_ready()
    N, K = e.shape
    dev = e.device
    if prev is not None:
        r = prev
        workspace = prev.workspace
        fn, name = lib.hgp_hmm_resmooth, "hgp_hmm_resmooth"
    else:
        r = HmmResult()
        r.alpha = torch.empty((N, K), dtype=F64, device=dev)
        r.beta = torch.empty((N, K), dtype=F64, device=dev)
        r.marg = torch.empty(N, dtype=F64, device=dev)
        r.z = torch.empty(N, dtype=I32, device=dev)
        r.zpair = torch.empty(N, dtype=I32, device=dev)
        r.boundary_out = torch.empty(2 * K, dtype=F64, device=dev)
        need = lib.hgp_hmm_workspace_bytes(N, K)
        if workspace is None or workspace.numel() < need:
            workspace = torch.empty(need, dtype=torch.uint8, device=dev)
        r.workspace = workspace
        fn, name = lib.hgp_hmm_smooth, "hgp_hmm_smooth"
    rounds = ctypes.c_int(0)
    check(fn(ptr(e), N, K, ptr(pi), ptr(PiT), ptr(Pi), ptr(Pc), ptr(boundary_in), int(has_prev), int(has_next),
             ptr(r.alpha), ptr(r.beta), ptr(r.marg), ptr(r.z), ptr(r.zpair), ptr(r.boundary_out), ptr(workspace),
             workspace.numel(), ctypes.byref(rounds), stream_ptr()), name)
    r.rounds = rounds.value if prev is None else max(r.rounds, rounds.value)
    return r


def suffstats(z, zpair, qbar, is_first_slice=True, workspace=None):
    """Returns (Nm [K], trans [K,K], start [K], Qem [1]) -- one packed f64 tensor is also returned
    so the multi-GPU path can all-reduce it in one call."""
    lib = _lib_ready()
    N, K = qbar.shape
    dev = qbar.device
    packed = torch.empty(K + K * K + K + 1, dtype=F64, device=dev)
    Nm = packed[:K]
    trans = packed[K:K + K * K].view(K, K)
    start = packed[K + K * K:K + K * K + K]
    Qem = packed[-1:]
    need = lib.hgp_suffstats_workspace_bytes(N, K)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    check(lib.hgp_suffstats(ptr(z), ptr(zpair), ptr(qbar), N, K, int(is_first_slice), ptr(Nm), ptr(trans), ptr(start),
                            ptr(Qem), ptr(workspace), workspace.numel(), stream_ptr()), "hgp_suffstats")
    return Nm, trans, start, Qem, packed


def gemm_batched(A, B, ia=None, ib=None, lowerA=False, transA=False):
    """C[j] = op(A[ia[j]]) @ B[ib[j]] for stacks of T x T matrices (FP64 tensor cores)."""
    lib = _lib_ready()
    A = _dev(A).contiguous()
    B = _dev(B).contiguous()
    T = A.shape[-1]
    J = ia.numel() if ia is not None else (ib.numel() if ib is not None else A.shape[0])
    C = torch.empty((J, T, T), dtype=F64, device=A.device)
    check(lib.hgp_gemm_batched(ptr(A), ptr(ia), ptr(B), ptr(ib), ptr(C), J, T, int(lowerA), int(transA), stream_ptr()),
          "hgp_gemm_batched")
    return C


def qlat_batched(A, Gamma, P, fmean, A_idx, G_idx, P_idx, fprev_idx, fcur_idx, gamma_scale=None, workspace=None):
    """Latent-transition scores (GPI_model.log_lat_error) for J members.  Returns (out [J], info [J])."""
    lib = _lib_ready()
    J = A_idx.numel()
    T = fmean.shape[1]
    dev = fmean.device
    out = torch.empty(J, dtype=F64, device=dev)
    info = torch.zeros(J, dtype=I32, device=dev)
    need = lib.hgp_qlat_workspace_bytes(J, T)
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=dev)
    check(lib.hgp_qlat_batched(ptr(A), ptr(Gamma), ptr(P), ptr(fmean), ptr(A_idx), ptr(G_idx), ptr(P_idx),
                               ptr(fprev_idx), ptr(fcur_idx), ptr(gamma_scale), J, T, ptr(out), ptr(info),
                               ptr(workspace), workspace.numel(), stream_ptr()), "hgp_qlat_batched")
    return out, info


def la_op(op, A, B=None, C=None, T=None):
    """Unit-test hook: run one CTA-level routine (see hgp_la_op).  Operands are modified in place."""
    lib = _lib_ready()
    T = T or A.shape[-1]
    piv = torch.zeros(T, dtype=I32, device=A.device)
    info = torch.zeros(1, dtype=I32, device=A.device)
    check(lib.hgp_la_op(int(op), ptr(A), ptr(B), ptr(C), ptr(piv), T, ptr(info), stream_ptr()), "hgp_la_op")
    return int(info)


def pred_dist_inducing(x_basis, x_post, mu, mu_idx, Sigma, sig_idx, kernel):
    """IterativeGaussianProcess.pred_dist kernel branch (GPI.py:470-501) for a batch of (grid, state) items.
    x_basis [nb]; x_post [n_items, nx] or [nx] (one shared grid); mu [*, nb]; Sigma [*, nb, nb]; kernel = (const,
    length, noise).  Returns (f [n_items, nx], cov [n_items, nx, nx], info [n_items])."""
    lib = _lib_ready()
    xb = _dev(x_basis).to(F64).contiguous()
    xp = _dev(x_post).to(F64).contiguous()
    mu_idx = _dev(mu_idx).to(I32).contiguous()
    sig_idx = _dev(sig_idx).to(I32).contiguous()
    n_items = mu_idx.numel()
    nb = xb.numel()
    shared = xp.dim() == 1
    nx = xp.shape[-1]
    if not shared and xp.shape[0] != n_items:
        raise _lib.HgpError("pred_dist_inducing: one grid per item expected")
    mu = _dev(mu).contiguous()
    Sigma = _dev(Sigma).contiguous()
    f = torch.empty((n_items, nx), dtype=F64, device=xb.device)
    cov = torch.empty((n_items, nx, nx), dtype=F64, device=xb.device)
    info = torch.zeros(n_items, dtype=I32, device=xb.device)
    work = torch.empty(int(lib.hgp_pred_dist_work_doubles(n_items, nb, nx)), dtype=F64, device=xb.device)
    c, ell, noise = (float(v) for v in kernel)
    check(lib.hgp_pred_dist_inducing(ptr(xb), nb, ptr(xp), 0 if shared else nx, nx, ptr(mu), ptr(mu_idx), ptr(Sigma),
                                     ptr(sig_idx), n_items, c, ell, noise, ptr(f), ptr(cov), ptr(work), ptr(info),
                                     stream_ptr()), "hgp_pred_dist_inducing")
    return f, cov, info


def chain_run(descs, T, pipeline=None):
    """descs: list of dicts with the fields of hgp_chain_desc (tensors or scalars).  Runs all chains in one launch.
    pipeline: None = choose (a four-CTA cluster per chain when the chains are few, long and replay whole member steps on
    the shared-memory path; one CTA per chain otherwise), 0 / 1 / 4 = force (see hgp_chain_run_ex)."""
    import ctypes as _ct
    lib = _lib_ready()
    n = len(descs)
    if pipeline is None:
        pipeline = 0
        eligible = (n > 0 and int(lib.hgp_chain_small_path(T)) and all(int(d.get("phases", 0)) in (0, 7, 15) for d in descs))
        forced = os.environ.get("HGP_CHAIN_PIPELINE")          # tests / A-B timing: 0, 1 or 4 wherever the mode applies
        if forced is not None:
            pipeline = int(forced) if eligible else 0
        elif (eligible and min(int(d["n_members"]) for d in descs) >= 4
              and n * int(lib.hgp_chain_pipeline_ctas()) <= torch.cuda.get_device_properties(0).multi_processor_count):
            pipeline = int(lib.hgp_chain_pipeline_ctas())
    arr = (_lib.ChainDesc * n)()
    keep = []
    for i, d in enumerate(descs):
        for name, _ty in _lib.ChainDesc._fields_:
            v = d.get(name, 0)
            if isinstance(v, torch.Tensor):
                keep.append(v)
                v = v.data_ptr()
            setattr(arr[i], name, v)
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    dev = host.cuda()
    with nvtx(f"chain_run[{n} chains, T={T}, pipeline={int(pipeline)}]"):
        check(lib.hgp_chain_run_ex(ptr(dev), n, T, int(pipeline), stream_ptr()), "hgp_chain_run_ex")
    torch.cuda.current_stream().synchronize()
    return keep


def warp_fit_batched(x_model, Y, y_model, noise, lam_s, lam_a, n_ctrl=8, lr=5e-2, train_iter=50, u0=None,
                     grad_scale=None, want_u=False, want_trace=False):
    """Warping_system.compute_warp_batch's optimisation (amtgp_warping_system.py:548-736) for every (beat, template)
    pair.  Y [N, T], y_model [R, T] -> x_warp, y_warp [R, N, T] (+ u [R, N, n_ctrl], loss trace [iters, R, N])."""
    lib = _lib_ready()
    x = _dev(x_model).to(F64).contiguous()
    Y = _dev(Y).to(F64).contiguous()
    Ym = _dev(y_model).to(F64).contiguous().reshape(-1, Y.shape[1])
    N, T = Y.shape
    R = Ym.shape[0]
    dev = Y.device
    xw = torch.empty((R, N, T), dtype=F64, device=dev)
    yw = torch.empty((R, N, T), dtype=F64, device=dev)
    u = torch.empty((R, N, n_ctrl), dtype=F64, device=dev) if want_u else None
    tr = torch.empty((train_iter, R, N), dtype=F64, device=dev) if want_trace else None
    if u0 is not None:
        u0 = _dev(u0).to(F64).contiguous().reshape(R, n_ctrl)
    if grad_scale is not None:
        grad_scale = _dev(grad_scale).to(F64).contiguous()
        if grad_scale.numel() != N:
            raise _lib.HgpError("warp_fit_batched: grad_scale must have one entry per beat")
    check(lib.hgp_warp_fit_batched(ptr(x), T, ptr(Y), N, ptr(Ym), R, ptr(u0), int(n_ctrl), int(train_iter), float(lr),
                                   float(noise), float(lam_s), float(lam_a), ptr(grad_scale), ptr(xw), ptr(yw), ptr(u),
                                   ptr(tr), stream_ptr()), "hgp_warp_fit_batched")
    return xw, yw, u, tr


def warp_prior_factor(x_model, rho, omega, diag_add, normalize_x=True):
    """Cholesky of the warp-prior covariance (WarpPriorAMTGP._ensure_cache, amtgp_warping_system.py:175-194):
    returns (W = L^-1 [1, T, T], logdet 0-d tensor)."""
    lib = _lib_ready()
    x = _dev(x_model).to(F64).contiguous()
    T = x.numel()
    K = torch.empty((1, T, T), dtype=F64, device=x.device)
    check(lib.hgp_warp_prior_cov(ptr(x), T, float(rho), float(omega), float(diag_add), int(bool(normalize_x)), ptr(K),
                                 stream_ptr()), "hgp_warp_prior_cov")
    Lf, info, logdet = chol_batched(K, jitter_scale=0.0, want_logdet=True)
    if int(info[0]):
        raise _lib.HgpError("warp prior covariance is not positive-definite")
    return tri_inverse_batched(Lf), logdet[0]


def warp_prior_score(W, logdet, x_warp):
    """WarpPriorAMTGP.log_sq_error_batch (:223-264): -0.5 (w^T K^-1 w + logdet + T log 2 pi) for rows of x_warp [B, T]."""
    xw = _dev(x_warp).to(F64).contiguous()
    B, T = xw.shape
    dev = xw.device
    zero_mu = torch.zeros((1, T), dtype=F64, device=dev)
    q = score_pairs(xw, zero_mu, W, torch.zeros((B, 1), dtype=I32, device=dev), torch.zeros(1, dtype=I32, device=dev))
    return q[:, 0] - 0.5 * logdet


def mniw_loglik_batched(M, M_idx, Sigma, S_idx, prior_mean, pm_idx, prior_rcov, pr_idx, prior_scale, ps_idx):
    """matrix_normal_inv_wishart.log_likelihood_MNIW (GPI_model.py:1346-1362) for J (parameter, prior) items selected by
    index from stacked [*, T, T] inputs.  Returns (out [J], info [J])."""
    lib = _lib_ready()
    T = M.shape[-1]
    dev = M.device
    idx = [_dev(i).to(I32).contiguous() for i in (M_idx, S_idx, pm_idx, pr_idx, ps_idx)]
    J = idx[0].numel()
    mats = [_dev(m).to(F64).contiguous() for m in (M, Sigma, prior_mean, prior_rcov, prior_scale)]
    out = torch.empty(J, dtype=F64, device=dev)
    info = torch.zeros(J, dtype=I32, device=dev)
    need = int(lib.hgp_mniw_workspace_bytes(J, T))
    ws = torch.empty(need, dtype=torch.uint8, device=dev)
    check(lib.hgp_mniw_loglik_batched(ptr(mats[0]), ptr(idx[0]), ptr(mats[1]), ptr(idx[1]), ptr(mats[2]), ptr(idx[2]),
                                      ptr(mats[3]), ptr(idx[3]), ptr(mats[4]), ptr(idx[4]), J, T, ptr(out), ptr(info),
                                      ptr(ws), need, stream_ptr()), "hgp_mniw_loglik_batched")
    return out, info


def hyperfit_batched(x, Y, noise_bounds, lr=0.1, max_iter=4000, min_iter=1000, atol=1e-4):
    """IterativeGaussianProcess.fit_torch's optimisation (GPI.py:610-698) for every row of Y [n_fits, T] at once.
    Returns a float64 CUDA tensor [n_fits, 8]: outputscale, lengthscale, noise, constant mean, last loss, iterations,
    Cholesky info, 0."""
    lib = _lib_ready()
    x = _dev(x).to(F64).contiguous().reshape(-1)
    Y = _dev(Y).to(F64).contiguous().reshape(-1, x.numel())
    n, T = Y.shape
    out = torch.zeros((n, 8), dtype=F64, device=Y.device)
    work = torch.empty(int(lib.hgp_hyperfit_work_doubles(n, T)), dtype=F64, device=Y.device)
    check(lib.hgp_hyperfit_batched(ptr(x), ptr(Y), n, T, float(noise_bounds[0]), float(noise_bounds[1]), float(lr),
                                   int(max_iter), int(min_iter), float(atol), ptr(out), ptr(work), stream_ptr()),
          "hgp_hyperfit_batched")
    return out


def rbf_kernel_matrix(xa, xb, const, length, diag_add=0.0):
    """ConstantKernel(const) * RBF(length) between two 1-D grids, sklearn's evaluation order; [na, nb] CUDA tensor."""
    lib = _lib_ready()
    xa = _dev(xa).to(F64).contiguous().reshape(-1)
    xb = _dev(xb).to(F64).contiguous().reshape(-1)
    K = torch.empty((xa.numel(), xb.numel()), dtype=F64, device=xa.device)
    check(lib.hgp_rbf_kernel_matrix(ptr(xa), xa.numel(), ptr(xb), xb.numel(), float(const), float(length),
                                    float(diag_add), ptr(K), stream_ptr()), "hgp_rbf_kernel_matrix")
    return K
