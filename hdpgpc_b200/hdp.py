"""E-step surface of the reference's `GPI_HDP` (reference hdpgpc/GPI_HDP.py) on the device:
lead weighting, HMM smoothing, hard responsibilities, sufficient statistics, and the
`cluster_new_batch(learning=False)` entry point -- same method names and argument meaning.

`EStepEngine` is the struct-of-arrays form of one E-step sweep: it owns the
device tables and runs score -> SNR -> lead weights -> HMM -> arg-max -> statistics for a beat
slice; `EStepEngine.sweep()` is what bench.py times.  Beats shard by contiguous time slice over
ranks (`torch.distributed`, NCCL): cluster tables are broadcast, the HMM boundary messages are
all-gathered, the statistics all-reduced.
"""
import os

import numpy as np
import torch
from scipy.special import digamma

from . import ops
from ._lib import HgpError
from .model import GPI_model, snr_state_index, state_index_map

F64 = torch.float64
I32 = torch.int32


# ------------------------------------------------------------------------------------------
# host-side K-sized operands, restated from the reference (these are O(K^2) scalars)
# ------------------------------------------------------------------------------------------
def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x, dtype=np.float64)


def _safe_exp_rows(x):
    with np.errstate(invalid="ignore", over="ignore"):
        e = np.exp(x - np.max(x, axis=1, keepdims=True))
    return np.nan_to_num(e, nan=1e-8)


def compute_trans_A(transTheta, K):
    """GPI_HDP.compute_trans_A (GPI_HDP.py:3527-3535)."""
    tt = _np(transTheta)
    tp = digamma(tt[:K, :K]) - digamma(np.sum(tt[:K, :K + 1], axis=1))[:, None]
    if tp.shape[0] == K:
        return tp
    out = np.full((K, K), -np.inf)
    out[:K - 1, :K - 1] = tp
    return out


def expected_log_pi(transTheta, startTheta, M):
    """startPi / transPi of cluster_new_batch(learning=False) (GPI_HDP.py:2989-2993)."""
    tt, st = _np(transTheta), _np(startTheta)
    transPi = digamma(tt[:M, :M]) - digamma(np.sum(tt[:M, :M + 1], axis=1) + 1e-5)[:, None]
    startPi = digamma(st[:M]) - digamma(np.sum(st[:M + 1]) + 1e-5)
    return startPi, transPi


def hmm_operands(transTheta, startPi, K):
    """pi, PiT (forward, GPI_HDP.py:3574-3585), Pi (backward, :3637-3643), Pc (pair coefficient,
    :3686-3687, no floor).  PiT is NOT Pi transposed: different max-shifts and floors."""
    sp = _np(startPi)
    pi = np.full(K, -np.inf)
    pi[:min(K, sp.shape[0])] = sp[:K]
    pi = np.exp(pi)
    tA = compute_trans_A(transTheta, K)
    PiT = _safe_exp_rows(tA.T.copy())
    PiT[PiT < 1e-6] += 1e-4
    pi[pi < 1e-10] += 1e-4
    Pc = _safe_exp_rows(tA.copy())
    Pi = Pc.copy()
    Pi[Pi < 1e-5] += 1e-4
    return pi, PiT, Pi, Pc


# ------------------------------------------------------------------------------------------
# struct-of-arrays E-step
# ------------------------------------------------------------------------------------------
class LeadTables:
    """Everything one lead needs on the device for a sweep over its beat slice."""

    def __init__(self, Y, mu, W, state_of, factor_of_state, mu_sm, snr_state_of, tile_path=None):
        self.Y = Y                                  # [N, T]
        self.mu = mu                                # [S, T]   emission means C_i f_i
        self.W = W                                  # [F, T, T] whitening factors L^{-1}
        self.state_of = state_of                    # [N, M] int32 (-1: cluster empty)
        self.factor_of_state = factor_of_state      # [S] int32
        self.mu_sm = mu_sm                          # [S2, T]  smoothed latent means (SNR)
        self.snr_state_of = snr_state_of            # [N, M] int32 or None (use_snr False)
        N, T = Y.shape
        M = state_of.shape[1]
        self.N, self.T, self.M = N, T, M
        self.Wpacked = None
        self.nu = None                              # [S, T] whitened means W mu (tile path)
        self.tile_state = None                      # [ceil(N / 64), M] uniform state per (tile, cluster) or -2
        self.factor_of_cluster = None
        self.pair_n = self.pair_m = None
        self._whiten_plan = None
        self._group_plan = None
        self.use_tiles = False
        self.block_path = T > 256                   # beats longer than the tile kernel's 256 rows: hgp_score_blocks
        if tile_path is None:
            tile_path = True
        if tile_path and N > 0:
            self._plan_tiles()

    def _plan_tiles(self):
        """Main factor per cluster = the factor most of its beats use; the rest go to the pair kernel."""
        f_of = self.factor_of_state.long()[self.state_of.clamp_min(0).long()]        # [N, M]
        f_of = torch.where(self.state_of >= 0, f_of, torch.full_like(f_of, -1))
        main = torch.mode(f_of, dim=0).values                                          # [M]
        # a column whose mode is -1 (empty cluster) scores 0 everywhere; any factor index will do
        main_c = main.clamp_min(0)
        exc = (f_of != main_c.unsqueeze(0)) & (f_of >= 0)
        n_exc = int(exc.sum())
        if n_exc * 2 > self.N * self.M:
            return  # mostly per-state covariances (estimation_limit=None regime): pair kernel for everything
        self.use_tiles = True
        self.factor_of_cluster = main_c.to(I32).contiguous()
        self._whiten()
        if not self.block_path:
            self.tile_state = ops.tile_uniform_states(self.state_of)
        if n_exc:
            nz = torch.nonzero(exc)
            self.pair_n = nz[:, 0].to(I32).contiguous()
            self.pair_m = nz[:, 1].to(I32).contiguous()

    def update_states(self, mu, Sigma, add_diag=None, mu_sm=None, W=None, info=None):
        """New cluster states on unchanged index maps (which state scores which beat, `factor_of_state`): what the
        reference re-derives inside every E-step -- one Cholesky per distinct covariance (GPI_model._chol_spd,
        GPI_model.py:83-87, called per group at :521-533) -- as one table build on the device: factorise, invert the
        factors, whiten the state means with them, re-pack the factors for the tensor cores.  mu [S, T], Sigma [F, T, T],
        add_diag [F] (the `first` jitter of the duplicated first-member factors, :527-529).  W / info: factors already
        inverted by the caller (EStepEngine.update_states factorises all leads in one launch)."""
        if W is None:
            _, W, info = ops.cholinv_batched(Sigma, add_diag=add_diag)
        self.W = W
        self.mu = mu
        if mu_sm is not None:
            self.mu_sm = mu_sm
        if self.use_tiles:
            self._whiten()
        return info

    def _whiten(self):
        """Packed factors and whitened state means nu = W_f mu.  Tables of at least a few tiles go through the tile
        kernel's own pipeline (ops.whiten_means_tiles; the work lists depend on factor_of_state only and are kept)."""
        if self.block_path:
            self.nu = ops.whiten_means(self.mu, self.W, self.factor_of_state)
            return
        self.Wpacked = ops.pack_factors(self.W)
        if self.mu.shape[0] < 4 * ops.tile_beats():
            self.nu = ops.whiten_means(self.mu, self.W, self.factor_of_state)
            return
        if self._whiten_plan is None:
            self._whiten_plan = ops.whiten_plan(self.factor_of_state)
        self.nu = ops.whiten_means_tiles(self.mu, self.W, self.Wpacked, self.factor_of_state, self._whiten_plan)

    def score(self, out, snr_out=None):
        """q (and, when snr_out is given, the SNR statistic) for this lead plane."""
        if self.use_tiles:
            fuse = snr_out is not None and self.snr_state_of is not None
            if self.block_path:
                ops.score_blocks(self.Y, self.nu, self.W, self.state_of, self.factor_of_cluster, out=out)
                if fuse:
                    self.snr(snr_out)
            else:
                ops.score_tiles(self.Y, self.nu, self.Wpacked, self.state_of, self.factor_of_cluster, out=out,
                                tile_state=self.tile_state, mu_sm=self.mu_sm if fuse else None,
                                snr_state_of=self.snr_state_of if fuse else None, snr_out=snr_out if fuse else None)
            if self.pair_n is not None:
                ops.score_pairs(self.Y, self.mu, self.W, self.state_of, self.factor_of_state, self.pair_n, self.pair_m,
                                out=out)
        else:
            # per-state covariances (estimation_limit=None): pairs grouped by factor, one factor read per group
            if self.T <= 512 and not os.environ.get("HGP_NO_GROUPS"):
                if self._group_plan is None:
                    self._group_plan = ops.group_plan(self.state_of, self.factor_of_state)
                ops.score_groups(self.Y, self.mu, self.W, self.state_of, self.factor_of_state, self._group_plan, out=out)
            else:
                ops.score_pairs(self.Y, self.mu, self.W, self.state_of, self.factor_of_state, out=out)
            if snr_out is not None and self.snr_state_of is not None:
                self.snr(snr_out)
        return out

    def snr(self, out):
        return ops.snr_states(self.Y, self.mu_sm, self.snr_state_of, out=out)

    def score_slice(self, n0, n1, out, snr_out=None):
        """Tile-path scores (and SNR) of beats [n0, n1) only; n0 must be a multiple of the 64-beat tile.  The exception
        pairs (score_exceptions) need the whole plane and run once all slices are in."""
        tile = ops.tile_beats()
        if not self.use_tiles or n0 % tile:
            raise HgpError("score_slice needs the tile path and a tile-aligned slice start")
        fuse = snr_out is not None and self.snr_state_of is not None
        if self.block_path:
            ops.score_blocks(self.Y[n0:n1], self.nu, self.W, self.state_of[n0:n1], self.factor_of_cluster, out=out[n0:n1])
            if fuse:
                ops.snr_states(self.Y[n0:n1], self.mu_sm, self.snr_state_of[n0:n1], out=snr_out[n0:n1])
            return
        ops.score_tiles(self.Y[n0:n1], self.nu, self.Wpacked, self.state_of[n0:n1], self.factor_of_cluster,
                        out=out[n0:n1], tile_state=self.tile_state[n0 // tile:],
                        mu_sm=self.mu_sm if fuse else None, snr_state_of=self.snr_state_of[n0:n1] if fuse else None,
                        snr_out=snr_out[n0:n1] if fuse else None)

    def score_exceptions(self, out):
        if self.use_tiles and self.pair_n is not None:
            ops.score_pairs(self.Y, self.mu, self.W, self.state_of, self.factor_of_state, self.pair_n, self.pair_m, out=out)


def sharded_hmm_exchange(smooth, K, rank, world, group, device, min_rounds=2):
    """Exact HMM smoothing over rank-sharded beats.  Every rank scans its slice from
    a guessed boundary message; the boundary messages (2K doubles per rank: alpha of the slice's last
    beat, beta (.) e of its first beat) are all-gathered; slices whose incoming message changed are
    scanned again; repeat until no boundary moves BITWISE.  The recursion is a deterministic function of
    the incoming message, so the fixed point equals the sequential scan bit for bit.

    smooth(boundary_in[2K], has_prev, has_next, prev) -> result with .boundary_out[2K] (device tensor); prev is the
    previous round's result, to be repaired in place from the new boundary instead of scanning the slice again.
    min_rounds: rounds before the first convergence check (every rank with a neighbour starts from a guessed message, so
    round 1 can never be the fixed point unless world == 1)."""
    dist = torch.distributed
    has_prev, has_next = rank > 0, rank < world - 1
    bin_ = torch.empty(2 * K, dtype=F64, device=device)
    bin_[:K] = 1.0 / K                     # guess for alpha of the previous rank's last beat
    bin_[K:] = 1.0                         # guess for (beta . e) of the next rank's first beat
    gathered_flat = torch.empty(world * 2 * K, dtype=F64, device=device)   # 1-D: accepted by nccl and gloo alike
    gathered = gathered_flat.view(world, 2 * K)
    rounds = 0
    hm = None
    while True:
        hm = smooth(bin_, has_prev, has_next, hm)
        dist.all_gather_into_tensor(gathered_flat, hm.boundary_out.contiguous(), group=group)
        new_in = bin_.clone()
        if has_prev:
            new_in[:K] = gathered[rank - 1, :K]
        if has_next:
            new_in[K:] = gathered[rank + 1, K:]
        changed = (new_in.view(torch.int64) != bin_.view(torch.int64)).any().to(torch.int32).reshape(1)
        rounds += 1
        # The first exchange always moves the boundary away from the guess, so its verdict is known without asking:
        # the flag is reduced and read on the host (one synchronisation) from the second round on only.
        if rounds >= min_rounds:
            dist.all_reduce(changed, op=dist.ReduceOp.MAX, group=group)
            if int(changed) == 0:
                return hm, rounds
        bin_ = new_in
        if rounds > world + 2:
            raise HgpError("sharded HMM boundary exchange did not converge")


class EStepEngine:
    """One E-step sweep over a (slice of a) beat sequence:  q[N,M,L] -> lead weights -> q-bar ->
    HMM forward/backward -> arg-max resp / respPair -> N_m, startStateCount, transStateCount, Q_em
    (= reference cluster_new_batch(learning=False), GPI_HDP.py:2975-3001, + the count lines :890-892)."""

    def __init__(self, leads, transTheta, startTheta, lead_w=None, group=None, sharded=None):
        """Sharding is OPT-IN: pass `group=` (a torch.distributed process group) or `sharded=True` (the default group) when
        the local beats are ONE CONTIGUOUS TIME SLICE of a sequence cut over the ranks in rank order.  Only then are HMM
        boundary messages exchanged, `startStateCount` taken from rank 0 and the statistics all-reduced.  An initialised
        process group alone changes nothing: every rank then holds (and smooths) a whole sequence of its own."""
        self.leads = leads
        self.L = len(leads)
        self.N, self.M = leads[0].N, leads[0].M
        self.device = leads[0].Y.device
        self.group = group
        self.rank, self.world = 0, 1
        if sharded is None:
            sharded = group is not None
        if sharded:
            if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
                raise HgpError("EStepEngine(sharded=True) needs an initialised torch.distributed process group")
            self.rank = torch.distributed.get_rank(group)
            self.world = torch.distributed.get_world_size(group)
        self.lead_w = lead_w
        self.use_snr = leads[0].snr_state_of is not None
        if not self.use_snr and lead_w is None:
            self.lead_w = torch.full((self.N, self.L), 1.0 / self.L, dtype=F64, device=self.device)
        self.set_hdp(transTheta, startTheta)
        dev = self.device
        self.q = torch.zeros((self.L, self.N, self.M), dtype=F64, device=dev)
        self.snr = torch.zeros((self.L, self.N, self.M), dtype=F64, device=dev) if self.use_snr else None
        self._hmm_ws = torch.empty(ops._lib.load().hgp_hmm_workspace_bytes(max(self.N, 1), self.M), dtype=torch.uint8,
                                   device=dev)
        self._stat_ws = torch.empty(ops._lib.load().hgp_suffstats_workspace_bytes(self.N, self.M), dtype=torch.uint8,
                                    device=dev)
        self.hmm_rounds = 0
        self.boundary_rounds = 0

    def set_hdp(self, transTheta, startTheta):
        M = self.M
        startPi, _ = expected_log_pi(transTheta, startTheta, M)
        pi, PiT, Pi, Pc = hmm_operands(transTheta, startPi, M)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)
        self.pi, self.PiT, self.Pi, self.Pc = up(pi), up(PiT), up(Pi), up(Pc)

    def update_states(self, tables):
        """Table build for new cluster states (see LeadTables.update_states): `tables[ld]` = dict(mu, Sigma, add_diag
        [, mu_sm]).  Raises LinAlgError for a covariance that is not positive definite (LAPACK-style info per matrix)."""
        # the factorisations of all leads in ONE launch each: a CTA per 256 x 256 factor is latency-bound and needs few
        # registers, so the factors of the second lead ride along on the same SMs instead of waiting for a second wave
        Sig = torch.cat([t["Sigma"] for t in tables], dim=0)
        add = None
        if any(t.get("add_diag") is not None for t in tables):
            add = torch.cat([t["add_diag"] if t.get("add_diag") is not None
                             else torch.zeros(t["Sigma"].shape[0], dtype=F64, device=Sig.device) for t in tables])
        with ops.nvtx("table_build"):
            _, W_all, info = ops.cholinv_batched(Sig, add_diag=add)
            off = 0
            for tb, t in zip(self.leads, tables):
                F = t["Sigma"].shape[0]
                tb.update_states(t["mu"], None, None, t.get("mu_sm"), W=W_all[off:off + F], info=info[off:off + F])
                off += F
        bad = torch.count_nonzero(info).reshape(1)
        self._table_info = bad          # checked lazily (check_tables) so that the build stays asynchronous
        return self

    def check_tables(self):
        bad = getattr(self, "_table_info", None)
        if bad is not None and int(bad.sum()):
            from .model import LinAlgError
            raise LinAlgError("linalg.cholesky: a cluster covariance is not positive-definite")

    # -- pieces --
    def score_all(self):
        for ld, tb in enumerate(self.leads):
            with ops.nvtx(f"score[lead {ld}]"):
                tb.score(self.q[ld], self.snr[ld] if self.use_snr else None)
        return self.q, self.snr

    def responsibilities(self):
        with ops.nvtx("lead_weights"):
            qbar, e, w, flags = ops.lead_weights(self.q, self.snr if self.use_snr else None, self.lead_w)
        with ops.nvtx("hmm_smooth"):
            if self.world == 1:
                hm = ops.hmm_smooth(e, self.pi, self.PiT, self.Pi, self.Pc, workspace=self._hmm_ws)
                self.boundary_rounds = 0
            else:
                hm = self._hmm_sharded(e)
        self.hmm_rounds = hm.rounds
        return qbar, e, w, hm

    def _hmm_sharded(self, e):
        smooth = lambda bin_, has_prev, has_next, prev=None: ops.hmm_smooth(
            e, self.pi, self.PiT, self.Pi, self.Pc, boundary_in=bin_, has_prev=has_prev, has_next=has_next,
            workspace=self._hmm_ws, prev=prev)
        hm, rounds = sharded_hmm_exchange(smooth, self.M, self.rank, self.world, self.group, self.device)
        self.boundary_rounds = rounds
        return hm

    def statistics(self, qbar, hm):
        Nm, trans, start, Qem, packed = ops.suffstats(hm.z, hm.zpair, qbar, is_first_slice=(self.rank == 0),
                                                      workspace=self._stat_ws)
        if self.world > 1:
            torch.distributed.all_reduce(packed, group=self.group)
        return dict(Nm=Nm, transStateCount=trans, startStateCount=start, Q_em=Qem, packed=packed)

    @staticmethod
    def slice_bounds(N, n_slices, growth=1.0, tile=None):
        """Tile-aligned slices of N beats whose sizes grow geometrically (growth = 1: equal slices).  Scoring a beat takes
        several times longer than copying it over PCIe, so a slice `growth` times longer than the one before still
        arrives under the scoring of its predecessor, while the exposed copy of the first slice shrinks."""
        tile = tile or ops.tile_beats()
        tiles = -(-N // tile)
        n_slices = max(1, min(int(n_slices), tiles))
        w = np.cumsum([float(growth) ** k for k in range(n_slices)])
        cuts = sorted({min(tiles, max(1, int(round(tiles * c / w[-1])))) for c in w[:-1]} | {tiles})
        edges = [0] + [min(N, c * tile) for c in cuts]
        return [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]

    def tune_slices(self, h2d_gbs, sweep_ms, head_start_ms=0.0, slice_overhead_ms=0.2, safety=1.15):
        """Choose the slice schedule of `sweep_from_host` from measured rates: the host-to-device rate this rank gets
        (GB/s, with every rank of the node copying at once), the time of a device-resident sweep, and the head start the
        copies have over the scoring (a table build issued before the call).  The candidates (2 to 8 slices, growth 1.25
        to 4) are played through a two-resource timeline -- a slice is scored when it has arrived and its predecessor is
        done -- with the copy rate derated by `safety` and `slice_overhead_ms` per slice for the extra launches and grid
        tails; the fastest wins.  With eight ranks sharing the host the copy rate halves (55 -> 29 GB/s per rank on the
        8 x B200 box): the 1 : 4 : 16 : 64 schedule of a lone rank then leaves the last slice waiting for its data
        (end to end 40.5 ms against 31.4 ms alone)."""
        N, L, T = self.N, self.L, self.leads[0].T
        copy_ms = safety * N * T * L * 8 / (float(h2d_gbs) * 1e9) * 1e3
        best = None
        for n in range(1, 9):
            for growth in ((1.0,) if n == 1 else (1.25, 1.5, 2.0, 2.5, 3.0, 4.0)):
                bounds = self.slice_bounds(N, n, growth)
                t_copy, t_score = 0.0, float(head_start_ms)
                for n0, n1 in bounds:
                    f = (n1 - n0) / N
                    t_copy += copy_ms * f
                    t_score = max(t_score, t_copy) + float(sweep_ms) * f + slice_overhead_ms
                if best is None or t_score < best[0] - 1e-9:
                    best = (t_score, len(bounds), growth)
        self._slice_cfg = (best[1], best[2])
        return self._slice_cfg

    def sweep_from_host(self, Y_host, n_slices=None, growth=None):
        """The end-to-end public call: beats arrive in host memory (pinned for an asynchronous copy) in the reference's
        [N, T, L] layout (tests/test_offline.py:31), labels and statistics go back to the host.  The beats are cut into
        tile-aligned slices (`slice_bounds`); the host-to-device copy of slice k+1 runs on a copy stream under the
        scoring of slice k.  Slice schedule: the arguments, else what `tune_slices` chose, else 4 slices growing 4x."""
        cfg = getattr(self, "_slice_cfg", None) or (4, 4.0)
        n_slices = cfg[0] if n_slices is None else n_slices
        growth = cfg[1] if growth is None else growth
        N, L, T = self.N, self.L, self.leads[0].T
        if tuple(Y_host.shape) != (N, T, L):
            raise HgpError(f"sweep_from_host: expected beats of shape {(N, T, L)}, got {tuple(Y_host.shape)}")
        if not all(tb.use_tiles for tb in self.leads):
            raise HgpError("sweep_from_host needs the tile path (one shared covariance per cluster for most beats)")
        dev = self.device
        if getattr(self, "_stage", None) is None:
            self._stage = torch.empty((N, T, L), dtype=F64, device=dev)
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._Yp = torch.empty((L, N, T), dtype=F64, device=dev)
            self._copy_stream.wait_stream(torch.cuda.current_stream())    # the fresh staging block may be a recycled one
            self._stage_free = None
        for ld, tb in enumerate(self.leads):
            tb.Y = self._Yp[ld]
        bounds = self.slice_bounds(N, n_slices, growth)
        main = torch.cuda.current_stream()
        # The copies wait for the previous call's last reader of the staging buffer only -- not for whatever else is queued
        # on the compute stream: a table build (update_states) issued just before this call then runs UNDER the first
        # copies instead of in front of them.
        if getattr(self, "_stage_free", None) is not None:
            self._copy_stream.wait_event(self._stage_free)
        events = []
        with torch.cuda.stream(self._copy_stream):
            for n0, n1 in bounds:
                self._stage[n0:n1].copy_(Y_host[n0:n1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
                events.append(ev)
        for (n0, n1), ev in zip(bounds, events):
            main.wait_event(ev)
            ops.pack_leads_slice(self._stage[n0:n1], self._Yp, n0)
            if n1 == N:
                self._stage_free = torch.cuda.Event()
                self._stage_free.record(main)
            for ld, tb in enumerate(self.leads):
                tb.score_slice(n0, n1, self.q[ld], self.snr[ld] if self.use_snr else None)
        for ld, tb in enumerate(self.leads):
            tb.score_exceptions(self.q[ld])
        qbar, e, w, hm = self.responsibilities()
        st = self.statistics(qbar, hm)
        st.update(z_host=hm.z.cpu(), stats_host=st["packed"].cpu(), z=hm.z, zpair=hm.zpair)
        return st

    def sweep(self):
        self.score_all()
        qbar, e, w, hm = self.responsibilities()
        st = self.statistics(qbar, hm)
        st.update(qbar=qbar, e=e, w=w, z=hm.z, zpair=hm.zpair, alpha=hm.alpha, beta=hm.beta, marg=hm.marg)
        return st


# ------------------------------------------------------------------------------------------
# the reference's method surface
# ------------------------------------------------------------------------------------------
class GPI_HDP:
    """E-step methods of the reference GPI_HDP operating on device-resident models.

    gpmodels[ld][m] are hdpgpc_b200.GPI_model objects (e.g. GPI_model.from_reference(ref_gp))."""

    def __init__(self, gpmodels, transTheta, startTheta, snr_norm=None, use_snr=True, device="cuda"):
        ops._lib.require_cuda()          # fail loudly: there is no CPU path
        self.gpmodels = gpmodels
        self.n_outputs = len(gpmodels)
        self.M = len(gpmodels[0])
        self.transTheta = _np(transTheta)
        self.startTheta = _np(startTheta)
        self.use_snr = use_snr
        self.device = torch.device(device)
        self.snr_norm = None if snr_norm is None else self._t(snr_norm)
        self.last_engine = None

    def _t(self, x):
        if isinstance(x, torch.Tensor):
            return x.to(device=self.device, dtype=F64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(self.device)

    # ---- GPI_HDP.compute_snr (GPI_HDP.py:732-748) ----
    def compute_snr(self, y_trains, gp):
        Y = self._t(y_trains)
        if Y.dim() == 3:
            Y = Y[:, :, 0].contiguous()
        N = Y.shape[0]
        if not self.use_snr:
            return torch.ones(N, dtype=F64, device=self.device)
        j = snr_state_index(gp.indexes, gp.f_star_sm.shape[0], N).astype(np.int32).reshape(N, 1)
        return ops.snr_states(Y, gp.f_star_sm, torch.from_numpy(j).to(self.device))[:, 0]

    # ---- GPI_HDP.weight_mean (:685-701), 3-D q (N, M, L) ----
    def weight_mean(self, q, snr=None):
        q = self._t(q)
        q_lnm = q.permute(2, 0, 1).contiguous()
        if snr is None:
            qbar, _, _, _ = ops.lead_weights(q_lnm, None, self.snr_norm)
        else:
            qbar, _, _, _ = ops.lead_weights(q_lnm, self._t(snr).permute(2, 0, 1).contiguous(), None)
        return qbar

    # ---- GPI_HDP.compute_snr_ini (:715-730), normalize_snr (:750-756) ----
    def compute_snr_ini(self, y_trains):
        """SNR of every beat against the mean beat of its lead, softmax over leads -> self.snr_norm [N, L]."""
        Y = self._t(y_trains)
        N, T, L = Y.shape
        if not self.use_snr:
            self.snr_norm = torch.ones((N, L), dtype=F64, device=self.device)
            return self.snr_norm
        Yp = ops.pack_leads(Y)
        snr = torch.empty((L, N, 1), dtype=F64, device=self.device)
        zero = torch.zeros((N, 1), dtype=torch.int32, device=self.device)
        for ld in range(L):
            ops.snr_states(Yp[ld], ops.mean_beat(Yp[ld]).reshape(1, T), zero, out=snr[ld])
        self.snr_norm = self.normalize_snr(snr.permute(1, 2, 0))
        return self.snr_norm

    def normalize_snr(self, snr):
        """softmax over leads of the per-beat maximum over clusters; snr [N, M, L] -> [N, L]."""
        s = self._t(snr).permute(2, 0, 1).contiguous()
        _, _, w, _ = ops.lead_weights(torch.zeros_like(s), s, None)
        return w

    # ---- GPI_HDP.full_LDS_elbo (:1838-1864), one_sample=False ----
    def full_LDS_elbo(self, gpmodels, sum_resp, one_sample=False):
        """Sum over the non-empty clusters of return_LDS_param_likelihood() * N_m / N, divided by their number."""
        from .model import lds_param_likelihood_batch
        sr = self._t(sum_resp).reshape(-1)
        live = [i for i in range(len(gpmodels)) if float(sr[i]) > 0]
        if not live:
            return torch.zeros(1, dtype=F64, device=self.device)
        lik = lds_param_likelihood_batch([gpmodels[i] for i in live])
        frac = sr[torch.as_tensor(live, device=self.device)] / torch.sum(sr)
        elb = torch.sum(lik * frac).reshape(1)
        return elb if one_sample else elb / len(live)

    # ---- data terms of GPI_HDP.compute_q_elbo (:1796-1836) and calcELBO_NonlinearTerms (:2682-2700) ----
    def elbo_data_terms(self, q, q_lat, snr, z, zpair):
        """Everything `compute_q_elbo` derives from per-beat quantities, on the device, for the hard assignment (z, zpair):

            Q_em   = sum_n qbar[n, z_n]       (:1805,  q_bas;  qbar = weight_mean(q, snr))
            Q_lat  = sum_n qbar_lat[n, z_n]   (:1806,  elbo_latent; qbar_lat = weight_mean(q_lat, snr))
            frac   = column sums of the lead weights, normalised, times T   (:1813-1818)
            elbo_LDS = sum_ld frac[ld] * full_LDS_elbo(gpmodels[ld], N_m)     (:1819-1821)
            entropy  = calcELBO_NonlinearTerms = H[q] -- exactly 0 for one-hot responsibilities: every term of
                       calc_Hstart / calc_Htable is 1 * log(1 + 1e-30) or 0 * log(1e-30)  (:2690-2700)

        q, q_lat, snr: [N, M, L]; z, zpair int32 [N].  Both sums run in the fixed order of hgp_suffstats, so accept /
        reject comparisons on them are reproducible.  The HDP terms (elbo_Linears, :1025-1074) are K-sized host
        scalars and stay with the caller.  Returns a dict of float64 CUDA tensors."""
        q_l = self._t(q).permute(2, 0, 1).contiguous()
        ql_l = self._t(q_lat).permute(2, 0, 1).contiguous()
        s_l = None if snr is None else self._t(snr).permute(2, 0, 1).contiguous()
        lw = None if snr is not None else self.snr_norm
        qbar, _, w, _ = ops.lead_weights(q_l, s_l, lw)
        qbar_lat, _, _, _ = ops.lead_weights(ql_l, s_l, lw)
        z = z.to(device=self.device, dtype=I32).contiguous()
        zpair = zpair.to(device=self.device, dtype=I32).contiguous()
        Nm, _, _, Q_em, _ = ops.suffstats(z, zpair, qbar)
        Q_em = Q_em.clone()
        _, _, _, Q_lat, _ = ops.suffstats(z, zpair, qbar_lat)
        wsum = torch.sum(w, dim=0)
        frac = wsum / torch.sum(wsum) * float(self.gpmodels[0][0].T if self.gpmodels[0] else 1)
        lds = torch.stack([self.full_LDS_elbo(self.gpmodels[ld], Nm).reshape(()) for ld in range(self.n_outputs)])
        return dict(Q_em=Q_em.reshape(()), Q_lat=Q_lat.clone().reshape(()), frac=frac, full_LDS=lds,
                    elbo_LDS=torch.sum(lds * frac), entropy=torch.zeros((), dtype=F64, device=self.device), Nm=Nm)

    # ---- GPI_HDP.LogLik (:632-661), axis=1 ----
    def LogLik(self, logSoftEv, axis=1):
        x = self._t(logSoftEv)
        c = torch.max(x, dim=axis)[0]
        if torch.any(torch.isinf(c)):
            return x, c
        return x - c.unsqueeze(axis), c

    def _operands(self, pi, K):
        pi_, PiT, Pi, Pc = hmm_operands(self.transTheta, _np(pi), K)
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)
        return up(pi_), up(PiT), up(Pi), up(Pc)

    def _smooth(self, pi, q):
        q = self._t(q)
        N, K = q.shape
        _, e, _, _ = ops.lead_weights(q.reshape(1, N, K), None, torch.ones((N, 1), dtype=F64, device=self.device))
        pi_, PiT, Pi, Pc = self._operands(pi, K)
        return ops.hmm_smooth(e, pi_, PiT, Pi, Pc)

    # ---- forward / backward / coupled_state_coef (:3546-3699).  trans_A is ignored exactly as the
    #      reference ignores it (recomputed from self.transTheta, :3580, :3637, :3686) ----
    def forward(self, pi=None, trans_A=None, q=None):
        hm = self._smooth(pi, q)
        self._last = (q, hm)
        return hm.alpha, hm.marg

    def backward(self, trans_A=None, q=None, margprob=None):
        if getattr(self, "_last", None) is not None and self._last[0] is q:
            return self._last[1].beta
        startPi, _ = expected_log_pi(self.transTheta, self.startTheta, q.shape[1])
        return self._smooth(startPi, q).beta

    def hard_assignments(self, pi, q):
        """(z, zpair): arg-max indices of _safe_exp(logresp) and _safe_exp(logrespPair) (:338-350)."""
        hm = self._smooth(pi, q)
        return hm.z, hm.zpair

    def _safe_exp(self, x):
        """GPI_HDP._safe_exp (:338-350): one-hot of the row arg-max (float64 for 2-D, float32 for 3-D)."""
        x = self._t(x)
        if x.dim() == 2:
            y = torch.zeros_like(x)
            y.scatter_(-1, x.argmax(dim=-1, keepdim=True), 1.0)
            return y
        N = x.shape[0]
        xf = x.reshape(N, -1)
        y = torch.zeros_like(xf, dtype=torch.float32)
        y.scatter_(1, xf.argmax(dim=-1, keepdim=True), 1.0)
        return y.reshape_as(x)

    # ---- build the struct-of-arrays tables for a batch scored against LAST states (i = -1) or
    #      against the time-indexed states of the training sequence ----
    def build_engine(self, y_trains, mode="last"):
        Y = self._t(y_trains)
        N, T, L = Y.shape
        if L != self.n_outputs:
            raise HgpError("y_trains has a different number of leads than the model")
        Yp = ops.pack_leads(Y)
        leads = []
        for ld in range(L):
            mus, Ws, fos, mus_sm = [], [], [], []
            s_off = f_off = sm_off = 0
            state_of = np.full((N, self.M), -1, dtype=np.int32)
            snr_of = np.zeros((N, self.M), dtype=np.int32)
            for m, gp in enumerate(self.gpmodels[ld]):
                tb = gp.tables()
                nS = tb["mu"].shape[0]
                mu_m = tb["mu"]
                fos_m = tb["factor_of_state"].copy()
                if mode == "last":
                    # log_sq_error(x, y, i=-1): last state; an empty cluster scores against its prior state 0
                    state_of[:, m] = s_off + (nS - 1)
                else:
                    if gp.N > 0:
                        i_vals, first = state_index_map(gp.indexes, N)
                        mu_m = torch.cat([mu_m, mu_m[1:2]], dim=0)
                        fos_m = np.concatenate([fos_m, [tb["first_factor"]]]).astype(np.int32)
                        state_of[:, m] = s_off + np.where(first, nS, i_vals)
                mus.append(mu_m)
                fos.append(fos_m + f_off)
                Ws.append(tb["W"])
                mus_sm.append(gp.f_star_sm)
                snr_of[:, m] = sm_off + snr_state_index(gp.indexes, gp.f_star_sm.shape[0], N)
                s_off += mu_m.shape[0]
                f_off += tb["W"].shape[0]
                sm_off += gp.f_star_sm.shape[0]
            dev = self.device
            leads.append(LeadTables(Yp[ld], torch.cat(mus), torch.cat(Ws), torch.from_numpy(state_of).to(dev),
                                    torch.from_numpy(np.concatenate(fos).astype(np.int32)).to(dev),
                                    torch.cat(mus_sm), torch.from_numpy(snr_of).to(dev) if self.use_snr else None))
        eng = EStepEngine(leads, self.transTheta, self.startTheta)
        self.last_engine = eng
        return eng

    # ---- GPI_HDP.cluster_new_batch(learning=False) (:2975-3001) ----
    def cluster_new_batch(self, x_trains, y_trains, learning=False, it_limit=None, warp=False):
        if learning or warp:
            raise HgpError("cluster_new_batch(learning=True / warp=True) is outside the built hot path")
        if len(self.gpmodels[0]) and self.gpmodels[0][0]._off_grid(x_trains) is not None:
            raise HgpError("cluster_new_batch on grids other than x_basis goes through GPI_model.compute_sq_err_all; "
                           "the fused sweep needs x_train == x_basis")
        eng = self.build_engine(y_trains, mode="last")
        out = eng.sweep()
        self.last_sweep = out
        return out["z"].long()
