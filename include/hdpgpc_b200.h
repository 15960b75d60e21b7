/*
 * hdpgpc_b200 -- C ABI of the B200-native HDP-GPC variational E-step hot path.
 *
 * The reference (AdrianPerezHerrero/HDP-GPC) is pure Python/torch and has no FFI; the drop-in
 * boundary is the Python method surface listed in SURVEY.md section 8b.  Each entry point
 * below names the reference method(s) whose arithmetic it replaces (paths are relative to
 * /root/reference/hdpgpc/hdpgpc/).  INTEGRATION.md shows the ctypes binding and the patch a
 * maintainer of the reference would apply.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host
 *   - matrices are row-major float64, contiguous; indices are int32; sizes are int64/int
 *   - every call takes the CUDA stream (as void*, a cudaStream_t) it must run on and is asynchronous
 *     (exception: hgp_hmm_smooth synchronises the stream once per repair round to read a flag)
 *   - no allocation inside: scratch is passed in, sized by the matching *_workspace_bytes call
 *   - return value: 0 = launched OK, < 0 = -(cudaError_t), > 0 = argument error code (HGP_E_*)
 *   - numerical failure (non-SPD matrix) is reported per matrix in an `info` array like LAPACK:
 *     0 = ok, j+1 = leading minor of order j+1 not positive; the Python layer turns it into the
 *     reference's torch.linalg.LinAlgError behaviour (GPI_model.py:1068-1071)
 */
#ifndef HDPGPC_B200_H
#define HDPGPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HGP_E_BADARG 1
#define HGP_E_UNSUPPORTED 2
#define HGP_E_WORKSPACE 3

/* Library / build information. */
int hgp_version(void);                 /* 100*major + minor */
const char* hgp_build_info(void);      /* arch, nvcc version, build flags */
const char* hgp_last_error(void);      /* text of the last failure on this thread */
/* Number of kernels this library has launched since load (all streams); bench.py reads it. */
int64_t hgp_launch_count(void);

/* ---- beat layout ---------------------------------------------------------------------
 * Reference beats are Y[N, T, L] (tests/test_offline.py:31; sliced per lead as
 * y_trains[:, :, [ld]] at GPI_HDP.py:2901, :2985).  Kernels want one contiguous [N, T] plane
 * per lead: Yp[L, N, T]. */
int hgp_pack_leads(const double* Y_ntl, int64_t N, int T, int L, double* Y_lnt, void* stream);
/* The same for a slice of n beats that has just arrived from the host: Y_ntl points at the slice, Y_lnt_at_slice at the
 * slice's first beat inside plane 0 of the full Yp[L, N, T] (plane_stride = N * T doubles).  Lets the host-to-device
 * copy of the next slice overlap the scoring of this one (EStepEngine.sweep_from_host). */
int hgp_pack_leads_slice(const double* Y_ntl, int64_t n, int T, int L, double* Y_lnt_at_slice, int64_t plane_stride,
                         void* stream);

/* ---- SPD factorisation: GPI_model._chol_spd (GPI_model.py:83-87) -------------------------
 * For f in [0, F):  M = Sigma[f] + add_diag[f] * I   (add_diag may be NULL; the `first` rule
 *                                                     GPI_model.py:271-273, :527-529)
 *                   Lfac[f] = chol( 0.5 (M + M^T) + jitter_scale * max(mean|diag M|, eps) * I )
 * Lfac is lower triangular with zeros above the diagonal.  logdet (may be NULL) receives
 * 2 * sum(log diag L) -- an extra output, never part of the emission score. */
int hgp_chol_batched(const double* Sigma, int64_t F, int T, const double* add_diag, double jitter_scale,
                     double* Lfac, double* logdet, int* info, void* stream);

/* W[f] = Lfac[f]^{-1} (lower triangular, zeros above the diagonal).  With it the Mahalanobis
 * term d^T Sigma^{-1} d of GPI_model.py:109-113 / :280-285 is |W d|^2 (one triangular product;
 * the reference's cholesky_solve does two). */
int hgp_tri_inverse_batched(const double* Lfac, int64_t F, int T, double* W, void* stream);
/* Both at once -- Lfac as hgp_chol_batched, W = Lfac^{-1} -- in one left-looking sweep per matrix on the tensor cores
 * (128 <= T <= 512; other sizes run the two calls above): the table build of a sweep (one factor per cluster covariance). */
int hgp_cholinv_batched(const double* Sigma, int64_t F, int T, const double* add_diag, double jitter_scale,
                        double* Lfac, double* W, double* logdet, int* info, void* stream);

/* Re-order W[F, T, T] into the DMMA-fragment-ordered, k-chunked stream the tile kernel reads
 * with 1-D bulk (TMA) copies.  Bytes per factor: hgp_packed_factor_bytes(T). */
int64_t hgp_packed_factor_bytes(int T);
int hgp_pack_factors(const double* W, int64_t F, int T, double* Wpacked, void* stream);

/* ---- emission scores: GPI_model.compute_sq_err_all (GPI_model.py:488-547),
 *      GPI_model.log_sq_error (:250-286), _gaussian_score_shared_cov (:92-113) ------------------
 * One lead plane.  For every beat n and cluster m:
 *     s = state_of[n*M + m]            (row of `mu`, the emission mean C_i f_i of GPI_model.observe
 *                                       :626-662; -1 => q = 0, "cluster has no members" :494-495)
 *     q[n*M + m] = -0.5 * | W_f (Y[n] - mu[s]) |^2 - 0.5 * T * log(2 pi)        (no log-det)
 *
 * hgp_score_tiles: f = factor_of_cluster[m]; all beats, W in packed form; tensor-core (DMMA) path.  Takes the
 *                  WHITENED state means nu[s] = W_f mu[s] (hgp_whiten_means) and evaluates |W_f Y[n] - nu[s]|^2: the
 *                  matrix product runs on the beats alone (one B operand for all clusters) and the mean is
 *                  subtracted in the epilogue.  tile_state (hgp_tile_uniform_states) tells the kernel which
 *                  (64-beat tile, cluster) items use one state for all their beats.  When snr != NULL the SNR lead
 *                  statistic of hgp_snr_states (below) is launched behind it on the same stream.
 * hgp_score_pairs: explicit list of (n, m) pairs with f = factor_of_state[s]; W in plain form;
 *                  used for the states whose covariance differs from the cluster's shared one
 *                  (per-state Sigma_i when estimation_limit=None, the `first` jitter) and as the
 *                  general path when every state has its own factor. */
int hgp_tile_beats(void);   /* beats per tile of hgp_score_tiles / hgp_tile_uniform_states (64) */
/* tile_state[tile*M + m] = the state index shared by all beats of tile `tile` for cluster m (-1: empty cluster),
 * or -2 when they differ.  tile_state: ceil(N / hgp_tile_beats()) * M ints. */
int hgp_tile_uniform_states(const int* state_of, int64_t N, int M, int* tile_state, void* stream);
/* nu[s] = W[factor_of_state[s]] mu[s] for s in [0, S)  (factor_of_state NULL: factor s). */
int hgp_whiten_means(const double* mu, const double* W, const int* factor_of_state, int64_t S, int T, double* nu,
                     void* stream);
/* The same product for a whole table build on the tile kernel's tensor-core pipeline (T <= 256, Wpacked from
 * hgp_pack_factors): items = n_items (tile, factor) int pairs -- every state s in [64 tile, 64 tile + 64) with
 * factor_of_state[s] == factor receives nu[s] = W_factor mu[s]; state_list = n_list states computed row by row from the
 * plain factors W instead (tiles that mix many factors).  A state must be covered by exactly one of the two lists. */
int hgp_whiten_means_tiles(const double* mu, const double* W, const double* Wpacked, const int* factor_of_state,
                           int64_t S, int T, const int* items, int64_t n_items, const int* state_list, int64_t n_list,
                           double* nu, void* stream);
int hgp_score_tiles(const double* Y, int64_t N, int T, const double* nu, const double* Wpacked,
                    const int* state_of, const int* tile_state, const int* factor_of_cluster, int M, double* q,
                    const double* mu_sm, const int* snr_state_of, double* snr, void* stream);
/* hgp_score_blocks: the same scores for beats longer than the tile kernel holds in registers (T > 256; any T works):
 *                   plain factors W[F, T, T], f = factor_of_cluster[m], whitened means nu as above; one CTA per (64-beat
 *                   tile, cluster), z never stored.  Same reference lines as hgp_score_tiles. */
int hgp_score_blocks(const double* Y, int64_t N, int T, const double* nu, const double* W, const int* state_of,
                     const int* factor_of_cluster, int M, double* q, void* stream);
int hgp_score_pairs(const double* Y, int64_t N, int T, const double* mu, const double* W,
                    const int* state_of, const int* factor_of_state, int M,
                    const int* pair_n, const int* pair_m, int64_t n_pairs, double* q, void* stream);
/* The same scores with the pairs GROUPED BY FACTOR (per-state covariances, estimation_limit = None): pair_n / pair_m
 * sorted so that pairs of one factor are adjacent, chunk_start[c] .. chunk_start[c + 1] (n_chunks + 1 entries) = the pairs
 * of chunk c, at most max_pairs of them (16 or 32 = hgp_score_groups_max_pairs(): two or four 8-pair tiles per chunk; 16
 * when factors rarely score more pairs -- idle tiles still cost tensor-pipe time), all with the same factor (every pair
 * must have a state >= 0).  One CTA streams the chunk's factor once and scores all its pairs on the tensor cores; T <= 512. */
int hgp_score_groups_max_pairs(void);
int hgp_score_groups(const double* Y, int64_t N, int T, const double* mu, const double* W, const int* state_of,
                     const int* factor_of_state, int M, const int* pair_n, const int* pair_m, const int* chunk_start,
                     int64_t n_chunks, int max_pairs, double* q, void* stream);

/* ---- lead weighting: GPI_HDP.compute_snr (GPI_HDP.py:732-748), weight_mean (:685-701),
 *      LogLik (:632-661) ------------------------------------------------------------------
 * snr[n*M + m] = 10 log10( (sum mu^2 + eps) / (sum (mu - y)^2 + eps) ),  mu = mu_sm[snr_state_of[n*M+m]]
 * (the smoothed latent mean f_star_sm[j], j = clip(find_closest_lower(n), 1, len-1)).
 * snr_state_of may be ANY map into mu_sm (-1 => 0).  The tensor-core path (T <= 256, M <= 128) holds one table row per
 * run of equal states per cluster inside a 64-beat tile: 128 rows for M <= 64, 192 for M <= 128 -- always enough for the
 * reference's rule, where an index only advances at the cluster's own members (<= M + 64 runs).  A tile with more runs
 * is summed directly (same result, slower); nothing is ever written out of bounds. */
int hgp_snr_states(const double* Y, int64_t N, int T, const double* mu_sm, const int* snr_state_of,
                   int M, double* snr, void* stream);
/* Mean beat of one lead plane (GPI_HDP.compute_snr_ini, GPI_HDP.py:715-730: the SNR of every beat against the mean
 * beat is hgp_snr_states with this single row as the table; softmax over leads is hgp_lead_weights' w).
 * work: hgp_mean_beat_work_doubles(T) doubles. */
int64_t hgp_mean_beat_work_doubles(int T);
int hgp_mean_beat(const double* Y, int64_t N, int T, double* mean, double* work, void* stream);
/* q, snr: [L, N, M] planes.  w[n, ld] = softmax_ld( max_m snr[ld, n, m] ) (or lead_w[N, L] when snr is
 * NULL: the saved self.snr_norm), qbar[n, m] = sum_ld q[ld, n, m] w[n, ld];
 * e[n, k] = nan_to_num(exp(qn - rowmax(qn)), 1e-8) with qn = qbar - rowmax(qbar), the subtraction
 * skipped for ALL rows when any row max is +-inf (LogLik's early return, :646-648).
 * flags[0]: bit 0 = "some row maximum is +-inf", bit 1 = "some row holds a NaN score" -- torch.max propagates NaN, so
 * every entry of such a row is nan_to_num(NaN) = 1e-8 exactly as in the reference.  wout (may be NULL) receives w[N, L]. */
int hgp_lead_weights(const double* q, const double* snr, const double* lead_w, int64_t N, int M, int L,
                     double* qbar, double* e, double* wout, int* flags, void* stream);

/* ---- HMM smoothing and hard responsibilities: GPI_HDP.forward (GPI_HDP.py:3546-3610),
 *      backward (:3612-3649), coupled_state_coef (:3651-3699), _safe_exp (:338-350) -------------
 * Operands pi[K], PiT[K,K], Pi[K,K], Pc[K,K] are built on the host exactly as the reference does
 * (digamma, max-shift, floors); the device scans.  The scan is chunk-parallel and EXACT: chunks
 * start from a guess, then are repaired from their predecessor's true boundary message until
 * every message is bit-identical to the sequential recursion.
 * alpha, beta: [N, K] outputs (normalised messages); marg[N] (may be NULL) = margPrObs of forward().
 * z[N] = argmax_k log(alpha beta),
 * zpair[N] = argmax over the flattened KxK pair coefficient (row 0 -> 0).
 * boundary_in (may be NULL): alpha_in[K] then beta_in[K] -- messages entering this slice from the
 * neighbouring ranks when the beat sequence is sharded (has_prev / has_next say which are valid).
 * boundary_out[2K]: alpha of the slice's last beat, then (beta (.) e) of the slice's first beat
 * -- what the neighbours need.  Returns the number of repair rounds used in rounds_host. */
int64_t hgp_hmm_workspace_bytes(int64_t N, int K);
int hgp_hmm_smooth(const double* e, int64_t N, int K, const double* pi, const double* PiT, const double* Pi,
                   const double* Pc, const double* boundary_in, int has_prev, int has_next,
                   double* alpha, double* beta, double* marg, int* z, int* zpair, double* boundary_out,
                   void* workspace, int64_t workspace_bytes, int* rounds_host, void* stream);
/* Same outputs as hgp_hmm_smooth for a CHANGED boundary_in, given that alpha / beta / marg / workspace still hold the
 * result of a previous call on the same e: only the chunks the new boundary messages actually alter are scanned again
 * (the boundary chunk first, then onwards until a recomputed message equals the stored one bitwise).  This is the
 * second and later rounds of the sharded boundary exchange: a full re-scan of the slice would double the HMM cost. */
int hgp_hmm_resmooth(const double* e, int64_t N, int K, const double* pi, const double* PiT, const double* Pi,
                   const double* Pc, const double* boundary_in, int has_prev, int has_next,
                   double* alpha, double* beta, double* marg, int* z, int* zpair, double* boundary_out,
                   void* workspace, int64_t workspace_bytes, int* rounds_host, void* stream);

/* ---- sufficient statistics: include_batch (GPI_HDP.py:890-892), compute_q_elbo (:1805) --------
 * Nm[K] = sum_n [z_n = k];  trans[K,K] = sum_n onehot(zpair_n) (counts, exact integers in f64);
 * start[K] = onehot(z_0) if is_first_slice else 0;  Qem[0] = sum_n qbar[n, z_n] (fixed-order sum). */
int64_t hgp_suffstats_workspace_bytes(int64_t N, int K);
int hgp_suffstats(const int* z, const int* zpair, const double* qbar, int64_t N, int K, int is_first_slice,
                  double* Nm, double* trans, double* start, double* Qem,
                  void* workspace, int64_t workspace_bytes, void* stream);

/* ---- emission means: GPI_model.observe (GPI_model.py:626-662) on the basis grid ---------------
 * mu[s] = C[c_idx[s]] @ f[f_idx[s]]  for s in [0, S)   (C: [nC, T, T], f: [nF, T]). */
int hgp_emission_means(const double* C, const double* f, const int* c_idx, const int* f_idx, int64_t S, int T,
                       double* mu, void* stream);

/* ---- latent transition score: GPI_model.compute_q_lat_all (GPI_model.py:549-559) ->
 *      log_lat_error (:288-323) ---------------------------------------------------------------
 * For j in [0, J):  r = f_cur[j] - A[j] f_prev[j];  Lg = _chol_spd(Gamma[j]);
 *   out[j] = -0.5 ( r^T Gamma^{-1} r + tr(A^T Gamma^{-1} A P_prev[j]) ) - 0.5 T log 2 pi
 * computed as |Lg^{-1} r|^2 + sum((X P) (.) X), X = Lg^{-1} A  (10/3 T^3 flops instead of 6.3 T^3).
 * A_idx/G_idx/P_idx/fprev_idx/fcur_idx select rows of the stacked inputs so histories are not copied. */
int64_t hgp_qlat_workspace_bytes(int64_t J, int T);
int hgp_qlat_batched(const double* A, const double* Gamma, const double* P, const double* fmean,
                     const int* A_idx, const int* G_idx, const int* P_idx, const int* fprev_idx,
                     const int* fcur_idx, const double* gamma_scale, int64_t J, int T,
                     double* out, int* info, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- batched T x T product on the FP64 tensor cores: C[j] = op(A[ia[j]]) B[ib[j]] -----------------
 * ia / ib (may be NULL: identity) select rows of the stacked inputs; lowerA skips the zero half of a lower-
 * triangular A; transA uses A^T.  Building block of hgp_qlat_batched and of the chain kernels
 * (covariance propagation A Sigma A^T + Gamma, GPI.py:134; MNIW moments, GPI_model.py:1318-1340). */
int hgp_gemm_batched(const double* A, const int* ia, const double* B, const int* ib, double* C, int64_t J, int T,
                     int lowerA, int transA, void* stream);

/* ---- cluster propagation (chain replay): GPI_model.full_pass_weighted (GPI_model.py:377-406) =
 *      per member Kalman update (IterativeGaussianProcess.posterior, GPI.py:72-151) + pair smoother
 *      (backward_notrange, GPI.py:272-300) + MNIW step (bayesian_new_params, GPI_model.py:966-1101;
 *      matrix_normal_inv_wishart.posterior :1300-1344), then the full RTS pass (backward, GPI.py:240-270).
 * One persistent CTA per (cluster, lead) chain; shared basis grid (x_train == x_basis), dynamic model.
 * Histories are pre-allocated by the caller with capacity n_members + 1; entry 0 of each is the chain's
 * initial state (GPI_model.initial_conditions, GPI_model.py:115-175).  All pointers are device pointers. */
typedef struct hgp_chain_desc {
    int n_members;            /* members to assimilate, ascending beat order */
    int first_is_prior;       /* 1: the first member sees the GP prior: P = cov, f* = 0, R = r_first I (GPI.py:136-139) */
    int annealing;            /* Gamma' += Gamma[0]/N^2, Sigma' += Sigma[0]/N^2 (GPI_model.py:1083-1091) */
    int estimation_limit;     /* <= 0: none */
    double r_first;           /* (c + noise) - c of the fitted kernel (kernel(x) - kernel(x, x), GPI.py:139) */
    const int* member_beats;  /* [n_members] row of Y per member */
    const double* Y;          /* beats plane [N][T] */
    double *f_star, *f_star_sm;                 /* [start_members + n_members + 1][T] */
    double *cov_f, *cov_f_sm, *A, *Gamma, *C, *Sigma;   /* [start_members + n_members + 1][T][T] */
    double *int_m_mean, *int_m_r_cov, *int_scale, *int_n0;   /* MNIW over (A, Gamma): in/out; n0 is a device scalar */
    double *obs_m_mean, *obs_m_r_cov, *obs_scale, *obs_n0;   /* MNIW over (C, Sigma) */
    double* work;             /* >= hgp_chain_work_doubles(T) */
    int* piv;                 /* >= T */
    int* status;              /* [2] out: first member (1-based) whose MNIW factorization failed or 0; parameter sets written */
    /* Online use: continue an existing chain and / or run only some of the per-member phases, so that the reference's
     * three seam calls GPI_model.include_weighted_sample (:353-375), backwards_pair (:705-724) and bayesian_new_params
     * (:966-1115) can be issued one by one (GPI_HDP.include_sample, GPI_HDP.py:2187-2192). */
    int start_members;        /* members already held by the histories before member 0 of this call (0: fresh chain) */
    int start_params;         /* index of the last parameter set (A, Gamma, C, Sigma) already written (0: fresh chain) */
    int phases;               /* bit 0 Kalman update, bit 1 pair smoother, bit 2 MNIW step, bit 3 full RTS pass; 0 = all */
    int reserved_;
    /* Optional (may be NULL), >= hgp_chain_rts_cache_doubles(T, start_members + n_members + 1) doubles: the shared-memory
     * path for small T (hgp_chain_small_path) keeps the smoother gain J_s, the predicted covariance P_s and A m_s of every
     * state it passes on the way forward, so that the full RTS pass re-uses them instead of re-deriving them (the pass
     * visits state s with exactly the parameters and the filtered covariance the pair smoother used). */
    double* rts_cache;
} hgp_chain_desc;
int64_t hgp_chain_desc_bytes(void);
int64_t hgp_chain_work_doubles(int T);
int64_t hgp_chain_rts_cache_doubles(int T, int n_states);   /* 0 when T is outside the shared-memory path */
int hgp_chain_small_path(int T);                            /* 1: hgp_chain_run uses the shared-memory kernel for this T */
int hgp_chain_run(const void* descs_device, int n_chains, int T, void* stream);
/* The same replay with the member step of every chain spread over a thread-block cluster: pipeline = 4 gives each chain
 * four CTAs (Kalman update | pair-smoother gain | MNIW over (A, Gamma) | MNIW over (C, Sigma), see hgp_chain.cu) -- for the
 * few long chains of a real fit, 4 n_chains <= SM count; pipeline = 1 runs the same re-organised step in one CTA;
 * pipeline = 0 is hgp_chain_run.  Shared-memory path only (hgp_chain_small_path(T)), whole member steps only (phases 0, 7
 * or 15 in every descriptor; status[0] = -2 otherwise). */
int hgp_chain_pipeline_ctas(void);
int hgp_chain_run_ex(const void* descs_device, int n_chains, int T, int pipeline, void* stream);
/* Unit-test hook for the CTA-level routines the chain kernel is built from (gemm variants, chol, trsm, LU solve). */
int hgp_la_op(int op, double* A, double* B, double* C, int* piv, int T, int* info, void* stream);

/* ---- emission distribution on a grid other than the basis grid: IterativeGaussianProcess.pred_dist
 *      (GPI.py:457-503, kernel branch :470-501; reached from GPI_model.observe :626-662 when x_train != x_basis)
 * Per item: kernel matrices K_bb, K_bx, K_xx of ConstantKernel(c)*RBF(l) (+ WhiteKernel(noise) on K_xx only),
 * L = chol(K_bb + 1e-4 mean|diag Sigma| I), W = K_bb^{-1} K_bx, f* = W^T mu,
 * cov = mean(diag Sigma) I if Sigma has a constant diagonal, else sym(K_xx - K_bx^T W + W^T Sigma W) + 1e-6 I.
 * x_post: [n_items] grids of nx points (stride x_post_stride doubles; 0 = one shared grid);
 * mu: rows of nb (emission means C f on the basis grid), Sigma: nb x nb stacks, both gathered by index. */
int64_t hgp_pred_dist_work_doubles(int64_t n_items, int nb, int nx);
int hgp_pred_dist_inducing(const double* x_basis, int nb, const double* x_post, int64_t x_post_stride, int nx,
                           const double* mu, const int* mu_idx, const double* Sigma, const int* sig_idx,
                           int64_t n_items, double kernel_const, double kernel_length, double kernel_noise,
                           double* f_out, double* cov_out, double* work, int* info, void* stream);

/* ---- kernel matrix K[i][j] = c exp(-0.5 (xa_i / l - xb_j / l)^2) + diag_add [i == j]: sklearn's
 *      ConstantKernel * RBF (+ WhiteKernel) as the reference evaluates it for the GP prior of a fresh cluster
 *      (GPI_model.py:50, :228-231) and inside the Kalman update (GPI.py:124-139). */
int hgp_rbf_kernel_matrix(const double* xa, int na, const double* xb, int nb, double kernel_const, double kernel_length,
                          double diag_add, double* K, void* stream);

/* ---- one-beat GP hyper-parameter fit, batched over beats: IterativeGaussianProcess.fit_torch (GPI.py:610-770,
 *      ExactGPModel branch: x_train == x_basis) with the gpytorch ExactGP objective restated in closed form
 *      (ConstantMean, ScaleKernel(RBF), GaussianLikelihood with Interval(noise_lo, noise_hi); Adam(lr) on -MLL / T;
 *      at most max_iter iterations, stop once more than min_iter losses exist and the last ten loss differences sum
 *      to within atol of zero, GPI.py:695-698).  One persistent CTA per fit.
 * x [T]; Y [n_fits][T]; out [n_fits][8] = outputscale, lengthscale, noise, constant mean, last loss, iterations,
 * Cholesky info (0 = ok), 0.  work: hgp_hyperfit_work_doubles(n_fits, T) doubles. */
int64_t hgp_hyperfit_work_doubles(int n_fits, int T);
int hgp_hyperfit_batched(const double* x, const double* Y, int n_fits, int T, double noise_lo, double noise_hi, double lr,
                         int max_iter, int min_iter, double atol, double* out, double* work, void* stream);

/* ---- MNIW log-likelihood of LDS parameters under the prior: matrix_normal_inv_wishart.log_likelihood_MNIW
 *      (GPI_model.py:1346-1362), the per-cluster ELBO term of return_LDS_param_likelihood (:459-486) ----
 * For j in [0, J):  L = chol(sym(Sigma[S_idx[j]]) + 1e-8 I);  D = M[M_idx[j]] - prior_mean[pm_idx[j]]
 *   out[j] = -0.5 sum((D R) (.) Sigma^{-1} D) - 0.5 tr(Sigma^{-1} S),  R = prior_rcov[pr_idx[j]], S = prior_scale[ps_idx[j]]
 * computed with W = L^{-1} as sum((X R) (.) X), X = W D, and sum((W S) (.) W) on the FP64 tensor cores. */
int64_t hgp_mniw_workspace_bytes(int64_t J, int T);
int hgp_mniw_loglik_batched(const double* M, const int* M_idx, const double* Sigma, const int* S_idx,
                            const double* prior_mean, const int* pm_idx, const double* prior_rcov, const int* pr_idx,
                            const double* prior_scale, const int* ps_idx, int64_t J, int T, double* out, int* info,
                            void* workspace, int64_t workspace_bytes, void* stream);

/* ---- batched alignment ("warp"): Warping_system.compute_warp_batch (amtgp_warping_system.py:548-736), D = 1.
 * Fits, for every (beat n, representative r) pair, the monotone warp g(t) (n_ctrl control values -> linear expansion ->
 * softplus -> cumulative sum -> [x_min, x_max]) with `train_iter` Adam(lr) steps on
 *   grad_scale[n] * (0.5 |y_n(g) - y_model_r|^2 / (noise + 1e-12) + lam_s |D2 (g - x)|^2 + lam_a |g - x|^2)
 * (grad_scale[n] = weight_n / (sum of weights of the reference's batch + 1e-12): the reference optimises the batch
 * mean, and Adam's epsilon makes the scale observable).  One warp per fit; closed-form gradient.
 * x_model [T]; Y [N][T]; y_model [R][T]; u0 [R][n_ctrl] warm start or NULL (zeros); grad_scale [N] or NULL (ones).
 * Outputs x_warp, y_warp [R][N][T] (g - x and the beat resampled on g); u_out [R][N][n_ctrl] or NULL;
 * loss_trace [train_iter][R][N] (unscaled loss before each step) or NULL. */
int hgp_warp_fit_batched(const double* x_model, int T, const double* Y, int64_t N, const double* y_model, int R,
                         const double* u0, int n_ctrl, int train_iter, double lr, double noise, double lam_s,
                         double lam_a, const double* grad_scale, double* x_warp, double* y_warp, double* u_out,
                         double* loss_trace, void* stream);

/* Covariance of the GP prior on warps: WarpPriorAMTGP._rbf_cov (amtgp_warping_system.py:160-173):
 * K = omega^2 exp(-0.5 (dx / rho)^2) + diag_add I on the grid normalised to [0, 1] when normalize_x != 0.
 * The prior score log_sq_error_batch (:223-264) is then hgp_chol_batched (log-det) + hgp_tri_inverse_batched +
 * hgp_score_pairs with a zero mean. */
int hgp_warp_prior_cov(const double* x, int T, double rho, double omega, double diag_add, int normalize_x, double* K,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HDPGPC_B200_H */
