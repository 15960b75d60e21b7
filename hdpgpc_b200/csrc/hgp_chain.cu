// Cluster propagation ("M-step" chain replay) on the device: reference GPI_model.full_pass_weighted
// (GPI_model.py:377-406) = per member  include_weighted_sample -> IterativeGaussianProcess.posterior (Kalman
// update in Joseph form, GPI.py:72-151)  +  backwards_pair -> backward_notrange (2-step RTS smoother,
// GPI.py:272-300)  +  bayesian_new_params -> matrix_normal_inv_wishart.posterior (MNIW 1-step update,
// GPI_model.py:966-1101, :1300-1344),  then backwards -> backward (full RTS pass, GPI.py:240-270).
//
// The chain is sequential in its members (parameters of step k depend on the smoothed state of step k-1), so
// the parallelism is (cluster, lead) chains x dense T x T algebra: ONE persistent CTA walks one chain through
// all of its steps with the CTA-level routines of hgp_cta_la.cuh (tensor-core GEMMs, blocked factorizations);
// the ~12 T x T temporaries and the state histories stay in L2/HBM, nothing returns to the host between steps.
// Shared basis grid only (x_train == x_basis: every shipped configuration); dynamic model (Gamma != 0).
#include "hgp_common.cuh"
#include "hgp_cta_la.cuh"
#include "hgp_smem_la.cuh"

using namespace hgp;

namespace {

struct Mniw {
    double* m_mean;
    double* m_r_cov;
    double* scale;
    double* n0;     // device scalar
};

// matrix_normal_inv_wishart.posterior for n_k = 1 with zero covariance terms and sse_matrix = I
// (GPI_model.py:1300-1344), in two halves so that a failed factorisation leaves the distribution untouched -- the
// reference catches torch.linalg.LinAlgError around BOTH posteriors and keeps the previous parameters of both
// (GPI_model.py:1068-1071).
//   compute: S2 = y2 y2^T + Sinv -> W1,  part_mean^T -> W3  (W0, W2: scratch; d is only read).  Returns chol info.
//   commit : m_mean, scale, m_r_cov, n0 updated from (W1, W3).
__device__ __noinline__ int mniw_posterior_compute(Mniw d, const double* y1, const double* y2, int T, double* W0, double* W1,
                                                   double* W2, double* W3, LaSmem& sm) {
    const int n = T * T;
    // Ls = chol(sym(m_r_cov) + 1e-2 * max(mean|diag scale|, eps) I)
    const double jitter = 1e-2 * fmax(la_mean_abs_diag(d.scale, T, sm), HGP_EPS);
    la_copy(W0, d.m_r_cov, n);
    la_symmetrize(W0, jitter, T);
    int info = la_chol(W0, T, sm);
    if (info) return info;                               // torch.linalg.cholesky raises here
    // Sinv = cholesky_solve(I, Ls)
    la_set_identity(W1, 1.0, T);
    la_trsm_lower(W0, W1, T, sm);
    la_trsm_lower_trans(W0, W1, T, sm);                 // W1 = Sinv
    // S2 = y2 y2^T + Sinv ; S1 = y1 y2^T + m_mean Sinv
    la_gemm(W2, d.m_mean, 0, W1, 0, T, 1.0, 0.0, nullptr, sm);   // W2 = m_mean Sinv
    la_rank1(W2, 1.0, y1, y2, T);                               // W2 = S1
    la_rank1(W1, 1.0, y2, y2, T);                               // W1 = S2
    // part_mean = cholesky_solve(S1^T, chol(sym(S2) + 1e-8 I))^T
    la_copy(W0, W1, n);
    la_symmetrize(W0, 1e-8, T);
    info = la_chol(W0, T, sm);
    if (info) return info;
    la_transpose(W3, W2, T);
    la_trsm_lower(W0, W3, T, sm);
    la_trsm_lower_trans(W0, W3, T, sm);                 // W3 = part_mean^T
    return 0;
}

__device__ __noinline__ void mniw_posterior_commit(Mniw d, const double* y1, const double* y2, int T, const double* S2,
                                                  const double* partT) {
    const int n = T * T;
    const double n0 = *d.n0;
    // new_m_mean = ((n0 - 2) m_mean + part_mean) / (n0 + 1 - 2) ; new_scale = ((n0 - 2) scale + e e^T) / (n0 - 1)
    const double a = (n0 - 2.0), den = (n0 + 1.0) - 2.0;
    for (int i = threadIdx.x; i < n; i += LA_THREADS) {
        const int r = i / T, c = i % T;
        d.m_mean[i] = (a * d.m_mean[i] + partT[(int64_t)c * T + r]) / den;
        const double e_r = y1[r] - y2[r], e_c = y1[c] - y2[c];
        d.scale[i] = (a * d.scale[i] + e_r * e_c) / den;
        d.m_r_cov[i] = S2[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) *d.n0 = n0 + 1.0;
    __syncthreads();
}

}  // namespace

namespace {

// B <- M^-T B for a matrix that is symmetric positive definite up to rounding -- S = C P C^T + R of the Kalman gain
// (GPI.py:145), P = A S0 A^T + Gamma of the smoother gains (GPI.py:267, :297): Cholesky of its symmetric part and two
// triangular solves (2.1 ms at T = 256) instead of the pivoted LU the reference's `solve` / `inv` imply (3.3 ms); the LU
// stays as the fallback for a matrix that has lost definiteness.  Wk: T x T scratch.
__device__ __noinline__ void chain_spd_solve(double* Wk, const double* M, double* B, int* piv, int T, LaSmem& sm) {
    la_copy(Wk, M, T * T);
    la_symmetrize(Wk, 0.0, T);
    if (la_chol(Wk, T, sm) == 0) {
        la_trsm_lower(Wk, B, T, sm);
        la_trsm_lower_trans(Wk, B, T, sm);
        return;
    }
    la_transpose(Wk, M, T);
    la_lu_factor(Wk, piv, T, sm);
    la_lu_solve(Wk, piv, B, T, sm);
}

__global__ void __launch_bounds__(LA_THREADS)
chain_kernel(const hgp_chain_desc* __restrict__ descs, int T) {
    __shared__ LaSmem sm;
    extern __shared__ double la_dyn[];
    if (threadIdx.x == 0) sm.big = T <= LA_SMEM_T ? la_dyn : nullptr;
    __syncthreads();
    const hgp_chain_desc d = descs[blockIdx.x];
    const int n = T * T;
    const int64_t tt = (int64_t)T * T;
    double* W0 = d.work;
    double* W1 = W0 + tt; double* W2 = W1 + tt; double* W3 = W2 + tt;
    double* W4 = W3 + tt; double* W5 = W4 + tt; double* W6 = W5 + tt; double* W7 = W6 + tt;
    double* v0 = W7 + tt; double* v1 = v0 + T; double* v2 = v1 + T;
    Mniw mi{d.int_m_mean, d.int_m_r_cov, d.int_scale, d.int_n0};
    Mniw mo{d.obs_m_mean, d.obs_m_r_cov, d.obs_scale, d.obs_n0};
    int fail = 0;
    int N = d.start_members;          // members assimilated
    int p = d.start_params;           // index of the last parameter set (A, Gamma, C, Sigma)
    const int phases = d.phases ? d.phases : 15;
    for (int k = 0; k < d.n_members; ++k) {
        const int s = d.start_members + k;    // last state index before this member
        const double* m = d.f_star_sm + (int64_t)s * T;
        const double* Sg = d.cov_f_sm + s * tt;
        const double* A = d.A + p * tt;
        const double* Gm = d.Gamma + p * tt;
        const double* C = d.C + p * tt;
        const double* R = d.Sigma + p * tt;
        const double* y = d.Y + (int64_t)d.member_beats[k] * T;
        double* m_new = d.f_star + (int64_t)(s + 1) * T;
        double* S_new = d.cov_f + (s + 1) * tt;
        // ---------------- Kalman update (GPI.py:104-150) ----------------
        const bool prior = d.first_is_prior && s == 0;
        const bool have_P = (phases & 1) && !prior;                      // W1 = A Sigma A^T + Gamma survives the update
        if (phases & 1) {
        la_gemv(v0, A, m, T, 0.0, nullptr);                              // v0 = A m  (x_basis_mean)
        const double* P;
        if (prior) {
            P = Sg;                                                      // P_t = cov_prior; f* = 0; R = r_first I
            for (int i = threadIdx.x; i < T; i += LA_THREADS) v1[i] = y[i];
            __syncthreads();
            la_gemm(W2, C, 0, P, 0, T, 1.0, 0.0, nullptr, sm);           // W2 = C P
            la_gemm(W3, W2, 0, C, 1, T, 1.0, 0.0, nullptr, sm);          // W3 = C P C^T
            la_add_diag(W3, d.r_first, T);                               //      + R
        } else {
            la_gemm(W0, A, 0, Sg, 0, T, 1.0, 0.0, nullptr, sm);          // W0 = A Sigma
            la_gemm(W1, W0, 0, A, 1, T, 1.0, 1.0, Gm, sm);               // W1 = A Sigma A^T + Gamma = P
            P = W1;
            la_gemv(v2, C, v0, T, 0.0, nullptr);                         // f* = C A m
            for (int i = threadIdx.x; i < T; i += LA_THREADS) v1[i] = y[i] - v2[i];
            __syncthreads();
            la_gemm(W2, C, 0, P, 0, T, 1.0, 0.0, nullptr, sm);           // W2 = C P
            la_gemm(W3, W2, 0, C, 1, T, 1.0, 1.0, R, sm);                // W3 = C P C^T + R = S
        }
        // K_t = solve(S^T, C P^T)^T
        la_gemm(W4, C, 0, P, 1, T, 1.0, 0.0, nullptr, sm);               // W4 = C P^T
        chain_spd_solve(W5, W3, W4, d.piv, T, sm);                       // W4 = S^-T (C P^T)
        la_transpose(W6, W4, T);                                         // W6 = K_t
        la_gemv(m_new, W6, v1, T, 1.0, v0);                              // m+ = A m + K (y - f*)
        // Joseph form: (I - K C) P (I - K C)^T + K R K^T
        la_gemm(W2, W6, 0, C, 0, T, -1.0, 0.0, nullptr, sm);             // W2 = -K C
        la_add_diag(W2, 1.0, T);                                         // W2 = I - K C
        la_gemm(W0, W2, 0, P, 0, T, 1.0, 0.0, nullptr, sm);              // W0 = IKC P   (P may be W1: W0 is free)
        la_gemm(W3, W0, 0, W2, 1, T, 1.0, 0.0, nullptr, sm);             // W3 = IKC P IKC^T
        if (prior) {
            la_gemm(S_new, W6, 0, W6, 1, T, d.r_first, 1.0, W3, sm);     // + r K K^T
        } else {
            la_gemm(W4, W6, 0, R, 0, T, 1.0, 0.0, nullptr, sm);          // W4 = K R
            la_gemm(S_new, W4, 0, W6, 1, T, 1.0, 1.0, W3, sm);           // S+ = W3 + K R K^T
        }
        la_copy(d.f_star_sm + (int64_t)(s + 1) * T, m_new, T);
        la_copy(d.cov_f_sm + (s + 1) * tt, S_new, n);
        }
        N += 1;
        // ---------------- pair smoother (GPI_model.py:705-716, GPI.py:294-299) ----------------
        if ((phases & 2) && N > 1) {
            const double* m0 = d.f_star + (int64_t)s * T;
            const double* S0 = d.cov_f + s * tt;
            if (!have_P) {                                               // the update of this member started from the same
                la_gemm(W0, A, 0, S0, 0, T, 1.0, 0.0, nullptr, sm);      // covariance with the same (A, Gamma): its P is
                la_gemm(W1, W0, 0, A, 1, T, 1.0, 1.0, Gm, sm);           // this P = A S0 A^T + Gamma (two products saved)
            }
            la_gemm(W4, A, 0, S0, 1, T, 1.0, 0.0, nullptr, sm);          // A S0^T
            chain_spd_solve(W5, W1, W4, d.piv, T, sm);                   // P^-T (A S0^T)
            la_transpose(W6, W4, T);                                     // J
            la_gemv(v2, A, m0, T, 0.0, nullptr);                         // A m0
            for (int i = threadIdx.x; i < T; i += LA_THREADS) v2[i] = m_new[i] - v2[i];
            __syncthreads();
            la_gemv(d.f_star_sm + (int64_t)s * T, W6, v2, T, 1.0, m0);   // m0 + J (m1 - A m0)
            if (!(phases & 8)) {                                         // the full RTS pass overwrites this covariance
                la_axpby(W2, 1.0, S_new, -1.0, W1, n);                   // S1 - P
                la_gemm(W0, W6, 0, W2, 0, T, 1.0, 0.0, nullptr, sm);     // J (S1 - P)
                la_gemm(d.cov_f_sm + s * tt, W0, 0, W6, 1, T, 1.0, 1.0, S0, sm);   // S0 + J (S1 - P) J^T
            }
        }
        // ---------------- MNIW step (GPI_model.py:966-1101) ----------------
        if (!(phases & 4)) continue;
        const bool below = d.estimation_limit <= 0 || N < d.estimation_limit;
        if (N > 1 && below) {
            // the reference also factorises P = A cov_ A^T + Gamma here and discards it (:990-998); only a
            // failure of that factorization is observable (keep previous parameters) -- P is SPD by construction
            // both posteriors are computed before either is committed: a failure in one keeps BOTH (W4..W7 are free
            // here: the Kalman / smoother temporaries are dead)
            const double* f1 = d.f_star_sm + (int64_t)(s + 1) * T;
            const double* f0 = d.f_star_sm + (int64_t)s * T;
            const int i1 = mniw_posterior_compute(mi, f1, f0, T, W0, W1, W2, W3, sm);
            const int i2 = i1 ? 0 : mniw_posterior_compute(mo, y, f1, T, W4, W5, W6, W7, sm);
            if (i1 || i2) {
                if (!fail) fail = k + 1;
            } else {
                mniw_posterior_commit(mi, f1, f0, T, W1, W3);
                mniw_posterior_commit(mo, y, f1, T, W5, W7);
            }
        }
        if (below) {
            double* An = d.A + (p + 1) * tt; double* Gn = d.Gamma + (p + 1) * tt;
            double* Cn = d.C + (p + 1) * tt; double* Sn = d.Sigma + (p + 1) * tt;
            const double gi = *d.int_n0, go = *d.obs_n0;
            const double fa = d.annealing ? 1.0 / ((double)N * (double)N) : 0.0;
            for (int i = threadIdx.x; i < n; i += LA_THREADS) {
                An[i] = d.int_m_mean[i];
                Cn[i] = d.obs_m_mean[i];
                const double g = (N > 1) ? d.int_scale[i] * gi / (gi - 2.0) : Gm[i];
                const double sg = (N > 1) ? d.obs_scale[i] * go / (go - 2.0) : R[i];
                Gn[i] = g + fa * d.Gamma[i];          // + Gamma[0] / N^2
                Sn[i] = sg + fa * d.Sigma[i];         // + Sigma[0] / N^2
            }
            __syncthreads();
            p += 1;
        }
    }
    // ---------------- full RTS pass (GPI_model.py:687-703, GPI.py:262-270) ----------------
    // means = f_star[1:], covars = cov_f[1:], A_prior = A[1:], Gamma_prior = Gamma[1:]
    const int Tn = d.start_members + d.n_members;
    const int nA = p;      // len(A[1:])
    if (!(phases & 8)) {
        if (threadIdx.x == 0) { d.status[0] = fail; d.status[1] = p + 1; }
        return;
    }
    if (Tn >= 1) {
        // the last state is its own smoothed value
        la_copy(d.f_star_sm + (int64_t)Tn * T, d.f_star + (int64_t)Tn * T, T);
        la_copy(d.cov_f_sm + Tn * tt, d.cov_f + Tn * tt, n);
    }
    for (int t = Tn - 2; t >= 0; --t) {
        const int ia = (t < nA ? t : nA - 1) + 1;
        const double* A = d.A + ia * tt;
        const double* Gm = d.Gamma + ia * tt;
        const double* mt = d.f_star + (int64_t)(t + 1) * T;
        const double* St = d.cov_f + (t + 1) * tt;
        const double* mn = d.f_star_sm + (int64_t)(t + 2) * T;       // already smoothed successor
        const double* Sn = d.cov_f_sm + (t + 2) * tt;
        la_gemm(W0, A, 0, St, 0, T, 1.0, 0.0, nullptr, sm);
        la_gemm(W1, W0, 0, A, 1, T, 1.0, 1.0, Gm, sm);               // P_t
        // J = covars[t] A^T inv(P): J^T = inv(P)^T (A covars[t]^T) -> solve P^T X = A St^T
        la_gemm(W4, A, 0, St, 1, T, 1.0, 0.0, nullptr, sm);
        chain_spd_solve(W5, W1, W4, d.piv, T, sm);
        la_transpose(W6, W4, T);                                     // J_t
        la_gemv(v2, A, mt, T, 0.0, nullptr);
        for (int i = threadIdx.x; i < T; i += LA_THREADS) v2[i] = mn[i] - v2[i];
        __syncthreads();
        la_gemv(d.f_star_sm + (int64_t)(t + 1) * T, W6, v2, T, 1.0, mt);
        la_axpby(W2, 1.0, Sn, -1.0, W1, n);
        la_gemm(W0, W6, 0, W2, 0, T, 1.0, 0.0, nullptr, sm);
        la_gemm(d.cov_f_sm + (t + 1) * tt, W0, 0, W6, 1, T, 1.0, 1.0, St, sm);
    }
    if (threadIdx.x == 0) { d.status[0] = fail; d.status[1] = p + 1; }
}

// ======================================================================================================
// Small systems (T <= 92: the MIT-BIH beat length is 90): the same chain on the shared-memory routines of
// hgp_smem_la.cuh.  Beyond moving the operands into shared memory the step is re-organised around three facts
// (each leaves the reference's arithmetic in place up to rounding):
//   * every system solved in the chain is symmetric positive definite -- S = C P C^T + R (Kalman gain, GPI.py:145),
//     P = A S0 A^T + Gamma (smoother gains, GPI.py:267, :297) -- so `solve` / `inv` become  chol -> L^{-1} -> two
//     triangular tensor-core products (sl_cholinv) instead of a pivoted LU with latency-bound substitutions;
//   * the pair smoother of member k (GPI_model.py:705-716) works on the covariance the Kalman update of the same
//     member started from, with the same (A, Gamma): its P = A S0 A^T + Gamma and A S0 ARE the update's -- reused;
//   * the full RTS pass (GPI.py:262-270) visits state i with the parameter set and the filtered covariance the pair
//     smoother of the next member used: its gain J_i and its P_i are exactly the ones already computed on the way
//     forward.  They are kept per member (`rts_cache`, 2 T^2 + T doubles per state), which turns a backward step
//     from five products, an LU and a solve into two products.
// One member costs 26 products (about half of them triangular) and 6 factorisations.

// G0 = A S,  P = G0 A^T + Gamma
__device__ __forceinline__ void small_predict(SlCtx& c, double* G0, double* P, const double* A, const double* Gm, const double* S) {
    SlEpi e0;
    sl_gemm(c, G0, A, 0, S, 0, e0);
    SlEpi e1;
    e1.beta = 1.0; e1.D = Gm;
    sl_gemm(c, P, G0, 0, A, 1, e1);
}
// X = G^T M^{-1} for SPD M (M is symmetrised first):  Linv = chol(M)^{-1},  Z = Linv G,  X = Z^T Linv.
// With G = C P this is the Kalman gain K = P C^T S^{-1}; with G = A S0 the smoother gain J = S0 A^T P^{-1}.
__device__ __forceinline__ int small_gain(SlCtx& c, double* X, const double* Mspd, const double* G, double* Linv, double* Z) {
    const int info = sl_cholinv(c, Linv, Mspd, 0.0);
    SlEpi e;
    sl_gemm(c, Z, Linv, 0, G, 0, e, SL_TRI_A);
    sl_gemm(c, X, Z, 1, Linv, 0, e, SL_TRI_B);
    return info;
}

__device__ __forceinline__ double small_mean_abs_diag(const double* A, int T, SlCtx& c) {
    double p = 0.0;
    for (int i = threadIdx.x; i < T; i += SL_THREADS) p += fabs(A[(int64_t)i * T + i]);
    p = warp_sum(p);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) c.red[threadIdx.x >> 5] = p;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < SL_THREADS / 32; ++w) tot += c.red[w];
    __syncthreads();
    return tot / T;
}

// matrix_normal_inv_wishart.posterior, n_k = 1 (GPI_model.py:1300-1344): S2 -> Ga, part_mean -> Gb; d is only read.
__device__ __noinline__ int small_mniw_compute(SlCtx& c, Mniw d, const double* y1, const double* y2, int T, double* Ga,
                                               double* Gb, double* Gc, double* Gd, double* Ge) {
    const int n = T * T;
    const double jitter = 1e-2 * fmax(small_mean_abs_diag(d.scale, T, c), HGP_EPS);
    int info = sl_cholinv(c, Gc, d.m_r_cov, jitter);                    // Ls^{-1}
    if (info) return info;
    SlEpi e;
    sl_gemm(c, Ga, Gc, 1, Gc, 0, e, SL_TRI_A | SL_TRI_B);               // Sinv = Ls^{-T} Ls^{-1}
    SlEpi e1;
    e1.s = 1.0; e1.u = y1; e1.v = y2;
    sl_gemm(c, Gd, d.m_mean, 0, Ga, 0, e1);                             // S1 = m_mean Sinv + y1 y2^T
    sl_invalidate(c, Ga);
    for (int i = threadIdx.x; i < n; i += SL_THREADS) Ga[i] += y2[i / T] * y2[i % T];      // S2 = Sinv + y2 y2^T
    __syncthreads();
    info = sl_cholinv(c, Gc, Ga, 1e-8);                                 // chol(sym(S2) + 1e-8 I)^{-1}
    if (info) return info;
    sl_gemm(c, Ge, Gd, 0, Gc, 1, e, SL_TRI_B);                          // S1 L2^{-T}
    sl_gemm(c, Gb, Ge, 0, Gc, 0, e, SL_TRI_B);                          // part_mean = S1 S2^{-1}
    return 0;
}
__device__ __noinline__ void small_mniw_commit(SlCtx& c, Mniw d, const double* y1, const double* y2, int T, const double* S2,
                                               const double* part) {
    const int n = T * T;
    const double n0 = *d.n0;
    const double a = (n0 - 2.0), den = (n0 + 1.0) - 2.0;
    sl_invalidate(c, d.m_mean); sl_invalidate(c, d.scale); sl_invalidate(c, d.m_r_cov);
    for (int i = threadIdx.x; i < n; i += SL_THREADS) {
        const int r = i / T, cc = i % T;
        d.m_mean[i] = (a * d.m_mean[i] + part[i]) / den;
        const double e_r = y1[r] - y2[r], e_c = y1[cc] - y2[cc];
        d.scale[i] = (a * d.scale[i] + e_r * e_c) / den;
        d.m_r_cov[i] = S2[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) *d.n0 = n0 + 1.0;
    __syncthreads();
}

__global__ void __launch_bounds__(SL_THREADS, 1)
chain_kernel_small(const hgp_chain_desc* __restrict__ descs, int T) {
    __shared__ SlCtx c;
    sl_init(c, T);
    const hgp_chain_desc d = descs[blockIdx.x];
    const int n = T * T;
    const int64_t tt = (int64_t)T * T;
    double* G[12];
    for (int i = 0; i < 12; ++i) G[i] = d.work + i * tt;
    double* v0 = d.work + 12 * tt; double* v1 = v0 + T; double* v2 = v1 + T;
    Mniw mi{d.int_m_mean, d.int_m_r_cov, d.int_scale, d.int_n0};
    Mniw mo{d.obs_m_mean, d.obs_m_r_cov, d.obs_scale, d.obs_n0};
    const int n_tot = d.start_members + d.n_members;
    double* Ph = d.rts_cache;                                            // [n_tot + 1][T][T]  P_s
    double* Jh = Ph ? Ph + (int64_t)(n_tot + 1) * tt : nullptr;          //                    J_s
    double* Vh = Ph ? Jh + (int64_t)(n_tot + 1) * tt : nullptr;          // [n_tot + 1][T]     A m_s
    int fail = 0;
    int N = d.start_members;
    int p = d.start_params;
    const int phases = d.phases ? d.phases : 15;
    for (int k = 0; k < d.n_members; ++k) {
        const int s = d.start_members + k;
        const double* m = d.f_star_sm + (int64_t)s * T;
        const double* Sg = d.cov_f_sm + s * tt;
        const double* A = d.A + p * tt;
        const double* Gm = d.Gamma + p * tt;
        const double* C = d.C + p * tt;
        const double* R = d.Sigma + p * tt;
        const double* y = d.Y + (int64_t)d.member_beats[k] * T;
        double* m_new = d.f_star + (int64_t)(s + 1) * T;
        double* S_new = d.cov_f + (s + 1) * tt;
        const bool prior = d.first_is_prior && s == 0;
        double* Pc = Ph ? Ph + s * tt : G[1];
        double* Am = Vh ? Vh + (int64_t)s * T : v0;
        bool have_P = false;
        // ---------------- Kalman update (GPI.py:104-150) ----------------
        if (phases & 1) {
            sl_gemv(Am, A, m, T, 0.0, nullptr);                          // A m
            const double* P;
            SlEpi e;
            if (prior) {
                P = Sg;                                                  // P_t = cov_prior; f* = 0; R = r_first I
                for (int i = threadIdx.x; i < T; i += SL_THREADS) v1[i] = y[i];
                __syncthreads();
                sl_gemm(c, G[2], C, 0, P, 0, e);                         // C P
                SlEpi es;
                es.diag_add = d.r_first;
                sl_gemm(c, G[3], G[2], 0, C, 1, es);                     // S = C P C^T + r I
            } else {
                small_predict(c, G[0], Pc, A, Gm, Sg);                   // G0 = A Sigma, P = A Sigma A^T + Gamma
                have_P = true;
                P = Pc;
                sl_gemv(v2, C, Am, T, 0.0, nullptr);                     // f* = C A m
                for (int i = threadIdx.x; i < T; i += SL_THREADS) v1[i] = y[i] - v2[i];
                __syncthreads();
                sl_gemm(c, G[2], C, 0, P, 0, e);                         // C P
                SlEpi es;
                es.beta = 1.0; es.D = R;
                sl_gemm(c, G[3], G[2], 0, C, 1, es);                     // S = C P C^T + R
            }
            small_gain(c, G[6], G[3], G[2], G[4], G[5]);                 // K = P C^T S^{-1}
            sl_gemv(m_new, G[6], v1, T, 1.0, Am);                        // m+ = A m + K (y - f*)
            // Joseph form: (I - K C) P (I - K C)^T + K R K^T
            SlEpi ei;
            ei.alpha = -1.0; ei.diag_add = 1.0;
            sl_gemm(c, G[7], G[6], 0, C, 0, ei);                         // I - K C
            sl_gemm(c, G[8], G[7], 0, P, 0, e);                          // (I - K C) P
            sl_gemm(c, G[9], G[8], 0, G[7], 1, e);                       // ... (I - K C)^T
            SlEpi ej;
            ej.beta = 1.0; ej.D = G[9];
            if (prior) {
                ej.alpha = d.r_first;
                sl_gemm(c, S_new, G[6], 0, G[6], 1, ej, 0, d.cov_f_sm + (s + 1) * tt);      // + r K K^T
            } else {
                sl_gemm(c, G[5], G[6], 0, R, 0, e);                      // K R
                sl_gemm(c, S_new, G[5], 0, G[6], 1, ej, 0, d.cov_f_sm + (s + 1) * tt);      // + K R K^T
            }
            sl_copy(d.f_star_sm + (int64_t)(s + 1) * T, m_new, T);
        }
        N += 1;
        // ---------------- pair smoother (GPI_model.py:705-716, GPI.py:294-299) ----------------
        if ((phases & 2) && N > 1) {
            const double* m0 = d.f_star + (int64_t)s * T;
            const double* S0 = d.cov_f + s * tt;
            if (!have_P) {                                               // phase called on its own (online seam)
                sl_gemv(Am, A, m0, T, 0.0, nullptr);
                small_predict(c, G[0], Pc, A, Gm, S0);
            }
            double* Jc = Jh ? Jh + s * tt : G[6];
            small_gain(c, Jc, Pc, G[0], G[4], G[5]);                     // J = S0 A^T P^{-1}
            for (int i = threadIdx.x; i < T; i += SL_THREADS) v2[i] = m_new[i] - Am[i];
            __syncthreads();
            sl_gemv(d.f_star_sm + (int64_t)s * T, Jc, v2, T, 1.0, m0);   // m0 + J (m1 - A m0)
            sl_invalidate(c, G[7]);
            sl_axpby(G[7], 1.0, S_new, -1.0, Pc, n);                     // S1 - P
            SlEpi e;
            sl_gemm(c, G[8], Jc, 0, G[7], 0, e);                         // J (S1 - P)
            SlEpi es;
            es.beta = 1.0; es.D = S0;
            sl_gemm(c, d.cov_f_sm + s * tt, G[8], 0, Jc, 1, es);         // S0 + J (S1 - P) J^T
        }
        // ---------------- MNIW step (GPI_model.py:966-1101) ----------------
        if (!(phases & 4)) continue;
        const bool below = d.estimation_limit <= 0 || N < d.estimation_limit;
        if (N > 1 && below) {
            const double* f1 = d.f_star_sm + (int64_t)(s + 1) * T;
            const double* f0 = d.f_star_sm + (int64_t)s * T;
            const int i1 = small_mniw_compute(c, mi, f1, f0, T, G[0], G[2], G[3], G[4], G[5]);
            const int i2 = i1 ? 0 : small_mniw_compute(c, mo, y, f1, T, G[6], G[7], G[8], G[9], G[10]);
            if (i1 || i2) {
                if (!fail) fail = k + 1;
            } else {
                small_mniw_commit(c, mi, f1, f0, T, G[0], G[2]);
                small_mniw_commit(c, mo, y, f1, T, G[6], G[7]);
            }
        }
        if (below) {
            double* An = d.A + (p + 1) * tt; double* Gn = d.Gamma + (p + 1) * tt;
            double* Cn = d.C + (p + 1) * tt; double* Sn = d.Sigma + (p + 1) * tt;
            const double gi = *d.int_n0, go = *d.obs_n0;
            const double fa = d.annealing ? 1.0 / ((double)N * (double)N) : 0.0;
            for (int i = threadIdx.x; i < n; i += SL_THREADS) {
                An[i] = d.int_m_mean[i];
                Cn[i] = d.obs_m_mean[i];
                const double g = (N > 1) ? d.int_scale[i] * gi / (gi - 2.0) : Gm[i];
                const double sg = (N > 1) ? d.obs_scale[i] * go / (go - 2.0) : R[i];
                Gn[i] = g + fa * d.Gamma[i];          // + Gamma[0] / N^2
                Sn[i] = sg + fa * d.Sigma[i];         // + Sigma[0] / N^2
            }
            __syncthreads();
            p += 1;
        }
    }
    // ---------------- full RTS pass (GPI_model.py:687-703, GPI.py:262-270) ----------------
    const int Tn = n_tot;
    const int nA = p;
    if (!(phases & 8)) {
        if (threadIdx.x == 0) { d.status[0] = fail; d.status[1] = p + 1; }
        return;
    }
    if (Tn >= 1) {
        sl_copy(d.f_star_sm + (int64_t)Tn * T, d.f_star + (int64_t)Tn * T, T);
        sl_invalidate(c, d.cov_f_sm + Tn * tt);
        sl_copy(d.cov_f_sm + Tn * tt, d.cov_f + Tn * tt, n);
    }
    const bool cached = Ph != nullptr && d.start_members == 0 && (phases & 3) == 3;
    for (int t = Tn - 2; t >= 0; --t) {
        const int i = t + 1;                                            // state index
        const int ia = (t < nA ? t : nA - 1) + 1;
        const double* A = d.A + ia * tt;
        const double* Gm = d.Gamma + ia * tt;
        const double* mt = d.f_star + (int64_t)i * T;
        const double* St = d.cov_f + i * tt;
        const double* mn = d.f_star_sm + (int64_t)(i + 1) * T;
        const double* Sn = d.cov_f_sm + (i + 1) * tt;
        const double *Pt, *Jt, *Amt;
        if (cached) {                                                   // the forward pass left P_i, J_i and A m_i behind
            Pt = Ph + i * tt; Jt = Jh + i * tt; Amt = Vh + (int64_t)i * T;
        } else {
            sl_gemv(v0, A, mt, T, 0.0, nullptr);
            small_predict(c, G[0], G[1], A, Gm, St);
            small_gain(c, G[6], G[1], G[0], G[4], G[5]);
            Pt = G[1]; Jt = G[6]; Amt = v0;
        }
        for (int j = threadIdx.x; j < T; j += SL_THREADS) v2[j] = mn[j] - Amt[j];
        __syncthreads();
        sl_gemv(d.f_star_sm + (int64_t)i * T, Jt, v2, T, 1.0, mt);
        sl_invalidate(c, G[7]);
        sl_axpby(G[7], 1.0, Sn, -1.0, Pt, n);
        SlEpi e;
        sl_gemm(c, G[8], Jt, 0, G[7], 0, e);
        SlEpi es;
        es.beta = 1.0; es.D = St;
        sl_gemm(c, d.cov_f_sm + i * tt, G[8], 0, Jt, 1, es);
    }
    if (threadIdx.x == 0) { d.status[0] = fail; d.status[1] = p + 1; }
}

// ======================================================================================================
// The same small-system chain as a PIPELINE over a thread-block cluster of four CTAs (one SM each) per chain.
// Real fits have a handful of long chains (MIT-BIH record 100: two chains, one of 2271 members), so one CTA per chain
// leaves 146 SMs idle while a member costs ~26 dependent products and 6 factorisations.  The member step is not one
// dependency chain, though:
//   K  (rank 0)  Kalman update: predict P, gain K, new mean m+ (published early), then the Joseph-form covariance;
//   J  (rank 1)  pair-smoother gain J = S0 A^T P^-1 as soon as P exists -- concurrently with the Kalman gain -- and the
//                smoothed mean of the previous state once m+ is there;
//   MI (rank 2)  MNIW posterior over (A, Gamma): its first factorisation, Sinv and m_mean Sinv depend on the PREVIOUS
//                step only and are done before the step's data arrive; the rank-1 terms, the second factorisation and
//                the two triangular products follow the smoothed mean;
//   MO (rank 3)  the same for (C, Sigma), which needs m+ only.
// The MNIW steps read MEANS only, so the Joseph form (five products) runs beside them, and the pair smoother's covariance
// -- overwritten by the full RTS pass whenever that pass follows (phase bit 3) -- is skipped in that case.  The critical
// path of a member drops from 26 products + 6 factorisations to 8 products + 2 factorisations.
// The CTAs hand matrices over through global memory (L2) and monotone flags (st.release.gpu / ld.acquire.gpu, value =
// member count); all cross-CTA reads use ld.global.cg.  A cluster launch guarantees that the four CTAs are co-resident,
// so spinning on a flag cannot deadlock.  NCTA = 1 runs the four roles one after the other in one CTA (same code, flags
// compiled out).  Only whole member steps (phases 7 or 15); the online single-phase calls stay on chain_kernel_small.
constexpr int PF_P = 0, PF_M = 1, PF_S = 2, PF_F0 = 3, PF_JD = 4, PF_A = 5, PF_C = 6, PF_II = 7, PF_IO = 8;
constexpr int PI_INFO_I = 16, PI_INFO_O = 17, PI_FAIL = 18;
constexpr int PIPE_FLAG_INTS = 32;
constexpr int PIPE_MATS = 26;

template <bool MULTI>
__device__ __forceinline__ void pipe_signal(int* flags, int f, int v) {
    __syncthreads();
    if (MULTI && threadIdx.x == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flags + f), "r"(v) : "memory");
    }
}
template <bool MULTI>
__device__ __forceinline__ void pipe_wait(const int* flags, int f, int v) {
    if (MULTI) {
        if (threadIdx.x == 0) {
            int cur;
            do {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(cur) : "l"(flags + f) : "memory");
            } while (cur < v);
        }
        __syncthreads();
    }
}
__device__ __forceinline__ int pipe_read(const int* flags, int i) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
    return v;
}

// First half of matrix_normal_inv_wishart.posterior (GPI_model.py:1300-1344): everything that depends on the
// distribution alone.  Sinv -> Ga, m_mean Sinv -> Gd, Gc scratch.
__device__ __noinline__ int pipe_mniw_prep(SlCtx& c, Mniw d, int T, double* Ga, double* Gc, double* Gd) {
    const double jitter = 1e-2 * fmax(small_mean_abs_diag(d.scale, T, c), HGP_EPS);
    const int info = sl_cholinv(c, Gc, d.m_r_cov, jitter);              // Ls^{-1}
    if (info) return info;
    SlEpi e;
    sl_gemm(c, Ga, Gc, 1, Gc, 0, e, SL_TRI_A | SL_TRI_B);               // Sinv = Ls^{-T} Ls^{-1}
    sl_gemm(c, Gd, d.m_mean, 0, Ga, 0, e);                              // m_mean Sinv
    return 0;
}
// Second half: S2 = Sinv + y2 y2^T -> Ga, S1 = m_mean Sinv + y1 y2^T -> Gd, part_mean = S1 (sym(S2) + 1e-8 I)^{-1} -> Gb.
__device__ __noinline__ int pipe_mniw_finish(SlCtx& c, const double* y1, const double* y2, int T, double* Ga, double* Gb,
                                             double* Gc, double* Gd, double* Ge) {
    const int n = T * T;
    sl_invalidate(c, Ga);
    sl_invalidate(c, Gd);
    for (int i = threadIdx.x; i < n; i += SL_THREADS) {
        const double a = __ldcg(y2 + i / T), b = __ldcg(y2 + i % T);
        Ga[i] += a * b;
        Gd[i] += __ldcg(y1 + i / T) * b;
    }
    __syncthreads();
    const int info = sl_cholinv(c, Gc, Ga, 1e-8);                       // chol(sym(S2) + 1e-8 I)^{-1}
    if (info) return info;
    SlEpi e;
    sl_gemm(c, Ge, Gd, 0, Gc, 1, e, SL_TRI_B);                          // S1 L2^{-T}
    sl_gemm(c, Gb, Ge, 0, Gc, 0, e, SL_TRI_B);                          // part_mean = S1 S2^{-1}
    return 0;
}
__device__ __noinline__ void pipe_mniw_commit(SlCtx& c, Mniw d, const double* y1, const double* y2, int T, const double* S2,
                                              const double* part) {
    const int n = T * T;
    const double n0 = *d.n0;
    const double a = (n0 - 2.0), den = (n0 + 1.0) - 2.0;
    sl_invalidate(c, d.m_mean); sl_invalidate(c, d.scale); sl_invalidate(c, d.m_r_cov);
    for (int i = threadIdx.x; i < n; i += SL_THREADS) {
        const int r = i / T, cc = i % T;
        d.m_mean[i] = (a * d.m_mean[i] + part[i]) / den;
        const double e_r = __ldcg(y1 + r) - __ldcg(y2 + r), e_c = __ldcg(y1 + cc) - __ldcg(y2 + cc);
        d.scale[i] = (a * d.scale[i] + e_r * e_c) / den;
        d.m_r_cov[i] = S2[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) *d.n0 = n0 + 1.0;
    __syncthreads();
}

template <int NCTA>
__global__ void __launch_bounds__(SL_THREADS, 1)
chain_kernel_pipe(const hgp_chain_desc* __restrict__ descs, int T) {
    constexpr bool MULTI = NCTA > 1;
    __shared__ SlCtx c;
    sl_init(c, T);
    int rank = 0, chain = blockIdx.x;
    if (MULTI) {
        asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
        asm("mov.u32 %0, %%clusterid.x;" : "=r"(chain));
    }
    const hgp_chain_desc d = descs[chain];
    const bool rK = !MULTI || rank == 0, rJ = !MULTI || rank == 1, rI = !MULTI || rank == 2, rO = !MULTI || rank == 3;
    const int n = T * T;
    const int64_t tt = (int64_t)T * T;
    double* G[PIPE_MATS];
    for (int i = 0; i < PIPE_MATS; ++i) G[i] = d.work + i * tt;
    double* v0 = d.work + PIPE_MATS * tt; double* v1 = v0 + T; double* vj = v1 + T;
    int* flags = reinterpret_cast<int*>(d.work + PIPE_MATS * tt + 8 * (int64_t)T);
    if (MULTI) {
        if (rank == 0 && threadIdx.x < PIPE_FLAG_INTS) flags[threadIdx.x] = 0;
        __threadfence();
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    Mniw mi{d.int_m_mean, d.int_m_r_cov, d.int_scale, d.int_n0};
    Mniw mo{d.obs_m_mean, d.obs_m_r_cov, d.obs_scale, d.obs_n0};
    const int n_tot = d.start_members + d.n_members;
    double* Ph = d.rts_cache;                                            // [n_tot + 1][T][T]  P_s
    double* Jh = Ph ? Ph + (int64_t)(n_tot + 1) * tt : nullptr;          //                    J_s
    double* Vh = Ph ? Jh + (int64_t)(n_tot + 1) * tt : nullptr;          // [n_tot + 1][T]     A m_s
    const int phases = d.phases ? d.phases : 15;
    if ((phases & 7) != 7) {                                             // whole member steps only (host side checks as well)
        if (rK && threadIdx.x == 0) { d.status[0] = -2; d.status[1] = d.start_params + 1; }
        return;
    }
    const bool rts = (phases & 8) != 0;
    int fail = 0;
    int N = d.start_members;
    int p = d.start_params;
    for (int k = 0; k < d.n_members; ++k) {
        const int s = d.start_members + k;
        const int kv = k + 1;
        const double* A = d.A + p * tt;
        const double* Gm = d.Gamma + p * tt;
        const double* C = d.C + p * tt;
        const double* R = d.Sigma + p * tt;
        const double* y = d.Y + (int64_t)d.member_beats[k] * T;
        double* m_new = d.f_star + (int64_t)(s + 1) * T;
        double* S_new = d.cov_f + (s + 1) * tt;
        const bool prior = d.first_is_prior && s == 0;
        double* Pc = Ph ? Ph + s * tt : G[10];
        double* Am = Vh ? Vh + (int64_t)s * T : v0;
        double* Jc = Jh ? Jh + s * tt : G[13];
        N += 1;
        const bool smooth = N > 1;
        const bool below = d.estimation_limit <= 0 || N < d.estimation_limit;
        const bool mniw = smooth && below;
        // ---------------- K: Kalman update (GPI.py:104-150) ----------------
        if (rK) {
            pipe_wait<MULTI>(flags, PF_A, k);                            // parameter set p is complete
            pipe_wait<MULTI>(flags, PF_C, k);
            pipe_wait<MULTI>(flags, PF_JD, k);                           // the smoother has read the previous step's G0, P, A m
            const double* m = d.f_star_sm + (int64_t)s * T;
            const double* Sg = d.cov_f_sm + s * tt;
            sl_gemv(Am, A, m, T, 0.0, nullptr);                          // A m
            const double* P;
            SlEpi e;
            if (prior) {
                P = Sg;                                                  // P_t = cov_prior; f* = 0; R = r_first I
                for (int i = threadIdx.x; i < T; i += SL_THREADS) v1[i] = y[i];
                __syncthreads();
                sl_gemm(c, G[2], C, 0, P, 0, e);                         // C P
                SlEpi es;
                es.diag_add = d.r_first;
                sl_gemm(c, G[3], G[2], 0, C, 1, es);                     // S = C P C^T + r I
            } else {
                small_predict(c, G[0], Pc, A, Gm, Sg);                   // G0 = A Sigma, P = A Sigma A^T + Gamma
                pipe_signal<MULTI>(flags, PF_P, kv);
                P = Pc;
                sl_gemv(vj + T, C, Am, T, 0.0, nullptr);                 // f* = C A m
                for (int i = threadIdx.x; i < T; i += SL_THREADS) v1[i] = y[i] - vj[T + i];
                __syncthreads();
                sl_gemm(c, G[2], C, 0, P, 0, e);                         // C P
                SlEpi es;
                es.beta = 1.0; es.D = R;
                sl_gemm(c, G[3], G[2], 0, C, 1, es);                     // S = C P C^T + R
            }
            small_gain(c, G[6], G[3], G[2], G[4], G[5]);                 // K = P C^T S^{-1}
            sl_gemv(m_new, G[6], v1, T, 1.0, Am);                        // m+ = A m + K (y - f*)
            sl_copy(d.f_star_sm + (int64_t)(s + 1) * T, m_new, T);
            pipe_signal<MULTI>(flags, PF_M, kv);
            // Joseph form: (I - K C) P (I - K C)^T + K R K^T
            SlEpi ei;
            ei.alpha = -1.0; ei.diag_add = 1.0;
            sl_gemm(c, G[7], G[6], 0, C, 0, ei);                         // I - K C
            sl_gemm(c, G[8], G[7], 0, P, 0, e);                          // (I - K C) P
            sl_gemm(c, G[9], G[8], 0, G[7], 1, e);                       // ... (I - K C)^T
            SlEpi ej;
            ej.beta = 1.0; ej.D = G[9];
            if (prior) {
                ej.alpha = d.r_first;
                sl_gemm(c, S_new, G[6], 0, G[6], 1, ej, 0, d.cov_f_sm + (s + 1) * tt);      // + r K K^T
            } else {
                sl_gemm(c, G[5], G[6], 0, R, 0, e);                      // K R
                sl_gemm(c, S_new, G[5], 0, G[6], 1, ej, 0, d.cov_f_sm + (s + 1) * tt);      // + K R K^T
            }
            pipe_signal<MULTI>(flags, PF_S, kv);
        }
        // ---------------- J: pair smoother (GPI_model.py:705-716, GPI.py:294-299) ----------------
        if (rJ) {
            if (smooth) {
                const double* m0 = d.f_star + (int64_t)s * T;
                const double* S0 = d.cov_f + s * tt;
                pipe_wait<MULTI>(flags, PF_P, kv);
                small_gain(c, Jc, Pc, G[0], G[11], G[12]);               // J = S0 A^T P^{-1}
                pipe_wait<MULTI>(flags, PF_M, kv);
                for (int i = threadIdx.x; i < T; i += SL_THREADS) vj[i] = __ldcg(m_new + i) - __ldcg(Am + i);
                __syncthreads();
                sl_gemv(d.f_star_sm + (int64_t)s * T, Jc, vj, T, 1.0, m0);   // m0 + J (m1 - A m0)
                pipe_signal<MULTI>(flags, PF_F0, kv);
                if (!rts) {                                              // the full RTS pass would overwrite it
                    pipe_wait<MULTI>(flags, PF_S, kv);
                    sl_invalidate(c, G[14]);
                    sl_axpby(G[14], 1.0, S_new, -1.0, Pc, n);            // S1 - P
                    SlEpi e;
                    sl_gemm(c, G[15], Jc, 0, G[14], 0, e);               // J (S1 - P)
                    SlEpi es;
                    es.beta = 1.0; es.D = S0;
                    sl_gemm(c, d.cov_f_sm + s * tt, G[15], 0, Jc, 1, es);    // S0 + J (S1 - P) J^T
                }
            }
            pipe_signal<MULTI>(flags, PF_JD, kv);
        }
        // ---------------- MI / MO: MNIW step (GPI_model.py:966-1101) ----------------
        const double* f1 = d.f_star_sm + (int64_t)(s + 1) * T;
        const double* f0 = d.f_star_sm + (int64_t)s * T;
        int i1 = 0, i2 = 0;
        if (rI && mniw) {
            i1 = pipe_mniw_prep(c, mi, T, G[16], G[18], G[19]);
            pipe_wait<MULTI>(flags, PF_F0, kv);
            pipe_wait<MULTI>(flags, PF_M, kv);
            if (!i1) i1 = pipe_mniw_finish(c, f1, f0, T, G[16], G[17], G[18], G[19], G[20]);
            if (MULTI) {
                if (threadIdx.x == 0) flags[PI_INFO_I] = i1;
                pipe_signal<MULTI>(flags, PF_II, kv);
            }
        }
        if (rO && mniw) {
            i2 = pipe_mniw_prep(c, mo, T, G[21], G[23], G[24]);
            pipe_wait<MULTI>(flags, PF_M, kv);
            if (!i2) i2 = pipe_mniw_finish(c, y, f1, T, G[21], G[22], G[23], G[24], G[25]);
            if (MULTI) {
                if (threadIdx.x == 0) flags[PI_INFO_O] = i2;
                pipe_signal<MULTI>(flags, PF_IO, kv);
            }
        }
        if (MULTI && mniw) {                                             // a failure on either side keeps BOTH (:1068-1071)
            if (rI) { pipe_wait<MULTI>(flags, PF_IO, kv); i2 = pipe_read(flags, PI_INFO_O); }
            if (rO) { pipe_wait<MULTI>(flags, PF_II, kv); i1 = pipe_read(flags, PI_INFO_I); }
        }
        if (mniw && (i1 || i2) && !fail) fail = k + 1;
        const double fa = d.annealing ? 1.0 / ((double)N * (double)N) : 0.0;
        if (rI) {
            if (mniw && !(i1 || i2)) pipe_mniw_commit(c, mi, f1, f0, T, G[16], G[17]);
            if (below) {
                double* An = d.A + (p + 1) * tt; double* Gn = d.Gamma + (p + 1) * tt;
                const double gi = *d.int_n0;
                for (int i = threadIdx.x; i < n; i += SL_THREADS) {
                    An[i] = d.int_m_mean[i];
                    const double g = (N > 1) ? d.int_scale[i] * gi / (gi - 2.0) : Gm[i];
                    Gn[i] = g + fa * d.Gamma[i];          // + Gamma[0] / N^2
                }
            }
            pipe_signal<MULTI>(flags, PF_A, kv);
        }
        if (rO) {
            if (mniw && !(i1 || i2)) pipe_mniw_commit(c, mo, y, f1, T, G[21], G[22]);
            if (below) {
                double* Cn = d.C + (p + 1) * tt; double* Sn = d.Sigma + (p + 1) * tt;
                const double go = *d.obs_n0;
                for (int i = threadIdx.x; i < n; i += SL_THREADS) {
                    Cn[i] = d.obs_m_mean[i];
                    const double sg = (N > 1) ? d.obs_scale[i] * go / (go - 2.0) : R[i];
                    Sn[i] = sg + fa * d.Sigma[i];         // + Sigma[0] / N^2
                }
            }
            pipe_signal<MULTI>(flags, PF_C, kv);
        }
        if (below) p += 1;
    }
    // ---------------- all roles join; rank 0 goes on alone ----------------
    if (MULTI) {
        const int fin = d.n_members + 1;
        if (rI) { if (threadIdx.x == 0) flags[PI_FAIL] = fail; pipe_signal<MULTI>(flags, PF_A, fin); }
        if (rO) pipe_signal<MULTI>(flags, PF_C, fin);
        if (rJ) pipe_signal<MULTI>(flags, PF_JD, fin);
        if (!rK) return;
        pipe_wait<MULTI>(flags, PF_A, fin);
        pipe_wait<MULTI>(flags, PF_C, fin);
        pipe_wait<MULTI>(flags, PF_JD, fin);
        fail = pipe_read(flags, PI_FAIL);
    }
    // ---------------- full RTS pass (GPI_model.py:687-703, GPI.py:262-270) ----------------
    const int Tn = n_tot;
    const int nA = p;
    if (!rts) {
        if (threadIdx.x == 0) { d.status[0] = fail; d.status[1] = p + 1; }
        return;
    }
    if (Tn >= 1) {
        sl_copy(d.f_star_sm + (int64_t)Tn * T, d.f_star + (int64_t)Tn * T, T);
        sl_invalidate(c, d.cov_f_sm + Tn * tt);
        sl_copy(d.cov_f_sm + Tn * tt, d.cov_f + Tn * tt, n);
    }
    const bool cached = Ph != nullptr && d.start_members == 0;
    for (int t = Tn - 2; t >= 0; --t) {
        const int i = t + 1;                                            // state index
        const int ia = (t < nA ? t : nA - 1) + 1;
        const double* A = d.A + ia * tt;
        const double* Gm = d.Gamma + ia * tt;
        const double* mt = d.f_star + (int64_t)i * T;
        const double* St = d.cov_f + i * tt;
        const double* mn = d.f_star_sm + (int64_t)(i + 1) * T;
        const double* Sn = d.cov_f_sm + (i + 1) * tt;
        const double *Pt, *Jt, *Amt;
        if (cached) {                                                   // the forward pass left P_i, J_i and A m_i behind
            Pt = Ph + i * tt; Jt = Jh + i * tt; Amt = Vh + (int64_t)i * T;
        } else {
            sl_gemv(v0, A, mt, T, 0.0, nullptr);
            small_predict(c, G[0], G[1], A, Gm, St);
            small_gain(c, G[6], G[1], G[0], G[4], G[5]);
            Pt = G[1]; Jt = G[6]; Amt = v0;
        }
        for (int j = threadIdx.x; j < T; j += SL_THREADS) vj[j] = mn[j] - __ldcg(Amt + j);
        __syncthreads();
        sl_gemv(d.f_star_sm + (int64_t)i * T, Jt, vj, T, 1.0, mt);
        sl_invalidate(c, G[7]);
        sl_axpby(G[7], 1.0, Sn, -1.0, Pt, n);
        SlEpi e;
        sl_gemm(c, G[8], Jt, 0, G[7], 0, e);
        SlEpi es;
        es.beta = 1.0; es.D = St;
        sl_gemm(c, d.cov_f_sm + i * tt, G[8], 0, Jt, 1, es);
    }
    if (threadIdx.x == 0) { d.status[0] = fail; d.status[1] = p + 1; }
}

// unit-test hook for the shared-memory routines
__global__ void __launch_bounds__(SL_THREADS, 1)
sl_op_kernel(int op, double* A, double* B, double* Cm, int T, int* info) {
    __shared__ SlCtx c;
    sl_init(c, T);
    int rc = 0;
    SlEpi e;
    switch (op) {
        case 10: sl_gemm(c, Cm, A, 0, B, 0, e); break;
        case 11: sl_gemm(c, Cm, A, 1, B, 0, e); break;
        case 12: sl_gemm(c, Cm, A, 0, B, 1, e); break;
        case 13: { SlEpi f; f.alpha = 2.0; f.beta = -1.0; f.D = Cm; f.diag_add = 0.5; sl_gemm(c, Cm, A, 1, B, 1, f); } break;
        case 14: rc = sl_cholinv(c, Cm, A, 0.25); if (threadIdx.x == 0) B[0] = c.logdet; break;   // Cm = chol(sym A + .25 I)^-1
        case 15: {   // Cm = A^T M^{-1} via small_gain (M = B SPD); scratch: the two T x T blocks behind Cm
            rc = small_gain(c, Cm, B, A, Cm + (int64_t)T * T, Cm + 2 * (int64_t)T * T);
        } break;
        case 16: sl_gemm(c, Cm, A, 0, B, 0, e, SL_TRI_A); break;      // A lower triangular
        case 17: sl_gemm(c, Cm, A, 1, B, 0, e, SL_TRI_A | SL_TRI_B); break;   // A^T B, both lower triangular
        case 18: sl_gemm(c, Cm, A, 0, B, 1, e, SL_TRI_B); break;      // A B^T, B lower triangular
        // steady-state cost of the routines inside one launch (tools/la_bench.py divides by 64)
        case 20: for (int i = 0; i < 64; ++i) sl_gemm(c, Cm, A, 0, B, 0, e); break;                 // operands stay cached
        case 21: for (int i = 0; i < 32; ++i) { sl_gemm(c, Cm, A, 0, B, 0, e); sl_gemm(c, Cm + (int64_t)T * T, B, 1, Cm, 0, e);
                                                 sl_invalidate(c, A); } break;                      // one operand miss per product
        case 22: for (int i = 0; i < 64; ++i) rc |= sl_cholinv(c, Cm, A, 0.25); break;
        case 23: for (int i = 0; i < 64; ++i) sl_gemv(Cm, A, B, T, 0.0, nullptr); break;
        case 24: for (int i = 0; i < 64; ++i) { sl_invalidate(c, Cm); sl_axpby(Cm, 1.0, A, -1.0, B, T * T); } break;
        default: rc = -1;
    }
    if (threadIdx.x == 0) info[0] = rc;
}

// unit-test hook for the CTA-level routines: op codes below
__global__ void __launch_bounds__(LA_THREADS)
la_op_kernel(int op, double* A, double* B, double* C, int* piv, int T, int* info) {
    __shared__ LaSmem sm;
    extern __shared__ double la_dyn[];
    if (threadIdx.x == 0) sm.big = T <= LA_SMEM_T ? la_dyn : nullptr;
    __syncthreads();
    int rc = 0;
    switch (op) {
        case 0: la_gemm(C, A, 0, B, 0, T, 1.0, 0.0, nullptr, sm); break;
        case 1: la_gemm(C, A, 1, B, 0, T, 1.0, 0.0, nullptr, sm); break;
        case 2: la_gemm(C, A, 0, B, 1, T, 1.0, 0.0, nullptr, sm); break;
        case 3: la_gemm(C, A, 1, B, 1, T, 2.0, -1.0, C, sm); break;
        case 4: rc = la_chol(A, T, sm); break;
        case 5: la_trsm_lower(A, B, T, sm); break;
        case 6: la_trsm_lower_trans(A, B, T, sm); break;
        case 7: la_lu_factor(A, piv, T, sm); la_lu_solve(A, piv, B, T, sm); break;
        case 8: la_symmetrize(A, 0.25, T); break;
        case 9: la_transpose(C, A, T); break;
        default: rc = -1;
    }
    if (threadIdx.x == 0) info[0] = rc;
}

}  // namespace

extern "C" int64_t hgp_chain_desc_bytes(void) { return (int64_t)sizeof(hgp_chain_desc); }
extern "C" int64_t hgp_chain_work_doubles(int T) { return PIPE_MATS * (int64_t)T * T + 8 * (int64_t)T + PIPE_FLAG_INTS; }
extern "C" int64_t hgp_chain_rts_cache_doubles(int T, int n_states) {
    return sl_supported(T) ? (int64_t)n_states * (2 * (int64_t)T * T + T) : 0;
}
extern "C" int hgp_chain_small_path(int T) { return sl_supported(T) && !getenv("HGP_CHAIN_V1") ? 1 : 0; }

extern "C" int hgp_chain_pipeline_ctas(void) { return 4; }

extern "C" int hgp_chain_run_ex(const void* descs_device, int n_chains, int T, int pipeline, void* stream) {
    HGP_REQUIRE(n_chains >= 0 && T > 0 && T <= 1024, "hgp_chain_run_ex: bad sizes");
    HGP_REQUIRE(pipeline == 0 || pipeline == 1 || pipeline == 4, "hgp_chain_run_ex: pipeline must be 0, 1 or 4");
    if (pipeline == 0) return hgp_chain_run(descs_device, n_chains, T, stream);
    if (!sl_supported(T)) { hgp_set_error("hgp_chain_run_ex: the pipelined chain needs the shared-memory path (T = %d)", T); return HGP_E_UNSUPPORTED; }
    if (n_chains == 0) return 0;
    const size_t sdyn = sl_dynamic_smem_bytes(T);
    const hgp_chain_desc* descs = reinterpret_cast<const hgp_chain_desc*>(descs_device);
    if (pipeline == 1) {
        cudaError_t e = cudaFuncSetAttribute(chain_kernel_pipe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sdyn);
        if (e != cudaSuccess) return hgp_status(e, "hgp_chain_run_ex: shared memory");
        chain_kernel_pipe<1><<<n_chains, SL_THREADS, sdyn, (cudaStream_t)stream>>>(descs, T);
        HGP_LAUNCH_CHECK("hgp_chain_run_ex");
        return 0;
    }
    cudaError_t e = cudaFuncSetAttribute(chain_kernel_pipe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sdyn);
    if (e != cudaSuccess) return hgp_status(e, "hgp_chain_run_ex: shared memory");
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(4u * (unsigned)n_chains, 1, 1);
    cfg.blockDim = dim3(SL_THREADS, 1, 1);
    cfg.dynamicSmemBytes = sdyn;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, chain_kernel_pipe<4>, descs, T);
    if (e != cudaSuccess) return hgp_status(e, "hgp_chain_run_ex: cluster launch");
    HGP_LAUNCH_CHECK("hgp_chain_run_ex");
    return 0;
}

extern "C" int hgp_chain_run(const void* descs_device, int n_chains, int T, void* stream) {
    HGP_REQUIRE(n_chains >= 0 && T > 0 && T <= 1024, "hgp_chain_run: bad sizes");
    if (n_chains == 0) return 0;
    if (hgp_chain_small_path(T)) {
        const size_t sdyn = sl_dynamic_smem_bytes(T);
        cudaError_t e = cudaFuncSetAttribute(chain_kernel_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sdyn);
        if (e != cudaSuccess) return hgp_status(e, "hgp_chain_run: shared memory (small path)");
        chain_kernel_small<<<n_chains, SL_THREADS, sdyn, (cudaStream_t)stream>>>(
            reinterpret_cast<const hgp_chain_desc*>(descs_device), T);
        HGP_LAUNCH_CHECK("hgp_chain_run");
        return 0;
    }
    const size_t dyn = la_dynamic_smem_bytes(T);
    if (dyn > 0) {
        cudaError_t e = cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return hgp_status(e, "hgp_chain_run: shared memory");
    }
    chain_kernel<<<n_chains, LA_THREADS, dyn, (cudaStream_t)stream>>>(reinterpret_cast<const hgp_chain_desc*>(descs_device), T);
    HGP_LAUNCH_CHECK("hgp_chain_run");
    return 0;
}

extern "C" int hgp_la_op(int op, double* A, double* B, double* C, int* piv, int T, int* info, void* stream) {
    HGP_REQUIRE(T > 0 && T <= 1024, "hgp_la_op: bad T");
    if (op >= 10) {      // shared-memory routines (hgp_smem_la.cuh)
        HGP_REQUIRE(sl_supported(T), "hgp_la_op: T outside the shared-memory path");
        const size_t sdyn = sl_dynamic_smem_bytes(T);
        cudaError_t e = cudaFuncSetAttribute(sl_op_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sdyn);
        if (e != cudaSuccess) return hgp_status(e, "hgp_la_op: shared memory");
        sl_op_kernel<<<1, SL_THREADS, sdyn, (cudaStream_t)stream>>>(op, A, B, C, T, info);
        HGP_LAUNCH_CHECK("hgp_la_op");
        return 0;
    }
    const size_t dyn = la_dynamic_smem_bytes(T);
    if (dyn > 0) {
        cudaError_t e = cudaFuncSetAttribute(la_op_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) return hgp_status(e, "hgp_la_op: shared memory");
    }
    la_op_kernel<<<1, LA_THREADS, dyn, (cudaStream_t)stream>>>(op, A, B, C, piv, T, info);
    HGP_LAUNCH_CHECK("hgp_la_op");
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// Emission distribution on a grid that differs from the basis grid: IterativeGaussianProcess.pred_dist
// (reference GPI.py:457-503, kernel branch :470-501) -- kernel-matrix construction, Cholesky, two triangular
// solves and the projected covariance, one CTA per (state, grid) item.  All matrices are embedded in D x D
// workspaces, D = max(n_b, n_x) (identity padding on the Cholesky diagonal, zeros elsewhere), so the square
// CTA-level routines apply unchanged.
//   K_bb = c exp(-(xb_i - xb_j)^2 / (2 l^2)),  K_bx likewise,  K_xx = the same + noise on the diagonal
//   L = chol(sym(K_bb) + 1e-4 max(mean|diag Sigma|, eps) I);  W = K_bb^{-1} K_bx;  f* = W^T mu
//   cov = mean(diag Sigma) I                      if all(isclose(diag Sigma, mean diag Sigma))   (:497-498)
//       = sym(K_xx - K_bx^T W + W^T Sigma W) + 1e-6 I   otherwise                                (:500-501)
namespace {

__global__ void __launch_bounds__(LA_THREADS)
pred_dist_kernel(const double* __restrict__ x_basis, int nb, const double* __restrict__ x_post, int64_t x_post_stride,
                 int nx, const double* __restrict__ mu, const int* __restrict__ mu_idx,
                 const double* __restrict__ Sigma, const int* __restrict__ sig_idx, double kc, double kl, double knoise,
                 double* __restrict__ fout, double* __restrict__ covout, double* __restrict__ work, int* __restrict__ info) {
    __shared__ LaSmem sm;
    __shared__ int s_const;
    if (threadIdx.x == 0) sm.big = nullptr;
    const int64_t it = blockIdx.x;
    const int D = max(nb, nx);
    const int64_t dd = (int64_t)D * D;
    double* Kbb = work + it * 5 * dd;     // -> chol factor
    double* Kbx = Kbb + dd;               // -> W
    double* Sg = Kbx + dd;                // Sigma padded
    double* T1 = Sg + dd;
    double* T2 = T1 + dd;
    const double* xp = x_post + it * x_post_stride;
    const double* S = Sigma + (int64_t)sig_idx[it] * nb * nb;
    const double* m = mu + (int64_t)mu_idx[it] * nb;
    const int tid = threadIdx.x;
    // mean of diag Sigma, constant-diagonal test
    double part = 0.0, parta = 0.0;
    for (int i = tid; i < nb; i += LA_THREADS) { part += S[(int64_t)i * nb + i]; parta += fabs(S[(int64_t)i * nb + i]); }
    part = warp_sum(part); parta = warp_sum(parta);
    if ((tid & 31) == 0) { sm.red[tid >> 5] = part; sm.red[8 + (tid >> 5)] = parta; }
    if (tid == 0) s_const = 1;
    __syncthreads();
    double dmean = 0.0, dmeana = 0.0;
    for (int w = 0; w < LA_THREADS / 32; ++w) { dmean += sm.red[w]; dmeana += sm.red[8 + w]; }
    dmean /= nb; dmeana /= nb;
    __syncthreads();
    for (int i = tid; i < nb; i += LA_THREADS) {
        const double d = S[(int64_t)i * nb + i];
        if (!(fabs(d - dmean) <= 1e-8 + 1e-5 * fabs(dmean))) s_const = 0;      // torch.isclose defaults
    }
    // kernel matrices (embedded)
    for (int idx = tid; idx < D * D; idx += LA_THREADS) {
        const int r = idx / D, c = idx % D;
        double kbb = (r == c) ? 1.0 : 0.0, kbx = 0.0, sg = 0.0;
        if (r < nb && c < nb) {
            const double d = x_basis[r] / kl - x_basis[c] / kl;
            kbb = kc * exp(-0.5 * d * d);
            sg = S[(int64_t)r * nb + c];
        }
        if (r < nb && c < nx) {
            const double d = x_basis[r] / kl - xp[c] / kl;
            kbx = kc * exp(-0.5 * d * d);
        }
        Kbb[idx] = kbb; Kbx[idx] = kbx; Sg[idx] = sg;
    }
    __syncthreads();
    const int is_const = s_const;
    const double jitter = 1e-4 * fmax(dmeana, HGP_EPS);
    for (int i = tid; i < nb; i += LA_THREADS) Kbb[(int64_t)i * D + i] += jitter;   // K_bb is exactly symmetric already
    __syncthreads();
    int rc = la_chol(Kbb, D, sm);
    la_copy(T1, Kbx, (int)dd);                       // keep K_bx
    la_trsm_lower(Kbb, Kbx, D, sm);
    la_trsm_lower_trans(Kbb, Kbx, D, sm);            // Kbx = W
    // f* = W^T mu
    {
        const int warp = tid >> 5, lane = tid & 31;
        for (int c = warp; c < nx; c += LA_THREADS / 32) {
            double acc = 0.0;
            for (int r = lane; r < nb; r += 32) acc += Kbx[(int64_t)r * D + c] * m[r];
            acc = warp_sum(acc);
            if (lane == 0) fout[it * nx + c] = acc;
        }
    }
    double* cov = covout + it * (int64_t)nx * nx;
    if (is_const) {
        for (int idx = tid; idx < nx * nx; idx += LA_THREADS) cov[idx] = (idx / nx == idx % nx) ? dmean : 0.0;
    } else {
        la_gemm(T2, Sg, 0, Kbx, 0, D, 1.0, 0.0, nullptr, sm);          // Sigma W
        la_gemm(Sg, Kbx, 1, T2, 0, D, 1.0, 0.0, nullptr, sm);          // W^T Sigma W      (Sg reused)
        la_gemm(T2, T1, 1, Kbx, 0, D, -1.0, 1.0, Sg, sm);              // - K_bx^T W + W^T Sigma W
        for (int idx = tid; idx < nx * nx; idx += LA_THREADS) {
            const int r = idx / nx, c = idx % nx;
            const double d = xp[r] / kl - xp[c] / kl;
            double kxx = (r == c) ? kc + knoise : kc * exp(-0.5 * d * d);   // kernel(x): RBF diagonal is exactly 1, + white noise
            const double a = kxx + T2[(int64_t)r * D + c];
            const double b = ((r == c) ? kc + knoise : kc * exp(-0.5 * d * d)) + T2[(int64_t)c * D + r];
            cov[idx] = 0.5 * (a + b) + ((r == c) ? 1e-6 : 0.0);
        }
    }
    if (tid == 0) info[it] = rc;
}

}  // namespace

extern "C" int64_t hgp_pred_dist_work_doubles(int64_t n_items, int nb, int nx) {
    const int64_t D = nb > nx ? nb : nx;
    return n_items * 5 * D * D;
}

extern "C" int hgp_pred_dist_inducing(const double* x_basis, int nb, const double* x_post, int64_t x_post_stride, int nx,
                                      const double* mu, const int* mu_idx, const double* Sigma, const int* sig_idx,
                                      int64_t n_items, double kernel_const, double kernel_length, double kernel_noise,
                                      double* f_out, double* cov_out, double* work, int* info, void* stream) {
    HGP_REQUIRE(n_items >= 0 && nb > 0 && nx > 0 && nb <= 1024 && nx <= 1024, "hgp_pred_dist_inducing: bad sizes");
    if (n_items == 0) return 0;
    pred_dist_kernel<<<(unsigned)n_items, LA_THREADS, 0, (cudaStream_t)stream>>>(
        x_basis, nb, x_post, x_post_stride, nx, mu, mu_idx, Sigma, sig_idx, kernel_const, kernel_length, kernel_noise,
        f_out, cov_out, work, info);
    HGP_LAUNCH_CHECK("hgp_pred_dist_inducing");
    return 0;
}
