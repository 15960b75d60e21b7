// Emission means (reference GPI_model.observe, GPI_model.py:626-662: mean = C_i f_i) as a batched
// GEMV: one CTA per state, one warp per output row, coalesced row reads (HBM-bound: T*T*8 bytes
// per state when every state has its own C).
#include "hgp_common.cuh"
#include "hgp_gemm.cuh"

namespace {

__global__ void __launch_bounds__(256)
emission_means_kernel(const double* __restrict__ C, const double* __restrict__ f, const int* __restrict__ c_idx,
                      const int* __restrict__ f_idx, int T, double* __restrict__ mu) {
    extern __shared__ double fs[];
    const int64_t s = blockIdx.x;
    const double* Cm = C + (int64_t)c_idx[s] * T * T;
    const double* fv = f + (int64_t)f_idx[s] * T;
    for (int t = threadIdx.x; t < T; t += blockDim.x) fs[t] = fv[t];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < T; r += blockDim.x >> 5) {
        const double* row = Cm + (int64_t)r * T;
        double acc = 0.0;
        for (int k = lane; k < T; k += 32) acc += row[k] * fs[k];
        acc = warp_sum(acc);
        if (lane == 0) mu[s * T + r] = acc;
    }
}


// Latent-transition score, final step per member: r = f_cur - A f_prev, z = W r (W = chol(Gamma)^{-1}),
// out = -0.5 (|z|^2 + trace) - 0.5 T log 2 pi with trace = sum of the GEMM tile partials (fixed order).
__global__ void __launch_bounds__(256)
qlat_finish_kernel(const double* __restrict__ A, const int* __restrict__ A_idx, const double* __restrict__ W,
                   const double* __restrict__ fmean, const int* __restrict__ fprev_idx,
                   const int* __restrict__ fcur_idx, const double* __restrict__ partial, int ntile2, int T,
                   const int* __restrict__ info, double* __restrict__ out) {
    extern __shared__ double sm[];
    double* fp = sm;          // f_prev
    double* r = sm + T;       // residual
    __shared__ double s_red[8];
    const int64_t j = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double* Aj = A + (int64_t)A_idx[j] * T * T;
    const double* Wj = W + j * (int64_t)T * T;
    const double* fprev = fmean + (int64_t)fprev_idx[j] * T;
    const double* fcur = fmean + (int64_t)fcur_idx[j] * T;
    for (int t = tid; t < T; t += 256) fp[t] = fprev[t];
    __syncthreads();
    for (int row = warp; row < T; row += 8) {
        const double* ar = Aj + (int64_t)row * T;
        double acc = 0.0;
        for (int k = lane; k < T; k += 32) acc += ar[k] * fp[k];
        acc = warp_sum(acc);
        if (lane == 0) r[row] = fcur[row] - acc;
    }
    __syncthreads();
    double mah = 0.0;
    for (int row = warp; row < T; row += 8) {
        const double* wr = Wj + (int64_t)row * T;
        double acc = 0.0;
        for (int k = lane; k <= row; k += 32) acc += wr[k] * r[k];
        acc = warp_sum(acc);
        mah += acc * acc;      // identical in every lane
    }
    if (lane == 0) s_red[warp] = mah;
    __syncthreads();
    if (tid == 0) {
        double m = 0.0;
        for (int w = 0; w < 8; ++w) m += s_red[w];
        double tr = 0.0;
        for (int t = 0; t < ntile2; ++t) tr += partial[j * ntile2 + t];
        out[j] = (info[j] == 0) ? (-0.5 * (m + tr) - 0.5 * (double)T * HGP_LOG2PI) : nan("");
    }
}

}  // namespace

extern "C" int hgp_emission_means(const double* C, const double* f, const int* c_idx, const int* f_idx, int64_t S,
                                  int T, double* mu, void* stream) {
    HGP_REQUIRE(S >= 0 && T > 0 && T <= 4096, "hgp_emission_means: bad sizes");
    if (S == 0) return 0;
    emission_means_kernel<<<(unsigned)S, 256, sizeof(double) * T, (cudaStream_t)stream>>>(C, f, c_idx, f_idx, T, mu);
    HGP_LAUNCH_CHECK("hgp_emission_means");
    return 0;
}

extern "C" int hgp_gemm_batched(const double* A, const int* ia, const double* B, const int* ib, double* C, int64_t J,
                                int T, int lowerA, int transA, void* stream) {
    HGP_REQUIRE(J >= 0 && T > 0 && T <= 4096, "hgp_gemm_batched: bad sizes");
    if (J == 0) return 0;
    const int nt = (T + hgp::GT - 1) / hgp::GT;
    const int64_t st = (int64_t)T * T;
    for (int64_t j0 = 0; j0 < J; j0 += 65535) {
        const int64_t jb = hgp_min64(65535, J - j0);
        dim3 grid(nt * nt, (unsigned)jb);
        hgp::gemm_tile_kernel<0><<<grid, 256, 0, (cudaStream_t)stream>>>(
            ia ? A : A + j0 * st, ia ? ia + j0 : nullptr, st, lowerA, transA, ib ? B : B + j0 * st, ib ? ib + j0 : nullptr,
            st, C + j0 * st, st, nullptr, 0, nullptr, T);
        HGP_LAUNCH_CHECK("hgp_gemm_batched");
    }
    return 0;
}

static int64_t qlat_sub_batch(int64_t J) { return hgp_min64(J, 256); }

extern "C" int64_t hgp_qlat_workspace_bytes(int64_t J, int T) {
    const int64_t jb = qlat_sub_batch(J);
    const int64_t nt = (T + hgp::GT - 1) / hgp::GT;
    return jb * (2 * (int64_t)T * T + nt * nt) * (int64_t)sizeof(double) + 256;
}

extern "C" int hgp_qlat_batched(const double* A, const double* Gamma, const double* P, const double* fmean,
                                const int* A_idx, const int* G_idx, const int* P_idx, const int* fprev_idx,
                                const int* fcur_idx, const double* gamma_scale, int64_t J, int T, double* out,
                                int* info, void* workspace, int64_t workspace_bytes, void* stream) {
    HGP_REQUIRE(J >= 0 && T > 0 && T <= 1024, "hgp_qlat_batched: need 0 < T <= 1024");
    if (J == 0) return 0;
    if (workspace_bytes < hgp_qlat_workspace_bytes(J, T)) { hgp_set_error("hgp_qlat_batched: workspace too small"); return HGP_E_WORKSPACE; }
    const int64_t jbmax = qlat_sub_batch(J);
    const int nt = (T + hgp::GT - 1) / hgp::GT;
    const int64_t st = (int64_t)T * T;
    double* W1 = reinterpret_cast<double*>(workspace);      // chol(Gamma), later X = W A
    double* W2 = W1 + jbmax * st;                           // W = chol(Gamma)^{-1}
    double* part = W2 + jbmax * st;
    cudaStream_t s = (cudaStream_t)stream;
    for (int64_t j0 = 0; j0 < J; j0 += jbmax) {
        const int64_t jb = hgp_min64(jbmax, J - j0);
        int rc = hgp_internal_chol(Gamma, G_idx + j0, gamma_scale ? gamma_scale + j0 : nullptr, jb, T, nullptr, 1e-8, W1,
                                   nullptr, info + j0, stream);
        if (rc) return rc;
        rc = hgp_tri_inverse_batched(W1, jb, T, W2, stream);
        if (rc) return rc;
        dim3 grid(nt * nt, (unsigned)jb);
        // X = W A   (W lower triangular)
        hgp::gemm_tile_kernel<0><<<grid, 256, 0, s>>>(W2, nullptr, st, 1, 0, A, A_idx + j0, st, W1, st, nullptr, 0, nullptr, T);
        HGP_LAUNCH_CHECK("hgp_qlat_batched: X = W A");
        // partial = sum (X P) .* X
        hgp::gemm_tile_kernel<1><<<grid, 256, 0, s>>>(W1, nullptr, st, 0, 0, P, P_idx + j0, st, nullptr, 0, W1, st, part, T);
        HGP_LAUNCH_CHECK("hgp_qlat_batched: trace");
        qlat_finish_kernel<<<(unsigned)jb, 256, 2 * sizeof(double) * T, s>>>(A, A_idx + j0, W2, fmean, fprev_idx + j0,
                                                                           fcur_idx + j0, part, nt * nt, T, info + j0,
                                                                           out + j0);
        HGP_LAUNCH_CHECK("hgp_qlat_batched: finish");
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// MNIW log-likelihood of LDS parameters under the prior (ELBO term of every non-empty cluster):
// matrix_normal_inv_wishart.log_likelihood_MNIW (reference GPI_model.py:1346-1362)
//   L = chol(sym(Sigma) + 1e-8 I);  D = M - m_mean
//   out = -0.5 sum((D R) (.) Sigma^{-1} D) - 0.5 tr(Sigma^{-1} S)          (R = m_r_cov, S = prior scale)
// computed with W = L^{-1}:  sum((X R) (.) X), X = W D   and   sum((W S) (.) W)   -- 2 + 2/3... T^3 flops on the
// FP64 tensor cores instead of two cholesky_solve with T right-hand sides.
namespace {

__global__ void mniw_diff_kernel(const double* __restrict__ M, const int* __restrict__ M_idx,
                                 const double* __restrict__ mean, const int* __restrict__ mean_idx, int64_t st,
                                 double* __restrict__ D) {
    const int64_t j = blockIdx.y;
    const double* a = M + (int64_t)M_idx[j] * st;
    const double* b = mean + (int64_t)mean_idx[j] * st;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < st; i += (int64_t)gridDim.x * blockDim.x)
        D[j * st + i] = a[i] - b[i];
}

__global__ void mniw_fill_kernel(double* p, int64_t n, double v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void mniw_finish_kernel(const double* __restrict__ p1, const double* __restrict__ p2, int ntile2,
                                   const int* __restrict__ info, int64_t J, double* __restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= J) return;
    double a = 0.0, b = 0.0;
    for (int t = 0; t < ntile2; ++t) { a += p1[j * ntile2 + t]; b += p2[j * ntile2 + t]; }
    out[j] = info[j] == 0 ? (-0.5 * a) + (-0.5 * b) : nan("");
}

}  // namespace

static int64_t mniw_sub_batch(int64_t J) { return hgp_min64(J, 128); }

extern "C" int64_t hgp_mniw_workspace_bytes(int64_t J, int T) {
    const int64_t jb = mniw_sub_batch(J);
    const int64_t nt = (T + hgp::GT - 1) / hgp::GT;
    return jb * (3 * (int64_t)T * T + 2 * nt * nt + 1) * (int64_t)sizeof(double) + 256;
}

extern "C" int hgp_mniw_loglik_batched(const double* M, const int* M_idx, const double* Sigma, const int* S_idx,
                                       const double* prior_mean, const int* pm_idx, const double* prior_rcov,
                                       const int* pr_idx, const double* prior_scale, const int* ps_idx, int64_t J, int T,
                                       double* out, int* info, void* workspace, int64_t workspace_bytes, void* stream) {
    HGP_REQUIRE(J >= 0 && T > 0 && T <= 1024, "hgp_mniw_loglik_batched: need 0 < T <= 1024");
    if (J == 0) return 0;
    if (workspace_bytes < hgp_mniw_workspace_bytes(J, T)) { hgp_set_error("hgp_mniw_loglik_batched: workspace too small"); return HGP_E_WORKSPACE; }
    const int64_t jbmax = mniw_sub_batch(J);
    const int nt = (T + hgp::GT - 1) / hgp::GT;
    const int64_t st = (int64_t)T * T;
    double* W1 = reinterpret_cast<double*>(workspace);      // chol(Sigma), later X = W D
    double* W2 = W1 + jbmax * st;                           // W
    double* Dm = W2 + jbmax * st;                           // D = M - m_mean
    double* p1 = Dm + jbmax * st;
    double* p2 = p1 + jbmax * nt * nt;
    double* add = p2 + jbmax * nt * nt;
    cudaStream_t s = (cudaStream_t)stream;
    mniw_fill_kernel<<<(unsigned)((jbmax + 127) / 128), 128, 0, s>>>(add, jbmax, 1e-8);
    HGP_LAUNCH_CHECK("hgp_mniw_loglik_batched: fill");
    for (int64_t j0 = 0; j0 < J; j0 += jbmax) {
        const int64_t jb = hgp_min64(jbmax, J - j0);
        int rc = hgp_internal_chol(Sigma, S_idx + j0, nullptr, jb, T, add, 0.0, W1, nullptr, info + j0, stream);
        if (rc) return rc;
        rc = hgp_tri_inverse_batched(W1, jb, T, W2, stream);
        if (rc) return rc;
        mniw_diff_kernel<<<dim3(32, (unsigned)jb), 256, 0, s>>>(M, M_idx + j0, prior_mean, pm_idx + j0, st, Dm);
        HGP_LAUNCH_CHECK("hgp_mniw_loglik_batched: diff");
        dim3 grid(nt * nt, (unsigned)jb);
        hgp::gemm_tile_kernel<0><<<grid, 256, 0, s>>>(W2, nullptr, st, 1, 0, Dm, nullptr, st, W1, st, nullptr, 0, nullptr, T);
        HGP_LAUNCH_CHECK("hgp_mniw_loglik_batched: X = W D");
        hgp::gemm_tile_kernel<1><<<grid, 256, 0, s>>>(W1, nullptr, st, 0, 0, prior_rcov, pr_idx + j0, st, nullptr, 0, W1, st, p1, T);
        HGP_LAUNCH_CHECK("hgp_mniw_loglik_batched: mean term");
        hgp::gemm_tile_kernel<1><<<grid, 256, 0, s>>>(W2, nullptr, st, 1, 0, prior_scale, ps_idx + j0, st, nullptr, 0, W2, st, p2, T);
        HGP_LAUNCH_CHECK("hgp_mniw_loglik_batched: scale term");
        mniw_finish_kernel<<<(unsigned)((jb + 127) / 128), 128, 0, s>>>(p1, p2, nt * nt, info + j0, jb, out + j0);
        HGP_LAUNCH_CHECK("hgp_mniw_loglik_batched: finish");
    }
    return 0;
}
