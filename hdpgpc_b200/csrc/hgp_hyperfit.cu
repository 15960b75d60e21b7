// One-beat GP hyper-parameter fit, batched over candidate beats: reference IterativeGaussianProcess.fit_torch
// (hdpgpc/GPI.py:610-770, ExactGPModel branch) = gpytorch ExactGP(ConstantMean, ScaleKernel(RBF)) +
// GaussianLikelihood(Interval noise) trained with Adam(lr) on -MLL / T until the loss curve is flat.
// gpytorch is a third-party dependency (pinned 1.13, pyproject.toml:29); its published parameterisation is
// restated in oracle/hyperfit.py ("parity unpinned") and this kernel follows that restatement:
//   c = raw_c, s = softplus(raw_s), l = softplus(raw_l), noise = lo + (hi - lo) sigmoid(raw_n), all raw = 0 at start
//   K = s exp(-0.5 d^2 / l^2) + noise I;  loss = (0.5 r^T K^-1 r + sum log L_ii + 0.5 T log 2 pi) / T,  r = y - c
//   d loss / d theta = -(1 / 2T) sum((alpha alpha^T - K^-1) (.) dK/dtheta),  alpha = K^-1 r   (closed form, no autograd)
// One persistent CTA per fit runs all iterations (<= max_iter, stop rule GPI.py:695-698 evaluated on the device) with
// the CTA-level Cholesky / triangular solves of hgp_cta_la.cuh; nothing returns to the host in between.
#include "hgp_common.cuh"
#include "hgp_cta_la.cuh"
#include "hgp_smem_la.cuh"

using namespace hgp;

namespace {

__device__ __forceinline__ double block_sum(double v, LaSmem& sm) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm.red[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < LA_THREADS / 32; ++w) tot += sm.red[w];
    return tot;
}

__device__ __forceinline__ double softplus_d(double x) { return x > 20.0 ? x : log1p(exp(x)); }
__device__ __forceinline__ double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }

__global__ void __launch_bounds__(LA_THREADS)
hyperfit_kernel(const double* __restrict__ x, const double* __restrict__ Y, int T, double lo, double hi, double lr,
                int max_iter, int min_iter, double atol, double* __restrict__ out, double* __restrict__ work) {
    __shared__ LaSmem sm;
    __shared__ double s_raw[4], s_m[4], s_v[4], s_hist[11];
    __shared__ int s_stop, s_count;
    const int tid = threadIdx.x;
    if (tid == 0) sm.big = nullptr;
    const int64_t fit = blockIdx.x;
    const int64_t tt = (int64_t)T * T;
    double* E = work + fit * (3 * tt + 2 * T);   // exp(-0.5 d^2 / l^2)
    double* Kc = E + tt;                          // K, then its Cholesky factor
    double* Ki = Kc + tt;                         // K^-1
    double* r = Ki + tt;                          // y - c
    double* al = r + T;                           // alpha
    const double* y = Y + fit * T;
    if (tid < 4) { s_raw[tid] = 0.0; s_m[tid] = 0.0; s_v[tid] = 0.0; }
    if (tid == 0) { s_stop = 0; s_count = 0; }
    __syncthreads();
    double b1p = 1.0, b2p = 1.0;
    int info_any = 0;
    int it = 0;
    for (; it < max_iter; ++it) {
        const double c = s_raw[0];
        const double sc = softplus_d(s_raw[1]);
        const double ell = softplus_d(s_raw[2]);
        const double sg_n = sigmoid_d(s_raw[3]);
        const double noise = lo + (hi - lo) * sg_n;
        const double ell2 = ell * ell;
        for (int idx = tid; idx < T * T; idx += LA_THREADS) {
            const int i = idx / T, j = idx % T;
            const double d = x[i] - x[j];
            const double e = exp(-0.5 * (d * d) / ell2);
            E[idx] = e;
            Kc[idx] = sc * e + (i == j ? noise : 0.0);
            Ki[idx] = (i == j) ? 1.0 : 0.0;
        }
        for (int i = tid; i < T; i += LA_THREADS) r[i] = y[i] - c;
        __syncthreads();
        info_any |= la_chol(Kc, T, sm);
        la_trsm_lower(Kc, Ki, T, sm);
        la_trsm_lower_trans(Kc, Ki, T, sm);                 // Ki = K^-1
        la_gemv(al, Ki, r, T, 0.0, nullptr);
        __syncthreads();
        // reductions
        double p_quad = 0.0, p_logd = 0.0, p_sa = 0.0, p_gn = 0.0;
        for (int i = tid; i < T; i += LA_THREADS) {
            p_quad += r[i] * al[i];
            p_logd += log(Kc[(int64_t)i * T + i]);
            p_sa += al[i];
            p_gn += al[i] * al[i] - Ki[(int64_t)i * T + i];
        }
        double p_gs = 0.0, p_gl = 0.0;
        for (int idx = tid; idx < T * T; idx += LA_THREADS) {
            const int i = idx / T, j = idx % T;
            const double d = x[i] - x[j];
            const double w = (al[i] * al[j] - Ki[idx]) * E[idx];
            p_gs += w;
            p_gl += w * (d * d);
        }
        const double quad = block_sum(p_quad, sm), logd = block_sum(p_logd, sm), sa = block_sum(p_sa, sm);
        const double gn = block_sum(p_gn, sm), gs = block_sum(p_gs, sm), gl = block_sum(p_gl, sm);
        b1p *= 0.9;
        b2p *= 0.999;
        if (tid == 0) {
            const double loss = (0.5 * quad + logd + 0.5 * (double)T * HGP_LOG2PI) / (double)T;
            const double k = -0.5 / (double)T;
            double g[4];
            g[0] = -sa / (double)T;
            g[1] = k * gs * sigmoid_d(s_raw[1]);
            g[2] = k * (sc * gl / (ell2 * ell)) * sigmoid_d(s_raw[2]);
            g[3] = k * gn * (hi - lo) * sg_n * (1.0 - sg_n);
            const double bc1 = 1.0 - b1p, bc2 = 1.0 - b2p;
            for (int p = 0; p < 4; ++p) {
                s_m[p] = 0.9 * s_m[p] + (1.0 - 0.9) * g[p];
                s_v[p] = 0.999 * s_v[p] + (1.0 - 0.999) * g[p] * g[p];
                s_raw[p] -= (lr / bc1) * s_m[p] / (sqrt(s_v[p]) / sqrt(bc2) + 1e-8);
            }
            // stop rule (GPI.py:695-698): more than min_iter losses and sum(last ten differences) within atol of zero
            for (int h = 0; h < 10; ++h) s_hist[h] = s_hist[h + 1];
            s_hist[10] = loss;
            s_count += 1;
            if (s_count > min_iter) {
                double dsum = 0.0;
                for (int h = 0; h < 10; ++h) dsum += s_hist[h + 1] - s_hist[h];
                if (fabs(dsum) <= atol) s_stop = 1;
            }
        }
        __syncthreads();
        if (s_stop) { ++it; break; }
    }
    if (tid == 0) {
        double* o = out + fit * 8;
        o[0] = softplus_d(s_raw[1]);                         // outputscale
        o[1] = softplus_d(s_raw[2]);                         // lengthscale (the reference then overwrites it with 1.2)
        o[2] = lo + (hi - lo) * sigmoid_d(s_raw[3]);         // noise
        o[3] = s_raw[0];                                     // constant mean
        o[4] = s_hist[10];                                   // last loss
        o[5] = (double)it;                                   // iterations run
        o[6] = (double)info_any;
        o[7] = 0.0;
    }
}

// Small systems (T <= 92, the MIT-BIH beat length): the whole iteration lives in shared memory (hgp_smem_la.cuh).
// Buffer 0 holds E = exp(-0.5 d^2 / l^2), buffer 1 holds K (destroyed by the factorisation, then K^-1 = L^-T L^-1 as a
// triangular tensor-core product), buffer 2 the inverse factor; nothing but x and y is read from global memory.
__device__ __forceinline__ double block_sum_sl(double v, SlCtx& c) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) c.red[threadIdx.x >> 5] = v;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < SL_THREADS / 32; ++w) tot += c.red[w];
    return tot;
}

__global__ void __launch_bounds__(SL_THREADS, 1)
hyperfit_kernel_small(const double* __restrict__ x, const double* __restrict__ Y, int T, double lo, double hi, double lr,
                      int max_iter, int min_iter, double atol, double* __restrict__ out) {
    __shared__ SlCtx c;
    __shared__ double s_raw[4], s_m[4], s_v[4], s_hist[11];
    __shared__ double s_x[96], s_r[96], s_al[96];
    __shared__ int s_stop, s_count;
    sl_init(c, T);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int LD = c.LD;
    const int64_t fit = blockIdx.x;
    const double* y = Y + fit * T;
    double* E = sl_dyn;
    double* K = sl_dyn + c.stride;
    if (tid < 4) { s_raw[tid] = 0.0; s_m[tid] = 0.0; s_v[tid] = 0.0; }
    if (tid == 0) { s_stop = 0; s_count = 0; }
    for (int i = tid; i < T; i += SL_THREADS) s_x[i] = x[i];
    __syncthreads();
    double b1p = 1.0, b2p = 1.0;
    int info_any = 0;
    int it = 0;
    for (; it < max_iter; ++it) {
        const double cm = s_raw[0];
        const double sc = softplus_d(s_raw[1]);
        const double ell = softplus_d(s_raw[2]);
        const double sg_n = sigmoid_d(s_raw[3]);
        const double noise = lo + (hi - lo) * sg_n;
        const double ell2 = ell * ell;
        for (int idx = tid; idx < T * T; idx += SL_THREADS) {
            const int i = idx / T, j = idx - i * T;
            const double d = s_x[i] - s_x[j];
            const double e = exp(-0.5 * (d * d) / ell2);
            E[i * LD + j] = e;
            K[i * LD + j] = sc * e + (i == j ? noise : 0.0);
        }
        for (int i = tid; i < T; i += SL_THREADS) s_r[i] = y[i] - cm;
        __syncthreads();
        info_any |= sl_cholinv_buf(c, 1, 2, 0.0);               // buffer 2 = L^-1, c.logdet = 2 sum log L_ii
        const double logd = 0.5 * c.logdet;
        sl_gemm_buf(c, 1, 2, 1, 2, 0, SL_TRI_A | SL_TRI_B);     // buffer 1 = K^-1 = L^-T L^-1
        // alpha = K^-1 r: one warp per row
        for (int r = warp; r < T; r += SL_THREADS / 32) {
            double a = 0.0;
            for (int k = lane; k < T; k += 32) a += K[r * LD + k] * s_r[k];
            a = warp_sum(a);
            if (lane == 0) s_al[r] = a;
        }
        __syncthreads();
        double p_quad = 0.0, p_sa = 0.0, p_gn = 0.0;
        for (int i = tid; i < T; i += SL_THREADS) {
            p_quad += s_r[i] * s_al[i];
            p_sa += s_al[i];
            p_gn += s_al[i] * s_al[i] - K[i * LD + i];
        }
        double p_gs = 0.0, p_gl = 0.0;
        for (int idx = tid; idx < T * T; idx += SL_THREADS) {
            const int i = idx / T, j = idx - i * T;
            const double d = s_x[i] - s_x[j];
            const double w = (s_al[i] * s_al[j] - K[i * LD + j]) * E[i * LD + j];
            p_gs += w;
            p_gl += w * (d * d);
        }
        const double quad = block_sum_sl(p_quad, c), sa = block_sum_sl(p_sa, c);
        const double gn = block_sum_sl(p_gn, c), gs = block_sum_sl(p_gs, c), gl = block_sum_sl(p_gl, c);
        b1p *= 0.9;
        b2p *= 0.999;
        if (tid == 0) {
            const double loss = (0.5 * quad + logd + 0.5 * (double)T * HGP_LOG2PI) / (double)T;
            const double k = -0.5 / (double)T;
            double g[4];
            g[0] = -sa / (double)T;
            g[1] = k * gs * sigmoid_d(s_raw[1]);
            g[2] = k * (sc * gl / (ell2 * ell)) * sigmoid_d(s_raw[2]);
            g[3] = k * gn * (hi - lo) * sg_n * (1.0 - sg_n);
            const double bc1 = 1.0 - b1p, bc2 = 1.0 - b2p;
            for (int p = 0; p < 4; ++p) {
                s_m[p] = 0.9 * s_m[p] + (1.0 - 0.9) * g[p];
                s_v[p] = 0.999 * s_v[p] + (1.0 - 0.999) * g[p] * g[p];
                s_raw[p] -= (lr / bc1) * s_m[p] / (sqrt(s_v[p]) / sqrt(bc2) + 1e-8);
            }
            for (int h = 0; h < 10; ++h) s_hist[h] = s_hist[h + 1];
            s_hist[10] = loss;
            s_count += 1;
            if (s_count > min_iter) {
                double dsum = 0.0;
                for (int h = 0; h < 10; ++h) dsum += s_hist[h + 1] - s_hist[h];
                if (fabs(dsum) <= atol) s_stop = 1;
            }
        }
        __syncthreads();
        if (s_stop) { ++it; break; }
    }
    if (tid == 0) {
        double* o = out + fit * 8;
        o[0] = softplus_d(s_raw[1]);
        o[1] = softplus_d(s_raw[2]);
        o[2] = lo + (hi - lo) * sigmoid_d(s_raw[3]);
        o[3] = s_raw[0];
        o[4] = s_hist[10];
        o[5] = (double)it;
        o[6] = (double)info_any;
        o[7] = 0.0;
    }
}

}  // namespace

extern "C" int64_t hgp_hyperfit_work_doubles(int n_fits, int T) { return (int64_t)n_fits * (3 * (int64_t)T * T + 2 * T); }

extern "C" int hgp_hyperfit_batched(const double* x, const double* Y, int n_fits, int T, double noise_lo, double noise_hi,
                                    double lr, int max_iter, int min_iter, double atol, double* out, double* work,
                                    void* stream) {
    HGP_REQUIRE(n_fits >= 0 && T > 0 && T <= 1024 && max_iter >= 0 && noise_hi >= noise_lo,
                "hgp_hyperfit_batched: bad arguments");
    if (n_fits == 0) return 0;
    if (sl_supported(T) && !getenv("HGP_HYPERFIT_V1")) {
        const size_t sdyn = sl_dynamic_smem_bytes(T);
        cudaError_t e = cudaFuncSetAttribute(hyperfit_kernel_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sdyn);
        if (e != cudaSuccess) return hgp_status(e, "hgp_hyperfit_batched: shared memory");
        hyperfit_kernel_small<<<(unsigned)n_fits, SL_THREADS, sdyn, (cudaStream_t)stream>>>(x, Y, T, noise_lo, noise_hi, lr,
                                                                                            max_iter, min_iter, atol, out);
        HGP_LAUNCH_CHECK("hgp_hyperfit_batched");
        return 0;
    }
    hyperfit_kernel<<<(unsigned)n_fits, LA_THREADS, 0, (cudaStream_t)stream>>>(x, Y, T, noise_lo, noise_hi, lr, max_iter,
                                                                             min_iter, atol, out, work);
    HGP_LAUNCH_CHECK("hgp_hyperfit_batched");
    return 0;
}
