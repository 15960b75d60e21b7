// Batched alignment ("warp") of beats against a cluster's representative beat, and the warp-prior covariance.
//
// Reference: Warping_system.compute_warp_batch (hdpgpc/amtgp_warping_system.py:548-736): per beat a monotone
// time-warp g(t) parameterised by n_ctrl unconstrained control values (linear expansion to T points, softplus,
// cumulative sum, normalisation to [x_min, x_max]), fitted by `train_iter` Adam steps on
//     0.5 * |y(g(t)) - y_model(t)|^2 / noise + lam_s * |D2 (g - x)|^2 + lam_a * |g - x|^2.
// The reference differentiates with autograd; here the gradient is written out in closed form and the whole fit of
// one (beat, representative) pair runs inside ONE WARP: lane l owns the contiguous samples [l*seg, (l+1)*seg), the
// two cumulative sums are a sequential pass over the lane's samples plus a warp shuffle scan over lanes (the very
// first forward pass is strictly sequential, see the comment in the loop), the
// control-point gradient is a warp reduction per control point, Adam state lives in the registers of lanes
// 0..n_ctrl-1.  No global memory traffic inside the optimisation loop: beat, template, grid and all per-sample
// intermediates stay in shared memory (5T + n_ctrl doubles per warp).
#include "hgp_common.cuh"

namespace {

constexpr int WARP_FITS_PER_CTA = 4;

__device__ __forceinline__ double warp_excl_scan(double v, double& total) {
    const int lane = threadIdx.x & 31;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - v;
}

__device__ __forceinline__ double warp_excl_scan_rev(double v) {   // sum over lanes > lane
    const int lane = threadIdx.x & 31;
    double inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double n = __shfl_down_sync(0xffffffffu, inc, o);
        if (lane + o < 32) inc += n;
    }
    return inc - v;
}

struct WarpFitArgs {
    const double* x;        // [T]
    const double* Y;        // [N][T]
    const double* Ym;       // [R][T]
    const double* u0;       // [R][n_ctrl] or null
    const double* gscale;   // [N] or null (1)
    int64_t N;
    int R, T, n_ctrl, iters;
    double lr, noise, lam_s, lam_a;
    double* xw;             // [R][N][T]
    double* yw;             // [R][N][T]
    double* u_out;          // [R][N][n_ctrl] or null
    double* loss_trace;     // [iters][R][N] or null
};

__global__ void __launch_bounds__(WARP_FITS_PER_CTA * 32)
warp_fit_kernel(WarpFitArgs a) {
    extern __shared__ double smem[];
    const int T = a.T, nc = a.n_ctrl;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // CTA-shared: grid x[T], control expansion weight lam[T], source index i0[T] (as int in a double slot pair)
    double* sx = smem;
    double* slam = sx + T;
    int* si0 = reinterpret_cast<int*>(slam + T);
    double* wbase = reinterpret_cast<double*>(si0 + ((T + 1) & ~1)) + (size_t)wid * (5 * T + ((nc + 1) & ~1));
    double* sy = wbase;            // the beat
    double* sym_ = sy + T;         // the representative beat (template)
    double* sr = sym_ + T;         // normalised cumulative sum r[t]
    double* sxw = sr + T;          // x_warp[t] = g[t] - x[t]
    double* sg = sxw + T;          // data gradient wrt g, later d/d inc, and sigmoid(uT) packed after use
    double* su = sg + T;           // control values

    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        sx[t] = a.x[t];
        // F.interpolate(mode="linear", align_corners=True): src = t (nc-1)/(T-1)
        const double scale = T > 1 ? (double)(nc - 1) / (double)(T - 1) : 0.0;
        const double src = scale * (double)t;
        int i0 = (int)floor(src);
        if (i0 > nc - 1) i0 = nc - 1;
        si0[t] = i0;
        slam[t] = src - (double)i0;
    }
    const int64_t fit = (int64_t)blockIdx.x * WARP_FITS_PER_CTA + wid;
    const int64_t n_fits = a.N * a.R;
    const bool active = fit < n_fits;
    const int r_id = active ? (int)(fit / a.N) : 0;
    const int64_t n_id = active ? fit % a.N : 0;
    if (active) {
        const double* yrow = a.Y + n_id * T;
        const double* mrow = a.Ym + (int64_t)r_id * T;
        for (int t = lane; t < T; t += 32) { sy[t] = yrow[t]; sym_[t] = mrow[t]; }
        if (lane < nc) su[lane] = a.u0 ? a.u0[(int64_t)r_id * nc + lane] : 0.0;
    }
    __syncthreads();
    if (!active) return;

    const int seg = (T + 31) / 32;
    const int t0 = min(lane * seg, T), t1 = min(t0 + seg, T);
    const double x_min = sx[0], x_max = sx[T - 1], span = x_max - x_min;
    const double inv_noise = 1.0 / (a.noise + 1e-12);
    const double gs = a.gscale ? a.gscale[n_id] : 1.0;
    double adam_m = 0.0, adam_v = 0.0, b1p = 1.0, b2p = 1.0;

    for (int it = 0; it <= a.iters; ++it) {
        const bool last = it == a.iters;
        // ---- control expansion, softplus, cumulative sum.
        // Iteration 0 reproduces torch.cumsum's sequential order bit for bit (and the forward pass below avoids FMA
        // contraction): when a beat is aligned against itself -- every cluster's representative beat is -- the
        // residual at the start is pure rounding of g(t) around the grid nodes, Adam's normalisation turns it into a
        // full-size first step, and only the same bits give the reference's answer.  Later iterations are smooth in
        // the rounding, so they use a lane-local pass plus a warp shuffle scan.
        for (int t = t0; t < t1; ++t) {
            const int i0 = si0[t];
            const int i1 = min(i0 + 1, nc - 1);
            const double lam = slam[t];
            const double uT = __dadd_rn(__dmul_rn(1.0 - lam, su[i0]), __dmul_rn(lam, su[i1]));
            const double e = exp(fmin(uT, 20.0));
            const double sp = uT > 20.0 ? uT : log1p(e);
            sr[t] = sp + 1e-6;
            sg[t] = uT;
        }
        double off = 0.0, total;
        if (it == 0) {
            __syncwarp();
            if (lane == 0) {
                double run = 0.0;
                for (int t = 0; t < T; ++t) { run = __dadd_rn(run, sr[t]); sr[t] = run; }
            }
            __syncwarp();
            total = sr[T - 1];
        } else {
            double run = 0.0;
            for (int t = t0; t < t1; ++t) { run += sr[t]; sr[t] = run; }
            off = warp_excl_scan(run, total);
        }
        const double first_inc = __shfl_sync(0xffffffffu, sr[min(t0, T - 1)], 0);   // s[0]
        const double den = __dadd_rn(__dsub_rn(total, first_inc), 1e-12);
        // ---- grid, warp offset, interpolation, data gradient
        double sse = 0.0, ap = 0.0;
        for (int t = t0; t < t1; ++t) {
            const double r = __ddiv_rn(__dsub_rn(off + sr[t], first_inc), den);
            const double g = __dadd_rn(x_min, __dmul_rn(span, r));
            sr[t] = r;
            sxw[t] = g - sx[t];
            ap += (g - sx[t]) * (g - sx[t]);
        }
        __syncwarp();
        double dr_sum = 0.0, drr_sum = 0.0, spen = 0.0;
        for (int t = t0; t < t1; ++t) {
            const double g = __dadd_rn(x_min, __dmul_rn(span, sr[t]));
            const double gq = fmin(fmax(g, x_min), x_max);
            // searchsorted(x, gq, right=False): first index with x[idx] >= gq; clamp to [1, T-1]
            int lo_i = 0, hi_i = T;
            while (lo_i < hi_i) {
                const int mid = (lo_i + hi_i) >> 1;
                if (sx[mid] < gq) lo_i = mid + 1; else hi_i = mid;
            }
            const int hi = min(max(lo_i, 1), T - 1), lo = hi - 1;
            const double dxi = __dadd_rn(sx[hi] - sx[lo], 1e-12);
            const double w = __ddiv_rn(gq - sx[lo], dxi);
            const double ylo = sy[lo], yhi = sy[hi];
            const double ywv = __dadd_rn(__dmul_rn(1.0 - w, ylo), __dmul_rn(w, yhi));
            const double resid = ywv - sym_[t];
            sse += resid * resid;
            if (last) {
                const int64_t o = ((int64_t)r_id * a.N + n_id) * T + t;
                a.xw[o] = sxw[t];
                a.yw[o] = ywv;
                continue;
            }
            const bool inside = (g >= x_min) && (g <= x_max);
            double dgv = inside ? resid * inv_noise * (yhi - ylo) / dxi : 0.0;
            // penalties: d2[k] = xw[k] - 2 xw[k+1] + xw[k+2], k in [0, T-3]
            double dpen = 2.0 * a.lam_a * sxw[t];
            double acc = 0.0;
            if (t + 2 < T) {
                const double d2 = sxw[t] - 2.0 * sxw[t + 1] + sxw[t + 2];
                acc += d2;
                spen += d2 * d2;
            }
            if (t >= 1 && t + 1 < T) acc -= 2.0 * (sxw[t - 1] - 2.0 * sxw[t] + sxw[t + 1]);
            if (t >= 2) acc += sxw[t - 2] - 2.0 * sxw[t - 1] + sxw[t];
            dpen += 2.0 * a.lam_s * acc;
            const double dr = span * gs * (dgv + dpen);
            dr_sum += dr;
            drr_sum += dr * sr[t];
            sr[t] = dr;      // r[t] is not needed past this point (own sample only)
        }
        if (last) break;
        if (a.loss_trace) {
            const double l_sse = warp_sum(sse), l_sp = warp_sum(spen), l_ap = warp_sum(ap);
            if (lane == 0)
                a.loss_trace[((int64_t)it * a.R + r_id) * a.N + n_id] =
                    0.5 * l_sse * inv_noise + a.lam_s * l_sp + a.lam_a * l_ap;
        }
        dr_sum = warp_sum(dr_sum);
        drr_sum = warp_sum(drr_sum);
        const double dden = -drr_sum / den;
        // ---- d/ds, reverse cumulative sum -> d/d inc, times sigmoid -> d/d uT
        double rrun = 0.0;
        for (int t = t1 - 1; t >= t0; --t) {
            double ds = sr[t] / den;
            if (t == 0) ds -= dr_sum / den + dden;
            if (t == T - 1) ds += dden;
            rrun += ds;
            sr[t] = rrun;
        }
        const double roff = warp_excl_scan_rev(rrun);
        for (int t = t0; t < t1; ++t) {
            const double uT = sg[t];
            sg[t] = (roff + sr[t]) / (1.0 + exp(-uT));
        }
        // ---- control-point gradient (warp reduction per control point) and Adam on lanes < n_ctrl
        double my_grad = 0.0;
        for (int c = 0; c < nc; ++c) {
            double part = 0.0;
            for (int t = t0; t < t1; ++t) {
                const int i0 = si0[t];
                const int i1 = min(i0 + 1, nc - 1);
                const double lam = slam[t];
                if (i0 == c) part += (1.0 - lam) * sg[t];
                if (i1 == c) part += lam * sg[t];
            }
            part = warp_sum(part);
            if (lane == c) my_grad = part;
        }
        b1p *= 0.9;
        b2p *= 0.999;
        __syncwarp();
        if (lane < nc) {
            // torch.optim.Adam (single-tensor path): denom = sqrt(v)/sqrt(bc2) + eps; p -= (lr/bc1) * m / denom
            adam_m = 0.9 * adam_m + (1.0 - 0.9) * my_grad;
            adam_v = 0.999 * adam_v + (1.0 - 0.999) * my_grad * my_grad;
            const double bc1 = 1.0 - b1p, bc2 = 1.0 - b2p;
            const double denom = sqrt(adam_v) / sqrt(bc2) + 1e-8;
            su[lane] -= (a.lr / bc1) * adam_m / denom;
        }
        __syncwarp();
    }
    if (a.u_out && lane < nc) a.u_out[((int64_t)r_id * a.N + n_id) * nc + lane] = su[lane];
}

__global__ void warp_prior_cov_kernel(const double* __restrict__ x, int T, double rho, double omega, double diag_add,
                                      int normalize, double* __restrict__ K) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * T) return;
    const int r = idx / T, c = idx % T;
    double xr = x[r], xc = x[c];
    if (normalize) {
        // WarpPriorAMTGP._rbf_cov (amtgp_warping_system.py:162-166)
        const double x0 = x[0];
        const double rng = fabs((x[T - 1] - x0) - (x[0] - x0)) + 1e-12;
        xr = (xr - x0) / rng;
        xc = (xc - x0) / rng;
    }
    const double dx = xr - xc;
    double v = (omega * omega) * exp(-0.5 * (dx * dx) / (rho * rho));
    if (r == c) v += diag_add;
    K[idx] = v;
}

}  // namespace

extern "C" int hgp_warp_fit_batched(const double* x_model, int T, const double* Y, int64_t N, const double* y_model, int R,
                                    const double* u0, int n_ctrl, int train_iter, double lr, double noise, double lam_s,
                                    double lam_a, const double* grad_scale, double* x_warp, double* y_warp, double* u_out,
                                    double* loss_trace, void* stream) {
    HGP_REQUIRE(T >= 3 && T <= 1024 && N >= 0 && R >= 0 && n_ctrl >= 2 && n_ctrl <= 32 && n_ctrl <= T && train_iter >= 0,
                "hgp_warp_fit_batched: bad sizes (3 <= T <= 1024, 2 <= n_ctrl <= min(32, T))");
    if (N == 0 || R == 0) return 0;
    WarpFitArgs a{x_model, Y, y_model, u0, grad_scale, N, R, T, n_ctrl, train_iter, lr, noise, lam_s, lam_a,
                  x_warp, y_warp, u_out, loss_trace};
    const size_t shared = sizeof(double) * (2 * (size_t)T) + sizeof(int) * (size_t)((T + 1) & ~1) +
                          sizeof(double) * WARP_FITS_PER_CTA * (5 * (size_t)T + ((n_ctrl + 1) & ~1));
    cudaError_t e = cudaFuncSetAttribute(warp_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shared);
    if (e != cudaSuccess) return hgp_status(e, "hgp_warp_fit_batched: shared memory");
    const int64_t n_fits = N * (int64_t)R;
    const int64_t grid = (n_fits + WARP_FITS_PER_CTA - 1) / WARP_FITS_PER_CTA;
    HGP_REQUIRE(grid < (1ll << 31), "hgp_warp_fit_batched: too many fits for one launch");
    warp_fit_kernel<<<(unsigned)grid, WARP_FITS_PER_CTA * 32, shared, (cudaStream_t)stream>>>(a);
    HGP_LAUNCH_CHECK("hgp_warp_fit_batched");
    return 0;
}

extern "C" int hgp_warp_prior_cov(const double* x, int T, double rho, double omega, double diag_add, int normalize_x,
                                  double* K, void* stream) {
    HGP_REQUIRE(T >= 1 && T <= 4096, "hgp_warp_prior_cov: bad T");
    const int n = T * T;
    warp_prior_cov_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, T, rho, omega, diag_add, normalize_x, K);
    HGP_LAUNCH_CHECK("hgp_warp_prior_cov");
    return 0;
}
