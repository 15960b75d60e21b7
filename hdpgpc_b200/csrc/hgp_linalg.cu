// Batched SPD factorisation, triangular inverse and operand packing (sm_100a).
//
// chol_batched restates GPI_model._chol_spd (reference GPI_model.py:83-87) for a batch of
// T x T float64 matrices, one CTA per matrix, blocked right-looking with the current panel in
// shared memory and the trailing matrix in L2.  tri_inverse_batched produces W = L^{-1} so the
// Mahalanobis term becomes a single triangular product (see include/hdpgpc_b200.h).
#include "hgp_common.cuh"
#include <stdlib.h>

namespace {

constexpr int NB = 16;           // panel width
constexpr int CHOL_THREADS = 256;

__global__ void pack_leads_kernel(const double* __restrict__ Y, int64_t N, int T, int L, double* __restrict__ out,
                                  int64_t plane_stride) {
    int64_t total = N * (int64_t)T * L;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // i indexes the output [L][N][T] (planes plane_stride doubles apart)
        int64_t t = i % T;
        int64_t n = (i / T) % N;
        int64_t ld = i / ((int64_t)T * N);
        out[ld * plane_stride + n * T + t] = Y[(n * T + t) * L + ld];
    }
}

__global__ void __launch_bounds__(CHOL_THREADS)
chol_kernel(const double* __restrict__ Sigma, const int* __restrict__ src_idx, const double* __restrict__ src_scale,
            int T, const double* __restrict__ add_diag, double jitter_scale,
            double* __restrict__ Lfac, double* __restrict__ logdet, int* __restrict__ info) {
    extern __shared__ double smem[];
    double* Dk = smem;                       // [NB][NB+1]
    double* P = smem + NB * (NB + 1);        // [T][NB+1]
    __shared__ double s_red[CHOL_THREADS / 32];
    __shared__ double s_jit;
    __shared__ int s_info;

    const int64_t f = blockIdx.x;
    const double* S = Sigma + (src_idx ? (int64_t)src_idx[f] : f) * (int64_t)T * T;
    const double sc = src_scale ? src_scale[f] : 1.0;     // matrix = sc * Sigma[src]
    double* A = Lfac + f * (int64_t)T * T;
    const int tid = threadIdx.x;
    const double add = add_diag ? add_diag[f] : 0.0;

    // diag mean of (S + add I)
    double part = 0.0;
    for (int i = tid; i < T; i += CHOL_THREADS) part += fabs(S[(int64_t)i * T + i] * sc + add);
    part = warp_sum(part);
    if ((tid & 31) == 0) s_red[tid >> 5] = part;
    if (tid == 0) s_info = 0;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < CHOL_THREADS / 32; ++w) tot += s_red[w];
        s_jit = jitter_scale * fmax(tot / T, HGP_EPS);
    }
    __syncthreads();
    const double jit = s_jit;
    for (int idx = tid; idx < T * T; idx += CHOL_THREADS) {
        int i = idx / T, j = idx % T;
        double v;
        if (j > i) v = 0.0;
        else if (j == i) v = (S[idx] * sc + add) + jit;   // sym() leaves the diagonal unchanged
        else v = 0.5 * (S[idx] * sc + S[(int64_t)j * T + i] * sc);
        A[idx] = v;
    }
    __syncthreads();

    double ld_acc = 0.0;
    for (int k0 = 0; k0 < T; k0 += NB) {
        const int nb = min(NB, T - k0);
        // diagonal block -> shared
        for (int idx = tid; idx < nb * nb; idx += CHOL_THREADS) {
            int i = idx / nb, j = idx % nb;
            Dk[i * (NB + 1) + j] = A[(int64_t)(k0 + i) * T + k0 + j];
        }
        __syncthreads();
        if (tid < 32) {
            for (int j = 0; j < nb; ++j) {
                double d = Dk[j * (NB + 1) + j];
                if (!(d > 0.0) && tid == 0 && s_info == 0) s_info = k0 + j + 1;
                double s = sqrt(d);
                __syncwarp();
                if (tid == 0) Dk[j * (NB + 1) + j] = s;
                if (tid > j && tid < nb) Dk[tid * (NB + 1) + j] /= s;
                __syncwarp();
                if (tid > j && tid < nb) {
                    double lij = Dk[tid * (NB + 1) + j];
                    for (int c = j + 1; c <= tid; ++c) Dk[tid * (NB + 1) + c] -= lij * Dk[c * (NB + 1) + j];
                }
                __syncwarp();
            }
        }
        __syncthreads();
        for (int idx = tid; idx < nb * nb; idx += CHOL_THREADS) {
            int i = idx / nb, j = idx % nb;
            if (j <= i) A[(int64_t)(k0 + i) * T + k0 + j] = Dk[i * (NB + 1) + j];
        }
        if (tid == 0) for (int j = 0; j < nb; ++j) ld_acc += log(Dk[j * (NB + 1) + j]);
        const int r0 = k0 + nb;      // first trailing row
        const int nt = T - r0;
        // panel solve: X L_kk^T = A[r0:, k0:k0+nb], one thread per row
        for (int r = tid; r < nt; r += CHOL_THREADS) {
            double x[NB];
            double* arow = A + (int64_t)(r0 + r) * T + k0;
#pragma unroll
            for (int c = 0; c < NB; ++c) {
                if (c < nb) {
                    double v = arow[c];
                    for (int p = 0; p < c; ++p) v -= x[p] * Dk[c * (NB + 1) + p];
                    x[c] = v / Dk[c * (NB + 1) + c];
                    arow[c] = x[c];
                    P[r * (NB + 1) + c] = x[c];
                }
            }
        }
        __syncthreads();
        // trailing update (lower part only): A[i][j] -= P[i] . P[j]
        const int ty = tid >> 4, tx = tid & 15;
        for (int ib = 0; ib < nt; ib += 16) {
            int i = ib + ty;
            for (int jb = 0; jb <= ib; jb += 16) {
                int j = jb + tx;
                if (i < nt && j <= i) {
                    double acc = 0.0;
#pragma unroll
                    for (int c = 0; c < NB; ++c)
                        if (c < nb) acc += P[i * (NB + 1) + c] * P[j * (NB + 1) + c];
                    A[(int64_t)(r0 + i) * T + r0 + j] -= acc;
                }
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (logdet) logdet[f] = 2.0 * ld_acc;
        info[f] = s_info;
    }
}

// W = L^{-1} by blocked forward substitution on the identity.  Column blocks of the inverse are independent of one
// another (column j of W only needs L and the rows of column j above it), so one matrix is spread over gridDim.y CTAs,
// each owning a group of 16-column blocks: the sweep over the row blocks is the serial part, and with 8 groups per
// 256 x 256 factor a table build of 128 factors fills the machine instead of occupying 128 of its 148 SMs for the
// whole serial sweep.
__global__ void __launch_bounds__(256)
tri_inverse_kernel(const double* __restrict__ Lfac, int T, double* __restrict__ Wout) {
    extern __shared__ double smem[];
    double* Lii = smem;                      // [NB][NB+1]
    double* Inv = Lii + NB * (NB + 1);       // [NB][NB+1]
    double* R = Inv + NB * (NB + 1);         // [nblk][NB][NB+1]
    const int64_t f = blockIdx.x;
    const double* Lm = Lfac + f * (int64_t)T * T;
    double* W = Wout + f * (int64_t)T * T;
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int nblk = (T + NB - 1) / NB;
    const int per = (nblk + gridDim.y - 1) / gridDim.y;
    const int bj_lo = blockIdx.y * per, bj_hi = min(nblk, bj_lo + per);     // this CTA's column blocks
    if (bj_lo >= nblk) return;
    const int c_lo = bj_lo * NB, c_hi = min(T, bj_hi * NB);

    for (int bi = 0; bi < nblk; ++bi) {
        const int r0 = bi * NB;
        const int nb = min(NB, T - r0);
        // zero the strict upper part of this block row inside the CTA's columns (columns > r0+row)
        for (int idx = tid; idx < nb * (c_hi - c_lo); idx += 256) {
            int i = idx / (c_hi - c_lo), j = c_lo + idx % (c_hi - c_lo);
            if (j > r0 + i) W[(int64_t)(r0 + i) * T + j] = 0.0;
        }
        if (bi < bj_lo) continue;            // rows above the group's first diagonal block hold zeros only
        if (ty < nb && tx < nb) Lii[ty * (NB + 1) + tx] = Lm[(int64_t)(r0 + ty) * T + r0 + tx];
        __syncthreads();
        // Inv = Lii^{-1}: thread c solves column c
        if (tid < nb) {
            const int c = tid;
            double x[NB];
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                if (i < nb) {
                    if (i < c) x[i] = 0.0;
                    else if (i == c) x[i] = 1.0 / Lii[i * (NB + 1) + i];
                    else {
                        double v = 0.0;
                        for (int p = c; p < i; ++p) v += Lii[i * (NB + 1) + p] * x[p];
                        x[i] = -v / Lii[i * (NB + 1) + i];
                    }
                    Inv[i * (NB + 1) + c] = x[i];
                }
            }
        }
        // R[bj] = sum_{k in [c0, r0)} L[r0+ty][k] * W[k][c0+tx]
        const int bj_end = min(bi, bj_hi);
        for (int bj = bj_lo; bj < bj_end; ++bj) {
            const int c0 = bj * NB;
            double acc = 0.0;
            if (ty < nb) {
                const double* lrow = Lm + (int64_t)(r0 + ty) * T;
                for (int k = c0; k < r0; ++k) acc += lrow[k] * W[(int64_t)k * T + c0 + tx];
            }
            R[(bj * NB + ty) * (NB + 1) + tx] = acc;
        }
        __syncthreads();
        for (int bj = bj_lo; bj < bj_end; ++bj) {
            const int c0 = bj * NB;
            if (ty < nb) {
                double acc = 0.0;
                for (int p = 0; p <= ty; ++p) acc += Inv[ty * (NB + 1) + p] * R[(bj * NB + p) * (NB + 1) + tx];
                W[(int64_t)(r0 + ty) * T + c0 + tx] = -acc;
            }
        }
        if (bi < bj_hi && ty < nb && tx < nb && tx <= ty) W[(int64_t)(r0 + ty) * T + r0 + tx] = Inv[ty * (NB + 1) + tx];
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// Factor AND inverse factor of one SPD matrix per CTA in a single left-looking sweep (T <= 512): what a table build needs
// per cluster covariance (hgp_cholinv_batched).  The two kernels above are latency chains: the right-looking Cholesky
// updates the whole trailing matrix in global memory with scalar FMAs once per 16-column panel (0.80 ms for the 128
// factors of a cfg4 table build), the substitution for L^-1 walks scalar k loops over global memory (0.94 ms).  Here
// every panel is two tensor-core products against what is already final:
//   U          = A[k0:, k0:k1] - L[k0:, 0:k0] L[k0:k1, 0:k0]^T       (the panel, never read again after this)
//   L[k0:k1,.] = chol(U[0:16]) (one warp, out of registers),  L[k1:, k0:k1] = U[16:] L_kk^-T
//   W[k0:k1, 0:k0] = -L_kk^-1 (L[k0:k1, 0:k0] W[0:k0, 0:k0]),   W[k0:k1, k0:k1] = L_kk^-1
// with the 16 x k0 row block L[k0:k1, 0:k0] staged in shared memory as the B operand of the first product and the A
// operand of the second; each matrix element is read O(T / 16) times through L2 and written once.  256 threads (the
// register-resident diagonal factorisation wants more than the 128 registers a 512-thread CTA leaves per thread), 16
// loads in flight per lane.
#ifndef HGP_CI_THREADS
#define HGP_CI_THREADS 256
#endif
constexpr int CI_THREADS = HGP_CI_THREADS;

// Pre-pass of the fused factorisation, fully parallel and bandwidth-bound: X = lower triangle of sym(A) with the diagonal
// rule of _chol_spd applied (A_ii + add + jitter, jitter from the mean |diagonal|), zeros above the diagonal of X and Y.
// The covariances come cold from HBM; read inside the latency-bound sweep they cost a DRAM round trip per row-tile trip
// (0.88 ms for 128 factors against 0.45 ms warm).  After this pass the sweep finds its panel entries in L2.
__global__ void __launch_bounds__(256)
cholinv_prepare_kernel(const double* __restrict__ Sigma, int T, const double* __restrict__ add_diag, double jitter_scale,
                       double* __restrict__ Lfac, double* __restrict__ Wout) {
    __shared__ double red[8];
    __shared__ double s_jit;
    const int64_t f = blockIdx.x;
    const double* S = Sigma + f * (int64_t)T * T;
    double* X = Lfac + f * (int64_t)T * T;
    double* Y = Wout + f * (int64_t)T * T;
    const int tid = threadIdx.x;
    const double add = add_diag ? add_diag[f] : 0.0;
    double part = 0.0;
    for (int i = tid; i < T; i += 256) part += fabs(S[(int64_t)i * T + i] + add);
    part = warp_sum(part);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += red[w];
        s_jit = jitter_scale * fmax(tot / T, HGP_EPS);
    }
    __syncthreads();
    const double jit = s_jit;
    // this CTA's slab of rows (gridDim.y slabs per matrix)
    const int rows_per = (T + gridDim.y - 1) / gridDim.y;
    const int r_lo = blockIdx.y * rows_per, r_hi = min(T, r_lo + rows_per);
    for (int idx = r_lo * T + tid; idx < r_hi * T; idx += 256) {
        const int i = idx / T, j = idx - i * T;
        double v = 0.0;
        if (j == i) v = (S[idx] + add) + jit;                 // sym() keeps the diagonal
        else if (j < i) v = 0.5 * (S[idx] + S[(int64_t)j * T + i]);
        X[idx] = v;
        if (j > i) Y[idx] = 0.0;
    }
}


constexpr int CI_NB = 16;
__host__ __device__ inline int ci_ldb(int T) { return ((T + 7) / 8) * 8 + 4; }      // == 4 (mod 8): conflict-free fragments
__host__ __device__ inline size_t ci_smem_bytes(int T) {
    const int TP = (T + 15) & ~15;
    return sizeof(double) * ((size_t)TP * 20 + 2 * (size_t)CI_NB * ci_ldb(T) + 2 * CI_NB * (CI_NB + 1) + 64);
}

#ifndef HGP_CI_MINBLOCKS
#define HGP_CI_MINBLOCKS 1
#endif
__global__ void __launch_bounds__(CI_THREADS, HGP_CI_MINBLOCKS)
cholinv_kernel(int T, double* __restrict__ Lfac, double* __restrict__ Wout, double* __restrict__ logdet,
               int* __restrict__ info) {
    extern __shared__ __align__(16) double ci_smem[];
    const int TP = (T + 15) & ~15;
    const int LDB = ci_ldb(T);
    double* Ps = ci_smem;                          // [TP][20]   the panel U (rows relative to k0)
    double* Bs = Ps + (size_t)TP * 20;             // [16][LDB]  L[k0:k1, 0:k0]
    double* Fs = Bs + CI_NB * LDB;                 // [16][LDB]  L[k0:k1, 0:k0] W[0:k0, 0:k0]
    double* Dk = Fs + CI_NB * LDB;                 // [16][17]   L_kk (+ 1 / diagonal in column 16)
    double* Wi = Dk + CI_NB * (CI_NB + 1);         // [16][17]   L_kk^-1
    double* red = Wi + CI_NB * (CI_NB + 1);
    __shared__ int s_info;
    __shared__ double s_logdet;

    const int64_t f = blockIdx.x;
    double* X = Lfac + f * (int64_t)T * T;              // in: lower triangle of the matrix to factorise (cholinv_prepare_kernel)
    double* Y = Wout + f * (int64_t)T * T;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int lr = lane >> 2, lk = lane & 3;
    if (tid == 0) { s_info = 0; s_logdet = 0.0; }
    (void)red;
    __syncthreads();

    for (int k0 = 0; k0 < T; k0 += CI_NB) {
        const int nb = min(CI_NB, T - k0), k1 = k0 + nb;
        // ---- 1. stage L[k0:k1, 0:k0] (zero rows below a short last block) ----
        //      one column per thread: the 16 loads of a column are in flight together (a load -> store loop would wait out
        //      one L2 latency per element)
        for (int c = tid; c < k0; c += CI_THREADS) {
            double v[CI_NB];
#pragma unroll
            for (int i = 0; i < CI_NB; ++i) v[i] = (i < nb) ? __ldcg(X + (int64_t)(k0 + i) * T + c) : 0.0;
#pragma unroll
            for (int i = 0; i < CI_NB; ++i) Bs[i * LDB + c] = v[i];
        }
        __syncthreads();
        // ---- 2. U = sym(A)[k0:, k0:k1] - L[k0:, 0:k0] Bs^T : one 8-row tile per warp trip, two 8-column tiles ----
        const int nrt = (T - k0 + 7) >> 3;
        for (int rt = warp; rt < nrt; rt += CI_THREADS / 32) {
            const int r = k0 + 8 * rt + lr;
            const double* xr = X + (int64_t)min(r, T - 1) * T;
            double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            // the panel's own entries (prepared lower triangle, still untouched at this panel): fetched before the k loop
            double a_rc[2][2];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int c = k0 + 8 * j + 2 * lk + e;
                    a_rc[j][e] = (r < T && c < k1 && c <= r) ? __ldcg(X + (int64_t)r * T + c) : 0.0;
                }
            for (int kk0 = 0; kk0 < k0; kk0 += 64) {              // 16 loads in flight per lane, then their 32 DMMAs
                double av[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int kk = kk0 + 4 * u + lk;
                    av[u] = (r < T && kk < k0) ? __ldcg(xr + kk) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int kk = kk0 + 4 * u;
                    if (kk < k0) {
                        dmma884(acc[0][0], acc[0][1], av[u], Bs[lr * LDB + kk + lk]);
                        dmma884(acc[1][0], acc[1][1], av[u], Bs[(8 + lr) * LDB + kk + lk]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int cl = 8 * j + 2 * lk + e, c = k0 + cl;
                    double v = 0.0;
                    if (r < T && c < k1 && c <= r) v = a_rc[j][e] - acc[j][e];
                    Ps[(8 * rt + lr) * 20 + cl] = v;
                }
        }
        __syncthreads();
        // ---- 3. warp 0: factor the diagonal block out of registers and invert it (as hgp_smem_la.cuh, sl_cholinv) ----
        if (warp == 0) {
            const int li = lane & 15;
            double row[CI_NB];
#pragma unroll
            for (int cc = 0; cc < CI_NB; ++cc) row[cc] = (li < nb && cc <= li) ? Ps[li * 20 + cc] : (li == cc ? 1.0 : 0.0);
            double dinv = 1.0;
            int bad = 0;
#pragma unroll
            for (int j = 0; j < CI_NB; ++j) {
                const double d = __shfl_sync(0xffffffffu, row[j], j);
                if (!(d > 0.0) && bad == 0) bad = k0 + j + 1;
                const double sq = sqrt(d);                       // the same operations as the two-kernel path: l = a / sqrt(d)
                double l = row[j] / sq;
                if (li == j) { l = sq; dinv = 1.0 / sq; }
                if (li >= j) row[j] = l;
#pragma unroll
                for (int cc = j + 1; cc < CI_NB; ++cc) {
                    const double lc = __shfl_sync(0xffffffffu, l, cc);
                    if (li >= cc) row[cc] -= l * lc;
                }
            }
            double dsel = 1.0;
#pragma unroll
            for (int cc = 0; cc < CI_NB; ++cc) if (cc == li) dsel = row[cc];
            double ldg = (li < nb && lane < 16) ? log(dsel) : 0.0;
            ldg = warp_sum(ldg);
            if (lane == 0) { s_logdet += 2.0 * ldg; if (bad && s_info == 0) s_info = bad; }
            if (lane < 16) {
#pragma unroll
                for (int cc = 0; cc < CI_NB; ++cc) Dk[li * (CI_NB + 1) + cc] = row[cc];
                Dk[li * (CI_NB + 1) + CI_NB] = dinv;
                if (li < nb) {
#pragma unroll
                    for (int cc = 0; cc < CI_NB; ++cc) if (cc <= li) X[(int64_t)(k0 + li) * T + k0 + cc] = row[cc];
                }
            }
            __syncwarp();
            if (lane < CI_NB) {
                double x[CI_NB];
#pragma unroll
                for (int i = 0; i < CI_NB; ++i) {
                    double v = (i == lane) ? 1.0 : 0.0;
#pragma unroll
                    for (int p = 0; p < CI_NB; ++p) if (p < i) v -= Dk[i * (CI_NB + 1) + p] * ((p >= lane) ? x[p] : 0.0);
                    x[i] = (i >= lane) ? v * Dk[i * (CI_NB + 1) + CI_NB] : 0.0;
                }
#pragma unroll
                for (int i = 0; i < CI_NB; ++i) {
                    Wi[i * (CI_NB + 1) + lane] = x[i];
                    if (i < nb && lane < nb && lane <= i) Y[(int64_t)(k0 + i) * T + k0 + lane] = x[i];
                }
            }
        }
        __syncthreads();
        // ---- 4. L[k1:, k0:k1] = U[16:] L_kk^-T, one thread per row ----
        for (int r = k1 + tid; r < T; r += CI_THREADS) {
            double x[CI_NB];
#pragma unroll
            for (int p = 0; p < CI_NB; ++p) x[p] = Ps[(r - k0) * 20 + p];
            double* xo = X + (int64_t)r * T + k0;
#pragma unroll
            for (int cc = 0; cc < CI_NB; ++cc) {
                double v = 0.0;
#pragma unroll
                for (int p = 0; p < CI_NB; ++p) if (p <= cc) v += x[p] * Wi[cc * (CI_NB + 1) + p];
                if (cc < nb) xo[cc] = v;
            }
        }
        // ---- 5. W[k0:k1, 0:k0] = -L_kk^-1 (Bs W[0:k0, 0:k0]): column tiles over the warps, k from the tile's diagonal ----
        const int nct = k0 >> 3;
        for (int ct = warp; ct < nct; ct += CI_THREADS / 32) {
            const int cs = 8 * ct;
            double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            for (int kk0 = cs; kk0 < k0; kk0 += 64) {
                double bv[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int kk = kk0 + 4 * u + lk;
                    bv[u] = (kk < k0) ? __ldcg(Y + (int64_t)kk * T + cs + lr) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const int kk = kk0 + 4 * u;
                    if (kk < k0) {
                        dmma884(acc[0][0], acc[0][1], Bs[lr * LDB + kk + lk], bv[u]);
                        dmma884(acc[1][0], acc[1][1], Bs[(8 + lr) * LDB + kk + lk], bv[u]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                Fs[(8 * i + lr) * LDB + cs + 2 * lk] = acc[i][0];
                Fs[(8 * i + lr) * LDB + cs + 2 * lk + 1] = acc[i][1];
            }
        }
        __syncthreads();
        for (int c = tid; c < k0; c += CI_THREADS) {
            double x[CI_NB];
#pragma unroll
            for (int p = 0; p < CI_NB; ++p) x[p] = Fs[p * LDB + c];
#pragma unroll
            for (int i = 0; i < CI_NB; ++i) {
                double v = 0.0;
#pragma unroll
                for (int p = 0; p < CI_NB; ++p) if (p <= i) v -= Wi[i * (CI_NB + 1) + p] * x[p];
                if (i < nb) Y[(int64_t)(k0 + i) * T + c] = v;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        if (logdet) logdet[f] = s_logdet;
        info[f] = s_info;
    }
}

// Packed factor stream for hgp_score_tiles.  Tp = T rounded up to 8, nrb = Tp/8 row blocks.
// chunk kc (k columns [8kc, 8kc+8)) holds row blocks rb = kc..nrb-1; block (kc, rb) is 32 lanes x 2
// doubles: element (lane, ks) = W[8 rb + lane/4][8 kc + 4 ks + lane%4], i.e. the A fragments of two
// consecutive DMMA.8x8x4 k-steps, so one LDS.128 per lane feeds two tensor-core instructions.
__global__ void pack_factors_kernel(const double* __restrict__ W, int T, int nrb, int64_t packed_doubles,
                                    double* __restrict__ out) {
    const int64_t f = blockIdx.y;
    const double* Wf = W + f * (int64_t)T * T;
    double* o = out + f * packed_doubles;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < packed_doubles;
         idx += (int64_t)gridDim.x * blockDim.x) {
        // locate chunk: offset(kc) = 64 * (kc*nrb - kc(kc-1)/2)
        int64_t blk = idx / 64;                 // global (kc, rb) block index
        int within = (int)(idx % 64);
        int kc = 0;
        // small linear search (nrb <= 32)
        int64_t base = 0;
        while (blk >= base + (nrb - kc)) { base += nrb - kc; ++kc; }
        int rb = kc + (int)(blk - base);
        int lane = within >> 1, ks = within & 1;
        int row = 8 * rb + (lane >> 2);
        int col = 8 * kc + 4 * ks + (lane & 3);
        double v = 0.0;
        if (row < T && col < T && col <= row) v = Wf[(int64_t)row * T + col];
        o[idx] = v;
    }
}

}  // namespace

extern "C" int hgp_pack_leads(const double* Y_ntl, int64_t N, int T, int L, double* Y_lnt, void* stream) {
    HGP_REQUIRE(N >= 0 && T > 0 && L > 0, "hgp_pack_leads: bad sizes");
    if (N == 0) return 0;
    int64_t total = N * (int64_t)T * L;
    int blocks = (int)hgp_min64((total + 255) / 256, 148 * 16);
    pack_leads_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(Y_ntl, N, T, L, Y_lnt, N * (int64_t)T);
    HGP_LAUNCH_CHECK("hgp_pack_leads");
    return 0;
}

extern "C" int hgp_pack_leads_slice(const double* Y_ntl, int64_t n, int T, int L, double* Y_lnt_at_slice,
                                    int64_t plane_stride, void* stream) {
    HGP_REQUIRE(n >= 0 && T > 0 && L > 0 && plane_stride >= n * (int64_t)T, "hgp_pack_leads_slice: bad sizes");
    if (n == 0) return 0;
    int64_t total = n * (int64_t)T * L;
    int blocks = (int)hgp_min64((total + 255) / 256, 148 * 16);
    pack_leads_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(Y_ntl, n, T, L, Y_lnt_at_slice, plane_stride);
    HGP_LAUNCH_CHECK("hgp_pack_leads_slice");
    return 0;
}

int hgp_internal_chol(const double* Sigma, const int* src_idx, const double* src_scale, int64_t F, int T,
                      const double* add_diag, double jitter_scale, double* Lfac, double* logdet, int* info,
                      void* stream) {
    HGP_REQUIRE(F >= 0 && T > 0 && T <= 1024, "hgp_chol_batched: need 0 < T <= 1024");
    if (F == 0) return 0;
    size_t smem = sizeof(double) * (NB * (NB + 1) + (size_t)T * (NB + 1));
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return hgp_status(e, "hgp_chol_batched: smem attribute");
    }
    chol_kernel<<<(unsigned)F, CHOL_THREADS, smem, (cudaStream_t)stream>>>(Sigma, src_idx, src_scale, T, add_diag,
                                                                          jitter_scale, Lfac, logdet, info);
    HGP_LAUNCH_CHECK("hgp_chol_batched");
    return 0;
}

extern "C" int hgp_chol_batched(const double* Sigma, int64_t F, int T, const double* add_diag, double jitter_scale,
                                double* Lfac, double* logdet, int* info, void* stream) {
    return hgp_internal_chol(Sigma, nullptr, nullptr, F, T, add_diag, jitter_scale, Lfac, logdet, info, stream);
}

extern "C" int hgp_tri_inverse_batched(const double* Lfac, int64_t F, int T, double* W, void* stream) {
    HGP_REQUIRE(F >= 0 && T > 0 && T <= 1024, "hgp_tri_inverse_batched: need 0 < T <= 1024");
    if (F == 0) return 0;
    int nblk = (T + NB - 1) / NB;
    size_t smem = sizeof(double) * (size_t)(2 + nblk) * NB * (NB + 1);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(tri_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return hgp_status(e, "hgp_tri_inverse_batched: smem attribute");
    }
    const int groups = nblk >= 16 ? 8 : (nblk >= 8 ? 4 : (nblk >= 4 ? 2 : 1));
    tri_inverse_kernel<<<dim3((unsigned)F, groups), 256, smem, (cudaStream_t)stream>>>(Lfac, T, W);
    HGP_LAUNCH_CHECK("hgp_tri_inverse_batched");
    return 0;
}

extern "C" int hgp_cholinv_batched(const double* Sigma, int64_t F, int T, const double* add_diag, double jitter_scale,
                                   double* Lfac, double* W, double* logdet, int* info, void* stream) {
    HGP_REQUIRE(F >= 0 && T > 0 && T <= 1024, "hgp_cholinv_batched: need 0 < T <= 1024");
    if (F == 0) return 0;
    // the fused sweep pays from T = 128 on (tools/table_bench.py: 2.7x at T = 256, 3.4x at T = 400; at T = 90 its six panels
    // are slower than the two small kernels) and holds its panels in shared memory up to T = 512
    if (((T < 128 || T > 512) && !getenv("HGP_CHOLINV_FUSED")) || getenv("HGP_CHOLINV_TWO_KERNELS")) {
        int rc = hgp_chol_batched(Sigma, F, T, add_diag, jitter_scale, Lfac, logdet, info, stream);
        if (rc) return rc;
        return hgp_tri_inverse_batched(Lfac, F, T, W, stream);
    }
    HGP_REQUIRE(T <= 512, "hgp_cholinv_batched: the fused kernel holds T <= 512");
    const size_t smem = ci_smem_bytes(T);
    cudaError_t e = cudaFuncSetAttribute(cholinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return hgp_status(e, "hgp_cholinv_batched: smem attribute");
    const int slabs = F >= 512 ? 1 : (F >= 64 ? 4 : 8);           // enough CTAs to stream the inputs at the HBM rate
    cholinv_prepare_kernel<<<dim3((unsigned)F, slabs), 256, 0, (cudaStream_t)stream>>>(Sigma, T, add_diag, jitter_scale, Lfac, W);
    HGP_LAUNCH_CHECK("hgp_cholinv_batched: prepare");
    cholinv_kernel<<<(unsigned)F, CI_THREADS, smem, (cudaStream_t)stream>>>(T, Lfac, W, logdet, info);
    HGP_LAUNCH_CHECK("hgp_cholinv_batched");
    return 0;
}

extern "C" int64_t hgp_packed_factor_bytes(int T) {
    int64_t nrb = (T + 7) / 8;
    return 512 * (nrb * (nrb + 1) / 2);
}

extern "C" int hgp_pack_factors(const double* W, int64_t F, int T, double* Wpacked, void* stream) {
    HGP_REQUIRE(F >= 0 && T > 0 && T <= 256, "hgp_pack_factors: need 0 < T <= 256");
    if (F == 0) return 0;
    int nrb = (T + 7) / 8;
    int64_t pd = hgp_packed_factor_bytes(T) / 8;
    dim3 grid((unsigned)hgp_min64((pd + 255) / 256, 64), (unsigned)F);
    pack_factors_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(W, T, nrb, pd, Wpacked);
    HGP_LAUNCH_CHECK("hgp_pack_factors");
    return 0;
}

// ---- kernel matrix of ConstantKernel(c) * RBF(l) (+ WhiteKernel on the diagonal when asked) -----------------------
// sklearn evaluates exp(-0.5 * sqeuclidean(x / l, y / l)) (reference call sites GPI_model.py:50,228; GPI.py:124-139);
// the division happens before the difference, and so it does here.
namespace {
__global__ void rbf_kernel_matrix_kernel(const double* __restrict__ xa, int na, const double* __restrict__ xb, int nb,
                                         double c, double ell, double diag_add, double* __restrict__ K) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= na * nb) return;
    const int i = idx / nb, j = idx % nb;
    const double d = xa[i] / ell - xb[j] / ell;
    K[idx] = c * exp(-0.5 * (d * d)) + ((i == j) ? diag_add : 0.0);
}
}  // namespace

extern "C" int hgp_rbf_kernel_matrix(const double* xa, int na, const double* xb, int nb, double kernel_const,
                                     double kernel_length, double diag_add, double* K, void* stream) {
    HGP_REQUIRE(na > 0 && nb > 0 && (int64_t)na * nb < (1ll << 31), "hgp_rbf_kernel_matrix: bad sizes");
    const int n = na * nb;
    rbf_kernel_matrix_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(xa, na, xb, nb, kernel_const, kernel_length,
                                                                              diag_add, K);
    HGP_LAUNCH_CHECK("hgp_rbf_kernel_matrix");
    return 0;
}

// ---- mean beat of a lead plane: torch.mean(y_trains[:, :, ld], dim=0) in GPI_HDP.compute_snr_ini (GPI_HDP.py:715-730) ----
// Two fixed-order stages (per-block partial sums over a contiguous slice of beats, then the blocks in order), so the
// result does not depend on scheduling.
namespace {
constexpr int MEAN_BLOCKS = 256;
__global__ void mean_beat_partial_kernel(const double* __restrict__ Y, int64_t N, int T, double* __restrict__ partial) {
    const int64_t per = (N + gridDim.x - 1) / gridDim.x;
    const int64_t n0 = blockIdx.x * per, n1 = hgp_min64(N, n0 + per);
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
        double acc = 0.0;
        for (int64_t n = n0; n < n1; ++n) acc += Y[n * T + t];
        partial[(int64_t)blockIdx.x * T + t] = acc;
    }
}
__global__ void mean_beat_finish_kernel(const double* __restrict__ partial, int nblocks, int64_t N, int T,
                                        double* __restrict__ mean) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double acc = 0.0;
    for (int b = 0; b < nblocks; ++b) acc += partial[(int64_t)b * T + t];
    mean[t] = acc / (double)N;
}
}  // namespace

extern "C" int64_t hgp_mean_beat_work_doubles(int T) { return (int64_t)MEAN_BLOCKS * T; }

extern "C" int hgp_mean_beat(const double* Y, int64_t N, int T, double* mean, double* work, void* stream) {
    HGP_REQUIRE(N > 0 && T > 0, "hgp_mean_beat: need N > 0, T > 0");
    const int nblocks = (int)hgp_min64(MEAN_BLOCKS, N);
    mean_beat_partial_kernel<<<nblocks, 256, 0, (cudaStream_t)stream>>>(Y, N, T, work);
    HGP_LAUNCH_CHECK("hgp_mean_beat: partial");
    mean_beat_finish_kernel<<<(T + 255) / 256, 256, 0, (cudaStream_t)stream>>>(work, nblocks, N, T, mean);
    HGP_LAUNCH_CHECK("hgp_mean_beat: finish");
    return 0;
}
