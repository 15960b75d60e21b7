// CTA-level dense linear algebra on T x T float64 matrices that live in global memory (L2-resident working
// set).  One CTA (256 threads) executes each routine cooperatively; every routine ends with __syncthreads(), so
// results are visible to the whole CTA.  Used by the persistent chain kernel (hgp_chain.cu), where one CTA walks
// one (cluster, lead) chain through its Kalman / smoother / MNIW steps without leaving the device.
// GEMMs run on the FP64 tensor cores (DMMA.8x8x4); factorizations are blocked with the panel in shared memory.
#pragma once
#include "hgp_common.cuh"

namespace hgp {

constexpr int LA_THREADS = 256;
constexpr int LA_NB = 16;

struct LaSmem {
    double As[64 * 20];            // GEMM A tile / panel scratch
    double Bs[16 * 72];            // GEMM B tile
    double Dk[LA_NB * (LA_NB + 1)];   // diagonal block
    double red[LA_THREADS / 32 + 8];
    int ipiv[LA_NB];
    int flag;
    double* big;                   // optional dynamic shared memory: 2 x T x T doubles when T <= LA_SMEM_T, else null
};

// Matrices up to this size are factorised / solved inside shared memory (the MIT-BIH shape is T = 90): a column step of
// the pivoted LU then costs shared-memory latencies instead of L2 round trips.
constexpr int LA_SMEM_T = 96;
__host__ __device__ inline size_t la_dynamic_smem_bytes(int T) {
    return T <= LA_SMEM_T ? 2 * (size_t)T * T * sizeof(double) : 0;
}

// ---- elementwise ------------------------------------------------------------------------------
__device__ __forceinline__ void la_copy(double* __restrict__ D, const double* __restrict__ S, int n) {
    for (int i = threadIdx.x; i < n; i += LA_THREADS) D[i] = S[i];
    __syncthreads();
}
__device__ __forceinline__ void la_transpose(double* __restrict__ D, const double* __restrict__ S, int T) {
    for (int i = threadIdx.x; i < T * T; i += LA_THREADS) { int r = i / T, c = i % T; D[(int64_t)c * T + r] = S[i]; }
    __syncthreads();
}
// D = a*X + b*Y (Y may be null)
__device__ __forceinline__ void la_axpby(double* D, double a, const double* X, double b, const double* Y, int n) {
    for (int i = threadIdx.x; i < n; i += LA_THREADS) D[i] = a * X[i] + (Y ? b * Y[i] : 0.0);
    __syncthreads();
}
__device__ __forceinline__ void la_add_diag(double* A, double v, int T) {
    for (int i = threadIdx.x; i < T; i += LA_THREADS) A[(int64_t)i * T + i] += v;
    __syncthreads();
}
__device__ __forceinline__ void la_set_identity(double* A, double v, int T) {
    for (int i = threadIdx.x; i < T * T; i += LA_THREADS) A[i] = (i / T == i % T) ? v : 0.0;
    __syncthreads();
}
// A = 0.5 (A + A^T) + add * I   (in place; element pairs handled by one thread)
__device__ __forceinline__ void la_symmetrize(double* A, double add, int T) {
    for (int i = threadIdx.x; i < T * T; i += LA_THREADS) {
        int r = i / T, c = i % T;
        if (c < r) {
            double v = 0.5 * (A[i] + A[(int64_t)c * T + r]);
            A[i] = v; A[(int64_t)c * T + r] = v;
        } else if (c == r) A[i] += add;
    }
    __syncthreads();
}
__device__ __forceinline__ double la_mean_abs_diag(const double* A, int T, LaSmem& sm) {
    double p = 0.0;
    for (int i = threadIdx.x; i < T; i += LA_THREADS) p += fabs(A[(int64_t)i * T + i]);
    p = warp_sum(p);
    if ((threadIdx.x & 31) == 0) sm.red[threadIdx.x >> 5] = p;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < LA_THREADS / 32; ++w) tot += sm.red[w];
    __syncthreads();
    return tot / T;
}
// y = op(A) x (+ y0), vectors in global memory; one warp per output row
__device__ __forceinline__ void la_gemv(double* y, const double* A, const double* x, int T, double beta, const double* y0,
                                        double alpha = 1.0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < T; r += LA_THREADS / 32) {
        double acc = 0.0;
        for (int k = lane; k < T; k += 32) acc += A[(int64_t)r * T + k] * x[k];
        acc = warp_sum(acc);
        if (lane == 0) y[r] = alpha * acc + (y0 ? beta * y0[r] : 0.0);
    }
    __syncthreads();
}
// A += s * u v^T
__device__ __forceinline__ void la_rank1(double* A, double s, const double* u, const double* v, int T) {
    for (int i = threadIdx.x; i < T * T; i += LA_THREADS) A[i] += s * u[i / T] * v[i % T];
    __syncthreads();
}

// ---- GEMM on the tensor cores -------------------------------------------------------------------
// C = alpha * op(A) op(B) + beta * D   (D may alias C or be null); opA/opB: 0 = as is, 1 = transposed.
// C must not alias A or B.
static __device__ __noinline__ void la_gemm(double* C, const double* A, int tA, const double* B,
                               int tB, int T, double alpha, double beta, const double* D, LaSmem& sm) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;
    const int nt = (T + 63) / 64;
    // The next K chunk is fetched from global memory (L2) into registers while the tensor cores work on the current
    // one: a single CTA has nothing else to hide the ~700-cycle load latency behind.
    double pa[4], pb[4];
    auto fetch = [&](int r0, int c0, int k0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = tid + u * LA_THREADS;
            int r, k;
            if (tA) { k = idx / 64; r = idx % 64; } else { r = idx / 16; k = idx % 16; }
            const int gr = r0 + r, gk = k0 + k;
            pa[u] = (gr < T && gk < T) ? (tA ? A[(int64_t)gk * T + gr] : A[(int64_t)gr * T + gk]) : 0.0;
            int kb, c;
            if (tB) { c = idx / 16; kb = idx % 16; } else { kb = idx / 64; c = idx % 64; }
            const int gkb = k0 + kb, gc = c0 + c;
            pb[u] = (gkb < T && gc < T) ? (tB ? B[(int64_t)gc * T + gkb] : B[(int64_t)gkb * T + gc]) : 0.0;
        }
    };
    auto stash = [&]() {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = tid + u * LA_THREADS;
            int r, k;
            if (tA) { k = idx / 64; r = idx % 64; } else { r = idx / 16; k = idx % 16; }
            sm.As[r * 20 + k] = pa[u];
            int kb, c;
            if (tB) { c = idx / 16; kb = idx % 16; } else { kb = idx / 64; c = idx % 64; }
            sm.Bs[kb * 72 + c] = pb[u];
        }
    };
    for (int tile = 0; tile < nt * nt; ++tile) {
        const int r0 = (tile / nt) * 64, c0 = (tile % nt) * 64;
        double acc[2][4][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        fetch(r0, c0, 0);
        for (int k0 = 0; k0 < T; k0 += 16) {
            stash();
            __syncthreads();
            if (k0 + 16 < T) fetch(r0, c0, k0 + 16);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                double a[2], bf[4];
#pragma unroll
                for (int i = 0; i < 2; ++i) a[i] = sm.As[(16 * wm + 8 * i + (lane >> 2)) * 20 + 4 * ks + (lane & 3)];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = sm.Bs[(4 * ks + (lane & 3)) * 72 + 32 * wn + 8 * j + (lane >> 2)];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bf[j]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int gr = r0 + 16 * wm + 8 * i + (lane >> 2), gc = c0 + 32 * wn + 8 * j + 2 * (lane & 3) + e;
                    if (gr < T && gc < T) {
                        const int64_t o = (int64_t)gr * T + gc;
                        C[o] = alpha * acc[i][j][e] + (D ? beta * D[o] : 0.0);
                    }
                }
    }
    __syncthreads();
}

// ---- Cholesky (lower, in place; strict upper part zeroed) ---------------------------------------------
// returns 0 or (index + 1) of the first non-positive pivot
static __device__ __noinline__ int la_chol(double* A, int T, LaSmem& sm) {
    const int tid = threadIdx.x;
    if (tid == 0) sm.flag = 0;
    for (int i = tid; i < T * T; i += LA_THREADS) if (i % T > i / T) A[i] = 0.0;
    __syncthreads();
    double* P = sm.As;   // panel rows are streamed through shared memory in slabs of 64
    for (int k0 = 0; k0 < T; k0 += LA_NB) {
        const int nb = min(LA_NB, T - k0);
        for (int idx = tid; idx < nb * nb; idx += LA_THREADS) sm.Dk[(idx / nb) * (LA_NB + 1) + idx % nb] = A[(int64_t)(k0 + idx / nb) * T + k0 + idx % nb];
        __syncthreads();
        if (tid < 32) {
            for (int j = 0; j < nb; ++j) {
                double d = sm.Dk[j * (LA_NB + 1) + j];
                if (!(d > 0.0) && tid == 0 && sm.flag == 0) sm.flag = k0 + j + 1;
                double s = sqrt(d);
                __syncwarp();
                if (tid == 0) sm.Dk[j * (LA_NB + 1) + j] = s;
                if (tid > j && tid < nb) sm.Dk[tid * (LA_NB + 1) + j] /= s;
                __syncwarp();
                if (tid > j && tid < nb) {
                    double lij = sm.Dk[tid * (LA_NB + 1) + j];
                    for (int c = j + 1; c <= tid; ++c) sm.Dk[tid * (LA_NB + 1) + c] -= lij * sm.Dk[c * (LA_NB + 1) + j];
                }
                __syncwarp();
            }
        }
        __syncthreads();
        for (int idx = tid; idx < nb * nb; idx += LA_THREADS) {
            int i = idx / nb, j = idx % nb;
            if (j <= i) A[(int64_t)(k0 + i) * T + k0 + j] = sm.Dk[i * (LA_NB + 1) + j];
        }
        const int r0 = k0 + nb, ntr = T - r0;
        for (int r = tid; r < ntr; r += LA_THREADS) {
            double x[LA_NB];
            double* arow = A + (int64_t)(r0 + r) * T + k0;
#pragma unroll
            for (int c = 0; c < LA_NB; ++c) {
                if (c < nb) {
                    double v = arow[c];
                    for (int p = 0; p < c; ++p) v -= x[p] * sm.Dk[c * (LA_NB + 1) + p];
                    x[c] = v / sm.Dk[c * (LA_NB + 1) + c];
                    arow[c] = x[c];
                }
            }
        }
        __syncthreads();
        // trailing update, lower part: A[i][j] -= L[i, k0:k0+nb] . L[j, k0:k0+nb], slabs of 64 rows of L in shared memory
        const int ty = tid >> 4, tx = tid & 15;
        for (int ib = 0; ib < ntr; ib += 16) {
            const int i = ib + ty;
            double li[LA_NB];
#pragma unroll
            for (int c = 0; c < LA_NB; ++c) li[c] = (i < ntr && c < nb) ? A[(int64_t)(r0 + i) * T + k0 + c] : 0.0;
            for (int jb = 0; jb <= ib; jb += 16) {
                // rows jb..jb+15 of the panel -> shared
                __syncthreads();
                {
                    const int jr = tid >> 4, c = tid & 15;
                    P[jr * (LA_NB + 1) + c] = (jb + jr < ntr && c < nb) ? A[(int64_t)(r0 + jb + jr) * T + k0 + c] : 0.0;
                }
                __syncthreads();
                const int j = jb + tx;
                if (i < ntr && j <= i) {
                    double acc = 0.0;
#pragma unroll
                    for (int c = 0; c < LA_NB; ++c) acc += li[c] * P[tx * (LA_NB + 1) + c];
                    A[(int64_t)(r0 + i) * T + r0 + j] -= acc;
                }
            }
        }
        __syncthreads();
    }
    return sm.flag;
}

// ---- triangular solves with T right-hand sides (in place on B) --------------------------------------
// Row-block update on the tensor cores: for i in [i0, i1) (at most 64 rows) and every column c,
//   C[i][c] += alpha * sum_{k in [k0, k1)} opA(i, k) * B[k][c],   opA(i, k) = transA ? A[k][i] : A[i][k].
// C and B may be the same matrix as long as the row ranges [i0, i1) and [k0, k1) do not overlap.
// Only columns c >= c_begin are touched.
static __device__ __noinline__ void la_rows_update(double* C, const double* A, int transA, const double* B, int T, int i0,
                                                   int i1, int k0, int k1, double alpha, LaSmem& sm, int c_begin = 0) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;
    const int nct = (T + 63) / 64;
    for (int ct = c_begin / 64; ct < nct; ++ct) {
        const int c0 = ct * 64;
        double acc[2][4][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int kk = k0; kk < k1; kk += 16) {
            for (int idx = tid; idx < 64 * 16; idx += LA_THREADS) {
                int r, k;
                if (transA) { k = idx / 64; r = idx % 64; } else { r = idx / 16; k = idx % 16; }
                const int gr = i0 + r, gk = kk + k;
                double v = 0.0;
                if (gr < i1 && gk < k1) v = transA ? A[(int64_t)gk * T + gr] : A[(int64_t)gr * T + gk];
                sm.As[r * 20 + k] = v;
            }
            for (int idx = tid; idx < 16 * 64; idx += LA_THREADS) {
                const int k = idx / 64, c = idx % 64;
                const int gk = kk + k, gc = c0 + c;
                sm.Bs[k * 72 + c] = (gk < k1 && gc < T) ? B[(int64_t)gk * T + gc] : 0.0;
            }
            __syncthreads();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                double a[2], bf[4];
#pragma unroll
                for (int i = 0; i < 2; ++i) a[i] = sm.As[(16 * wm + 8 * i + (lane >> 2)) * 20 + 4 * ks + (lane & 3)];
#pragma unroll
                for (int j = 0; j < 4; ++j) bf[j] = sm.Bs[(4 * ks + (lane & 3)) * 72 + 32 * wn + 8 * j + (lane >> 2)];
#pragma unroll
                for (int i = 0; i < 2; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bf[j]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int gr = i0 + 16 * wm + 8 * i + (lane >> 2), gc = c0 + 32 * wn + 8 * j + 2 * (lane & 3) + e;
                    if (gr < i1 && gc < T && gc >= c_begin) C[(int64_t)gr * T + gc] += alpha * acc[i][j][e];
                }
    }
    __syncthreads();
}

// Blocked substitution, two levels: 64-row outer blocks whose coupling to the already-solved rows is one tensor-core
// row-block update, 16-row inner blocks solved by one thread per column of B with the diagonal block in shared memory.
//   mode 0: B <- L^{-1} B   (L lower, forward)           mode 1: B <- L^{-T} B  (L lower, backward)
//   mode 2: B <- L1^{-1} B  (unit lower part of an LU)   mode 3: B <- U^{-1} B  (upper part of an LU, backward)
static __device__ __noinline__ void la_trsm_blocked(const double* __restrict__ M, double* B, int T, int mode, LaSmem& sm) {
    const int tid = threadIdx.x;
    const bool backward = (mode == 1 || mode == 3), trans = (mode == 1), unit = (mode == 2);
    const int nouter = (T + 63) / 64;
    for (int ob = 0; ob < nouter; ++ob) {
        const int R0 = (backward ? nouter - 1 - ob : ob) * 64, R1 = min(T, R0 + 64);
        if (!backward && R0 > 0) la_rows_update(B, M, 0, B, T, R0, R1, 0, R0, -1.0, sm);
        if (backward && R1 < T) la_rows_update(B, M, trans, B, T, R0, R1, R1, T, -1.0, sm);
        const int ninner = (R1 - R0 + LA_NB - 1) / LA_NB;
        for (int ib = 0; ib < ninner; ++ib) {
            const int r0 = R0 + (backward ? ninner - 1 - ib : ib) * LA_NB, nb = min(LA_NB, R1 - r0), r1 = r0 + nb;
            // coupling to the rows of this outer block that are already solved
            const int ka = backward ? r1 : R0, kb = backward ? R1 : r0;
            if (kb > ka) {
                for (int idx = tid; idx < nb * T; idx += LA_THREADS) {
                    const int i = r0 + idx / T, c = idx % T;
                    double acc = 0.0;
                    for (int k = ka; k < kb; ++k)
                        acc += (trans ? M[(int64_t)k * T + i] : M[(int64_t)i * T + k]) * B[(int64_t)k * T + c];
                    B[(int64_t)i * T + c] -= acc;
                }
            }
            // diagonal block, as coef(i, p) = Dk[i][p]
            for (int idx = tid; idx < nb * nb; idx += LA_THREADS) {
                const int i = idx / nb, p = idx % nb;
                sm.Dk[i * (LA_NB + 1) + p] = trans ? M[(int64_t)(r0 + p) * T + r0 + i] : M[(int64_t)(r0 + i) * T + r0 + p];
            }
            __syncthreads();
            for (int c = tid; c < T; c += LA_THREADS) {
                double x[LA_NB];
                if (!backward) {
#pragma unroll
                    for (int i = 0; i < LA_NB; ++i) {
                        if (i < nb) {
                            double v = B[(int64_t)(r0 + i) * T + c];
                            for (int p = 0; p < i; ++p) v -= sm.Dk[i * (LA_NB + 1) + p] * x[p];
                            x[i] = unit ? v : v / sm.Dk[i * (LA_NB + 1) + i];
                            B[(int64_t)(r0 + i) * T + c] = x[i];
                        }
                    }
                } else {
#pragma unroll
                    for (int ii = 0; ii < LA_NB; ++ii) {
                        const int i = nb - 1 - ii;
                        if (i >= 0) {
                            double v = B[(int64_t)(r0 + i) * T + c];
                            for (int p = i + 1; p < nb; ++p) v -= sm.Dk[i * (LA_NB + 1) + p] * x[p];
                            x[i] = v / sm.Dk[i * (LA_NB + 1) + i];
                            B[(int64_t)(r0 + i) * T + c] = x[i];
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
}

// B <- L^{-1} B  (forward substitution), L lower triangular
static __device__ __forceinline__ void la_trsm_lower(const double* __restrict__ L, double* B, int T, LaSmem& sm) {
    la_trsm_blocked(L, B, T, 0, sm);
}
// B <- L^{-T} B  (backward substitution with the transpose of a lower-triangular L)
static __device__ __forceinline__ void la_trsm_lower_trans(const double* __restrict__ L, double* B, int T, LaSmem& sm) {
    la_trsm_blocked(L, B, T, 1, sm);
}

// Unblocked LU for small matrices (T < LA_LU_BLOCKED_MIN: the MIT-BIH shape T = 90), where the barrier count of
// the column loop is lower than the fixed costs of the blocked algorithm below.
constexpr int LA_LU_BLOCKED_MIN = 160;
static __device__ __noinline__ void la_lu_factor_small(double* A, int* piv, int T, LaSmem& sm) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int k = 0; k < T; ++k) {
        // pivot search in column k, rows k..T-1 (first maximum of |a|, like LAPACK idamax)
        double best = -1.0; int bi = k;
        for (int i = k + tid; i < T; i += LA_THREADS) {
            double v = fabs(A[(int64_t)i * T + k]);
            if (v > best) { best = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, best, o);
            int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        __syncthreads();   // previous column's readers of sm.red / sm.ipiv are done
        if (lane == 0) { sm.red[warp] = best; sm.ipiv[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            double b = sm.red[0]; int p = sm.ipiv[0];
            for (int w = 1; w < LA_THREADS / 32; ++w) {
                const double ov = sm.red[w]; const int oi = sm.ipiv[w];
                if (ov > b || (ov == b && oi < p)) { b = ov; p = oi; }
            }
            piv[k] = p;
            sm.flag = p;
        }
        __syncthreads();
        const int p = sm.flag;
        if (p != k) {
            for (int c = tid; c < T; c += LA_THREADS) {
                double t = A[(int64_t)k * T + c]; A[(int64_t)k * T + c] = A[(int64_t)p * T + c]; A[(int64_t)p * T + c] = t;
            }
        }
        __syncthreads();
        const double inv = 1.0 / A[(int64_t)k * T + k];
        // multipliers
        for (int i = k + 1 + tid; i < T; i += LA_THREADS) A[(int64_t)i * T + k] *= inv;
        __syncthreads();
        // trailing update: A[i][j] -= A[i][k] * A[k][j]
        for (int i = k + 1 + (tid >> 5); i < T; i += LA_THREADS / 32) {
            const double lik = A[(int64_t)i * T + k];
            for (int j = k + 1 + (tid & 31); j < T; j += 32) A[(int64_t)i * T + j] -= lik * A[(int64_t)k * T + j];
        }
        __syncthreads();
    }
}

// ---- LU with partial pivoting (in place) and solve with T right-hand sides -----------------------------
// piv[k] = row swapped with k at step k (global memory, T ints).  Right-looking, panels of 16 columns:
//   * panel (rows k0.., 16 columns = one 128-byte line per row, L1-resident): factorised by ONE warp without any
//     CTA barrier -- pivot search by shuffle reduction (first maximum of |a|, like LAPACK idamax), swap, scale and the
//     rank-1 updates inside the panel;
//   * the 16 row interchanges are applied to the columns outside the panel, one thread per column;
//   * U12 = L11^{-1} A12 (unit lower 16 x 16 block in shared memory, one thread per column);
//   * A22 -= L21 U12 on the tensor cores (la_rows_update, one call per 64-row block).
static __device__ __noinline__ void la_lu_factor_blocked(double* A, int* piv, int T, LaSmem& sm);
static __device__ __noinline__ void la_lu_factor(double* A, int* piv, int T, LaSmem& sm) {
    if (T <= LA_SMEM_T && sm.big) {
        // small matrix: stage it in shared memory and run the blocked algorithm there (generic pointers): the panel
        // warp and the trailing update then pay shared-memory latencies instead of L2 round trips
        double* S = sm.big;
        for (int i = threadIdx.x; i < T * T; i += LA_THREADS) S[i] = A[i];
        __syncthreads();
        la_lu_factor_blocked(S, piv, T, sm);
        for (int i = threadIdx.x; i < T * T; i += LA_THREADS) A[i] = S[i];
        __syncthreads();
        return;
    }
    if (T < LA_LU_BLOCKED_MIN) { la_lu_factor_small(A, piv, T, sm); return; }
    la_lu_factor_blocked(A, piv, T, sm);
}
static __device__ __noinline__ void la_lu_factor_blocked(double* A, int* piv, int T, LaSmem& sm) {
    const int tid = threadIdx.x, lane = tid & 31;
    for (int k0 = 0; k0 < T; k0 += LA_NB) {
        const int nb = min(LA_NB, T - k0), k1 = k0 + nb;
        if (tid < 32) {
            for (int j = 0; j < nb; ++j) {
                const int col = k0 + j;
                double best = -1.0; int bi = col;
                for (int i = col + lane; i < T; i += 32) {
                    const double v = fabs(A[(int64_t)i * T + col]);
                    if (v > best) { best = v; bi = i; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
                }
                if (lane == 0) { piv[col] = bi; sm.ipiv[j] = bi; }
                if (bi != col && lane < nb) {       // swap inside the panel only; the rest of the rows follows below
                    double* ra = A + (int64_t)col * T + k0 + lane;
                    double* rb = A + (int64_t)bi * T + k0 + lane;
                    const double t = *ra; *ra = *rb; *rb = t;
                }
                __syncwarp();
                const double inv = 1.0 / A[(int64_t)col * T + col];
                for (int i = col + 1 + lane; i < T; i += 32) {
                    double* row = A + (int64_t)i * T;
                    const double l = row[col] * inv;
                    row[col] = l;
                    for (int c = col + 1; c < k1; ++c) row[c] -= l * A[(int64_t)col * T + c];
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // row interchanges outside the panel, then the unit-lower 16 x 16 block for the U12 solve
        for (int c = tid; c < T; c += LA_THREADS) {
            if (c >= k0 && c < k1) continue;
            for (int j = 0; j < nb; ++j) {
                const int p = sm.ipiv[j];
                if (p != k0 + j) {
                    const double t = A[(int64_t)(k0 + j) * T + c];
                    A[(int64_t)(k0 + j) * T + c] = A[(int64_t)p * T + c];
                    A[(int64_t)p * T + c] = t;
                }
            }
        }
        for (int idx = tid; idx < nb * nb; idx += LA_THREADS)
            sm.Dk[(idx / nb) * (LA_NB + 1) + idx % nb] = A[(int64_t)(k0 + idx / nb) * T + k0 + idx % nb];
        __syncthreads();
        if (k1 < T) {
            for (int c = k1 + tid; c < T; c += LA_THREADS) {
                double x[LA_NB];
#pragma unroll
                for (int i = 0; i < LA_NB; ++i) {
                    if (i < nb) {
                        double v = A[(int64_t)(k0 + i) * T + c];
                        for (int p = 0; p < i; ++p) v -= sm.Dk[i * (LA_NB + 1) + p] * x[p];
                        x[i] = v;
                        A[(int64_t)(k0 + i) * T + c] = v;
                    }
                }
            }
            __syncthreads();
            for (int r0 = k1; r0 < T; r0 += 64)
                la_rows_update(A, A, 0, A, T, r0, min(T, r0 + 64), k0, k1, -1.0, sm, k1);
        }
    }
}
// B <- A^{-1} B using the factorization above (B: T x T, in place)
static __device__ __noinline__ void la_lu_solve(const double* __restrict__ LU, const int* __restrict__ piv, double* B, int T, LaSmem& sm) {
    const int tid = threadIdx.x;
    if (T <= LA_SMEM_T && sm.big) {
        double* SL = sm.big;
        double* SB = sm.big + (size_t)T * T;
        for (int i = tid; i < T * T; i += LA_THREADS) { SL[i] = LU[i]; SB[i] = B[i]; }
        __syncthreads();
        for (int c = tid; c < T; c += LA_THREADS)
            for (int k = 0; k < T; ++k) {
                const int p = piv[k];
                if (p != k) { const double t = SB[k * T + c]; SB[k * T + c] = SB[p * T + c]; SB[p * T + c] = t; }
            }
        __syncthreads();
        la_trsm_blocked(SL, SB, T, 2, sm);     // unit lower
        la_trsm_blocked(SL, SB, T, 3, sm);     // upper
        for (int i = tid; i < T * T; i += LA_THREADS) B[i] = SB[i];
        __syncthreads();
        return;
    }
    for (int k = 0; k < T; ++k) {   // apply the row interchanges
        const int p = piv[k];
        if (p != k)
            for (int c = tid; c < T; c += LA_THREADS) {
                double t = B[(int64_t)k * T + c]; B[(int64_t)k * T + c] = B[(int64_t)p * T + c]; B[(int64_t)p * T + c] = t;
            }
        __syncthreads();
    }
    if (T >= LA_LU_BLOCKED_MIN) {
        la_trsm_blocked(LU, B, T, 2, sm);     // unit lower
        la_trsm_blocked(LU, B, T, 3, sm);     // upper
    } else {
    // forward (unit lower) and backward (upper): one thread per column, rows streamed
        for (int c = tid; c < T; c += LA_THREADS) {
            for (int i = 1; i < T; ++i) {
                double v = B[(int64_t)i * T + c];
                const double* lr = LU + (int64_t)i * T;
                for (int k = 0; k < i; ++k) v -= lr[k] * B[(int64_t)k * T + c];
                B[(int64_t)i * T + c] = v;
            }
            for (int i = T - 1; i >= 0; --i) {
                double v = B[(int64_t)i * T + c];
                const double* ur = LU + (int64_t)i * T;
                for (int k = i + 1; k < T; ++k) v -= ur[k] * B[(int64_t)k * T + c];
                B[(int64_t)i * T + c] = v / ur[i];
            }
        }
    }
    __syncthreads();
    (void)sm;
}

}  // namespace hgp
