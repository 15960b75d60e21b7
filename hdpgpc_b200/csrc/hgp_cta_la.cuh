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
};

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
    for (int tile = 0; tile < nt * nt; ++tile) {
        const int r0 = (tile / nt) * 64, c0 = (tile % nt) * 64;
        double acc[2][4][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int k0 = 0; k0 < T; k0 += 16) {
            for (int idx = tid; idx < 64 * 16; idx += LA_THREADS) {
                int r, k;
                if (tA) { k = idx / 64; r = idx % 64; } else { r = idx / 16; k = idx % 16; }
                const int gr = r0 + r, gk = k0 + k;
                double v = 0.0;
                if (gr < T && gk < T) v = tA ? A[(int64_t)gk * T + gr] : A[(int64_t)gr * T + gk];
                sm.As[r * 20 + k] = v;
            }
            for (int idx = tid; idx < 16 * 64; idx += LA_THREADS) {
                int k, c;
                if (tB) { c = idx / 16; k = idx % 16; } else { k = idx / 64; c = idx % 64; }
                const int gk = k0 + k, gc = c0 + c;
                double v = 0.0;
                if (gk < T && gc < T) v = tB ? B[(int64_t)gc * T + gk] : B[(int64_t)gk * T + gc];
                sm.Bs[k * 72 + c] = v;
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
// B <- L^{-1} B  (forward substitution), L lower triangular
static __device__ __noinline__ void la_trsm_lower(const double* __restrict__ L, double* B, int T, LaSmem& sm) {
    const int tid = threadIdx.x;
    for (int r0 = 0; r0 < T; r0 += LA_NB) {
        const int nb = min(LA_NB, T - r0);
        // rows r0..r0+nb: B[i][c] -= sum_{k<r0} L[i][k] B[k][c]
        for (int idx = tid; idx < nb * T; idx += LA_THREADS) {
            const int i = r0 + idx / T, c = idx % T;
            double acc = 0.0;
            const double* lr = L + (int64_t)i * T;
            for (int k = 0; k < r0; ++k) acc += lr[k] * B[(int64_t)k * T + c];
            B[(int64_t)i * T + c] -= acc;
        }
        for (int idx = tid; idx < nb * nb; idx += LA_THREADS) sm.Dk[(idx / nb) * (LA_NB + 1) + idx % nb] = L[(int64_t)(r0 + idx / nb) * T + r0 + idx % nb];
        __syncthreads();
        // diagonal block: one thread per column of B
        for (int c = tid; c < T; c += LA_THREADS) {
            double x[LA_NB];
#pragma unroll
            for (int i = 0; i < LA_NB; ++i) {
                if (i < nb) {
                    double v = B[(int64_t)(r0 + i) * T + c];
                    for (int p = 0; p < i; ++p) v -= sm.Dk[i * (LA_NB + 1) + p] * x[p];
                    x[i] = v / sm.Dk[i * (LA_NB + 1) + i];
                    B[(int64_t)(r0 + i) * T + c] = x[i];
                }
            }
        }
        __syncthreads();
    }
}
// B <- L^{-T} B  (backward substitution with the transpose of a lower-triangular L)
static __device__ __noinline__ void la_trsm_lower_trans(const double* __restrict__ L, double* B, int T, LaSmem& sm) {
    const int tid = threadIdx.x;
    const int nblk = (T + LA_NB - 1) / LA_NB;
    for (int b = nblk - 1; b >= 0; --b) {
        const int r0 = b * LA_NB, nb = min(LA_NB, T - r0), r1 = r0 + nb;
        // rows r0..r1: B[i][c] -= sum_{k>=r1} L[k][i] B[k][c]
        for (int idx = tid; idx < nb * T; idx += LA_THREADS) {
            const int i = r0 + idx / T, c = idx % T;
            double acc = 0.0;
            for (int k = r1; k < T; ++k) acc += L[(int64_t)k * T + i] * B[(int64_t)k * T + c];
            B[(int64_t)i * T + c] -= acc;
        }
        for (int idx = tid; idx < nb * nb; idx += LA_THREADS) sm.Dk[(idx / nb) * (LA_NB + 1) + idx % nb] = L[(int64_t)(r0 + idx / nb) * T + r0 + idx % nb];
        __syncthreads();
        for (int c = tid; c < T; c += LA_THREADS) {
            double x[LA_NB];
#pragma unroll
            for (int ii = 0; ii < LA_NB; ++ii) {
                const int i = nb - 1 - ii;
                if (i >= 0) {
                    double v = B[(int64_t)(r0 + i) * T + c];
                    for (int p = i + 1; p < nb; ++p) v -= sm.Dk[p * (LA_NB + 1) + i] * x[p];
                    x[i] = v / sm.Dk[i * (LA_NB + 1) + i];
                    B[(int64_t)(r0 + i) * T + c] = x[i];
                }
            }
        }
        __syncthreads();
    }
}

// ---- LU with partial pivoting (in place) and solve with T right-hand sides -----------------------------
// piv[k] = row swapped with k at step k (global memory, T ints).  Unblocked column loop with the trailing
// update spread over the CTA: 2/3 T^3 flops, the matrix stays in L2.
static __device__ __noinline__ void la_lu_factor(double* A, int* piv, int T, LaSmem& sm) {
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
// B <- A^{-1} B using the factorization above (B: T x T, in place)
static __device__ __noinline__ void la_lu_solve(const double* __restrict__ LU, const int* __restrict__ piv, double* B, int T, LaSmem& sm) {
    const int tid = threadIdx.x;
    for (int k = 0; k < T; ++k) {   // apply the row interchanges
        const int p = piv[k];
        if (p != k)
            for (int c = tid; c < T; c += LA_THREADS) {
                double t = B[(int64_t)k * T + c]; B[(int64_t)k * T + c] = B[(int64_t)p * T + c]; B[(int64_t)p * T + c] = t;
            }
        __syncthreads();
    }
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
    __syncthreads();
    (void)sm;
}

}  // namespace hgp
