// CTA-level dense linear algebra for SMALL systems (T <= 92: the MIT-BIH beat length is 90) with the operands resident
// in shared memory.  The routines of hgp_cta_la.cuh keep their matrices in global memory (L2) and are bound by L2 round
// trips and CTA barriers at this size (a 90 x 90 x 90 product took 33 us, a pivoted LU + solve 230 us on one SM).  Here
//   * three T x T operand buffers live in shared memory as a tiny write-through cache of global matrices ("tags"):
//     a product loads the operands it does not already hold, runs entirely out of shared memory on the FP64 tensor
//     cores (DMMA.8x8x4, no barrier inside the k loop), keeps its result in the third buffer and writes it through to
//     its global home -- so chains of products (A S A^T + G, ...) never wait for L2 between steps;
//   * every solve against a symmetric positive-definite matrix is  chol -> L^{-1} -> products:  `sl_cholinv` is a
//     blocked right-looking Cholesky (16-column panels: diagonal block and its inverse by one warp out of registers,
//     panel by plain FMAs, trailing update on the tensor cores) that produces L^{-1} row block by row block in the same
//     sweep, so that X = M^{-1} B is two triangular tensor-core products instead of two latency-bound substitutions.
// Used by the chain kernel (hgp_chain.cu, chain_kernel_small) and the hyper-parameter fit.
//
// Global operands are read with ld.global.cg (L2 only): the pipelined chain kernel hands matrices and vectors from one CTA
// of a cluster to another through global memory, and a line left in a reader's L1 by an earlier step would be stale.
//
// All buffer accesses go through `sl_dyn`, the kernel's dynamic shared memory, so that they compile to LDS / STS (a
// pointer fetched from a struct would be a generic pointer: LD.E + address translation in the inner loops).
#pragma once
#include "hgp_common.cuh"

namespace hgp {

extern __shared__ __align__(16) double sl_dyn[];     // the three operand buffers; kernels using this header own no other

constexpr int SL_THREADS = 256;
constexpr int SL_NB = 16;            // Cholesky panel width
constexpr int SL_MT = 3, SL_NT = 6;  // tiles per warp of the 4 x 2 warp grid: up to 12 x 12 tiles of 8 x 8 (T <= 96)

struct SlCtx {
    const double* tag[3];            // global matrix currently held by buffer i (nullptr: none)
    unsigned stamp[3];
    unsigned clock;
    int T, TP, LD, K4, stride;       // TP = T rounded up to 16 (rows), LD = row stride (== 4 mod 8, >= K4), K4 = T rounded
                                     // up to 4, stride = TP * LD doubles per buffer
    int flag;
    double logdet;
    double diag[SL_NB * (SL_NB + 1)];
    double winv[SL_NB * (SL_NB + 1)];
    double red[SL_THREADS / 32];
};

__host__ __device__ inline int sl_ld(int T) { return ((T + 3) / 8) * 8 + 4; }
__host__ __device__ inline int sl_tp(int T) { return (T + 15) & ~15; }   // rows: whole 16-row panels
__host__ __device__ inline size_t sl_dynamic_smem_bytes(int T) { return 3 * (size_t)sl_tp(T) * sl_ld(T) * sizeof(double); }
// largest supported system: 12 x 12 tiles and three buffers inside the 227 KB a CTA may own (static part ~5 KB)
__host__ __device__ inline bool sl_supported(int T) {
    return T >= 4 && sl_tp(T) <= 96 && sl_dynamic_smem_bytes(T) + 6 * 1024 <= 227 * 1024;
}

__device__ __forceinline__ void sl_init(SlCtx& c, int T) {
    const int TP = sl_tp(T), LD = sl_ld(T);
    if (threadIdx.x == 0) {
        c.T = T; c.TP = TP; c.LD = LD; c.K4 = (T + 3) & ~3; c.stride = TP * LD; c.clock = 0; c.flag = 0; c.logdet = 0.0;
        for (int b = 0; b < 3; ++b) { c.tag[b] = nullptr; c.stamp[b] = 0; }
    }
    for (int i = threadIdx.x; i < 3 * TP * LD; i += SL_THREADS) sl_dyn[i] = 0.0;      // pads stay zero for ever
    __syncthreads();
}

// ---- the operand cache -------------------------------------------------------------------------------------------
// All threads take the same decisions (tags are read after a barrier); every public routine starts with a barrier.
__device__ __forceinline__ int sl_find(const SlCtx& c, const double* g) {
    int b = -1;
#pragma unroll
    for (int i = 0; i < 3; ++i) if (c.tag[i] == g) b = i;
    return b;
}
__device__ __forceinline__ int sl_victim(const SlCtx& c, int pinA, int pinB) {
    int best = -1;
    unsigned bs = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (i == pinA || i == pinB) continue;
        const unsigned s = c.tag[i] ? c.stamp[i] : 0u;
        if (best < 0 || s < bs) { best = i; bs = s; }
    }
    return best;
}
__device__ __forceinline__ void sl_touch(SlCtx& c, int b, const double* g) {     // thread 0 only, followed by a barrier
    c.tag[b] = g;
    c.stamp[b] = ++c.clock;
}
// a buffer other than the pinned ones, its tag cleared (content undefined, pads zero)
__device__ __forceinline__ int sl_alloc(SlCtx& c, int pinA, int pinB) {
    const int b = sl_victim(c, pinA, pinB);
    __syncthreads();
    if (threadIdx.x == 0) c.tag[b] = nullptr;
    __syncthreads();
    return b;
}
__device__ __forceinline__ void sl_invalidate(SlCtx& c, const double* g) {
    __syncthreads();
    if (threadIdx.x == 0)
        for (int i = 0; i < 3; ++i) if (c.tag[i] == g) c.tag[i] = nullptr;
    __syncthreads();
}
// global T x T (row-major, dense) -> buffer b; all loads of a thread are in flight before the first store
__device__ __forceinline__ void sl_load(SlCtx& c, int b, const double* __restrict__ g) {
    const int T = c.T, LD = c.LD;
    double* dst = sl_dyn + b * c.stride;
    if ((T & 1) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0) {
        const int h = T >> 1, n2 = T * h;
        const double2* g2 = reinterpret_cast<const double2*>(g);
        constexpr int U = 8;
        for (int i0 = threadIdx.x; i0 < n2; i0 += SL_THREADS * U) {
            double2 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { const int i = i0 + u * SL_THREADS; if (i < n2) v[u] = __ldcg(g2 + i); }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * SL_THREADS;
                if (i < n2) { const int r = i / h, cc = i - r * h; *reinterpret_cast<double2*>(dst + r * LD + 2 * cc) = v[u]; }
            }
        }
    } else {
        constexpr int U = 8;
        const int n = T * T;
        for (int i0 = threadIdx.x; i0 < n; i0 += SL_THREADS * U) {
            double v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { const int i = i0 + u * SL_THREADS; if (i < n) v[u] = __ldcg(g + i); }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * SL_THREADS;
                if (i < n) { const int r = i / T, cc = i - r * T; dst[r * LD + cc] = v[u]; }
            }
        }
    }
}
// the buffer that holds g (loaded on a miss); pinA / pinB are never evicted
__device__ __forceinline__ int sl_get(SlCtx& c, const double* g, int pinA = -1, int pinB = -1) {
    __syncthreads();
    int b = sl_find(c, g);
    if (b >= 0) {
        __syncthreads();
        if (threadIdx.x == 0) c.stamp[b] = ++c.clock;
        __syncthreads();
        return b;
    }
    b = sl_victim(c, pinA, pinB);
    __syncthreads();
    sl_load(c, b, g);
    if (threadIdx.x == 0) sl_touch(c, b, g);
    __syncthreads();
    return b;
}

// ---- vectors (global memory; every load of a thread is issued before the first reduction) ------------------------
// y = alpha * A x + beta * y0 (y0 may be null); A global dense T x T, T <= 96
__device__ __forceinline__ void sl_gemv(double* __restrict__ y, const double* __restrict__ A, const double* __restrict__ x,
                                        int T, double beta, const double* __restrict__ y0, double alpha = 1.0) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double xv[3], acc[12];
#pragma unroll
    for (int q = 0; q < 3; ++q) { const int k = lane + 32 * q; xv[q] = (k < T) ? __ldcg(x + k) : 0.0; }
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const int r = warp + 8 * i;
        double a = 0.0;
        if (r < T) {
#pragma unroll
            for (int q = 0; q < 3; ++q) { const int k = lane + 32 * q; if (k < T) a += __ldcg(A + (int64_t)r * T + k) * xv[q]; }
        }
        acc[i] = a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < 12; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const int r = warp + 8 * i;
        if (r < T && lane == 0) y[r] = alpha * acc[i] + (y0 ? beta * __ldcg(y0 + r) : 0.0);
    }
    __syncthreads();
}
// D = a X + b Y elementwise over n doubles (global)
__device__ __forceinline__ void sl_axpby(double* __restrict__ D, double a, const double* __restrict__ X, double b,
                                         const double* __restrict__ Y, int n) {
    constexpr int U = 8;
    for (int i0 = threadIdx.x; i0 < n; i0 += SL_THREADS * U) {
        double xv[U], yv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int i = i0 + u * SL_THREADS; if (i < n) { xv[u] = __ldcg(X + i); yv[u] = __ldcg(Y + i); } }
#pragma unroll
        for (int u = 0; u < U; ++u) { const int i = i0 + u * SL_THREADS; if (i < n) D[i] = a * xv[u] + b * yv[u]; }
    }
    __syncthreads();
}
__device__ __forceinline__ void sl_copy(double* __restrict__ D, const double* __restrict__ S, int n) {
    constexpr int U = 8;
    for (int i0 = threadIdx.x; i0 < n; i0 += SL_THREADS * U) {
        double v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { const int i = i0 + u * SL_THREADS; if (i < n) v[u] = __ldcg(S + i); }
#pragma unroll
        for (int u = 0; u < U; ++u) { const int i = i0 + u * SL_THREADS; if (i < n) D[i] = v[u]; }
    }
    __syncthreads();
}

// ---- products ------------------------------------------------------------------------------------------------------
constexpr int SL_TRI_A = 1;   // A is lower triangular as stored (op(A) = A^T is then upper): skip the zero part of k
constexpr int SL_TRI_B = 2;   // B is lower triangular as stored

struct SlEpi {                // C = alpha * op(A) op(B) + beta * D + diag_add * I + s * u v^T
    double alpha = 1.0, beta = 0.0, diag_add = 0.0, s = 0.0;
    const double* D = nullptr;   // global, dense T x T
    const double* u = nullptr;   // global vectors
    const double* v = nullptr;
};

// the k loop of one warp: acc[i][j] += op(A)[r0 + 8 i ..][k] op(B)[k][c0 + 8 j ..], fragments straight from shared memory
template <int TA, int TB, bool FULL>
__device__ __forceinline__ void sl_mma_core(double (&acc)[SL_MT][SL_NT][2], int ia, int ib, int stride, int LD, int r0, int c0,
                                            int mcnt, int ncnt, int klo, int khi, int lane) {
    const int lr = lane >> 2, lk = lane & 3;
    const double* A = sl_dyn + ia * stride;
    const double* B = sl_dyn + ib * stride;
    const double* ap = TA ? A + (klo + lk) * LD + r0 + lr : A + (r0 + lr) * LD + klo + lk;
    const double* bp = TB ? B + (c0 + lr) * LD + klo + lk : B + (klo + lk) * LD + c0 + lr;
    const int sa = TA ? 4 * LD : 4, sb = TB ? 4 : 4 * LD;      // step of one k block
    const int ra = TA ? 8 : 8 * LD, cb = TB ? 8 * LD : 8;      // step of one tile
#pragma unroll 2
    for (int k = klo; k < khi; k += 4) {
        double a[SL_MT], b[SL_NT];
#pragma unroll
        for (int i = 0; i < SL_MT; ++i) if (FULL || i < mcnt) a[i] = ap[i * ra];
#pragma unroll
        for (int j = 0; j < SL_NT; ++j) if (FULL || j < ncnt) b[j] = bp[j * cb];
#pragma unroll
        for (int i = 0; i < SL_MT; ++i)
            if (FULL || i < mcnt) {
#pragma unroll
                for (int j = 0; j < SL_NT; ++j) if (FULL || j < ncnt) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        ap += sa;
        bp += sb;
    }
}

// C (global, dense) = epilogue(op(A) op(B)); A, B global dense T x T.  The result also stays cached in shared memory.
// `Calso`: a second global destination for the same values (the chain stores a new covariance in two histories).
static __device__ __noinline__ void sl_gemm(SlCtx& c, double* Cg, const double* Ag, int tA, const double* Bg, int tB,
                                            const SlEpi& ep, int tri = 0, double* Calso = nullptr) {
    const int ia = sl_get(c, Ag);
    const int ib = (Bg == Ag) ? ia : sl_get(c, Bg, ia);
    const int ic = sl_alloc(c, ia, ib);
    const int T = c.T, LD = c.LD, TP = c.TP, stride = c.stride;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;
    const int ntile = TP >> 3;
    const int mtw = (ntile + 3) >> 2, ntw = (ntile + 1) >> 1;
    const int mt0 = wm * mtw, nt0 = wn * ntw;
    const int mcnt = max(0, min(mtw, ntile - mt0)), ncnt = max(0, min(ntw, ntile - nt0));
    double acc[SL_MT][SL_NT][2];
#pragma unroll
    for (int i = 0; i < SL_MT; ++i)
#pragma unroll
        for (int j = 0; j < SL_NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    if (mcnt > 0 && ncnt > 0) {
        const int r0 = mt0 * 8, r1 = (mt0 + mcnt) * 8, c0 = nt0 * 8, c1 = (nt0 + ncnt) * 8;
        int klo = 0, khi = c.K4;
        if (tri & SL_TRI_A) { if (tA) klo = max(klo, r0); else khi = min(khi, r1); }
        if (tri & SL_TRI_B) { if (tB) khi = min(khi, c1); else klo = max(klo, c0); }
        klo &= ~3;
        khi = min(c.K4, (khi + 3) & ~3);
        const bool full = (mcnt == SL_MT && ncnt == SL_NT);      // T = 90: every warp owns 3 x 6 tiles
        if (full) {
            if (!tA && !tB) sl_mma_core<0, 0, true>(acc, ia, ib, stride, LD, r0, c0, mcnt, ncnt, klo, khi, lane);
            else if (tA && !tB) sl_mma_core<1, 0, true>(acc, ia, ib, stride, LD, r0, c0, mcnt, ncnt, klo, khi, lane);
            else if (!tA && tB) sl_mma_core<0, 1, true>(acc, ia, ib, stride, LD, r0, c0, mcnt, ncnt, klo, khi, lane);
            else sl_mma_core<1, 1, true>(acc, ia, ib, stride, LD, r0, c0, mcnt, ncnt, klo, khi, lane);
        } else {
            if (!tA && !tB) sl_mma_core<0, 0, false>(acc, ia, ib, stride, LD, r0, c0, mcnt, ncnt, klo, khi, lane);
            else if (tA && !tB) sl_mma_core<1, 0, false>(acc, ia, ib, stride, LD, r0, c0, mcnt, ncnt, klo, khi, lane);
            else if (!tA && tB) sl_mma_core<0, 1, false>(acc, ia, ib, stride, LD, r0, c0, mcnt, ncnt, klo, khi, lane);
            else sl_mma_core<1, 1, false>(acc, ia, ib, stride, LD, r0, c0, mcnt, ncnt, klo, khi, lane);
        }
        // epilogue: lane holds C[r][cc], C[r][cc + 1].  The D operand of a row of tiles is fetched before the first store.
        const int lr = lane >> 2, lk = lane & 3;
        double* Cs = sl_dyn + ic * stride;
        const bool pair_ok = (T & 1) == 0 && (reinterpret_cast<uintptr_t>(Cg) & 15) == 0 &&
                             (!ep.D || (reinterpret_cast<uintptr_t>(ep.D) & 15) == 0) &&
                             (!Calso || (reinterpret_cast<uintptr_t>(Calso) & 15) == 0);
#pragma unroll
        for (int i = 0; i < SL_MT; ++i) {
            if (i >= mcnt) continue;
            const int r = r0 + 8 * i + lr;
            if (r >= T) continue;
            const double ur = (ep.u ? ep.s * __ldcg(ep.u + r) : 0.0);
            double2 dv[SL_NT];
            if (ep.D) {
#pragma unroll
                for (int j = 0; j < SL_NT; ++j) {
                    const int cc = c0 + 8 * j + 2 * lk;
                    dv[j] = make_double2(0.0, 0.0);
                    if (j < ncnt && cc < T) {
                        if (pair_ok) dv[j] = __ldcg(reinterpret_cast<const double2*>(ep.D + (int64_t)r * T + cc));
                        else { dv[j].x = __ldcg(ep.D + (int64_t)r * T + cc); if (cc + 1 < T) dv[j].y = __ldcg(ep.D + (int64_t)r * T + cc + 1); }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < SL_NT; ++j) {
                if (j >= ncnt) continue;
                const int cc = c0 + 8 * j + 2 * lk;
                if (cc >= T) continue;
                double2 v;
                v.x = ep.alpha * acc[i][j][0];
                v.y = ep.alpha * acc[i][j][1];
                if (ep.D) { v.x += ep.beta * dv[j].x; v.y += ep.beta * dv[j].y; }
                if (cc == r) v.x += ep.diag_add;
                if (cc + 1 == r) v.y += ep.diag_add;
                if (ep.u) { v.x += ur * __ldcg(ep.v + cc); if (cc + 1 < T) v.y += ur * __ldcg(ep.v + cc + 1); }
                if (pair_ok) {                       // T even: cc + 1 < T, 16-byte aligned everywhere (LD is even as well)
                    *reinterpret_cast<double2*>(Cs + r * LD + cc) = v;
                    *reinterpret_cast<double2*>(Cg + (int64_t)r * T + cc) = v;
                    if (Calso) *reinterpret_cast<double2*>(Calso + (int64_t)r * T + cc) = v;
                } else {
                    Cs[r * LD + cc] = v.x;
                    Cg[(int64_t)r * T + cc] = v.x;
                    if (Calso) Calso[(int64_t)r * T + cc] = v.x;
                    if (cc + 1 < T) {
                        Cs[r * LD + cc + 1] = v.y;
                        Cg[(int64_t)r * T + cc + 1] = v.y;
                        if (Calso) Calso[(int64_t)r * T + cc + 1] = v.y;
                    }
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) if (c.tag[i] == Cg || (Calso && c.tag[i] == Calso)) c.tag[i] = nullptr;
        sl_touch(c, ic, Cg);
    }
    __syncthreads();
}

// ---- Cholesky + inverse of the factor ---------------------------------------------------------------------------------
// Linv (global, dense, lower triangular, upper part zero) = chol(0.5 (M + M^T) + add_diag I)^{-1}; M global dense SPD.
// Returns 0 or (index + 1) of the first non-positive pivot (Linv is then garbage).  c.logdet = log det of the
// factorised matrix (2 sum log L_jj).
// Buffer form: X = buffer ix (destroyed), Y = buffer iy receives the inverse factor; tags are the caller's business.
static __device__ __noinline__ int sl_cholinv_buf(SlCtx& c, int ix, int iy, double add_diag) {
    const int T = c.T, LD = c.LD, TP = c.TP;
    double* X = sl_dyn + ix * c.stride;
    double* Y = sl_dyn + iy * c.stride;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int lr = lane >> 2, lk = lane & 3;
    // X <- sym(X) + add_diag I (both triangles); Y <- 0
    for (int i = tid; i < T * T; i += SL_THREADS) {
        const int r = i / T, cc = i - r * T;
        if (cc < r) {
            const double v = 0.5 * (X[r * LD + cc] + X[cc * LD + r]);
            X[r * LD + cc] = v; X[cc * LD + r] = v;
        } else if (cc == r) X[r * LD + r] += add_diag;
        Y[r * LD + cc] = 0.0;
    }
    if (tid == 0) { c.flag = 0; c.logdet = 0.0; }
    __syncthreads();
    for (int k0 = 0; k0 < T; k0 += SL_NB) {
        const int nb = min(SL_NB, T - k0), k1 = k0 + nb;
        const int nb4 = (nb + 3) & ~3;                // columns k0 .. k0 + nb4 exist (pads are zero)
        // (a) warp 0: factor the diagonal block and invert it.  Lane i (< 16; the upper half-warp mirrors the lower) keeps
        //     row i of the block in registers: a column step is one shuffle for the pivot, a reciprocal square root, and a
        //     fully unrolled rank-1 update whose multipliers travel by shuffle -- no shared-memory round trip on the
        //     serial path.  A short last block is padded with an identity tail.
        if (warp == 0) {
            const int li = lane & 15;
            double row[SL_NB];
#pragma unroll
            for (int cc = 0; cc < SL_NB; ++cc)
                row[cc] = (li < nb && cc <= li) ? X[(k0 + li) * LD + k0 + cc] : (li == cc ? 1.0 : 0.0);
            double dinv = 1.0;
            int bad = 0;
#pragma unroll
            for (int j = 0; j < SL_NB; ++j) {
                const double d = __shfl_sync(0xffffffffu, row[j], j);
                if (!(d > 0.0) && bad == 0) bad = k0 + j + 1;
                const double rs = rsqrt(d);
                double l = row[j] * rs;
                if (li == j) { l = d * rs; dinv = rs; }
                if (li >= j) row[j] = l;
#pragma unroll
                for (int cc = j + 1; cc < SL_NB; ++cc) {
                    const double lc = __shfl_sync(0xffffffffu, l, cc);
                    if (li >= cc) row[cc] -= l * lc;
                }
            }
            // log det: the diagonal entry is picked with a static scan (a runtime index would spill the row to local memory)
            double dsel = 1.0;
#pragma unroll
            for (int cc = 0; cc < SL_NB; ++cc) if (cc == li) dsel = row[cc];
            double ldg = (li < nb && lane < 16) ? log(dsel) : 0.0;
            ldg = warp_sum(ldg);
            if (lane == 0) { c.logdet += 2.0 * ldg; if (bad && c.flag == 0) c.flag = bad; }
            double* D = c.diag;
            if (lane < 16) {
#pragma unroll
                for (int cc = 0; cc < SL_NB; ++cc) D[li * (SL_NB + 1) + cc] = row[cc];
                D[li * (SL_NB + 1) + SL_NB] = dinv;              // column 16 of the padded block: 1 / L[i][i]
            }
            __syncwarp();
            // inverse: lane j owns column j of Winv; coefficients are broadcast reads of the factor just stored
            if (lane < SL_NB) {
                double x[SL_NB];
#pragma unroll
                for (int i = 0; i < SL_NB; ++i) {
                    double v = (i == lane) ? 1.0 : 0.0;
#pragma unroll
                    for (int p = 0; p < SL_NB; ++p) if (p < i) v -= D[i * (SL_NB + 1) + p] * ((p >= lane) ? x[p] : 0.0);
                    x[i] = (i >= lane) ? v * D[i * (SL_NB + 1) + SL_NB] : 0.0;
                }
#pragma unroll
                for (int i = 0; i < SL_NB; ++i) c.winv[i * (SL_NB + 1) + lane] = x[i];
            }
        }
        __syncthreads();
        // (b) warps 0-3: panel L21 = X[k1:, k0:k1] Winv^T, one thread per row, in place
        //     warps 4-7: F = -Winv X[k0:k1, 0:k0], one thread per column, in place (row block k of L is dead afterwards)
        if (warp < 4) {
            for (int r = k1 + tid; r < T; r += 128) {
                double x[SL_NB], o[SL_NB];
#pragma unroll
                for (int p = 0; p < SL_NB; ++p) x[p] = (p < nb4) ? X[r * LD + k0 + p] : 0.0;
#pragma unroll
                for (int cc = 0; cc < SL_NB; ++cc) {
                    double v = 0.0;
#pragma unroll
                    for (int p = 0; p < SL_NB; ++p) if (p <= cc) v += x[p] * c.winv[cc * (SL_NB + 1) + p];
                    o[cc] = v;
                }
#pragma unroll
                for (int cc = 0; cc < SL_NB; ++cc) if (cc < nb) X[r * LD + k0 + cc] = o[cc];
            }
        } else {
            for (int cc = tid - 128; cc < k0; cc += 128) {
                double x[SL_NB], o[SL_NB];
#pragma unroll
                for (int p = 0; p < SL_NB; ++p) x[p] = (p < nb) ? X[(k0 + p) * LD + cc] : 0.0;
#pragma unroll
                for (int i = 0; i < SL_NB; ++i) {
                    double v = 0.0;
#pragma unroll
                    for (int p = 0; p < SL_NB; ++p) if (p <= i) v -= c.winv[i * (SL_NB + 1) + p] * x[p];
                    o[i] = v;
                }
#pragma unroll
                for (int i = 0; i < SL_NB; ++i) if (i < nb) X[(k0 + i) * LD + cc] = o[i];
            }
        }
        __syncthreads();
        // (c) trailing update X[k1:, k1:] -= L21 L21^T on the tensor cores (8 x 8 tiles of the lower half), the tile
        //     pairs dealt round-robin to the warps, up to seven independent accumulators per warp in flight
        if (k1 < T) {
            const int t0 = k1 >> 3, nt = (TP >> 3) - t0;      // k1 is a multiple of 16 here
            const int npairs = nt * (nt + 1) / 2;
            constexpr int PW = 7;                             // nt <= 10: at most 55 pairs over 8 warps
            double acc[PW][2];
            int rr[PW], cc0[PW];
#pragma unroll
            for (int q = 0; q < PW; ++q) {
                acc[q][0] = acc[q][1] = 0.0;
                const int p = warp + 8 * q;
                int ti = 0;
                while ((ti + 1) * (ti + 2) / 2 <= p) ++ti;
                const int tj = p - ti * (ti + 1) / 2;
                rr[q] = (t0 + ti) * 8; cc0[q] = (t0 + tj) * 8;
            }
#pragma unroll
            for (int kk = 0; kk < SL_NB; kk += 4) {
                if (kk < nb4) {
#pragma unroll
                    for (int q = 0; q < PW; ++q)
                        if (warp + 8 * q < npairs)
                            dmma884(acc[q][0], acc[q][1], X[(rr[q] + lr) * LD + k0 + kk + lk], X[(cc0[q] + lr) * LD + k0 + kk + lk]);
                }
            }
#pragma unroll
            for (int q = 0; q < PW; ++q) {
                // rows / columns >= T are padding (their products are zero): never written, a tile may reach past the stride
                if (warp + 8 * q < npairs && rr[q] + lr < T) {
                    double* dst = X + (rr[q] + lr) * LD + cc0[q] + 2 * lk;
                    if (cc0[q] + 2 * lk < T) dst[0] -= acc[q][0];
                    if (cc0[q] + 2 * lk + 1 < T) dst[1] -= acc[q][1];
                }
            }
        }
        // (d) row block k of the inverse: Y[k0:k1, 0:k0] = F Y[0:k0, 0:k0] (tensor cores), Y[k0:k1, k0:k1] = Winv.
        //     Warp w owns column tiles w and w + 8; even and odd k blocks accumulate separately (four chains per tile pair).
        {
            const int ntn = k0 >> 3;                          // k0 is a multiple of 16
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int tj = warp + 8 * q;
                if (tj >= ntn) continue;
                double a[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
                const int cs = tj * 8;
                for (int k = cs; k < k0; k += 8) {            // Y[0:k0, 0:k0] is lower triangular: rows k >= column
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int kk = k + 4 * h;
                        const double bv = Y[(kk + lk) * LD + cs + lr];
                        dmma884(a[h][0][0], a[h][0][1], X[(k0 + lr) * LD + kk + lk], bv);
                        dmma884(a[h][1][0], a[h][1][1], X[(k0 + 8 + lr) * LD + kk + lk], bv);
                    }
                }
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const int r = k0 + 8 * hh + lr;
                    if (r < k1) {
                        Y[r * LD + cs + 2 * lk] = a[0][hh][0] + a[1][hh][0];
                        Y[r * LD + cs + 2 * lk + 1] = a[0][hh][1] + a[1][hh][1];
                    }
                }
            }
            for (int i = tid; i < nb * nb; i += SL_THREADS) {
                const int r = i / nb, cc = i - r * nb;
                Y[(k0 + r) * LD + k0 + cc] = (cc <= r) ? c.winv[r * (SL_NB + 1) + cc] : 0.0;
            }
        }
        __syncthreads();
    }
    return c.flag;
}

static __device__ __noinline__ int sl_cholinv(SlCtx& c, double* Linv_g, const double* Mg, double add_diag) {
    const int ix = sl_get(c, Mg);
    const int iy = sl_alloc(c, ix, -1);
    if (threadIdx.x == 0) c.tag[ix] = nullptr;               // X is destroyed by the factorisation
    __syncthreads();
    const int info = sl_cholinv_buf(c, ix, iy, add_diag);
    // write through
    {
        const int T = c.T, LD = c.LD, tid = threadIdx.x;
        const double* Y = sl_dyn + iy * c.stride;
        constexpr int U = 4;
        const int n = T * T;
        for (int i0 = tid; i0 < n; i0 += SL_THREADS * U) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * SL_THREADS;
                if (i < n) { const int r = i / T, cc = i - r * T; Linv_g[i] = Y[r * LD + cc]; }
            }
        }
        if (tid == 0) {
            for (int i = 0; i < 3; ++i) if (c.tag[i] == Linv_g) c.tag[i] = nullptr;
            sl_touch(c, iy, Linv_g);
        }
    }
    __syncthreads();
    return info;
}

// Buffer form of the product: buffer ic = op(buffer ia) op(buffer ib) (ic must differ from ia and ib); nothing goes to
// global memory, tags are the caller's business.  Used where a whole iteration lives in shared memory (hyper-fit).
static __device__ __noinline__ void sl_gemm_buf(SlCtx& c, int ic, int ia, int tA, int ib, int tB, int tri = 0) {
    __syncthreads();
    const int T = c.T, LD = c.LD, TP = c.TP, stride = c.stride;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp >> 1, wn = warp & 1;
    const int ntile = TP >> 3;
    const int mtw = (ntile + 3) >> 2, ntw = (ntile + 1) >> 1;
    const int mt0 = wm * mtw, nt0 = wn * ntw;
    const int mcnt = max(0, min(mtw, ntile - mt0)), ncnt = max(0, min(ntw, ntile - nt0));
    double acc[SL_MT][SL_NT][2];
#pragma unroll
    for (int i = 0; i < SL_MT; ++i)
#pragma unroll
        for (int j = 0; j < SL_NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    if (mcnt > 0 && ncnt > 0) {
        const int r0 = mt0 * 8, r1 = (mt0 + mcnt) * 8, c0 = nt0 * 8, c1 = (nt0 + ncnt) * 8;
        int klo = 0, khi = c.K4;
        if (tri & SL_TRI_A) { if (tA) klo = max(klo, r0); else khi = min(khi, r1); }
        if (tri & SL_TRI_B) { if (tB) khi = min(khi, c1); else klo = max(klo, c0); }
        klo &= ~3;
        khi = min(c.K4, (khi + 3) & ~3);
        if (!tA && !tB) sl_mma_core<0, 0, false>(acc, ia, ib, stride, LD, r0, c0, mcnt, ncnt, klo, khi, lane);
        else if (tA && !tB) sl_mma_core<1, 0, false>(acc, ia, ib, stride, LD, r0, c0, mcnt, ncnt, klo, khi, lane);
        else if (!tA && tB) sl_mma_core<0, 1, false>(acc, ia, ib, stride, LD, r0, c0, mcnt, ncnt, klo, khi, lane);
        else sl_mma_core<1, 1, false>(acc, ia, ib, stride, LD, r0, c0, mcnt, ncnt, klo, khi, lane);
        const int lr = lane >> 2, lk = lane & 3;
        double* Cs = sl_dyn + ic * stride;
#pragma unroll
        for (int i = 0; i < SL_MT; ++i) {
            if (i >= mcnt) continue;
            const int r = r0 + 8 * i + lr;
            if (r >= T) continue;
#pragma unroll
            for (int j = 0; j < SL_NT; ++j) {
                if (j >= ncnt) continue;
                const int cc = c0 + 8 * j + 2 * lk;
                if (cc < T) Cs[r * LD + cc] = acc[i][j][0];
                if (cc + 1 < T) Cs[r * LD + cc + 1] = acc[i][j][1];
            }
        }
    }
    __syncthreads();
}

}  // namespace hgp
