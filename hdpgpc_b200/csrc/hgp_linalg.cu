// Batched SPD factorisation, triangular inverse and operand packing (sm_100a).
//
// chol_batched restates GPI_model._chol_spd (reference GPI_model.py:83-87) for a batch of
// T x T float64 matrices, one CTA per matrix, blocked right-looking with the current panel in
// shared memory and the trailing matrix in L2.  tri_inverse_batched produces W = L^{-1} so the
// Mahalanobis term becomes a single triangular product (see include/hdpgpc_b200.h).
#include "hgp_common.cuh"

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
