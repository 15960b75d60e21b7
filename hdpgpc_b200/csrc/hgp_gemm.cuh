// Batched T x T float64 GEMM on the FP64 tensor cores (DMMA.8x8x4), CTA tile 64 x 64, K chunks of 16.
// Building block of the latent-transition score (hgp_qlat.cu) and of the chain kernels.
#pragma once
#include "hgp_common.cuh"

namespace hgp {

constexpr int GT = 64;        // CTA tile (rows and columns)
constexpr int GK = 16;        // K chunk
constexpr int GA_PITCH = 20;  // As[64][20]: rows r and r+4 share banks -> 2 wavefronts per fragment load (optimal)
constexpr int GB_PITCH = 72;  // Bs[16][72]

// MODE 0: C[b] = op(A[ia[b]]) * B[ib[b]]            (transA: use A^T)
// MODE 1: partial[b * ntiles + tile] = sum( (A * B) .* E[b] )  over the tile (fixed order), nothing stored
// lowerA: A is lower triangular (zeros above the diagonal are skipped in the K loop)
template <int MODE>
__global__ void __launch_bounds__(256)
gemm_tile_kernel(const double* __restrict__ A, const int* __restrict__ ia, int64_t strideA, int lowerA, int transA,
                 const double* __restrict__ B, const int* __restrict__ ib, int64_t strideB,
                 double* __restrict__ C, int64_t strideC, const double* __restrict__ E, int64_t strideE,
                 double* __restrict__ partial, int T) {
    __shared__ double As[GT * GA_PITCH];
    __shared__ double Bs[GK * GB_PITCH];
    __shared__ double s_red[8];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nt = (T + GT - 1) / GT;
    const int tile = blockIdx.x;
    const int rt = tile / nt, ct = tile % nt;
    const int64_t b = blockIdx.y;
    const double* Ab = A + (ia ? (int64_t)ia[b] : b) * strideA;
    const double* Bb = B + (ib ? (int64_t)ib[b] : b) * strideB;
    const int r0 = rt * GT, c0 = ct * GT;
    const int wm = warp >> 1, wn = warp & 1;   // warp tile: rows [16 wm, 16 wm + 16), cols [32 wn, 32 wn + 32)
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int kmax = lowerA ? min(T, r0 + GT) : T;
    for (int k0 = 0; k0 < kmax; k0 += GK) {
        // A tile: 64 x 16
        for (int idx = tid; idx < GT * GK; idx += 256) {
            int r, k;
            if (transA) { k = idx / GT; r = idx % GT; } else { r = idx / GK; k = idx % GK; }
            const int gr = r0 + r, gk = k0 + k;
            double v = 0.0;
            if (gr < T && gk < T) v = transA ? Ab[(int64_t)gk * T + gr] : Ab[(int64_t)gr * T + gk];
            As[r * GA_PITCH + k] = v;
        }
        // B tile: 16 x 64
        for (int idx = tid; idx < GK * GT; idx += 256) {
            const int k = idx / GT, c = idx % GT;
            const int gk = k0 + k, gc = c0 + c;
            Bs[k * GB_PITCH + c] = (gk < T && gc < T) ? Bb[(int64_t)gk * T + gc] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < GK / 4; ++ks) {
            double a[2], bf[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = As[(16 * wm + 8 * i + (lane >> 2)) * GA_PITCH + 4 * ks + (lane & 3)];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = Bs[(4 * ks + (lane & 3)) * GB_PITCH + 32 * wn + 8 * j + (lane >> 2)];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bf[j]);
        }
        __syncthreads();
    }
    // epilogue: element (row = r0 + 16 wm + 8 i + lane/4, col = c0 + 32 wn + 8 j + 2 (lane%4) + e)
    if (MODE == 0) {
        double* Cb = C + b * strideC;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int gr = r0 + 16 * wm + 8 * i + (lane >> 2), gc = c0 + 32 * wn + 8 * j + 2 * (lane & 3) + e;
                    if (gr < T && gc < T) Cb[(int64_t)gr * T + gc] = acc[i][j][e];
                }
    } else {
        const double* Eb = E + b * strideE;
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int gr = r0 + 16 * wm + 8 * i + (lane >> 2), gc = c0 + 32 * wn + 8 * j + 2 * (lane & 3) + e;
                    if (gr < T && gc < T) s += acc[i][j][e] * Eb[(int64_t)gr * T + gc];
                }
        s = warp_sum(s);
        if (lane == 0) s_red[warp] = s;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int w = 0; w < 8; ++w) tot += s_red[w];
            partial[b * (int64_t)(nt * nt) + tile] = tot;
        }
    }
}

}  // namespace hgp
