// Emission means (reference GPI_model.observe, GPI_model.py:626-662: mean = C_i f_i) as a batched
// GEMV: one CTA per state, one warp per output row, coalesced row reads (HBM-bound: T*T*8 bytes
// per state when every state has its own C).
#include "hgp_common.cuh"

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

}  // namespace

extern "C" int hgp_emission_means(const double* C, const double* f, const int* c_idx, const int* f_idx, int64_t S,
                                  int T, double* mu, void* stream) {
    HGP_REQUIRE(S >= 0 && T > 0 && T <= 4096, "hgp_emission_means: bad sizes");
    if (S == 0) return 0;
    emission_means_kernel<<<(unsigned)S, 256, sizeof(double) * T, (cudaStream_t)stream>>>(C, f, c_idx, f_idx, T, mu);
    HGP_LAUNCH_CHECK("hgp_emission_means");
    return 0;
}
