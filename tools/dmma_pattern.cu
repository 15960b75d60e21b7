// Replays the tile kernel's per-chunk DMMA pattern from registers only: 4 row blocks x 8 n-tiles x 2 k-steps,
// 64 accumulators per lane, distinct A per block and B per n-tile.  2 warps per sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(double* out, int iters, const double* in) {
    double acc[4][8][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) acc[j][nt][0] = acc[j][nt][1] = 0.0;
    double2 a[4], b[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] = make_double2(in[threadIdx.x + j], in[threadIdx.x + 7 + j]);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) b[nt] = make_double2(in[threadIdx.x + 32 + nt], in[threadIdx.x + 64 + nt]);
    const int nblk = MODE == 2 ? 2 : 4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < nblk) {
                if (MODE == 0 || MODE == 2) {        // ks-outer within block (kernel order)
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) dmma(acc[j][nt][0], acc[j][nt][1], a[j].x, b[nt].x);
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) dmma(acc[j][nt][0], acc[j][nt][1], a[j].y, b[nt].y);
                } else {                              // back-to-back dependent pairs
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) { dmma(acc[j][nt][0], acc[j][nt][1], a[j].x, b[nt].x); dmma(acc[j][nt][0], acc[j][nt][1], a[j].y, b[nt].y); }
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) s += acc[j][nt][0] + acc[j][nt][1];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}
template <int MODE> double run(int sms, double* out, double* in, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, 256>>>(out, iters, in); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<sms, 256>>>(out, iters, in); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int per = (MODE == 2 ? 32 : 64);
    return 2.0 * 256 * per * (double)iters * 8.0 * sms / ms * 1e-9;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; double *out, *in; cudaMalloc(&out, sizeof(double) * sms * 256); cudaMalloc(&in, 8 * 1024);
    cudaMemset(in, 0, 8 * 1024);
    printf("{\"pattern_ks_outer_tflops\": %.2f, \"pattern_dependent_pairs_tflops\": %.2f, \"pattern_2blocks_tflops\": %.2f}\n",
           run<0>(sms, out, in, 20000), run<1>(sms, out, in, 20000), run<2>(sms, out, in, 20000));
    return 0;
}
