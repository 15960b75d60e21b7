// Does non-tensor instruction issue overlap with DMMA execution on the same SM sub-partition?
// 8 warps/CTA, 1 CTA/SM: warps 0-3 (one per sub-partition) run a DMMA loop; warps 4-7 run nothing / an
// integer ALU loop / an LDS loop / an FP64 FMA loop.  Reports DMMA TFLOP/s for each companion.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 1) k(double* out, int iters, int mode) {
    __shared__ double sm[1024];
    int warp = threadIdx.x >> 5;
    sm[threadIdx.x] = threadIdx.x; sm[threadIdx.x + 256] = 1.0; __syncthreads();
    if (warp < 4) {
        double c[8][2]; double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-7;
#pragma unroll
        for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
        out[blockIdx.x * 256 + threadIdx.x] = s;
    } else {
        if (mode == 0) return;
        // run roughly as long as the DMMA warps: iters*8 DMMAs * 16 clk = iters*128 clk
        long long t_end = clock64() + (long long)iters * 128;
        if (mode == 1) {            // integer ALU: 1 instr/clk attempts
            unsigned x = threadIdx.x, y = 12345;
            while (clock64() < t_end) {
#pragma unroll
                for (int i = 0; i < 64; ++i) { x = x * 1664525u + y; y ^= x >> 3; }
            }
            out[blockIdx.x * 256 + threadIdx.x] = x + y;
        } else if (mode == 2) {     // shared loads
            double acc = 0; int idx = threadIdx.x & 255;
            while (clock64() < t_end) {
#pragma unroll
                for (int i = 0; i < 32; ++i) { acc += sm[(idx + i * 32) & 1023]; }
            }
            out[blockIdx.x * 256 + threadIdx.x] = acc;
        } else if (mode == 3) {     // FP64 FMA
            double x = threadIdx.x, y = 1.0000001;
            while (clock64() < t_end) {
#pragma unroll
                for (int i = 0; i < 32; ++i) x = fma(x, y, 1e-9);
            }
            out[blockIdx.x * 256 + threadIdx.x] = x;
        } else if (mode == 4) {     // sparse integer: ~1 instr per 4 clk
            unsigned x = threadIdx.x;
            while (clock64() < t_end) { x = x * 1664525u + 1013904223u; __nanosleep(0); }
            out[blockIdx.x * 256 + threadIdx.x] = x;
        }
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; double* out; cudaMalloc(&out, sizeof(double) * sms * 256);
    const int iters = 40000;
    const char* names[] = {"none", "int_alu", "lds", "dfma", "sparse_int"};
    printf("{");
    for (int mode = 0; mode < 5; ++mode) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<<<sms, 256>>>(out, iters, mode); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<<<sms, 256>>>(out, iters, mode); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%s\"dmma_tflops_with_%s\": %.2f", mode ? ", " : "", names[mode], 2.0 * 256 * 8 * iters * 4.0 * sms / ms * 1e-9);
    }
    printf("}\n");
    return 0;
}
