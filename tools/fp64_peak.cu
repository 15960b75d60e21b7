// FP64 throughput probes for B200 (sm_100a): vector DFMA and tensor DMMA (mma.sync f64 shapes).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_peak tools/fp64_peak.cu
// Used once per round to fix the FP64 roofline denominator next to cuBLAS DGEMM (bench.py --probe-fp64).
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m8n8k4: 256 FMA per warp instruction
template <int NT>
__global__ void dmma884_kernel(double* out, int iters) {
    double c[NT][2];
    double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-7;
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k16: 2048 FMA per warp instruction (sm_90+)
template <int NT>
__global__ void dmma16816_kernel(double* out, int iters) {
    double c[NT][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-6 + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = 1.0 + threadIdx.x * 1e-7 + i;
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                           "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m16n8k4: 512 FMA per warp instruction (sm_90+)
template <int NT>
__global__ void dmma1684_kernel(double* out, int iters) {
    double c[NT][4];
    double a[2] = {threadIdx.x * 1e-6, 0.5}, b = 1.0 + threadIdx.x * 1e-7;
#pragma unroll
    for (int i = 0; i < NT; ++i) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NT; ++i)
            asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NT; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d", p.name, sms);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
    const int iters = 20000;
    for (int warps : {4, 8, 16, 32}) {
        int threads = warps * 32, blocks = sms * 2;
        float ms = time_ms([&] { dfma_kernel<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf(", \"dfma_w%d_tflops\": %.3f", warps, 2.0 * 8 * iters * (double)threads * blocks / ms * 1e-9);
    }
    for (int warps : {4, 8, 16}) {
        int threads = warps * 32, blocks = sms * 2;
        float ms = time_ms([&] { dmma884_kernel<8><<<blocks, threads>>>(out, iters / 4); }, 5);
        printf(", \"dmma884_w%d_tflops\": %.3f", warps, 2.0 * 256 * 8 * (iters / 4) * (double)warps * blocks / ms * 1e-9);
        ms = time_ms([&] { dmma1684_kernel<8><<<blocks, threads>>>(out, iters / 4); }, 5);
        printf(", \"dmma1684_w%d_tflops\": %.3f", warps, 2.0 * 512 * 8 * (iters / 4) * (double)warps * blocks / ms * 1e-9);
        ms = time_ms([&] { dmma16816_kernel<4><<<blocks, threads>>>(out, iters / 8); }, 5);
        printf(", \"dmma16816_w%d_tflops\": %.3f", warps, 2.0 * 2048 * 4 * (iters / 8) * (double)warps * blocks / ms * 1e-9);
    }
    printf("}\n");
    CK(cudaDeviceSynchronize());
    return 0;
}
