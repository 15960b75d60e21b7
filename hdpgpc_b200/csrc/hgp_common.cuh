// Shared helpers for the hdpgpc_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <atomic>
#include <stdio.h>

#include "../../include/hdpgpc_b200.h"

#define HGP_STR_(x) #x
#define HGP_STR(x) HGP_STR_(x)
#define HGP_LOG2PI 1.8378770664093454835606594728112  // log(2*pi)
#define HGP_EPS 2.220446049250313e-16                  // finfo(float64).eps

extern std::atomic<int64_t> g_hgp_launches;
void hgp_set_error(const char* fmt, ...);

// chol of sc[f] * Sigma[src_idx[f]] (+ add_diag[f] I); src_idx / src_scale may be NULL
int hgp_internal_chol(const double* Sigma, const int* src_idx, const double* src_scale, int64_t F, int T,
                      const double* add_diag, double jitter_scale, double* Lfac, double* logdet, int* info,
                      void* stream);

static inline int hgp_status(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    hgp_set_error("%s: %s", what, cudaGetErrorString(e));
    return -(int)e;
}

#define HGP_LAUNCH_CHECK(name)                                   \
    do {                                                         \
        g_hgp_launches.fetch_add(1, std::memory_order_relaxed);  \
        cudaError_t e_ = cudaGetLastError();                     \
        if (e_ != cudaSuccess) return hgp_status(e_, name);      \
    } while (0)

#define HGP_REQUIRE(cond, msg)                                   \
    do {                                                         \
        if (!(cond)) { hgp_set_error("%s", msg); return HGP_E_BADARG; } \
    } while (0)

__host__ __device__ inline int64_t hgp_min64(int64_t a, int64_t b) { return a < b ? a : b; }
__host__ __device__ inline int64_t hgp_max64(int64_t a, int64_t b) { return a > b ? a : b; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// D = A(8x4, row) * B(4x8, col) + D on the FP64 tensor core (SASS: DMMA.8x8x4).
// lane holds a = A[lane/4][lane%4], b = B[lane%4][lane/4], c0/c1 = C[lane/4][2*(lane%4) + {0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
