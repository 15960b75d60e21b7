// Emission scores of the E-step (reference GPI_model.compute_sq_err_all, GPI_model.py:488-547;
// log_sq_error :250-286; _gaussian_score_shared_cov :92-113) and the SNR lead statistic
// (GPI_HDP.compute_snr, GPI_HDP.py:732-748), hand-written for sm_100a.
//
// score_tiles_kernel -- the hot kernel.  For a tile of 64 consecutive beats and a run of clusters it
// evaluates  z = W_m (y_n - mu_{s(n,m)})  as a lower-triangular matrix product on the FP64 tensor
// cores (DMMA.8x8x4) and reduces |z|^2 per beat in the epilogue.  Warp-specialised:
//   * producer warpgroup (4 warps, registers released with setmaxnreg.dec): streams W_m in
//     fragment-ordered k-chunks with 1-D bulk async copies (TMA, completion on an mbarrier) through a
//     4-stage shared-memory ring, and builds the matching 8 x 64 slice of D = Y - mu (beat tile
//     resident in shared memory, means gathered through L2) directly in B-fragment order;
//   * 2 consumer warpgroups (8 warps, setmaxnreg.inc): row blocks of 8 are dealt to warps in a mirrored
//     order so that the triangular
//     shrinkage (row block rb only needs k-chunks kc <= rb) stays balanced over the 4 SM
//     sub-partitions; each warp keeps its 32 x 64 slice of z in registers (64 f64 accumulators/lane).
#include "hgp_common.cuh"
#include <type_traits>

namespace {

// ------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier + bulk async copy
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

constexpr int BT = 64;            // beats per tile
constexpr int NCW = 8;            // consumer warps (warpgroups 0 and 1)
constexpr int NPW = 4;            // producer warps (warpgroup 2)
constexpr int STAGES = 4;
constexpr int MAX_NRB = 32;       // T <= 256
constexpr int TILE_THREADS = (NCW + NPW) * 32;
constexpr int W_STAGE_BYTES = MAX_NRB * 512;   // 16 KB
constexpr int D_STAGE_BYTES = 8 * BT * 8;      // 4 KB
// Register re-balancing between the warpgroups (setmaxnreg): the kernel launches with 168
// registers/thread (3 warps per SM sub-partition); consumers grow, producers shrink.
// Per sub-partition: 2 consumer warps x 32 x 216 + 1 producer warp x 32 x 72 = 16128 <= 16384.
#ifndef HGP_PF
#define HGP_PF 2
#endif
constexpr int PF = HGP_PF;     // producer prefetch depth (k-chunks of gathered means in flight)
#define HGP_CONSUMER_REGS 216
#define HGP_PRODUCER_REGS 72

__device__ __forceinline__ int chunk_of(int i, int nrb) { (void)nrb; return i; }   // k-chunks in ascending order


struct TileSmem {
    // offsets (bytes) into dynamic shared memory
    int ytile, wst, dst, red, bars, total;
};
__host__ __device__ inline TileSmem tile_smem_layout(int Tp) {
    TileSmem s;
    int yp = Tp + 2;
    s.ytile = 0;
    s.wst = ((BT * yp * 8) + 127) / 128 * 128;
    s.dst = s.wst + STAGES * W_STAGE_BYTES;
    s.red = s.dst + STAGES * D_STAGE_BYTES;
    s.bars = s.red + 2 * NCW * BT * 8;
    s.total = s.bars + 2 * STAGES * 8;
    return s;
}

// Beat tile -> shared memory, zero padded (rows n >= N, samples t >= T), by all warps of the CTA.
__device__ __forceinline__ void load_beat_tile(double* Ytile, const double* __restrict__ Y, int64_t N, int T, int YP,
                                               int64_t n0, int warp, int lane) {
    const bool vec = ((T & 1) == 0) && ((reinterpret_cast<uintptr_t>(Y) & 15) == 0);
    for (int c = warp; c < BT; c += NCW + NPW) {
        const int64_t n = n0 + c;
        double* dst = Ytile + c * YP;
        if (n < N && vec) {
            const double2* src = reinterpret_cast<const double2*>(Y + n * T);
            for (int t2 = lane; t2 < YP / 2; t2 += 32)
                reinterpret_cast<double2*>(dst)[t2] = (2 * t2 < T) ? src[t2] : make_double2(0.0, 0.0);
        } else {
            const double* src = Y + n * T;
            for (int t = lane; t < YP; t += 32) dst[t] = (n < N && t < T) ? src[t] : 0.0;
        }
    }
}

// One k-chunk for a warp whose row blocks j >= J0 are active (T = 256 fast path): no per-block tests,
// so the A/B fragment loads of the whole chunk are issued up front and the 16 * (4 - J0) tensor-core
// instructions follow as one straight-line stream.
template <int J0>
__device__ __forceinline__ void fast_chunk(double (&acc)[4][8][2], uint32_t it, int kc, int rb0, int rb1, int rb2,
                                           int rb3, int lane, const unsigned char* Wst, const unsigned char* Dst,
                                           uint64_t* full_bar, uint64_t* empty_bar) {
    const int stage = it & (STAGES - 1);
    mbar_wait(&full_bar[stage], (it / STAGES) & 1);
    if (J0 < 4) {
        const double2* ds = reinterpret_cast<const double2*>(Dst + stage * D_STAGE_BYTES) + lane;
        const double2* ws = reinterpret_cast<const double2*>(Wst + stage * W_STAGE_BYTES) + lane - kc * 32;
        double2 b[8];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) b[nt] = ds[nt * 32];
        double2 a[4];
        if (J0 <= 0) a[0] = ws[rb0 * 32];
        if (J0 <= 1) a[1] = ws[rb1 * 32];
        if (J0 <= 2) a[2] = ws[rb2 * 32];
        if (J0 <= 3) a[3] = ws[rb3 * 32];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j >= J0) {
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) dmma884(acc[j][nt][0], acc[j][nt][1], a[j].x, b[nt].x);
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) dmma884(acc[j][nt][0], acc[j][nt][1], a[j].y, b[nt].y);
            }
        }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
}

__global__ void __launch_bounds__(TILE_THREADS, 1)
score_tiles_kernel(const double* __restrict__ Y, int64_t N, int T, const double* __restrict__ mu,
                   const double* __restrict__ Wpacked, int64_t packed_doubles, const int* __restrict__ state_of,
                   const int* __restrict__ factor_of_cluster, int M, int m_per_item, int m_splits, int64_t n_items,
                   double* __restrict__ q, const double* __restrict__ mu_sm, const int* __restrict__ snr_state_of,
                   double* __restrict__ snr) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int nrb = (T + 7) / 8;
    const int Tp = nrb * 8;
    const int YP = Tp + 2;
    const TileSmem lay = tile_smem_layout(Tp);
    double* Ytile = reinterpret_cast<double*>(smem_raw + lay.ytile);
    unsigned char* Wst = smem_raw + lay.wst;
    unsigned char* Dst = smem_raw + lay.dst;
    double* red = reinterpret_cast<double*>(smem_raw + lay.red);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + lay.bars);
    uint64_t* empty_bar = full_bar + STAGES;

    const int tid = threadIdx.x;
    // warp index through a shuffle: tells the compiler it is warp-uniform, so the role / row-block branches
    // need no reconvergence (no WARPSYNC / BSSY around the tensor-core blocks)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1 + NPW);   // arrive.expect_tx (W bytes) + one arrive per producer warp (D part)
            mbar_init(&empty_bar[s], NCW);      // one arrive per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t it = 0;   // running (m, kc) step counter; identical in producers and consumers
    const double half_T_log2pi = 0.5 * (double)T * HGP_LOG2PI;

    if (warp >= NCW) {
        // ===================================== producers =====================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HGP_PRODUCER_REGS));
        const int pw = warp - NCW;            // this warp builds n-tiles 2*pw and 2*pw+1 of every D chunk
        const int kk = lane & 3, nn = lane >> 2;
        for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int64_t tile = item / m_splits;
            const int m_begin = (int)(item % m_splits) * m_per_item;
            const int m_end = min(M, m_begin + m_per_item);
            const int64_t n0 = tile * BT;
            load_beat_tile(Ytile, Y, N, T, YP, n0, warp, lane);
            __syncthreads();   // tile complete (all 12 warps load it)
            const double* y0 = Ytile + ((2 * pw) * 8 + nn) * YP + kk;
            const double* y1 = y0 + 8 * YP;
            // State indices are fetched two clusters ahead and the mean rows of cluster m+1 are pulled into L2
            // while cluster m is produced: the per-chunk gathers below then hit L2 instead of HBM.
            const int64_t na = n0 + (2 * pw) * 8 + nn, nb = na + 8;
            auto load_state = [&](int m, int& sa, int& sb) {
                sa = (m < m_end && na < N) ? state_of[na * M + m] : -1;
                sb = (m < m_end && nb < N) ? state_of[nb * M + m] : -1;
            };
            auto prefetch_rows = [&](int sa, int sb) {
                // the 4 lanes of a column (kk = 0..3) split its row into 128-byte lines
                const int lines = (T * 8 + 127) / 128;
                for (int l = kk; l < lines; l += 4) {
                    if (sa >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(mu + (int64_t)sa * T + l * 16));
                    if (sb >= 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(mu + (int64_t)sb * T + l * 16));
                }
            };
            int sa, sb, sa1, sb1, sa2, sb2;
            load_state(m_begin, sa, sb);
            load_state(m_begin + 1, sa1, sb1);
            prefetch_rows(sa, sb);
            for (int m = m_begin; m < m_end; ++m) {
                const unsigned char* Wp =
                    reinterpret_cast<const unsigned char*>(Wpacked + (int64_t)factor_of_cluster[m] * packed_doubles);
                load_state(m + 2, sa2, sb2);
                prefetch_rows(sa1, sb1);
                const double* ma = (sa >= 0) ? mu + (int64_t)sa * T + kk : nullptr;
                const double* mb = (sb >= 0) ? mu + (int64_t)sb * T + kk : nullptr;
                // The means are gathered through L2 (~650 cycles): keep PF chunks of them in flight in a
                // register ring so the producer never waits on a load it issued less than PF steps ago.
                double ring[PF][4];
                auto load_chunk = [&](int i, double (&dst)[4]) {
                    dst[0] = dst[1] = dst[2] = dst[3] = 0.0;
                    if (i < nrb) {
                        const int o = 8 * chunk_of(i, nrb);
                        const bool w0 = o + kk < T, w1 = o + kk + 4 < T;
                        if (ma) { if (w0) dst[0] = __ldg(ma + o); if (w1) dst[1] = __ldg(ma + o + 4); }
                        if (mb) { if (w0) dst[2] = __ldg(mb + o); if (w1) dst[3] = __ldg(mb + o + 4); }
                    }
                };
#pragma unroll
                for (int p = 0; p < PF; ++p) load_chunk(p, ring[p]);
                for (int i0 = 0; i0 < nrb; i0 += PF) {
#pragma unroll
                    for (int p = 0; p < PF; ++p) {
                        const int i = i0 + p;
                        if (i < nrb) {
                            const int kc = chunk_of(i, nrb);
                            const int stage = it % STAGES;
                            const uint32_t phase = (it / STAGES) & 1;
                            const uint32_t bytes = (uint32_t)(nrb - kc) * 512u;
                            const uint32_t off = 512u * (uint32_t)(kc * nrb - (kc * (kc - 1)) / 2);
                            mbar_wait(&empty_bar[stage], phase ^ 1);
                            if (pw == 0 && lane == 0) {
                                mbar_arrive_expect_tx(&full_bar[stage], bytes);
                                bulk_g2s(Wst + stage * W_STAGE_BYTES, Wp + off, bytes, &full_bar[stage]);
                            }
                            double2* dstage = reinterpret_cast<double2*>(Dst + stage * D_STAGE_BYTES) + (2 * pw) * 32 + lane;
                            double2 va, vb;   // rows t >= T: y = 0 and mu = 0
                            va.x = ma ? y0[kc * 8] - ring[p][0] : 0.0;
                            va.y = ma ? y0[kc * 8 + 4] - ring[p][1] : 0.0;
                            vb.x = mb ? y1[kc * 8] - ring[p][2] : 0.0;
                            vb.y = mb ? y1[kc * 8 + 4] - ring[p][3] : 0.0;
                            dstage[0] = va;
                            dstage[32] = vb;
                            load_chunk(i + PF, ring[p]);
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&full_bar[stage]);
                            ++it;
                        }
                    }
                }
                sa = sa1; sb = sb1; sa1 = sa2; sb1 = sb2;
            }
            __syncthreads();   // consumers are done with this item; the beat tile may be rewritten
        }
    } else {
        // ===================================== consumers =====================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(HGP_CONSUMER_REGS));
        uint32_t epi = 0;  // running epilogue counter (double-buffers `red`)
        for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int64_t tile = item / m_splits;
            const int m_begin = (int)(item % m_splits) * m_per_item;
            const int m_end = min(M, m_begin + m_per_item);
            const int64_t n0 = tile * BT;
            load_beat_tile(Ytile, Y, N, T, YP, n0, warp, lane);
            __syncthreads();   // tile complete
            for (int m = m_begin; m < m_end; ++m) {
                double acc[4][8][2];
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) acc[j][nt][0] = acc[j][nt][1] = 0.0;
                // epilogue operand fetched early so its latency hides under the chunk loop
                int st_epi = -1;
                if (tid < BT && n0 + tid < N) st_epi = state_of[(n0 + tid) * M + m];

                if (nrb == MAX_NRB) {
                    // T = 256 fast path.  This warp's row blocks rb_0 < rb_1 < rb_2 < rb_3 retire one after the
                    // other as k advances, so the k-loop splits into phases with a FIXED active set.
                    const int rb0 = warp, rb1 = 15 - warp, rb2 = 16 + warp, rb3 = 31 - warp;
                    int kc = 0;
#pragma unroll 1
                    for (; kc <= rb0; ++kc, ++it) fast_chunk<0>(acc, it, kc, rb0, rb1, rb2, rb3, lane, Wst, Dst, full_bar, empty_bar);
#pragma unroll 1
                    for (; kc <= rb1; ++kc, ++it) fast_chunk<1>(acc, it, kc, rb0, rb1, rb2, rb3, lane, Wst, Dst, full_bar, empty_bar);
#pragma unroll 1
                    for (; kc <= rb2; ++kc, ++it) fast_chunk<2>(acc, it, kc, rb0, rb1, rb2, rb3, lane, Wst, Dst, full_bar, empty_bar);
#pragma unroll 1
                    for (; kc <= rb3; ++kc, ++it) fast_chunk<3>(acc, it, kc, rb0, rb1, rb2, rb3, lane, Wst, Dst, full_bar, empty_bar);
#pragma unroll 1
                    for (; kc < MAX_NRB; ++kc, ++it) fast_chunk<4>(acc, it, kc, rb0, rb1, rb2, rb3, lane, Wst, Dst, full_bar, empty_bar);
                } else {
                    // generic path (T < 256): row blocks tested per chunk
                    for (int kc = 0; kc < nrb; ++kc, ++it) {
                        const int stage = it & (STAGES - 1);
                        mbar_wait(&full_bar[stage], (it / STAGES) & 1);
                        if (31 - warp >= kc) {   // rb_3 = 31 - w is this warp's largest block
                            const double2* ds = reinterpret_cast<const double2*>(Dst + stage * D_STAGE_BYTES) + lane;
                            double2 b[8];
#pragma unroll
                            for (int nt = 0; nt < 8; ++nt) b[nt] = ds[nt * 32];
                            const double2* ws = reinterpret_cast<const double2*>(Wst + stage * W_STAGE_BYTES) + lane;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int rb = (j & 1) ? (8 * j + 7 - warp) : (8 * j + warp);
                                // a real branch (16 tensor instructions per body): predicated-off DMMAs would
                                // still occupy the FP64 tensor pipe and forfeit the triangular saving
                                if (rb >= kc && rb < nrb) {
                                    const double2 a = ws[(rb - kc) * 32];
#pragma unroll
                                    for (int nt = 0; nt < 8; ++nt) dmma884(acc[j][nt][0], acc[j][nt][1], a.x, b[nt].x);
#pragma unroll
                                    for (int nt = 0; nt < 8; ++nt) dmma884(acc[j][nt][0], acc[j][nt][1], a.y, b[nt].y);
                                }
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty_bar[stage]);
                    }
                }
                // ---- epilogue: |z|^2 per beat ----
                double* rbuf = red + (epi & 1) * (NCW * BT);
                ++epi;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        double v = 0.0;
#pragma unroll
                        for (int j = 0; j < 4; ++j) v += acc[j][nt][e] * acc[j][nt][e];
                        v += __shfl_xor_sync(0xffffffffu, v, 4);
                        v += __shfl_xor_sync(0xffffffffu, v, 8);
                        v += __shfl_xor_sync(0xffffffffu, v, 16);
                        if (lane < 4) rbuf[warp * BT + nt * 8 + 2 * lane + e] = v;
                    }
                }
                consumer_bar();
                if (tid < BT) {
                    const int64_t n = n0 + tid;
                    if (n < N) {
                        double s = 0.0;
#pragma unroll
                        for (int w = 0; w < NCW; ++w) s += rbuf[w * BT + tid];
                        q[n * M + m] = (st_epi >= 0) ? (-0.5 * s - half_T_log2pi) : 0.0;
                    }
                }
            }
            __syncthreads();   // matches the producers' end-of-item barrier
        }
    }
}

// ------------------------------------------------------------------------------------------
// generic pair kernel: one warp per (n, m) pair, arbitrary factor per state
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
score_pairs_kernel(const double* __restrict__ Y, int64_t N, int T, const double* __restrict__ mu,
                   const double* __restrict__ W, const int* __restrict__ state_of,
                   const int* __restrict__ factor_of_state, int M, const int* __restrict__ pair_n,
                   const int* __restrict__ pair_m, int64_t n_pairs, double* __restrict__ q) {
    extern __shared__ double dsm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* d = dsm + warp * T;
    const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; p < n_pairs; p += wstride) {
        int64_t n;
        int m;
        if (pair_n) { n = pair_n[p]; m = pair_m[p]; }
        else { n = p / M; m = (int)(p % M); }
        const int s = state_of[n * M + m];
        if (s < 0) {
            if (lane == 0) q[n * M + m] = 0.0;
            continue;
        }
        const int64_t f = factor_of_state ? factor_of_state[s] : s;
        const double* Wf = W + f * (int64_t)T * T;
        const double* yrow = Y + n * T;
        const double* mrow = mu + (int64_t)s * T;
        __syncwarp();
        for (int t = lane; t < T; t += 32) d[t] = yrow[t] - mrow[t];
        __syncwarp();
        double acc = 0.0;
        for (int r = 0; r < T; ++r) {
            const double* wr = Wf + (int64_t)r * T;
            double part = 0.0;
            for (int k = lane; k <= r; k += 32) part += wr[k] * d[k];
            part = warp_sum(part);
            acc += part * part;
        }
        if (lane == 0) q[n * M + m] = -0.5 * acc - 0.5 * (double)T * HGP_LOG2PI;
    }
}

// ------------------------------------------------------------------------------------------
// SNR statistic: one warp per beat, looping clusters; smoothed means gathered through L1/L2
// ------------------------------------------------------------------------------------------
template <int NREG>   // NREG = ceil(T / 32) <= 8: the beat lives in registers
__global__ void __launch_bounds__(256)
snr_states_reg_kernel(const double* __restrict__ Y, int64_t N, int T, const double* __restrict__ mu_sm,
                      const int* __restrict__ snr_state_of, int M, double* __restrict__ snr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; n < N; n += wstride) {
        double y[NREG];
#pragma unroll
        for (int k = 0; k < NREG; ++k) y[k] = (lane + 32 * k < T) ? Y[n * T + lane + 32 * k] : 0.0;
        for (int m0 = 0; m0 < M; m0 += 2) {
            // two clusters per step: 2 * NREG independent gathers in flight per lane
            const int s0 = snr_state_of[n * M + m0];
            const int s1 = (m0 + 1 < M) ? snr_state_of[n * M + m0 + 1] : -1;
            const double* r0 = mu_sm + (int64_t)max(s0, 0) * T;
            const double* r1 = mu_sm + (int64_t)max(s1, 0) * T;
            double v0[NREG], v1[NREG];
#pragma unroll
            for (int k = 0; k < NREG; ++k) {
                const int t = lane + 32 * k;
                v0[k] = (s0 >= 0 && t < T) ? __ldg(r0 + t) : 0.0;
                v1[k] = (s1 >= 0 && t < T) ? __ldg(r1 + t) : 0.0;
            }
            double sig0 = 0.0, noi0 = 0.0, sig1 = 0.0, noi1 = 0.0;
#pragma unroll
            for (int k = 0; k < NREG; ++k) {
                const bool in = lane + 32 * k < T;
                const double d0 = in ? v0[k] - y[k] : 0.0, d1 = in ? v1[k] - y[k] : 0.0;
                sig0 += v0[k] * v0[k]; noi0 += d0 * d0;
                sig1 += v1[k] * v1[k]; noi1 += d1 * d1;
            }
            sig0 = warp_sum(sig0); noi0 = warp_sum(noi0);
            sig1 = warp_sum(sig1); noi1 = warp_sum(noi1);
            if (lane == 0) {
                snr[n * M + m0] = (s0 >= 0) ? 10.0 * log10((sig0 + HGP_EPS) / (noi0 + HGP_EPS)) : 0.0;
                if (m0 + 1 < M) snr[n * M + m0 + 1] = (s1 >= 0) ? 10.0 * log10((sig1 + HGP_EPS) / (noi1 + HGP_EPS)) : 0.0;
            }
        }
    }
}

__global__ void __launch_bounds__(256)
snr_states_kernel(const double* __restrict__ Y, int64_t N, int T, const double* __restrict__ mu_sm,
                  const int* __restrict__ snr_state_of, int M, double* __restrict__ snr) {
    extern __shared__ double ysm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* y = ysm + warp * T;
    const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; n < N; n += wstride) {
        __syncwarp();
        for (int t = lane; t < T; t += 32) y[t] = Y[n * T + t];
        __syncwarp();
        for (int m = 0; m < M; ++m) {
            const int s = snr_state_of[n * M + m];
            double sig = 0.0, noi = 0.0;
            if (s >= 0) {
                const double* mr = mu_sm + (int64_t)s * T;
                for (int t = lane; t < T; t += 32) {
                    const double mv = mr[t];
                    const double dv = mv - y[t];
                    sig += mv * mv;
                    noi += dv * dv;
                }
            }
            sig = warp_sum(sig);
            noi = warp_sum(noi);
            if (lane == 0) snr[n * M + m] = (s >= 0) ? 10.0 * log10((sig + HGP_EPS) / (noi + HGP_EPS)) : 0.0;
        }
    }
}

}  // namespace

extern "C" int hgp_score_tiles(const double* Y, int64_t N, int T, const double* mu, const double* Wpacked,
                               const int* state_of, const int* factor_of_cluster, int M, double* q,
                               const double* mu_sm, const int* snr_state_of, double* snr, void* stream) {
    HGP_REQUIRE(N >= 0 && M >= 0, "hgp_score_tiles: bad sizes");
    HGP_REQUIRE(snr == nullptr || (mu_sm != nullptr && snr_state_of != nullptr), "hgp_score_tiles: snr needs mu_sm and snr_state_of");
    if (T <= 0 || T > 256) { hgp_set_error("hgp_score_tiles: need 0 < T <= 256 (got %d)", T); return HGP_E_UNSUPPORTED; }
    if (N == 0 || M == 0) return 0;
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (n_sm <= 0) n_sm = 148;
    }
    const int nrb = (T + 7) / 8;
    const TileSmem lay = tile_smem_layout(nrb * 8);
    cudaError_t e = cudaFuncSetAttribute(score_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total);
    if (e != cudaSuccess) return hgp_status(e, "hgp_score_tiles: smem attribute");
    const int64_t n_tiles = (N + BT - 1) / BT;
    // split the cluster range so that the item count is >= ~24 waves of the persistent grid (tail < 4 %)
    int m_splits = 1;
    while (n_tiles * m_splits < 24 * (int64_t)n_sm && m_splits * 2 <= M && M / (m_splits * 2) >= 4) m_splits *= 2;
    const int m_per_item = (M + m_splits - 1) / m_splits;
    m_splits = (M + m_per_item - 1) / m_per_item;
    const int64_t n_items = n_tiles * m_splits;
    const int grid = (int)hgp_min64(n_items, n_sm);
    score_tiles_kernel<<<grid, TILE_THREADS, lay.total, (cudaStream_t)stream>>>(
        Y, N, T, mu, Wpacked, hgp_packed_factor_bytes(T) / 8, state_of, factor_of_cluster, M, m_per_item, m_splits,
        n_items, q, mu_sm, snr_state_of, snr);
    HGP_LAUNCH_CHECK("hgp_score_tiles");
    if (snr) return hgp_snr_states(Y, N, T, mu_sm, snr_state_of, M, snr, stream);
    return 0;
}

extern "C" int hgp_score_pairs(const double* Y, int64_t N, int T, const double* mu, const double* W,
                               const int* state_of, const int* factor_of_state, int M, const int* pair_n,
                               const int* pair_m, int64_t n_pairs, double* q, void* stream) {
    HGP_REQUIRE(N >= 0 && T > 0 && M >= 0 && n_pairs >= 0, "hgp_score_pairs: bad sizes");
    HGP_REQUIRE(T <= 1024, "hgp_score_pairs: need T <= 1024");
    HGP_REQUIRE((pair_n == nullptr) == (pair_m == nullptr), "hgp_score_pairs: pair_n/pair_m must both be given");
    if (n_pairs == 0) return 0;
    const int warps = 8;
    size_t smem = sizeof(double) * warps * T;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(score_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return hgp_status(e, "hgp_score_pairs: smem attribute");
    }
    int blocks = (int)hgp_min64((n_pairs + warps - 1) / warps, 148 * 8);
    score_pairs_kernel<<<blocks, warps * 32, smem, (cudaStream_t)stream>>>(Y, N, T, mu, W, state_of, factor_of_state,
                                                                          M, pair_n, pair_m, n_pairs, q);
    HGP_LAUNCH_CHECK("hgp_score_pairs");
    return 0;
}

extern "C" int hgp_snr_states(const double* Y, int64_t N, int T, const double* mu_sm, const int* snr_state_of, int M,
                              double* snr, void* stream) {
    HGP_REQUIRE(N >= 0 && T > 0 && M >= 0 && T <= 1024, "hgp_snr_states: bad sizes");
    if (N == 0 || M == 0) return 0;
    const int warps = 8;
    size_t smem = sizeof(double) * warps * T;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(snr_states_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return hgp_status(e, "hgp_snr_states: smem attribute");
    }
    int blocks = (int)hgp_min64((N + warps - 1) / warps, 148 * 8);
    if (T <= 128) {
        snr_states_reg_kernel<4><<<blocks, warps * 32, 0, (cudaStream_t)stream>>>(Y, N, T, mu_sm, snr_state_of, M, snr);
    } else if (T <= 256) {
        snr_states_reg_kernel<8><<<blocks, warps * 32, 0, (cudaStream_t)stream>>>(Y, N, T, mu_sm, snr_state_of, M, snr);
    } else {
        snr_states_kernel<<<blocks, warps * 32, smem, (cudaStream_t)stream>>>(Y, N, T, mu_sm, snr_state_of, M, snr);
    }
    HGP_LAUNCH_CHECK("hgp_snr_states");
    return 0;
}
