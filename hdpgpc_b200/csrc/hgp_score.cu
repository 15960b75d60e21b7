// Emission scores of the E-step (reference GPI_model.compute_sq_err_all, GPI_model.py:488-547;
// log_sq_error :250-286; _gaussian_score_shared_cov :92-113) and the SNR lead statistic
// (GPI_HDP.compute_snr, GPI_HDP.py:732-748), hand-written for sm_100a.
//
// score_tiles_kernel -- the hot kernel.  For a tile of 64 consecutive beats and a run of clusters it evaluates
//     z = W_m y_n - nu_{s(n,m)},   nu_s = W_m mu_s  (whitened state means, hgp_whiten_means),   q = -|z|^2/2 - T log(2 pi)/2
// as a lower-triangular matrix product on the FP64 tensor cores (DMMA.8x8x4) with the mean subtracted in the
// epilogue.  Why not z = W (y - mu): building D = Y - mu per (tile, cluster) needs 16384 scalar FP64 subtractions per
// item, and scalar FP64 instructions share the pipe with DMMA -- each one queues ~150 cycles behind the tensor
// stream and the producer warps fell behind (measured: 13 % of the kernel).  With the product taken on y alone the
// B operand is the SAME for every cluster, so the beat tile is stored once per item directly in B-fragment order
// and nothing but the factors W_m moves through the pipeline; the subtraction shrinks to 64 per lane per item
// against a per-item vector nu (all beats of a tile almost always score against the same state of a cluster,
// hgp_tile_uniform_states) or, for the few (tile, cluster) items that contain members of the cluster, a gathered
// nu row per beat.  |W y| is 1e2-1e3 times |z|, so the cancellation costs ~3 of the 16 digits: q agrees with the
// reference's cholesky_solve form to ~1e-13 relative (tests: 1e-8 required).
//   * producer warpgroup: an elected thread streams W_m in fragment-ordered k-chunks with 1-D bulk async copies (TMA,
//     completion on an mbarrier) through a 5-stage shared-memory ring; one stage carries TWO k-chunks, s and
//     nrb-1-s, so that every stage holds the same amount of triangular work;
//   * 8 consumer warps: row blocks of 8 are dealt to warps in a mirrored order (w, 15-w, 16+w, 31-w) so the
//     triangular shrinkage stays balanced over the 4 SM sub-partitions; each warp keeps its 32 x 64 slice of z in
//     registers (64 f64 accumulators per lane).
#include "hgp_common.cuh"
#include <stdlib.h>

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
// every warp but the one that streams the factors: beat-tile hand-over at the item boundaries
__device__ __forceinline__ void tile_bar() { asm volatile("bar.sync 2, 352;" ::: "memory"); }

#ifdef HGP_NO_MMA   // experiment: everything but the tensor-core instruction (operands stay live)
__device__ __forceinline__ void tile_mma(double& c0, double& c1, double a, double b) { asm volatile("" ::"d"(a), "d"(b), "d"(c0), "d"(c1)); }
#else
__device__ __forceinline__ void tile_mma(double& c0, double& c1, double a, double b) { dmma884(c0, c1, a, b); }
#endif

constexpr int BT = 64;            // beats per tile
constexpr int NCW = 8;            // consumer warps
constexpr int NPW = 4;            // producer warpgroup: one elected thread issues the bulk copies; the other warps only hold
                                  // registers for the consumers (a sub-partition's registers are shared by ITS warps: a
                                  // CTA of 9 warps would cap everyone at 168, the donors let the consumers have 232)
#define HGP_CONSUMER_REGS 232
#define HGP_PRODUCER_REGS 40
constexpr int STAGES = 5;
constexpr int MAX_NRB = 32;       // T <= 256
constexpr int TILE_THREADS = (NCW + NPW) * 32;
// One pipeline step carries TWO k-chunks, s and nrb-1-s: the triangular product needs nrb-kc row blocks of chunk kc,
// so the pair always adds up to nrb+1 blocks -- every step has the same tensor-core work and the same stage size,
// instead of steps that shrink to one row block at the end of the k-loop while the per-step cost stays the same.
constexpr int W_STAGE_BYTES = (MAX_NRB + 1) * 512;   // 16.5 KB
constexpr int Y_CHUNK_DOUBLES = 8 * BT;              // one 8 x 64 slice of the beat tile in B-fragment order (4 KB)

struct TileSmem {
    // offsets (bytes) into dynamic shared memory
    int ytile, wst, red, bars, total;
};
__host__ __device__ inline TileSmem tile_smem_layout(int nrb) {
    TileSmem s;
    s.ytile = 0;
    s.wst = nrb * Y_CHUNK_DOUBLES * 8;
    s.red = s.wst + STAGES * W_STAGE_BYTES;
    s.bars = s.red + 2 * NCW * BT * 8 + BT * 4;      // + the tile's state indices (mixed-state epilogue)
    s.total = s.bars + (2 * STAGES + 4) * 8;     // full/empty per stage + epilogue full/free per reduction buffer
    return s;
}

// Beat tile -> shared memory in B-fragment order, zero padded (rows n >= N, samples t >= T), by 11 of the 12 warps.
// Sample t of beat c goes to chunk t/8, double2 slot (c/8)*32 + (c%8)*4 + t%4, component (t%8)/4: the lane that owns
// column c%8 and k-index t%4 of n-tile c/8 finds both of its k-steps in one 16-byte load.
__device__ __forceinline__ void load_beat_tile(double* Yfrag, const double* __restrict__ Y, int64_t N, int T, int nrb,
                                               int64_t n0, int warp, int lane) {
    // loaders: every warp but the factor streamer (warp NCW)
    for (int c = warp < NCW ? warp : warp - 1; c < BT; c += NCW + NPW - 1) {
        const int64_t n = n0 + c;
        const double* src = Y + n * T;
        double* base = Yfrag + ((c >> 3) * 32 + (c & 7) * 4) * 2;
        // all (up to eight) loads of the row first: with a run-time trip count the loop would wait out one memory
        // latency per 32 samples
        double v[MAX_NRB / 4];
#pragma unroll
        for (int k = 0; k < MAX_NRB / 4; ++k) {
            const int t = lane + 32 * k;
            v[k] = (n < N && t < T) ? __ldg(src + t) : 0.0;
        }
#pragma unroll
        for (int k = 0; k < MAX_NRB / 4; ++k) {
            const int t = lane + 32 * k;
            if (t < nrb * 8) base[(t >> 3) * Y_CHUNK_DOUBLES + (t & 3) * 2 + ((t >> 2) & 1)] = v[k];
        }
    }
}

// One pipeline step (k-chunks s and 31-s) of the T = 256 fast path for a warp whose row blocks j >= JA are active in
// chunk s and j >= JB in chunk 31-s: no per-block tests, the fragment loads are issued up front and the tensor-core
// instructions follow as one straight-line stream (64 to 80 DMMAs per step).
template <int JA, int JB>
__device__ __forceinline__ void fast_step(double (&acc)[4][8][2], uint32_t it, int s, int rb0, int rb1, int rb2,
                                          int rb3, int lane, const unsigned char* Wst, const double* Yfrag,
                                          uint64_t* full_bar, uint64_t* empty_bar) {
    const int stage = it % STAGES;
    const double2* ya = reinterpret_cast<const double2*>(Yfrag + s * Y_CHUNK_DOUBLES) + lane;
    double2 b[8];
#ifdef HGP_NO_FRAGLOAD
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) b[nt] = make_double2(1.0 + nt, 2.0 + lane);
    (void)ya;
#else
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) b[nt] = ya[nt * 32];      // the beat fragments do not depend on the pipeline
#endif
#ifndef HGP_NO_BAR
    mbar_wait(&full_bar[stage], (it / STAGES) & 1);
#endif
    const double2* ws = reinterpret_cast<const double2*>(Wst + stage * W_STAGE_BYTES) + lane;
    {
        // chunk s: block rb sits at (rb - s) * 512 bytes
        const double2* wa = ws - s * 32;
        double2 a[4];
#ifdef HGP_NO_FRAGLOAD
        a[0] = a[1] = a[2] = a[3] = make_double2(0.5 * lane, 0.25 * s);
        (void)wa;
#else
        if (JA <= 0) a[0] = wa[rb0 * 32];
        if (JA <= 1) a[1] = wa[rb1 * 32];
        if (JA <= 2) a[2] = wa[rb2 * 32];
        if (JA <= 3) a[3] = wa[rb3 * 32];
#endif
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j >= JA) {
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) tile_mma(acc[j][nt][0], acc[j][nt][1], a[j].x, b[nt].x);
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) tile_mma(acc[j][nt][0], acc[j][nt][1], a[j].y, b[nt].y);
            }
        }
    }
    if (JB < 4) {
        // chunk 31 - s follows the 32 - s blocks of chunk s: block rb sits at (32 - s + rb - (31 - s)) = (rb + 1) * 512
        const double2* yb = reinterpret_cast<const double2*>(Yfrag + (MAX_NRB - 1 - s) * Y_CHUNK_DOUBLES) + lane;
        double2 a[4];
#ifdef HGP_NO_FRAGLOAD
        a[2] = a[3] = make_double2(0.125 * lane, 0.75 * s);
        (void)yb;
#else
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) b[nt] = yb[nt * 32];
        if (JB <= 2) a[2] = ws[(rb2 + 1) * 32];
        if (JB <= 3) a[3] = ws[(rb3 + 1) * 32];
#endif
#pragma unroll
        for (int j = 2; j < 4; ++j) {
            if (j >= JB) {
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) tile_mma(acc[j][nt][0], acc[j][nt][1], a[j].x, b[nt].x);
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) tile_mma(acc[j][nt][0], acc[j][nt][1], a[j].y, b[nt].y);
            }
        }
    }
#ifndef HGP_NO_BAR
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
#endif
}

// Work items of the persistent loop: item < n_coarse is (tile, run of m_per_item clusters).  The items that would form
// the last, partially filled round of the grid are cut into `fine` pieces each, so that the tail costs a fraction of a
// coarse item instead of a whole one (6252 coarse items over 148 SMs: 43 rounds for 42.24 rounds of work).
struct TileItem { int64_t tile; int m_begin, m_end; };
__device__ __forceinline__ TileItem decode_item(int64_t item, int64_t n_coarse, int m_splits, int m_per_item, int fine, int M) {
    int64_t coarse = item;
    int part = 0, parts = 1;
    if (item >= n_coarse) {
        coarse = n_coarse + (item - n_coarse) / fine;
        part = (int)((item - n_coarse) % fine);
        parts = fine;
    }
    TileItem r;
    r.tile = coarse / m_splits;
    const int mb = (int)(coarse % m_splits) * m_per_item;
    const int me = min(M, mb + m_per_item);
    const int len = (me - mb + parts - 1) / parts;
    r.m_begin = min(me, mb + part * len);
    r.m_end = min(me, r.m_begin + len);
    return r;
}

// WHITEN = true turns the same pipeline into the table build's  nu_s = W_f mu_s  (hgp_whiten_means_tiles): the "beats" are
// the rows of the state-mean table, a work item is one (64-state tile, factor f) pair from an explicit list, and the
// epilogue stores z itself -- transposed back to [state][row] -- for the states of the tile that use factor f, instead
// of reducing |z - nu|^2.  The reducer warp and the epilogue barriers stay idle.
struct WhitenArgs {
    const int2* items;                 // (tile, factor) per work item
    const int* factor_of_state;        // [S]
    double* out;                       // [S][T]
};

template <bool WHITEN>
__device__ __forceinline__ TileItem tile_item(int64_t item, int64_t n_coarse, int m_splits, int m_per_item, int fine, int M,
                                              const WhitenArgs& wa) {
    if (WHITEN) {
        const int2 w = wa.items[item];
        TileItem r;
        r.tile = w.x; r.m_begin = w.y; r.m_end = w.y + 1;
        return r;
    }
    return decode_item(item, n_coarse, m_splits, m_per_item, fine, M);
}

template <bool WHITEN>
__global__ void __launch_bounds__(TILE_THREADS, 1)
score_tiles_kernel(const double* __restrict__ Y, int64_t N, int T, const double* __restrict__ nu,
                   const double* __restrict__ Wpacked, int64_t packed_doubles, const int* __restrict__ state_of,
                   const int* __restrict__ tile_state, const int* __restrict__ factor_of_cluster, int M, int m_per_item,
                   int m_splits, int64_t n_items, int64_t n_coarse, int fine, double* __restrict__ q, WhitenArgs wa) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int nrb = (T + 7) / 8;
    const TileSmem lay = tile_smem_layout(nrb);
    double* Yfrag = reinterpret_cast<double*>(smem_raw + lay.ytile);
    unsigned char* Wst = smem_raw + lay.wst;
    double* red = reinterpret_cast<double*>(smem_raw + lay.red);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + lay.bars);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* epi_full = empty_bar + STAGES;     // [2]: all consumer warps have written their partial |z|^2
    uint64_t* epi_free = epi_full + 2;           // [2]: the reducer warp is done with the buffer

    const int tid = threadIdx.x;
    // warp index through a shuffle: tells the compiler it is warp-uniform, so the role / row-block branches
    // need no reconvergence (no WARPSYNC / BSSY around the tensor-core blocks)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);         // the producer's arrive.expect_tx; the bulk copies complete the bytes
            mbar_init(&empty_bar[s], NCW);      // one arrive per consumer warp
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&epi_full[b], NCW);
            mbar_init(&epi_free[b], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t it = 0;   // running pipeline-step counter; identical in producer and consumers
    const double half_T_log2pi = 0.5 * (double)T * HGP_LOG2PI;
    const int n_steps = (nrb + 1) >> 1;

    if (warp >= NCW) {
        // ===================================== producer =====================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(HGP_PRODUCER_REGS));
        if (warp == NCW) {
            // The factor stream does not depend on the beat tile: this warp stays out of the item barriers and keeps
            // the ring full across item boundaries, bounded only by the consumers' empty-barrier arrivals.
            if (lane == 0) {
                for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
                    const TileItem wi = tile_item<WHITEN>(item, n_coarse, m_splits, m_per_item, fine, M, wa);
                    for (int m = wi.m_begin; m < wi.m_end; ++m) {
                        const unsigned char* Wp = reinterpret_cast<const unsigned char*>(
                            Wpacked + (int64_t)(WHITEN ? m : factor_of_cluster[m]) * packed_doubles);
                        for (int st = 0; st < n_steps; ++st, ++it) {
                            const int ka = st, kb = nrb - 1 - st;
                            const bool two = kb > ka;
                            const int stage = it % STAGES;
                            const uint32_t bytes_a = (uint32_t)(nrb - ka) * 512u;
                            const uint32_t bytes_b = two ? (uint32_t)(nrb - kb) * 512u : 0u;
                            const uint32_t off_a = 512u * (uint32_t)(ka * nrb - (ka * (ka - 1)) / 2);
                            const uint32_t off_b = 512u * (uint32_t)(kb * nrb - (kb * (kb - 1)) / 2);
                            mbar_wait(&empty_bar[stage], ((it / STAGES) & 1) ^ 1);
                            mbar_arrive_expect_tx(&full_bar[stage], bytes_a + bytes_b);
                            bulk_g2s(Wst + stage * W_STAGE_BYTES, Wp + off_a, bytes_a, &full_bar[stage]);
                            if (two) bulk_g2s(Wst + stage * W_STAGE_BYTES + bytes_a, Wp + off_b, bytes_b, &full_bar[stage]);
                        }
                    }
                }
            }
            return;
        }
        for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            const TileItem wi = tile_item<WHITEN>(item, n_coarse, m_splits, m_per_item, fine, M, wa);
            const int64_t tile = wi.tile;
            const int m_begin = wi.m_begin, m_end = wi.m_end;
            load_beat_tile(Yfrag, Y, N, T, nrb, tile * BT, warp, lane);
            tile_bar();   // tile complete (every warp but the factor streamer loads it)
            if (!WHITEN && warp == NCW + 1) {
                // Reducer: the consumers never meet at a CTA barrier for the per-beat sum over the eight warps' partial
                // |z|^2 -- they drop their partials into a double-buffered array, arrive on an mbarrier and move on to
                // the next cluster; this (otherwise idle, register-donating) warp adds them up in a fixed order and
                // writes the scores.  Without the barrier the two consumer warps of a sub-partition are free to drift
                // apart, so their epilogues stop coinciding and the tensor pipe keeps working through them.
                const int64_t n0 = tile * BT;
                for (int m = m_begin; m < m_end; ++m, ++it) {
                    const int b = it & 1;
                    const int64_t na = n0 + lane, nb = na + 32;
                    const int sa = (na < N) ? state_of[na * M + m] : -1;
                    const int sb = (nb < N) ? state_of[nb * M + m] : -1;
                    mbar_wait(&epi_full[b], (it >> 1) & 1);
                    const double* rbuf = red + b * (NCW * BT);
                    double ra = 0.0, rb = 0.0;
#pragma unroll
                    for (int w = 0; w < NCW; ++w) { ra += rbuf[w * BT + lane]; rb += rbuf[w * BT + lane + 32]; }
                    if (na < N) q[na * M + m] = (sa >= 0) ? (-0.5 * ra - half_T_log2pi) : 0.0;
                    if (nb < N) q[nb * M + m] = (sb >= 0) ? (-0.5 * rb - half_T_log2pi) : 0.0;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&epi_free[b]);
                }
            }
            tile_bar();   // consumers are done with this item; the beat tile may be rewritten
        }
    } else {
        // ===================================== consumers =====================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(HGP_CONSUMER_REGS));
        uint32_t epi = 0;  // running epilogue counter (double-buffers `red`)
        const int rb_of[4] = {warp, 15 - warp, 16 + warp, 31 - warp};
        const int qrow = lane >> 2;
        for (int64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
            const TileItem wi = tile_item<WHITEN>(item, n_coarse, m_splits, m_per_item, fine, M, wa);
            const int64_t tile = wi.tile;
            const int m_begin = wi.m_begin, m_end = wi.m_end;
            const int64_t n0 = tile * BT;
            load_beat_tile(Yfrag, Y, N, T, nrb, n0, warp, lane);
            tile_bar();   // tile complete
            for (int m = m_begin; m < m_end; ++m) {
                double acc[4][8][2];
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) acc[j][nt][0] = acc[j][nt][1] = 0.0;
                // epilogue operands fetched early so their latency hides under the k-loop:
                // the tile's uniform state of this cluster (-1: empty cluster, -2: several states) and its nu rows
                const int st_u = WHITEN ? -1 : tile_state[tile * M + m];
                double nu_r[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int row = rb_of[j] * 8 + qrow;
                    nu_r[j] = (st_u >= 0 && row < T) ? __ldg(nu + (int64_t)st_u * T + row) : 0.0;
                }
                int st_epi = -1;
                if (!WHITEN && st_u == -2 && tid < BT && n0 + tid < N) st_epi = state_of[(n0 + tid) * M + m];   // mixed-state epilogue only

                if (nrb == MAX_NRB) {
                    // T = 256 fast path.  This warp's row blocks are rb_0 = w < rb_1 = 15-w < rb_2 = 16+w < rb_3 = 31-w;
                    // in step s (chunks s and 31-s) chunk s needs the blocks >= s and chunk 31-s the blocks >= 31-s,
                    // which gives five phases with a FIXED active set each (4 or 5 (block, chunk) products per step).
                    const int rb0 = warp, rb1 = 15 - warp, rb2 = 16 + warp, rb3 = 31 - warp;
                    int s = 0;
#pragma unroll 1
                    for (; s < rb0; ++s, ++it) fast_step<0, 4>(acc, it, s, rb0, rb1, rb2, rb3, lane, Wst, Yfrag, full_bar, empty_bar);
                    fast_step<0, 3>(acc, it, s, rb0, rb1, rb2, rb3, lane, Wst, Yfrag, full_bar, empty_bar);   // s = w
                    ++s; ++it;
#pragma unroll 1
                    for (; s < rb1; ++s, ++it) fast_step<1, 3>(acc, it, s, rb0, rb1, rb2, rb3, lane, Wst, Yfrag, full_bar, empty_bar);
                    if (s == rb1) {                                                                          // s = 15 - w
                        fast_step<1, 2>(acc, it, s, rb0, rb1, rb2, rb3, lane, Wst, Yfrag, full_bar, empty_bar);
                        ++s; ++it;
                    }
#pragma unroll 1
                    for (; s < MAX_NRB / 2; ++s, ++it) fast_step<2, 2>(acc, it, s, rb0, rb1, rb2, rb3, lane, Wst, Yfrag, full_bar, empty_bar);
                } else {
                    // generic path (T < 256): row blocks tested per chunk
                    for (int s = 0; s < n_steps; ++s, ++it) {
                        const int stage = it % STAGES;
                        mbar_wait(&full_bar[stage], (it / STAGES) & 1);
                        const int ka = s, kb = nrb - 1 - s;
#pragma unroll 1
                        for (int half = 0; half < 2; ++half) {
                            const int kc = half ? kb : ka;
                            if (half && kb <= ka) break;
                            if (31 - warp < kc) continue;   // rb_3 = 31 - w is this warp's largest block
                            const double2* ys = reinterpret_cast<const double2*>(Yfrag + kc * Y_CHUNK_DOUBLES) + lane;
                            double2 b[8];
#pragma unroll
                            for (int nt = 0; nt < 8; ++nt) b[nt] = ys[nt * 32];
                            const double2* ws = reinterpret_cast<const double2*>(Wst + stage * W_STAGE_BYTES) + lane +
                                                (half ? (nrb - ka) * 32 : 0);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int rb = (j & 1) ? (8 * j + 7 - warp) : (8 * j + warp);
                                // a real branch (16 tensor instructions per body): predicated-off DMMAs would
                                // still occupy the FP64 tensor pipe and forfeit the triangular saving
                                if (rb >= kc && rb < nrb) {
                                    const double2 a = ws[(rb - kc) * 32];
#pragma unroll
                                    for (int nt = 0; nt < 8; ++nt) tile_mma(acc[j][nt][0], acc[j][nt][1], a.x, b[nt].x);
#pragma unroll
                                    for (int nt = 0; nt < 8; ++nt) tile_mma(acc[j][nt][0], acc[j][nt][1], a.y, b[nt].y);
                                }
                            }
                        }
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&empty_bar[stage]);
                    }
                }
                if (WHITEN) {
                    // ---- table build: nu[state][row] = z for the states of the tile that use this factor ----
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int64_t st = n0 + nt * 8 + 2 * (lane & 3) + e;
                            if (st < N && wa.factor_of_state[st] == m) {
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const int row = rb_of[j] * 8 + qrow;
                                    if (row < T) wa.out[st * T + row] = acc[j][nt][e];
                                }
                            }
                        }
                    }
                    continue;
                }
                // ---- epilogue: z = acc - nu, |z|^2 per beat ----
                double* rbuf = red + (epi & 1) * (NCW * BT);
                mbar_wait(&epi_free[epi & 1], ((epi >> 1) & 1) ^ 1);    // the reducer has consumed this buffer's last use
                if (st_u >= -1) {
                    // every beat of the tile scores against the same state: one nu value per row block
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            double r = 0.0;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const double d = acc[j][nt][e] - nu_r[j];
                                r += d * d;
                            }
                            r += __shfl_xor_sync(0xffffffffu, r, 4);
                            r += __shfl_xor_sync(0xffffffffu, r, 8);
                            r += __shfl_xor_sync(0xffffffffu, r, 16);
                            if (lane < 4) rbuf[warp * BT + nt * 8 + 2 * lane + e] = r;
                        }
                    }
                } else {
                    // the tile holds members of this cluster: one nu row per beat, gathered through L2 (16 loads in
                    // flight per lane).  The state indices come from the st_epi loads through shared memory; this
                    // branch is CTA-uniform, so the extra barrier is safe.
                    int* sstate = reinterpret_cast<int*>(red + 2 * NCW * BT);
                    if (tid < BT) sstate[tid] = st_epi;
                    consumer_bar();
#pragma unroll
                    for (int np = 0; np < 4; ++np) {
                        double g[2][2][4];
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int sb = sstate[(2 * np + h) * 8 + 2 * (lane & 3) + e];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const int row = rb_of[j] * 8 + qrow;
                                    g[h][e][j] = (sb >= 0 && row < T) ? __ldg(nu + (int64_t)sb * T + row) : 0.0;
                                }
                            }
#pragma unroll
                        for (int h = 0; h < 2; ++h)
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int nt = 2 * np + h;
                                double r = 0.0;
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const double d = acc[j][nt][e] - g[h][e][j];
                                    r += d * d;
                                }
                                r += __shfl_xor_sync(0xffffffffu, r, 4);
                                r += __shfl_xor_sync(0xffffffffu, r, 8);
                                r += __shfl_xor_sync(0xffffffffu, r, 16);
                                if (lane < 4) rbuf[warp * BT + nt * 8 + 2 * lane + e] = r;
                            }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&epi_full[epi & 1]);
                ++epi;
            }
            tile_bar();   // matches the donor warps' end-of-item barrier
        }
    }
}

// ------------------------------------------------------------------------------------------
// Block path for beats longer than the tile kernel's register-resident 256 rows (256 < T <= 1024): the same product
// z = W y - nu, one CTA per (64-beat tile, cluster) walking the row tiles of W one after the other.  A plain
// shared-memory DMMA tile (64 x 64, K chunks of 16, the CTA tile of hgp_gemm.cuh) with the beats as the B operand read
// straight from their row-major layout (B[k][c] = Y[n0 + c][k]), the triangular K loop cut at the row tile's diagonal,
// and the whitened mean subtracted / squared / summed per beat in registers, so z never goes to memory.  Every
// (tile, cluster) item is computed the same way whatever N is: a sliced sweep stays bitwise equal to the resident one.
constexpr int SB = 64, SBK = 32, SB_PITCH = 36;        // pitch == 4 (mod 8): fragment loads of a half-warp hit 16 banks
constexpr int SB_TILE = SB * SB_PITCH;                   // doubles per operand tile
__host__ __device__ inline size_t sb_smem_bytes() { return sizeof(double) * 4 * SB_TILE; }   // A and B, double-buffered
__global__ void __launch_bounds__(256)
score_blocks_kernel(const double* __restrict__ Y, int64_t N, int T, const double* __restrict__ nu,
                    const double* __restrict__ W, const int* __restrict__ state_of,
                    const int* __restrict__ factor_of_cluster, int M, double* __restrict__ q) {
    extern __shared__ __align__(16) double sb_smem[];    // [2][A tile | B tile]: both stored [row or beat][k]
    __shared__ double s_col[4][SB];
    __shared__ int s_state[SB];
    __shared__ int s_any;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n0 = (int64_t)blockIdx.x * SB;
    const int m = blockIdx.y;
    if (tid == 0) s_any = 0;
    __syncthreads();
    if (tid < SB) {
        const int st = (n0 + tid < N) ? state_of[(n0 + tid) * M + m] : -1;
        s_state[tid] = st;
        if (st >= 0) s_any = 1;
    }
    __syncthreads();
    if (!s_any) {                          // empty cluster (or a tile without a scored beat): q = 0 (GPI_model.py:494-495)
        if (tid < SB && n0 + tid < N) q[(n0 + tid) * M + m] = 0.0;
        return;
    }
    const double* Wm = W + (int64_t)factor_of_cluster[m] * T * T;
    const int wm = warp >> 1, wn = warp & 1;   // warp tile: rows [16 wm, 16 wm + 16), beats [32 wn, 32 wn + 32)
    int sc[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) sc[j][e] = s_state[32 * wn + 8 * j + 2 * (lane & 3) + e];
    double colsum[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) colsum[j][0] = colsum[j][1] = 0.0;
    const int nrt = (T + SB - 1) / SB;
    // thread (ld_r, ld_k): rows / beats ld_r + 8 u, samples ld_k of the 32-wide chunk -- consecutive lanes walk a row
    const int ld_k = tid & 31, ld_r = tid >> 5;
    for (int rt = 0; rt < nrt; ++rt) {
        const int r0 = rt * SB;
        double acc[2][4][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        const int kmax = min(T, r0 + SB);      // W is lower triangular: columns past the row tile's diagonal are zero
        // this warp's 16 rows end at their own diagonal: chunks past it (and row groups past the last row of a short tail
        // tile) multiply zeros only and are skipped behind a warp-uniform branch -- 97 % of the executed products are
        // useful at T = 400 instead of the 72 % of a 64-row granularity
        const int kend_w = (r0 + 16 * wm < T) ? min(kmax, r0 + 16 * wm + 16) : 0;
        // operands of chunk c + 1 travel global -> registers under the DMMAs of chunk c and land in the OTHER shared
        // buffer: one barrier per 32-wide chunk (64 DMMAs per warp between barriers)
        double ra[8], rb[8];
        auto gload = [&](int k0) {
            const int gk = k0 + ld_k;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int gr = r0 + ld_r + 8 * u;
                ra[u] = (gr < T && gk <= gr) ? __ldg(Wm + (int64_t)gr * T + gk) : 0.0;
                const int64_t n = n0 + ld_r + 8 * u;
                rb[u] = (gk < T && n < N) ? __ldg(Y + n * T + gk) : 0.0;
            }
        };
        auto sstore = [&](int buf) {
            double* As = sb_smem + buf * 2 * SB_TILE;
            double* Bs = As + SB_TILE;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                As[(ld_r + 8 * u) * SB_PITCH + ld_k] = ra[u];
                Bs[(ld_r + 8 * u) * SB_PITCH + ld_k] = rb[u];
            }
        };
        __syncthreads();                       // the previous row tile's readers are done with both buffers
        gload(0);
        sstore(0);
        int buf = 0;
        for (int k0 = 0; k0 < kmax; k0 += SBK, buf ^= 1) {
            __syncthreads();                   // chunk k0 is in `buf`; everybody has left buffer buf ^ 1 (chunk k0 - 32)
            const bool more = k0 + SBK < kmax;
            if (more) gload(k0 + SBK);
            if (k0 < kend_w) {
                const double* As = sb_smem + buf * 2 * SB_TILE;
                const double* Bs = As + SB_TILE;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (k0 + 16 * h < kend_w) {          // kend_w is a multiple of 16: a real branch around 32 DMMAs
#pragma unroll
                        for (int ks = 4 * h; ks < 4 * h + 4; ++ks) {
                            double a[2], bf[4];
#pragma unroll
                            for (int i = 0; i < 2; ++i)
                                a[i] = As[(16 * wm + 8 * i + (lane >> 2)) * SB_PITCH + 4 * ks + (lane & 3)];
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                bf[j] = Bs[(32 * wn + 8 * j + (lane >> 2)) * SB_PITCH + 4 * ks + (lane & 3)];
#pragma unroll
                            for (int i = 0; i < 2; ++i)
#pragma unroll
                                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], bf[j]);
                        }
                    }
                }
            }
            if (more) sstore(buf ^ 1);
        }
        // element (row = r0 + 16 wm + 8 i + lane/4, beat = 32 wn + 8 j + 2 (lane%4) + e): z = acc - nu[state][row]
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int gr = r0 + 16 * wm + 8 * i + (lane >> 2);
            if (gr < T) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int st = sc[j][e];
                        if (st >= 0) {
                            const double d = acc[i][j][e] - __ldg(nu + (int64_t)st * T + gr);
                            colsum[j][e] += d * d;
                        }
                    }
            }
        }
    }
    // per-beat sum over the rows: the eight row groups of a warp (shuffles), then the four row warps (fixed order)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double r = colsum[j][e];
            r += __shfl_xor_sync(0xffffffffu, r, 4);
            r += __shfl_xor_sync(0xffffffffu, r, 8);
            r += __shfl_xor_sync(0xffffffffu, r, 16);
            if (lane < 4) s_col[wm][32 * wn + 8 * j + 2 * lane + e] = r;
        }
    __syncthreads();
    if (tid < SB && n0 + tid < N) {
        const double tot = ((s_col[0][tid] + s_col[1][tid]) + s_col[2][tid]) + s_col[3][tid];
        q[(n0 + tid) * M + m] = (s_state[tid] >= 0) ? (-0.5 * tot - 0.5 * (double)T * HGP_LOG2PI) : 0.0;
    }
}

// Uniform state of every (64-beat tile, cluster): the common state index if all beats of the tile (inside [0, N))
// score against the same state of the cluster (-1 = empty cluster included), -2 otherwise.
__global__ void tile_uniform_states_kernel(const int* __restrict__ state_of, int64_t N, int M, int64_t n_tiles,
                                           int* __restrict__ out) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_tiles * M) return;
    const int64_t tile = idx / M;
    const int m = (int)(idx % M);
    const int64_t n0 = tile * BT, n1 = hgp_min64(N, n0 + BT);
    const int s0 = state_of[n0 * M + m];
    int u = s0;
    for (int64_t n = n0 + 1; n < n1; ++n)
        if (state_of[n * M + m] != s0) { u = -2; break; }
    out[idx] = u;
}

// nu[s] = W[factor_of_state[s]] mu[s]  (W lower triangular): one CTA per state, one warp per output row.
__global__ void __launch_bounds__(256)
whiten_means_kernel(const double* __restrict__ mu, const double* __restrict__ W, const int* __restrict__ factor_of_state,
                    const int* __restrict__ state_list, int T, double* __restrict__ nu) {
    extern __shared__ double msm[];
    const int64_t s = state_list ? state_list[blockIdx.x] : blockIdx.x;
    const double* Wf = W + (int64_t)(factor_of_state ? factor_of_state[s] : s) * T * T;
    for (int t = threadIdx.x; t < T; t += blockDim.x) msm[t] = mu[s * T + t];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < T; r += blockDim.x >> 5) {
        const double* wr = Wf + (int64_t)r * T;
        double acc = 0.0;
        for (int k = lane; k <= r; k += 32) acc += wr[k] * msm[k];
        acc = warp_sum(acc);
        if (lane == 0) nu[s * T + r] = acc;
    }
}

// The same product on the tensor cores.  The rows of a state table are grouped by cluster, so 64 consecutive states
// almost always share one factor: nu[s0 .. s0 + 64][:] = Mu_tile W_f^T is then a 64 x T x T lower-triangular product
// (the k loop of column block c stops at its diagonal), Mu rows as the A operand and rows of W as the (transposed) B
// operand of DMMA.8x8x4, the next k chunk prefetched into registers under the current one.  Tiles that straddle a
// cluster boundary (and tables with one factor per state, factor_of_state == NULL) take the row-by-row path.
// A table build re-whitens every state mean of a sweep (cfg4: 100k states x 256^2 per lead -- 6.5 GFLOP, 3.3 ms with one
// warp per output element); as a dense contraction it costs a fraction of a millisecond.
__global__ void __launch_bounds__(256, 3)      // latency-bound k loop (two barriers per 16-wide chunk): three CTAs per SM
whiten_tiles_kernel(const double* __restrict__ mu, const double* __restrict__ W, const int* __restrict__ factor_of_state,
                    int64_t S, int T, double* __restrict__ nu) {
    __shared__ double As[64 * 20];
    __shared__ double Bs[64 * 20];              // [column][k]: stored and read along k, conflict-free like As
    __shared__ int s_run[65], s_nrun;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // last tile first: the table ends with the duplicated first-member states, one factor each -- the slow tile must
    // not be the tail of the launch
    const int64_t s0 = ((int64_t)gridDim.x - 1 - blockIdx.x) * 64;
    const int ns = (int)hgp_min64(64, S - s0);
    if (tid == 0) {                             // runs of equal factor inside the tile
        int n = 0;
        s_run[0] = 0;
        for (int i = 1; i < ns; ++i)
            if (factor_of_state[s0 + i] != factor_of_state[s0 + i - 1]) s_run[++n] = i;
        s_run[++n] = ns;
        s_nrun = n;
    }
    __syncthreads();
    const int nrun = s_nrun;
    if (nrun > 4) {
        // (nearly) one factor per state: one warp per state, four output rows in flight
        for (int sl = warp; sl < ns; sl += 8) {
            const int64_t st = s0 + sl;
            const double* Wf = W + (int64_t)factor_of_state[st] * T * T;
            const double* m = mu + st * T;
            for (int r0 = 0; r0 < T; r0 += 4) {
                double acc[4] = {0.0, 0.0, 0.0, 0.0};
                for (int k = lane; k <= min(T - 1, r0 + 3); k += 32) {
                    const double mv = m[k];
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (r0 + q < T && k <= r0 + q) acc[q] += Wf[(int64_t)(r0 + q) * T + k] * mv;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
                const double v = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
                if (lane < 4 && r0 + lane < T) nu[st * T + r0 + lane] = v;
            }
        }
        return;
    }
    const int wm = warp >> 1, wn = warp & 1;
    const int nct = (T + 63) / 64;
    double pa[4], pb[4];
    for (int run = 0; run < nrun; ++run) {
        const int ra = s_run[run], rb = s_run[run + 1];       // states [ra, rb) of the tile share one factor
        const double* Wf = W + (int64_t)factor_of_state[s0 + ra] * T * T;
        for (int ct = 0; ct < nct; ++ct) {
            const int c0 = ct * 64;
            const int kend = min(T, c0 + 64);            // W[r][k] = 0 for k > r, r < c0 + 64
            auto fetch = [&](int k0) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = tid + u * 256;
                    const int r = idx >> 4, k = idx & 15;                  // A: 64 states x 16 k
                    pa[u] = (r >= ra && r < rb && k0 + k < kend) ? mu[(s0 + r) * T + k0 + k] : 0.0;
                    const int cc = idx >> 4, kb = idx & 15;                // B[kb][cc] = W[c0 + cc][k0 + kb]
                    pb[u] = (c0 + cc < T && k0 + kb < kend) ? Wf[(int64_t)(c0 + cc) * T + k0 + kb] : 0.0;
                }
            };
            double acc[2][4][2];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            fetch(0);
            for (int k0 = 0; k0 < kend; k0 += 16) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = tid + u * 256;
                    As[(idx >> 4) * 20 + (idx & 15)] = pa[u];
                    Bs[(idx >> 4) * 20 + (idx & 15)] = pb[u];
                }
                __syncthreads();
                if (k0 + 16 < kend) fetch(k0 + 16);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    double a[2], bf[4];
#pragma unroll
                    for (int i = 0; i < 2; ++i) a[i] = As[(16 * wm + 8 * i + (lane >> 2)) * 20 + 4 * ks + (lane & 3)];
#pragma unroll
                    for (int j = 0; j < 4; ++j) bf[j] = Bs[(32 * wn + 8 * j + (lane >> 2)) * 20 + 4 * ks + (lane & 3)];
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
                        const int r = 16 * wm + 8 * i + (lane >> 2), cc = c0 + 32 * wn + 8 * j + 2 * (lane & 3) + e;
                        if (r >= ra && r < rb && cc < T) nu[(s0 + r) * T + cc] = acc[i][j][e];
                    }
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
    // One CTA per pair (persistent over pairs): the residual d = y - mu sits in shared memory, the eight warps take the
    // rows r = w, w + 8, ... of the triangular product, two rows per trip so that two row loads are in flight, and the
    // eight partial sums are added in a fixed order.  (One warp per pair left a short exception list -- the 64 `first`
    // pairs of a cfg4 lead plane -- latency-bound at 0.35 ms; a CTA per pair is 8x shorter and moves the same bytes
    // when every state has its own factor.)
    extern __shared__ double dsm[];
    __shared__ double s_part[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* d = dsm;
    for (int64_t p = blockIdx.x; p < n_pairs; p += gridDim.x) {
        int64_t n;
        int m;
        if (pair_n) { n = pair_n[p]; m = pair_m[p]; }
        else { n = p / M; m = (int)(p % M); }
        const int s = state_of[n * M + m];
        if (s < 0) {
            if (threadIdx.x == 0) q[n * M + m] = 0.0;
            continue;
        }
        const int64_t f = factor_of_state ? factor_of_state[s] : s;
        const double* Wf = W + f * (int64_t)T * T;
        const double* yrow = Y + n * T;
        const double* mrow = mu + (int64_t)s * T;
        __syncthreads();            // previous pair's readers of d / s_part are done
        for (int t = threadIdx.x; t < T; t += blockDim.x) d[t] = yrow[t] - mrow[t];
        __syncthreads();
        double acc = 0.0;
        for (int r = warp; r < T; r += 16) {
            const int r2 = r + 8;
            const double* w0 = Wf + (int64_t)r * T;
            const double* w1 = Wf + (int64_t)r2 * T;
            double p0 = 0.0, p1 = 0.0;
            for (int k = lane; k <= r; k += 32) p0 += w0[k] * d[k];
            if (r2 < T)
                for (int k = lane; k <= r2; k += 32) p1 += w1[k] * d[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                p0 += __shfl_xor_sync(0xffffffffu, p0, o);
                p1 += __shfl_xor_sync(0xffffffffu, p1, o);
            }
            acc += p0 * p0 + p1 * p1;
        }
        if (lane == 0) s_part[warp] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += s_part[w];
            q[n * M + m] = -0.5 * tot - 0.5 * (double)T * HGP_LOG2PI;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Per-state covariances (estimation_limit = None, regime R2): pairs GROUPED BY FACTOR.  The pair kernel above reads a
// whole factor per (beat, cluster) pair -- 4 T^2 bytes for T^2 + 3T flops -- although all beats between two consecutive
// members of a cluster score against the same state, i.e. the same factor (round 1: 13 M pairs/s at T = 256 while
// pulling 3.5 TB/s of factors through L2).  Here the pair list arrives sorted by factor and cut into chunks of at most
// 32 pairs that share one; a CTA builds the chunk's residuals d = y - mu as the B operand in shared memory (up to four
// 8-pair tiles) and streams the factor's rows ONCE from global memory as DMMA A fragments, 16 loads in flight per
// lane, row tiles dealt to the warps in mirrored order so the triangular k loops balance.  A pair's score does not
// depend on its chunk mates (every output column of the product is computed on its own, in a fixed k order).
constexpr int SG_MAXP = 32;
__host__ __device__ inline int sg_ldd(int T) { return ((T + 7) / 8) * 8 + 4; }

// NT = 8-pair tiles per chunk (2 or 4: chunks of <= 16 or <= 32 pairs).  A DMMA that is merely predicated off still occupies
// the FP64 tensor pipe (ncu on the first version: pipe 55 % busy, much of it on tiles without pairs), so the plan picks
// NT = 2 when factors rarely score more than 16 pairs.
template <int NT>
__global__ void __launch_bounds__(256)
score_groups_kernel(const double* __restrict__ Y, int T, const double* __restrict__ mu, const double* __restrict__ W,
                    const int* __restrict__ state_of, const int* __restrict__ factor_of_state, int M,
                    const int* __restrict__ pair_n, const int* __restrict__ pair_m, const int* __restrict__ chunk_start,
                    int64_t n_chunks, double* __restrict__ q) {
    extern __shared__ __align__(16) double sg_smem[];
    __shared__ double s_part[8][8 * NT];
    __shared__ int s_factor, s_pn[8 * NT], s_pm[8 * NT];
    const int LDD = sg_ldd(T);
    double* D = sg_smem;                              // [32][LDD] residuals of the chunk's pairs
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int lr = lane >> 2, lk = lane & 3;
    const int nrt = (T + 7) >> 3;
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int p0 = chunk_start[chunk];
        const int cnt = min(8 * NT, chunk_start[chunk + 1] - p0);
        const int ntl = (cnt + 7) >> 3;
        __syncthreads();                              // the previous chunk's readers of D / s_part are done
        // residuals of the chunk's pairs.  Warp w builds rows w, w + 8, w + 16, w + 24: the index chains (pair -> state ->
        // rows of Y and mu) of its rows are fetched side by side, and the samples travel eight 32-wide slices at a time --
        // all loads of a batch before its first store (a load -> store loop waits out one memory latency per slice)
        {
            const double* yrow[NT];
            const double* mrow[NT];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int c = warp + 8 * j;
                yrow[j] = mrow[j] = nullptr;
                if (c < cnt) {
                    const int n = pair_n[p0 + c], m = pair_m[p0 + c];
                    const int st = state_of[(int64_t)n * M + m];
                    yrow[j] = Y + (int64_t)n * T;
                    mrow[j] = mu + (int64_t)st * T;
                    if (lane == 0) { s_pn[c] = n; s_pm[c] = m; }
                    if (c == 0 && lane == 0) s_factor = factor_of_state ? factor_of_state[st] : st;
                }
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int c = warp + 8 * j;
                if (c >= 8 * ntl) continue;
                double* drow = D + c * LDD;
                for (int t0 = 0; t0 < LDD; t0 += 256) {
                    double yv[8], mv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int t = t0 + 32 * u + lane;
                        const bool live = yrow[j] != nullptr && t < T;
                        yv[u] = live ? __ldg(yrow[j] + t) : 0.0;
                        mv[u] = live ? __ldg(mrow[j] + t) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int t = t0 + 32 * u + lane;
                        if (t < LDD) drow[t] = yv[u] - mv[u];
                    }
                }
            }
        }
        __syncthreads();
        const double* Wf = W + (int64_t)s_factor * T * T;
        double colsum[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) colsum[nt][0] = colsum[nt][1] = 0.0;
        // The warp's (row tile, 64-wide k batch) units form one sequence; the 16 loads of unit u + 1 are issued before the
        // DMMAs of unit u, so the factor streams with two batches in flight per lane instead of one memory latency per batch.
        // Row tiles are dealt in mirrored order (i, i ^ 7 in alternate groups of eight): long and short k loops alternate.
        const int i_end = (nrt + 7) & ~7;
        auto tile_of = [&](int i) { return ((i >> 3) & 1) ? (i ^ 7) : i; };
        auto load_unit = [&](int i, int kk0, double (&av)[16]) {
            const int rt = tile_of(i);
            const int r = 8 * rt + lr;
            const int kend = min(T, 8 * rt + 8);
            const double* wr = Wf + (int64_t)min(r, T - 1) * T;
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int kk = kk0 + 4 * u + lk;
                av[u] = (rt < nrt && r < T && kk < kend) ? __ldg(wr + kk) : 0.0;
            }
        };
        auto advance = [&](int& i, int& kk0) {          // next unit of this warp (i >= i_end: none left)
            kk0 += 64;
            if (tile_of(i) >= nrt || kk0 >= min(T, 8 * tile_of(i) + 8)) { i += 8; kk0 = 0; }
            while (i < i_end && tile_of(i) >= nrt) i += 8;
        };
        int ci = warp, ckk = 0;
        while (ci < i_end && tile_of(ci) >= nrt) ci += 8;
        double avA[16], avB[16];
        if (ci < i_end) load_unit(ci, ckk, avA);
        double acc[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = 0.0;
        auto consume = [&](int i, int kk0, const double (&av)[16]) {
            const int kend = min(T, 8 * tile_of(i) + 8);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const int kk = kk0 + 4 * u;
                if (kk < kend) {
                    const double* bp = D + lr * LDD + kk + lk;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        if (nt < ntl) dmma884(acc[nt][0], acc[nt][1], av[u], bp[nt * 8 * LDD]);
                }
            }
            if (kk0 + 64 >= kend) {                      // last batch of the row tile: z is complete
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    colsum[nt][0] += acc[nt][0] * acc[nt][0];
                    colsum[nt][1] += acc[nt][1] * acc[nt][1];
                    acc[nt][0] = acc[nt][1] = 0.0;
                }
            }
        };
        while (ci < i_end) {
            int ni = ci, nkk = ckk;
            advance(ni, nkk);
            if (ni < i_end) load_unit(ni, nkk, avB);
            consume(ci, ckk, avA);
            if (ni >= i_end) break;
            ci = ni; ckk = nkk;
            advance(ni, nkk);
            if (ni < i_end) load_unit(ni, nkk, avA);
            consume(ci, ckk, avB);
            ci = ni; ckk = nkk;
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double v = colsum[nt][e];
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                if (lane < 4) s_part[warp][nt * 8 + 2 * lane + e] = v;
            }
        __syncthreads();
        if (tid < cnt) {
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += s_part[w][tid];
            q[(int64_t)s_pn[tid] * M + s_pm[tid]] = -0.5 * tot - 0.5 * (double)T * HGP_LOG2PI;
        }
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

// SNR statistic, tile form: consecutive beats almost always score against the SAME state of a cluster (the state index
// only advances when that cluster gains a member), so a (32-beat tile, cluster) pair touches 1 + (#members of the
// cluster inside the tile) distinct rows of the smoothed-mean table.  One CTA keeps the beat tile in shared memory; each
// warp walks its clusters down the tile with the current state row in registers and reloads it (through L2) only when
// the index changes: (M + 32) row reads per tile instead of 32 M.
constexpr int SNR_BT = 32;
template <int NREG>
__global__ void __launch_bounds__(256)
snr_tiles_kernel(const double* __restrict__ Y, int64_t N, int T, const double* __restrict__ mu_sm,
                 const int* __restrict__ snr_state_of, int M, double* __restrict__ snr) {
    extern __shared__ double ytile[];      // [SNR_BT][T]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int64_t n_tiles = (N + SNR_BT - 1) / SNR_BT;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t n0 = tile * SNR_BT;
        const int nb = (int)hgp_min64(SNR_BT, N - n0);
        __syncthreads();
        for (int i = threadIdx.x; i < nb * T; i += blockDim.x) ytile[i] = Y[n0 * T + i];
        __syncthreads();
        for (int m = warp; m < M; m += nwarps) {
            const int my_s = (lane < nb) ? snr_state_of[(n0 + lane) * M + m] : -1;
            int s_cur = -2;
            double v[NREG], sig = 0.0;
            double my_sig = 0.0, my_noi = 0.0;       // lane b keeps the sums of beat b: one log10 per lane, not per beat
            auto reload = [&](int s) {
                s_cur = s;
                const double* r = mu_sm + (int64_t)max(s, 0) * T;
                double p = 0.0;
#pragma unroll
                for (int k = 0; k < NREG; ++k) {
                    const int t = lane + 32 * k;
                    v[k] = (s >= 0 && t < T) ? __ldg(r + t) : 0.0;
                    p += v[k] * v[k];
                }
                sig = warp_sum(p);
            };
            int b = 0;
            while (b < nb) {
              // four beats per trip while they share the state: four independent reduction chains in flight
              bool four = false;
              if (b + 4 <= nb) {
                const int s0 = __shfl_sync(0xffffffffu, my_s, b), s1 = __shfl_sync(0xffffffffu, my_s, b + 1);
                const int s2 = __shfl_sync(0xffffffffu, my_s, b + 2), s3 = __shfl_sync(0xffffffffu, my_s, b + 3);
                four = (s0 == s1 && s0 == s2 && s0 == s3);
                if (four && s0 != s_cur) reload(s0);
              }
              if (four) {
                double n0_ = 0.0, n1_ = 0.0, n2_ = 0.0, n3_ = 0.0;
                const double* y = ytile + b * T;
#pragma unroll
                for (int k = 0; k < NREG; ++k) {
                    const int t = lane + 32 * k;
                    if (t < T) {
                        const double d0 = v[k] - y[t], d1 = v[k] - y[T + t], d2 = v[k] - y[2 * T + t], d3 = v[k] - y[3 * T + t];
                        n0_ += d0 * d0; n1_ += d1 * d1; n2_ += d2 * d2; n3_ += d3 * d3;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    n0_ += __shfl_xor_sync(0xffffffffu, n0_, o); n1_ += __shfl_xor_sync(0xffffffffu, n1_, o);
                    n2_ += __shfl_xor_sync(0xffffffffu, n2_, o); n3_ += __shfl_xor_sync(0xffffffffu, n3_, o);
                }
                if (lane >= b && lane < b + 4) {
                    my_sig = sig;
                    my_noi = lane == b ? n0_ : lane == b + 1 ? n1_ : lane == b + 2 ? n2_ : n3_;
                }
                b += 4;
              } else {
                const int s = __shfl_sync(0xffffffffu, my_s, b);
                if (s != s_cur) reload(s);
                const double* y = ytile + b * T;
                double noi = 0.0;
#pragma unroll
                for (int k = 0; k < NREG; ++k) {
                    const int t = lane + 32 * k;
                    const double d = (t < T) ? v[k] - y[t] : 0.0;
                    noi += d * d;
                }
                noi = warp_sum(noi);
                if (lane == b) { my_sig = sig; my_noi = noi; }
                ++b;
              }
            }
            if (lane < nb)
                snr[(n0 + lane) * M + m] = (my_s >= 0) ? 10.0 * log10((my_sig + HGP_EPS) / (my_noi + HGP_EPS)) : 0.0;
        }
    }
}

// ------------------------------------------------------------------------------------------
// SNR statistic on the tensor cores.  sum (mu - y)^2 = sum mu^2 - 2 mu.y + sum y^2: the cross term of the (at most
// M + 64) distinct state rows a 64-beat tile can meet against the tile's beats is a dense product V Y^T -- the same B
// operand as the score kernel (beat tile in B-fragment order), a gathered, non-triangular A operand.  The scalar
// kernel above spends 77 warp instructions per (beat, cluster) pair, most of them in 32-lane reductions; here a pair
// costs 1/8 of a DMMA plus one log10.
//   * rows: thread m scans cluster m down the tile and emits one row per distinct state (the index only advances
//     at the cluster's own members, so there are at most M + 64 rows);
//   * A operand: streamed in groups of 32 columns; warp w always supplies row w of every 8-row block (a coalesced
//     256-byte read per row), scatters it into A-fragment order in shared memory and accumulates sum mu^2 of its rows
//     on the way (fixed order); the next group is prefetched into registers under the current group's DMMAs;
//   * warp w multiplies row blocks w, w + 8 (, w + 16) against all eight n-tiles;
//   * epilogue: element (row, beat) is the pair (beat, cluster of the row) iff the beat's state for that cluster is
//     the row's state; noise = sum mu^2 - 2 cross + sum y^2, recomputed directly in the rare case where the expansion
//     would cancel more than eight digits (SNR > 80 dB).
constexpr int SNRM_BT = 64;
// doubles behind the beat tile: the four A stages, or what the epilogue's [NROW][65] cross-term exchange needs beyond the
// beat tile it overlays, whichever is larger
__host__ __device__ inline size_t snr_ast_doubles(int nrb, int rbw) {
    const size_t stages = 4 * 8 * (size_t)rbw * 64, tile = (size_t)nrb * 512, exch = 64 * (size_t)rbw * 65;
    return exch > tile + stages ? exch - tile : stages;
}
template <int RBW>   // row blocks per warp: 2 (M <= 64) or 3 (M <= 128)
__global__ void __launch_bounds__(256, 1)
snr_mma_kernel(const double* __restrict__ Y, int64_t N, int T, const double* __restrict__ mu_sm,
               const int* __restrict__ snr_state_of, int M, double* __restrict__ snr) {
    constexpr int NRB = 8 * RBW;            // row blocks per tile
    constexpr int NROW = 8 * NRB;           // rows (padded)
    extern __shared__ __align__(16) unsigned char smraw[];
    const int nrb = (T + 7) / 8;            // k-chunks
    double* Yfrag = reinterpret_cast<double*>(smraw);                    // [nrb][512]
    double* Ast = Yfrag + nrb * Y_CHUNK_DOUBLES;                         // [4][NRB][64] (+ room for the epilogue's exchange)
    double* ysq = Ast + snr_ast_doubles(nrb, RBW);                       // [64]
    double* row_sig = ysq + SNRM_BT;                                     // [NROW]
    int* row_s = reinterpret_cast<int*>(row_sig + NROW);                 // [NROW]
    int* row_m = row_s + NROW;                                           // [NROW]
    int* cnt = row_m + NROW;                                             // [M + 1]
    int* sstate = cnt + ((M + 2) & ~1);                                  // [64][M]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_tiles = (N + SNRM_BT - 1) / SNRM_BT;
    const int n_groups = (nrb + 3) / 4;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t n0 = tile * SNRM_BT;
        const int nb = (int)hgp_min64(SNRM_BT, N - n0);
        __syncthreads();                     // previous tile's readers are done
        // ---- beat tile in B-fragment order + sum y^2 per beat; state indices of the tile
        //      (a warp's eight rows in two passes of four: 32 loads in flight per lane, not one memory latency per row)
#pragma unroll 1
        for (int c0 = warp; c0 < SNRM_BT; c0 += 32) {
            double v[4][8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = c0 + 8 * q;
                const int64_t n = n0 + c;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int t = lane + 32 * k;
                    v[q][k] = (c < nb && t < T) ? __ldg(Y + n * T + t) : 0.0;
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = c0 + 8 * q;
                double* base = Yfrag + ((c >> 3) * 32 + (c & 7) * 4) * 2;
                double p = 0.0;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int t = lane + 32 * k;
                    if (t < nrb * 8) base[(t >> 3) * Y_CHUNK_DOUBLES + (t & 3) * 2 + ((t >> 2) & 1)] = v[q][k];
                    p += v[q][k] * v[q][k];
                }
                p = warp_sum(p);
                if (lane == 0) ysq[c] = p;
            }
        }
        //      state indices: eight loads in flight per thread
        for (int i0 = tid; i0 < SNRM_BT * M; i0 += 256 * 8) {
            int sv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + 256 * u;
                sv[u] = (i < SNRM_BT * M && i / M < nb) ? snr_state_of[n0 * M + i] : -3;   // -3: outside the sequence
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + 256 * u;
                if (i < SNRM_BT * M) sstate[i] = sv[u];
            }
        }
        __syncthreads();
        // ---- rows: one per distinct state of every cluster
        if (tid < M) {
            int c = 0, prev = -2;
            for (int b = 0; b < nb; ++b) {
                const int sb = sstate[b * M + tid];
                if (sb != prev) { ++c; prev = sb; }
            }
            cnt[tid] = c;
        }
        __syncthreads();
        if (warp == 0) {                     // exclusive prefix sum over the clusters (M <= 128: four per lane + a warp scan)
            int c4[4], tot = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int m = 4 * lane + u; c4[u] = (m < M) ? cnt[m] : 0; tot += c4[u]; }
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            int run = incl - tot;
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int m = 4 * lane + u; if (m < M) cnt[m] = run; run += c4[u]; }
            if (lane == 31) cnt[M] = incl;
        }
        __syncthreads();
        const int n_rows = cnt[M];
        if (n_rows > NROW) {
            // An arbitrary snr_state_of can change state more often down a tile than the A operand has rows (the M + 64
            // bound only holds for the reference's rule: an index advances at the cluster's own members).  Such a tile is
            // summed directly, one warp per pair -- block-uniform branch, nothing has been written to row_s / row_m yet.
            for (int p = warp; p < nb * M; p += 8) {
                const int b = p / M, m = p - b * M;
                const int s = sstate[b * M + m];
                double sig = 0.0, noi = 0.0;
                if (s >= 0) {
                    const double* mr = mu_sm + (int64_t)s * T;
                    const double* yr = Y + (n0 + b) * T;
                    for (int t = lane; t < T; t += 32) {
                        const double mv = __ldg(mr + t);
                        const double d = mv - __ldg(yr + t);
                        sig += mv * mv;
                        noi += d * d;
                    }
                }
                sig = warp_sum(sig);
                noi = warp_sum(noi);
                if (lane == 0) snr[(n0 + b) * M + m] = (s >= 0) ? 10.0 * log10((sig + HGP_EPS) / (noi + HGP_EPS)) : 0.0;
            }
            continue;
        }
        if (tid < M) {
            int r = cnt[tid], prev = -2;
            for (int b = 0; b < nb; ++b) {
                const int sb = sstate[b * M + tid];
                if (sb != prev) { row_s[r] = sb; row_m[r] = tid; ++r; prev = sb; }
                sstate[b * M + tid] = r - 1;         // from here on: the ROW of the pair (its state is row_s[row])
            }
        }
        for (int r = n_rows + tid; r < NROW; r += 256) { row_s[r] = -1; row_m[r] = -1; }
        __syncthreads();

        // the next tile's beats start their way into L2 under this tile's product
        if (tile + gridDim.x < n_tiles) {
            const int64_t nn0 = (tile + gridDim.x) * SNRM_BT;
            const char* nxt = reinterpret_cast<const char*>(Y + nn0 * T);
            const int64_t bytes = hgp_min64(SNRM_BT, N - nn0) * (int64_t)T * 8;
            for (int64_t o = (int64_t)tid * 128; o < bytes; o += 256 * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + o));
        }
        // ---- product: acc[j][nt] for row blocks rb = warp + 8 j
        double acc[RBW][8][2];
#pragma unroll
        for (int j = 0; j < RBW; ++j)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) acc[j][nt][0] = acc[j][nt][1] = 0.0;
        double pre[NRB], rsq[NRB];
#pragma unroll
        for (int i = 0; i < NRB; ++i) rsq[i] = 0.0;
        auto fetch = [&](int g) {
            const int col = g * 32 + lane;
#pragma unroll
            for (int i = 0; i < NRB; ++i) {
                if (8 * i >= n_rows) break;          // row blocks past the tile's rows are never multiplied
                const int sr = row_s[warp + 8 * i];
                pre[i] = (sr >= 0 && col < T) ? __ldg(mu_sm + (int64_t)sr * T + col) : 0.0;
            }
        };
        fetch(0);
        for (int g = 0; g < n_groups; ++g) {
            // scatter the prefetched group into A-fragment order: row-in-block = warp, column lane
            {
                const int j = lane >> 3, cc = lane & 7;
                double* dst = Ast + ((j * NRB) * 32 + warp * 4 + (cc & 3)) * 2 + (cc >> 2);
#pragma unroll
                for (int i = 0; i < NRB; ++i) {
                    if (8 * i >= n_rows) break;
                    dst[i * 64] = pre[i];
                    rsq[i] += pre[i] * pre[i];
                }
            }
            __syncthreads();
            if (g + 1 < n_groups) fetch(g + 1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kc = g * 4 + j;
                if (kc < nrb) {
                    const double2* ys = reinterpret_cast<const double2*>(Yfrag + kc * Y_CHUNK_DOUBLES) + lane;
                    double2 b[8];
#pragma unroll
                    for (int nt = 0; nt < 8; ++nt) b[nt] = ys[nt * 32];
#pragma unroll
                    for (int q2 = 0; q2 < RBW; ++q2) {
                        if (8 * (warp + 8 * q2) >= n_rows) break;        // warp-uniform: an empty row block (typically
                                                                         // half of them: M rows + one per member in the tile)
                        const double2 a = reinterpret_cast<const double2*>(Ast)[(j * NRB + warp + 8 * q2) * 32 + lane];
#pragma unroll
                        for (int nt = 0; nt < 8; ++nt) dmma884(acc[q2][nt][0], acc[q2][nt][1], a.x, b[nt].x);
#pragma unroll
                        for (int nt = 0; nt < 8; ++nt) dmma884(acc[q2][nt][0], acc[q2][nt][1], a.y, b[nt].y);
                    }
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < NRB; ++i) {
            const double tot = warp_sum(rsq[i]);
            if (lane == 0) row_sig[warp + 8 * i] = tot;
        }
        __syncthreads();
        // ---- epilogue.  The cross terms leave the registers through shared memory (over the beat tile and the A stages, both
        //      dead now), so that every (beat, cluster) pair is finished by exactly one thread with its neighbours in m: no
        //      divergent search for "which beats of my row block belong to my row", one log10 per pair instead of one
        //      predicated pass per accumulator, and the stores to snr[n][m] are coalesced (ncu: a third of the kernel sat here).
        double* Cs = reinterpret_cast<double*>(smraw);                     // [NROW][65]
#pragma unroll
        for (int j = 0; j < RBW; ++j) {
            const int r = (warp + 8 * j) * 8 + (lane >> 2);
            if (r < n_rows) {
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    Cs[r * 65 + nt * 8 + 2 * (lane & 3)] = acc[j][nt][0];
                    Cs[r * 65 + nt * 8 + 2 * (lane & 3) + 1] = acc[j][nt][1];
                }
            }
        }
        __syncthreads();
        for (int p = tid; p < nb * M; p += 256) {
            const int c = p / M;
            const int r = sstate[p];
            const int sr = row_s[r];
            double out = 0.0;
            if (sr >= 0) {
                const double sig = row_sig[r];
                double noi = sig - 2.0 * Cs[r * 65 + c] + ysq[c];
                if (noi < 1e-8 * (sig + ysq[c])) {
                    // the expansion would lose more than eight digits (the dB value would be off by more than
                    // 4e-8 absolute at 80 dB): direct sum for this pair
                    const double* mr = mu_sm + (int64_t)sr * T;
                    const double* yr = Y + (n0 + c) * T;
                    noi = 0.0;
                    for (int t = 0; t < T; ++t) { const double d = mr[t] - yr[t]; noi += d * d; }
                }
                out = 10.0 * log10((sig + HGP_EPS) / (noi + HGP_EPS));
            }
            snr[n0 * M + p] = out;
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

extern "C" int hgp_tile_beats(void) { return BT; }

extern "C" int hgp_tile_uniform_states(const int* state_of, int64_t N, int M, int* tile_state, void* stream) {
    HGP_REQUIRE(N >= 0 && M >= 0, "hgp_tile_uniform_states: bad sizes");
    if (N == 0 || M == 0) return 0;
    const int64_t n_tiles = (N + BT - 1) / BT;
    const int64_t n = n_tiles * M;
    tile_uniform_states_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(state_of, N, M, n_tiles,
                                                                                            tile_state);
    HGP_LAUNCH_CHECK("hgp_tile_uniform_states");
    return 0;
}

extern "C" int hgp_whiten_means(const double* mu, const double* W, const int* factor_of_state, int64_t S, int T,
                                double* nu, void* stream) {
    HGP_REQUIRE(S >= 0 && T > 0 && T <= 4096, "hgp_whiten_means: bad sizes");
    if (S == 0) return 0;
    HGP_REQUIRE(S < (1ll << 31), "hgp_whiten_means: too many states for one launch");
    if (factor_of_state && S >= 64 && !getenv("HGP_WHITEN_SCALAR"))
        whiten_tiles_kernel<<<(unsigned)((S + 63) / 64), 256, 0, (cudaStream_t)stream>>>(mu, W, factor_of_state, S, T, nu);
    else
        whiten_means_kernel<<<(unsigned)S, 256, sizeof(double) * T, (cudaStream_t)stream>>>(mu, W, factor_of_state, nullptr, T, nu);
    HGP_LAUNCH_CHECK("hgp_whiten_means");
    return 0;
}

// The table build's whitening on the score kernel's own pipeline (T <= 256): `items` lists the (64-state tile, factor)
// pairs to compute -- every state of tile `x` whose factor is `y` receives nu = W_y mu -- and `state_list` the states
// left to the row-by-row kernel (tiles with many factors: the duplicated first-member states at the end of a table).
// Both lists come from the planner on the host side (hdp.LeadTables) and depend on factor_of_state only.
extern "C" int hgp_whiten_means_tiles(const double* mu, const double* W, const double* Wpacked, const int* factor_of_state,
                                      int64_t S, int T, const int* items, int64_t n_items, const int* state_list,
                                      int64_t n_list, double* nu, void* stream) {
    HGP_REQUIRE(S >= 0 && T > 0 && n_items >= 0 && n_list >= 0, "hgp_whiten_means_tiles: bad sizes");
    HGP_REQUIRE(factor_of_state != nullptr, "hgp_whiten_means_tiles: factor_of_state required");
    if (T > 256) { hgp_set_error("hgp_whiten_means_tiles: need T <= 256 (got %d)", T); return HGP_E_UNSUPPORTED; }
    if (n_items > 0) {
        static int n_sm = 0;
        if (n_sm == 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
            if (n_sm <= 0) n_sm = 148;
        }
        const int nrb = (T + 7) / 8;
        const TileSmem lay = tile_smem_layout(nrb);
        cudaError_t e = cudaFuncSetAttribute(score_tiles_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total);
        if (e != cudaSuccess) return hgp_status(e, "hgp_whiten_means_tiles: smem attribute");
        const int grid = (int)hgp_min64(n_items, n_sm);
        score_tiles_kernel<true><<<grid, TILE_THREADS, lay.total, (cudaStream_t)stream>>>(
            mu, S, T, nullptr, Wpacked, hgp_packed_factor_bytes(T) / 8, nullptr, nullptr, nullptr, 0, 1, 1, n_items, n_items, 1,
            nullptr, WhitenArgs{reinterpret_cast<const int2*>(items), factor_of_state, nu});
        HGP_LAUNCH_CHECK("hgp_whiten_means_tiles");
    }
    if (n_list > 0) {
        whiten_means_kernel<<<(unsigned)n_list, 256, sizeof(double) * T, (cudaStream_t)stream>>>(mu, W, factor_of_state,
                                                                                              state_list, T, nu);
        HGP_LAUNCH_CHECK("hgp_whiten_means_tiles: listed states");
    }
    return 0;
}

extern "C" int hgp_score_tiles(const double* Y, int64_t N, int T, const double* nu, const double* Wpacked,
                               const int* state_of, const int* tile_state, const int* factor_of_cluster, int M, double* q,
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
    const TileSmem lay = tile_smem_layout(nrb);
    cudaError_t e = cudaFuncSetAttribute(score_tiles_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lay.total);
    if (e != cudaSuccess) return hgp_status(e, "hgp_score_tiles: smem attribute");
    const int64_t n_tiles = (N + BT - 1) / BT;
    // split the cluster range so that the item count is >= ~24 waves of the persistent grid (tail < 4 %)
    int m_splits = 1;
    while (n_tiles * m_splits < 24 * (int64_t)n_sm && m_splits * 2 <= M && M / (m_splits * 2) >= 4) m_splits *= 2;
    const int m_per_item = (M + m_splits - 1) / m_splits;
    m_splits = (M + m_per_item - 1) / m_per_item;
    const int64_t n_all = n_tiles * m_splits;
    const int grid = (int)hgp_min64(n_all, n_sm);
    // the coarse items of the last, partially filled round are cut into `fine` pieces (see decode_item)
    int fine = 1;
    int64_t n_coarse = n_all;
    const int64_t rem = n_all % grid;
    if (n_all > grid && rem != 0 && m_per_item >= 8) {
        fine = 4;
        n_coarse = n_all - rem;
    }
    const int64_t n_items = n_coarse + (n_all - n_coarse) * fine;
    score_tiles_kernel<false><<<grid, TILE_THREADS, lay.total, (cudaStream_t)stream>>>(
        Y, N, T, nu, Wpacked, hgp_packed_factor_bytes(T) / 8, state_of, tile_state, factor_of_cluster, M, m_per_item,
        m_splits, n_items, n_coarse, fine, q, WhitenArgs{nullptr, nullptr, nullptr});
    HGP_LAUNCH_CHECK("hgp_score_tiles");
    if (snr) return hgp_snr_states(Y, N, T, mu_sm, snr_state_of, M, snr, stream);
    return 0;
}

extern "C" int hgp_score_blocks(const double* Y, int64_t N, int T, const double* nu, const double* W, const int* state_of,
                                const int* factor_of_cluster, int M, double* q, void* stream) {
    HGP_REQUIRE(N >= 0 && M >= 0 && T > 0, "hgp_score_blocks: bad sizes");
    if (M > 65535) { hgp_set_error("hgp_score_blocks: M <= 65535 (got %d)", M); return HGP_E_UNSUPPORTED; }
    if (N == 0 || M == 0) return 0;
    const dim3 grid((unsigned)((N + SB - 1) / SB), (unsigned)M);
    cudaError_t e = cudaFuncSetAttribute(score_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb_smem_bytes());
    if (e != cudaSuccess) return hgp_status(e, "hgp_score_blocks: smem attribute");
    score_blocks_kernel<<<grid, 256, sb_smem_bytes(), (cudaStream_t)stream>>>(Y, N, T, nu, W, state_of, factor_of_cluster, M, q);
    HGP_LAUNCH_CHECK("hgp_score_blocks");
    return 0;
}

extern "C" int hgp_score_pairs(const double* Y, int64_t N, int T, const double* mu, const double* W,
                               const int* state_of, const int* factor_of_state, int M, const int* pair_n,
                               const int* pair_m, int64_t n_pairs, double* q, void* stream) {
    HGP_REQUIRE(N >= 0 && T > 0 && M >= 0 && n_pairs >= 0, "hgp_score_pairs: bad sizes");
    HGP_REQUIRE(T <= 1024, "hgp_score_pairs: need T <= 1024");
    HGP_REQUIRE((pair_n == nullptr) == (pair_m == nullptr), "hgp_score_pairs: pair_n/pair_m must both be given");
    if (n_pairs == 0) return 0;
    const size_t smem = sizeof(double) * T;
    const int blocks = (int)hgp_min64(n_pairs, 148 * 8);
    score_pairs_kernel<<<blocks, 256, smem, (cudaStream_t)stream>>>(Y, N, T, mu, W, state_of, factor_of_state,
                                                                          M, pair_n, pair_m, n_pairs, q);
    HGP_LAUNCH_CHECK("hgp_score_pairs");
    return 0;
}

extern "C" int hgp_score_groups_max_pairs(void) { return SG_MAXP; }

extern "C" int hgp_score_groups(const double* Y, int64_t N, int T, const double* mu, const double* W, const int* state_of,
                                const int* factor_of_state, int M, const int* pair_n, const int* pair_m,
                                const int* chunk_start, int64_t n_chunks, int max_pairs, double* q, void* stream) {
    HGP_REQUIRE(max_pairs == 16 || max_pairs == 32, "hgp_score_groups: max_pairs (pairs per chunk) must be 16 or 32");
    HGP_REQUIRE(N >= 0 && T > 0 && M >= 0 && n_chunks >= 0, "hgp_score_groups: bad sizes");
    HGP_REQUIRE(T <= 512, "hgp_score_groups: need T <= 512");
    HGP_REQUIRE(pair_n != nullptr && pair_m != nullptr && chunk_start != nullptr, "hgp_score_groups: pair lists required");
    if (n_chunks == 0) return 0;
    const size_t smem = sizeof(double) * (size_t)max_pairs * sg_ldd(T);
    auto kern = max_pairs == 16 ? score_groups_kernel<2> : score_groups_kernel<4>;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return hgp_status(e, "hgp_score_groups: smem attribute");
    }
    const int blocks = (int)hgp_min64(n_chunks, 148 * 4);
    kern<<<blocks, 256, smem, (cudaStream_t)stream>>>(Y, T, mu, W, state_of, factor_of_state, M, pair_n, pair_m, chunk_start,
                                                       n_chunks, q);
    HGP_LAUNCH_CHECK("hgp_score_groups");
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
    if (T <= 256 && M <= 128 && M >= 1 && !getenv("HGP_SNR_SCALAR")) {   // never a function of N: a sliced sweep must take the same path
        const int nrbk = (T + 7) / 8;
        const int rbw = M <= 64 ? 2 : 3;
        const int nrow = 64 * rbw;
        const size_t msm = sizeof(double) * ((size_t)nrbk * Y_CHUNK_DOUBLES + snr_ast_doubles(nrbk, rbw) + SNRM_BT + nrow) +
                           sizeof(int) * (2 * (size_t)nrow + ((M + 2) & ~1) + (size_t)SNRM_BT * M) + 16;
        auto kern = rbw == 2 ? snr_mma_kernel<2> : snr_mma_kernel<3>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msm);
        if (e != cudaSuccess) return hgp_status(e, "hgp_snr_states: smem attribute");
        const int mblocks = (int)hgp_min64((N + SNRM_BT - 1) / SNRM_BT, 148);
        kern<<<mblocks, 256, msm, (cudaStream_t)stream>>>(Y, N, T, mu_sm, snr_state_of, M, snr);
    } else if (T <= 256) {
        const size_t tsm = sizeof(double) * SNR_BT * T;          // <= 64 KB: three CTAs per SM
        const int tblocks = (int)hgp_min64((N + SNR_BT - 1) / SNR_BT, 148 * 3);
        auto kern = T <= 128 ? snr_tiles_kernel<4> : snr_tiles_kernel<8>;
        if (tsm > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm);
            if (e != cudaSuccess) return hgp_status(e, "hgp_snr_states: smem attribute");
        }
        kern<<<tblocks, warps * 32, tsm, (cudaStream_t)stream>>>(Y, N, T, mu_sm, snr_state_of, M, snr);
    } else {
        snr_states_kernel<<<blocks, warps * 32, smem, (cudaStream_t)stream>>>(Y, N, T, mu_sm, snr_state_of, M, snr);
    }
    HGP_LAUNCH_CHECK("hgp_snr_states");
    return 0;
}
