// Lead weighting, HMM smoothing, hard responsibilities and sufficient statistics of the E-step
// (reference GPI_HDP.weight_mean GPI_HDP.py:685-701, LogLik :632-661, forward :3546-3610,
// backward :3612-3649, coupled_state_coef :3651-3699, _safe_exp :338-350, counts :890-892).
//
// The forward/backward recursions are sequential in the beat index.  They are parallelised EXACTLY:
// the sequence is cut into chunks, every chunk is first scanned from a guessed boundary message,
// then repaired from its predecessor's true boundary message until the recomputed message is
// bit-identical to the stored one (from there on the stored tail is already exact, because the
// recursion is a deterministic function of the previous message).  Rounds repeat until no chunk
// boundary changes.  With peaked emission likelihoods the chain forgets its start within a few
// beats, so one repair round of a few steps per chunk is the typical cost; the worst case degrades
// to the sequential scan but never to a different answer.
#include "hgp_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// lead weights -> qbar -> e
// ------------------------------------------------------------------------------------------
constexpr int MAX_LEADS = 16;

__global__ void __launch_bounds__(256)
lead_weights_kernel(const double* __restrict__ q, const double* __restrict__ snr, const double* __restrict__ lead_w,
                    int64_t N, int M, int L, double* __restrict__ qbar, double* __restrict__ e,
                    double* __restrict__ wout, int* __restrict__ flags) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int64_t plane = N * (int64_t)M;
    for (int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; n < N; n += wstride) {
        double w[MAX_LEADS];
        if (snr) {
            double mxall = -INFINITY;
            for (int ld = 0; ld < L; ++ld) {
                double mx = -INFINITY;
                for (int m = lane; m < M; m += 32) mx = fmax(mx, snr[ld * plane + n * M + m]);
                mx = warp_max(mx);
                w[ld] = mx;
                mxall = fmax(mxall, mx);
            }
            double den = 0.0;
            for (int ld = 0; ld < L; ++ld) { w[ld] = exp(w[ld] - mxall); den += w[ld]; }
            for (int ld = 0; ld < L; ++ld) w[ld] /= den;
        } else {
            for (int ld = 0; ld < L; ++ld) w[ld] = lead_w[n * L + ld];
        }
        if (wout && lane == 0)
            for (int ld = 0; ld < L; ++ld) wout[n * L + ld] = w[ld];
        double rmax = -INFINITY;
        int has_nan = 0;
        for (int m = lane; m < M; m += 32) {
            double acc = 0.0;
            for (int ld = 0; ld < L; ++ld) acc += q[ld * plane + n * M + m] * w[ld];
            qbar[n * M + m] = acc;
            has_nan |= isnan(acc);
            rmax = fmax(rmax, acc);
        }
        rmax = warp_max(rmax);
        // torch.max propagates NaN (fmax does not): a row with ANY NaN score has a NaN maximum in LogLik / safe_exp
        // (GPI_HDP.py:646, :3577), so every entry of the row becomes nan_to_num(NaN) = 1e-8.  flags bit 1 = row of NaNs.
        has_nan = __any_sync(0xffffffffu, has_nan);
        if (lane == 0 && isinf(rmax)) atomicOr(flags, 1);
        if (lane == 0 && has_nan) atomicOr(flags, 2);
        for (int m = lane; m < M; m += 32) {
            double v = exp(qbar[n * M + m] - rmax);
            if (has_nan || isnan(v)) v = 1e-8;            // torch.nan_to_num(..., 1e-8)
            else if (isinf(v)) v = 1.7976931348623157e308;
            e[n * M + m] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------
// chunked exact scan
// ------------------------------------------------------------------------------------------
#ifndef HGP_HMM_CHUNK
#define HGP_HMM_CHUNK 256
#endif
constexpr int CHUNK = HGP_HMM_CHUNK;   // beats per chunk

// One CTA per chunk, 4*KP threads: thread (k = tid/4, p = tid%4) owns elements [p*KP/4, (p+1)*KP/4)
// of row k of the transition operand in registers.
template <int KP, bool BACKWARD>
__global__ void __launch_bounds__(4 * KP)
hmm_scan_kernel(const double* __restrict__ e, int64_t N, int K, const double* __restrict__ pi,
                const double* __restrict__ Mat, const double* __restrict__ boundary, int has_boundary,
                double* __restrict__ msgs, double* __restrict__ marg, const double* __restrict__ ends_in,
                double* __restrict__ ends_out, int* __restrict__ changed, int repair, int rebase) {
    constexpr int SEG = KP / 4;
    constexpr int NW = (4 * KP + 31) / 32;
    __shared__ __align__(16) double s_prev[KP];      // forward: alpha_{t-1};  backward: u_{t+1} = beta_{t+1} * e_{t+1}
    __shared__ double s_wsum[NW];
    const int tid = threadIdx.x;
    const int k = tid >> 2, p = tid & 3;
    const int lane = tid & 31, warp = tid >> 5;
    const int64_t C = (N + CHUNK - 1) / CHUNK;
    const int64_t c = blockIdx.x;
    const int64_t t0 = c * CHUNK, t1 = hgp_min64(N, t0 + CHUNK);

    // The four p lanes of a row read s_prev segments that lie SEG * 8 bytes apart -- the same banks -- and a 16-byte
    // shared load is served a quarter-warp (two rows x four segments) at a time: read in segment order every load costs
    // four wavefronts per quarter-warp and the scan is bound by the shared-memory pipe (ncu: 16-way conflict on every
    // LDS.128, short-scoreboard stalls on 12 of 19 cycles per issue).  Lane p therefore walks its segment in the order
    // c ^ p (16-byte chunks), which puts the four lanes in different banks; `row` is stored in the same order.
    constexpr int NCH = SEG / 2;                         // 16-byte chunks per segment
    constexpr int SKEW = (NCH >= 4) ? (NCH - 1) : 0;
    double row[SEG];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const int cc = c ^ (p & SKEW);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int col = p * SEG + 2 * cc + h;
            row[2 * c + h] = (k < K && col < K) ? Mat[(int64_t)k * K + col] : 0.0;
        }
    }

    // the chunk whose start is exact by construction
    const bool exact_start = BACKWARD ? (c == C - 1) : (c == 0);
    // rebase: the message ENTERING the slice changed since the stored solution was computed (sharded scan: the
    // neighbour rank's boundary arrived), so the boundary chunk is repaired like any other chunk
    if (repair && exact_start && !rebase) {
        if (tid < K) ends_out[c * K + tid] = ends_in[c * K + tid];
        return;
    }

    int64_t t = BACKWARD ? t1 - 1 : t0;
    const int64_t step = BACKWARD ? -1 : 1;
    int64_t remaining = t1 - t0;
    bool skip_first_matvec = false;   // true for the very first beat of the global sequence

    // ---- starting message ----
    if (tid < KP) {
        double v = 0.0;
        if (tid < K) {
            if (!BACKWARD) {
                if (exact_start) {
                    if (has_boundary) v = boundary[tid];                 // alpha of the previous rank's last beat
                    else v = pi[tid];                                     // handled below (no matvec for t = 0)
                } else if (repair) v = ends_in[(c - 1) * K + tid];
                else v = 1.0 / (double)K;                                 // guess
            } else {
                if (exact_start) {
                    if (has_boundary) v = boundary[tid];                 // (beta (.) e) of the next rank's first beat
                    else v = 0.0;                                         // handled below (beta_{N-1} = 1)
                } else if (repair) v = ends_in[(c + 1) * K + tid];
                else v = e[t1 * K + tid];                                 // guess beta_{t1} = 1  ->  u = e_{t1}
            }
        }
        s_prev[tid] = v;
    }
    if (exact_start && !has_boundary) skip_first_matvec = true;
    __syncthreads();

    bool converged = false;
    double e_next = 0.0;
    if (p == 0 && k < K) e_next = e[t * K + k];
    double last_val = 0.0;   // row owner's last message value (normalised)

    for (; remaining > 0; --remaining, t += step) {
        const double e_t = e_next;
        if (remaining > 1 && p == 0 && k < K) e_next = e[(t + step) * K + k];

        double v;
        if (skip_first_matvec) {
            // forward t = 0: alpha_0 = pi (.) e_0 (GPI_HDP.py:3598); backward t = N-1: beta = 1 (:3642)
            v = BACKWARD ? 1.0 : s_prev[k < KP ? k : 0] * e_t;
            if (k >= K) v = 0.0;
        } else {
            double part = 0.0, part1 = 0.0;              // two chains: the matvec is a latency chain of DFMAs
            const double2* sp = reinterpret_cast<const double2*>(s_prev + p * SEG);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const double2 pv = sp[c ^ (p & SKEW)];
                part += row[2 * c] * pv.x;
                part1 += row[2 * c + 1] * pv.y;
            }
            part += part1;
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            v = BACKWARD ? part : part * e_t;
        }
        double tot = 1.0;
        if (!(BACKWARD && skip_first_matvec)) {
            // forward: sum over all states (:3601); backward: all but the last state (:3646)
            double contrib = (p == 0 && k < (BACKWARD ? K - 1 : K)) ? v : 0.0;
            contrib = warp_sum(contrib);
            if (lane == 0) s_wsum[warp] = contrib;   // last step's readers are behind its closing barrier
            __syncthreads();
            tot = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) tot += s_wsum[w];
            // reciprocal + multiply: v is often far below 1e-37 (peaked responsibilities), which sends the IEEE
            // division of v / tot down its slow path on every step; tot itself is O(1)
            v = v * (1.0 / tot);
        } else {
            __syncthreads();
        }
        skip_first_matvec = false;

        // ---- compare / store ----
        int same = 1;
        if (p == 0 && k < K) {
            if (repair) {
                const double old = msgs[t * K + k];
                same = (__double_as_longlong(old) == __double_as_longlong(v));
                if (!same) msgs[t * K + k] = v;
            } else {
                msgs[t * K + k] = v;
            }
            last_val = v;
            s_prev[k] = BACKWARD ? v * e_t : v;
            if (!BACKWARD && marg && k == 0) marg[t] = tot;   // margPrObs[t] of forward() (:3601)
        }
        if (repair) {
            if (__syncthreads_and(same)) { converged = true; break; }
        } else {
            __syncthreads();
        }
    }

    // ---- boundary message handed to the neighbour chunk ----
    if (p == 0 && k < K) {
        if (repair && converged) {
            ends_out[c * K + k] = ends_in[c * K + k];
        } else {
            const double out = s_prev[k];
            if (repair) {
                const double old = ends_in[c * K + k];
                if (__double_as_longlong(old) != __double_as_longlong(out)) atomicOr(changed, 1);
            }
            ends_out[c * K + k] = out;
        }
    }
    (void)last_val;
}

// ------------------------------------------------------------------------------------------
// The same scan for many states (K > 32): EIGHT chunks per CTA on the FP64 tensor cores.
// One step of one chunk is a K x K matrix-vector product -- 16K multiply-adds behind one another's shuffles and two CTA
// barriers; at K = 128 the one-chunk kernel holds its rows in 98 registers x 512 threads (one CTA per SM) and took 2300
// cycles per step, 9x the FP64 pipe's time for the arithmetic (cfg5: 2 x 16.4 ms per sweep).  Eight chunks side by
// side turn the step into a K x K x 8 product: the transition operand stays in registers as DMMA A fragments (warp w
// owns rows 8w .. 8w+7, K/4 fragments per lane), the eight previous messages are the B operand in shared memory, and
// lane (lr, lk) of warp w ends up with row 8w+lr of chunks 2lk and 2lk+1 -- the barriers, the normalisation and the
// message stores are paid once per eight chunk-steps.  The arithmetic of a step (four DMMA accumulation chains added in a
// fixed order, sums over warps in warp order) depends on nothing but the incoming message, so the fixed-point argument
// of the chunk protocol holds unchanged; which kernel runs depends on K only, never on N, so a sharded scan stays
// bitwise equal to the unsharded one.
template <int KP, bool BACKWARD>
__global__ void __launch_bounds__(4 * KP)
hmm_scan8_kernel(const double* __restrict__ e, int64_t N, int K, const double* __restrict__ pi,
                 const double* __restrict__ Mat, const double* __restrict__ boundary, int has_boundary,
                 double* __restrict__ msgs, double* __restrict__ marg, const double* __restrict__ ends_in,
                 double* __restrict__ ends_out, int* __restrict__ changed, int repair, int rebase) {
    constexpr int NWARP = KP / 8;
    constexpr int LDP = KP + 4;                       // == 4 (mod 8): the B-fragment loads of a half-warp hit 16 banks
    constexpr int NKS = KP / 4;
    __shared__ __align__(16) double s_prev[8 * LDP];  // forward: alpha_{t-1};  backward: u_{t+1} = beta_{t+1} * e_{t+1}
    __shared__ __align__(16) double s_wsum[NWARP * 8];
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int lr = lane >> 2, lk = lane & 3;
    const int row = 8 * warp + lr;
    const int64_t C = (N + CHUNK - 1) / CHUNK;
    const int64_t cbase = (int64_t)blockIdx.x * 8;

    double a[NKS];
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) {
        const int col = 4 * ks + lk;
        a[ks] = (row < K && col < K) ? Mat[(int64_t)row * K + col] : 0.0;
    }

    // ---- the eight chunks: starting messages (all threads) and this lane's two chunks ----
    int maxlen = 0;
    for (int n = 0; n < 8; ++n) {
        const int64_t c = cbase + n;
        if (c >= C) break;
        const bool exact = BACKWARD ? (c == C - 1) : (c == 0);
        if (repair && exact && !rebase) continue;                        // nothing to repair: ends copied below
        const int64_t t0 = c * CHUNK, t1 = hgp_min64(N, t0 + CHUNK);
        maxlen = max(maxlen, (int)(t1 - t0));
    }
    for (int idx = tid; idx < 8 * KP; idx += 4 * KP) {
        const int n = idx / KP, k = idx - n * KP;
        const int64_t c = cbase + n;
        double v = 0.0;
        if (c < C && k < K) {
            const bool exact = BACKWARD ? (c == C - 1) : (c == 0);
            const int64_t t1 = hgp_min64(N, c * CHUNK + CHUNK);
            if (!BACKWARD) {
                if (exact) v = has_boundary ? boundary[k] : pi[k];
                else if (repair) v = ends_in[(c - 1) * K + k];
                else v = 1.0 / (double)K;                                 // guess
            } else {
                if (exact) v = has_boundary ? boundary[k] : 0.0;
                else if (repair) v = ends_in[(c + 1) * K + k];
                else v = e[t1 * K + k];                                   // guess beta_{t1} = 1  ->  u = e_{t1}
            }
        }
        s_prev[n * LDP + k] = v;
    }
    int len[2];
    int64_t tcur[2];
    bool first_plain[2];        // the very first beat of the global sequence: no matrix product (GPI_HDP.py:3598, :3642)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int64_t c = cbase + 2 * lk + j;
        len[j] = 0; tcur[j] = 0; first_plain[j] = false;
        if (c < C) {
            const bool exact = BACKWARD ? (c == C - 1) : (c == 0);
            const int64_t t0 = c * CHUNK, t1 = hgp_min64(N, t0 + CHUNK);
            if (!(repair && exact && !rebase)) len[j] = (int)(t1 - t0);
            tcur[j] = BACKWARD ? t1 - 1 : t0;
            first_plain[j] = exact && !has_boundary;
        }
    }
    const int64_t tstep = BACKWARD ? -1 : 1;
    __syncthreads();

    double e_next[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) e_next[j] = (len[j] > 0 && row < K) ? e[tcur[j] * K + row] : 0.0;

    int stop_step = maxlen;      // repair: the step at which every chunk still running reproduced its stored message
    for (int step = 0; step < maxlen; ++step) {
        double e_t[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            e_t[j] = e_next[j];
            if (step + 1 < len[j] && row < K) e_next[j] = e[(tcur[j] + tstep) * K + row];
        }
        // ---- (transition operand) x (eight previous messages): four accumulation chains, added in a fixed order ----
        double acc[4][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q][0] = acc[q][1] = 0.0;
        const double* bp = s_prev + lr * LDP + lk;
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) dmma884(acc[ks & 3][0], acc[ks & 3][1], a[ks], bp[4 * ks]);
        double v[2], contrib[2];
        bool plain[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double d = (acc[0][j] + acc[1][j]) + (acc[2][j] + acc[3][j]);
            plain[j] = first_plain[j] && step == 0;
            if (plain[j]) v[j] = BACKWARD ? 1.0 : s_prev[(2 * lk + j) * LDP + (row < KP ? row : 0)] * e_t[j];
            else v[j] = BACKWARD ? d : d * e_t[j];
            if (row >= K || step >= len[j]) v[j] = 0.0;
            // forward: sum over all states (:3601); backward: all but the last state (:3646)
            contrib[j] = (row < (BACKWARD ? K - 1 : K)) ? v[j] : 0.0;
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
            contrib[0] += __shfl_xor_sync(0xffffffffu, contrib[0], o);
            contrib[1] += __shfl_xor_sync(0xffffffffu, contrib[1], o);
        }
        if (lr == 0) *reinterpret_cast<double2*>(&s_wsum[warp * 8 + 2 * lk]) = make_double2(contrib[0], contrib[1]);
        __syncthreads();          // partial sums complete; every warp has read s_prev
        double tot[2] = {0.0, 0.0};
#pragma unroll
        for (int w = 0; w < NWARP; ++w) {
            const double2 p2 = *reinterpret_cast<const double2*>(&s_wsum[w * 8 + 2 * lk]);
            tot[0] += p2.x;
            tot[1] += p2.y;
        }
        int same = 1;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (step < len[j]) {
                if (BACKWARD && plain[j]) tot[j] = 1.0;
                else v[j] = v[j] * (1.0 / tot[j]);       // reciprocal + multiply: see the one-chunk kernel
                if (row < K) {
                    const int64_t t = tcur[j];
                    if (repair) {
                        const double old = msgs[t * K + row];
                        if (__double_as_longlong(old) != __double_as_longlong(v[j])) { same = 0; msgs[t * K + row] = v[j]; }
                    } else {
                        msgs[t * K + row] = v[j];
                    }
                    s_prev[(2 * lk + j) * LDP + row] = BACKWARD ? v[j] * e_t[j] : v[j];
                    if (!BACKWARD && marg && row == 0) marg[t] = tot[j];     // margPrObs[t] of forward() (:3601)
                }
                tcur[j] += tstep;
            }
        }
        if (repair) {
            if (__syncthreads_and(same)) { stop_step = step; break; }
        } else {
            __syncthreads();
        }
    }

    // ---- boundary messages handed to the neighbour chunks ----
    for (int idx = tid; idx < 8 * KP; idx += 4 * KP) {
        const int n = idx / KP, k = idx - n * KP;
        const int64_t c = cbase + n;
        if (c >= C || k >= K) continue;
        const bool exact = BACKWARD ? (c == C - 1) : (c == 0);
        const int len_n = (int)(hgp_min64(N, c * CHUNK + CHUNK) - c * CHUNK);
        // a chunk that was still running when the loop stopped has a stored tail that is already exact; a (shorter) chunk
        // that had run to its end before that hands over the message it ended with
        if (repair && ((exact && !rebase) || stop_step < len_n)) {
            ends_out[c * K + k] = ends_in[c * K + k];
        } else {
            const double out = s_prev[n * LDP + k];
            if (repair) {
                const double old = ends_in[c * K + k];
                if (__double_as_longlong(old) != __double_as_longlong(out)) atomicOr(changed, 1);
            }
            ends_out[c * K + k] = out;
        }
    }
}

// ------------------------------------------------------------------------------------------
// hard responsibilities: one warp per beat
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
hard_resp_kernel(const double* __restrict__ e, const double* __restrict__ alpha, const double* __restrict__ beta,
                 const double* __restrict__ Pc, const double* __restrict__ alpha_in, int has_prev, int64_t N, int K,
                 int* __restrict__ z, int* __restrict__ zpair) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* a_prev = sm + warp * 2 * K;
    double* b_cur = a_prev + K;
    const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; t < N; t += wstride) {
        __syncwarp();
        // z_t = argmax_k alpha*beta (first maximum)
        double bv = -1.0;
        int bi = 0x7fffffff;
        for (int k = lane; k < K; k += 32) {
            const double al = alpha[t * K + k], be = beta[t * K + k];
            const double v = al * be;
            if (v > bv) { bv = v; bi = k; }
            b_cur[k] = e[t * K + k] * be;
            if (t > 0) a_prev[k] = alpha[(t - 1) * K + k];
            else if (has_prev) a_prev[k] = alpha_in[k];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) z[t] = (bi == 0x7fffffff) ? 0 : bi;
        __syncwarp();
        // zpair_t = argmax_{i,j} (alpha_{t-1,i} * b_j) * Pc_ij ; first beat of the whole sequence -> 0
        if (t == 0 && !has_prev) {
            if (lane == 0) zpair[t] = 0;
            continue;
        }
        double pv = 0.0;
        int pidx = 0x7fffffff;
        for (int i = 0; i < K; ++i) {
            const double ai = a_prev[i];
            if (!(ai > 0.0)) continue;   // zero rows can only tie at value 0, which resolves to index 0 anyway
            const double* prow = Pc + (int64_t)i * K;
            for (int j = lane; j < K; j += 32) {
                const double v = (ai * b_cur[j]) * prow[j];
                if (v > pv) { pv = v; pidx = i * K + j; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, pv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, pidx, o);
            if (ov > pv || (ov == pv && oi < pidx)) { pv = ov; pidx = oi; }
        }
        if (lane == 0) zpair[t] = (pidx == 0x7fffffff) ? 0 : pidx;
    }
}

// ------------------------------------------------------------------------------------------
// sufficient statistics
// ------------------------------------------------------------------------------------------
constexpr int STAT_BLOCK = 256;

__global__ void __launch_bounds__(STAT_BLOCK)
stats_count_kernel(const int* __restrict__ z, const int* __restrict__ zpair, const double* __restrict__ qbar,
                   int64_t N, int K, int* __restrict__ counts, double* __restrict__ partial) {
    __shared__ double s_red[STAT_BLOCK / 32];
    // contiguous slab per block => the Q_em summation order is fixed by (N, grid) only
    const int64_t per = (N + gridDim.x - 1) / gridDim.x;
    const int64_t b0 = blockIdx.x * per, b1 = hgp_min64(N, b0 + per);
    double acc = 0.0;
    for (int64_t n = b0 + threadIdx.x; n < b1; n += STAT_BLOCK) {
        const int zn = z[n];
        atomicAdd(&counts[zn], 1);
        atomicAdd(&counts[K + zpair[n]], 1);
        acc += qbar[n * K + zn];
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < STAT_BLOCK / 32; ++w) tot += s_red[w];
        partial[blockIdx.x] = tot;
    }
}

__global__ void stats_finish_kernel(const int* __restrict__ counts, const double* __restrict__ partial, int nblocks,
                                    const int* __restrict__ z, int64_t N, int K, int is_first, int pair0_is_dummy,
                                    double* __restrict__ Nm, double* __restrict__ trans, double* __restrict__ start,
                                    double* __restrict__ Qem) {
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        Nm[i] = (double)counts[i];
        start[i] = (is_first && N > 0 && z[0] == i) ? 1.0 : 0.0;
    }
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) trans[i] = (double)counts[K + i];
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int b = 0; b < nblocks; ++b) tot += partial[b];
        Qem[0] = tot;
    }
    (void)pair0_is_dummy;
}

// The forward and the backward scan are independent: the backward one runs on a side stream (fork / join with events)
// so that the two latency-bound chunk scans share the SMs instead of following each other.
struct HmmSide {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    int device = -1;
};
static HmmSide* hmm_side() {
    static thread_local HmmSide side;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    if (side.device != dev) {
        if (side.stream) { cudaStreamDestroy(side.stream); cudaEventDestroy(side.fork); cudaEventDestroy(side.join); }
        side = HmmSide();
        if (cudaStreamCreateWithFlags(&side.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&side.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&side.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        side.device = dev;
    }
    return &side;
}

// forward / backward scan launch: the eight-chunk tensor-core kernel for KP >= 64, one chunk per CTA below
template <int KP, bool BACKWARD>
static void launch_scan(const double* e, int64_t N, int K, const double* pi, const double* Mat, const double* boundary,
                        int has_boundary, double* msgs, double* marg, const double* ends_in, double* ends_out, int* changed,
                        int repair, int rebase, cudaStream_t st) {
    const int64_t C = (N + CHUNK - 1) / CHUNK;
    if (KP >= 64) {
        hmm_scan8_kernel<(KP >= 64 ? KP : 64), BACKWARD><<<(unsigned)((C + 7) / 8), 4 * KP, 0, st>>>(
            e, N, K, pi, Mat, boundary, has_boundary, msgs, marg, ends_in, ends_out, changed, repair, rebase);
    } else {
        hmm_scan_kernel<KP, BACKWARD><<<(unsigned)C, 4 * KP, 0, st>>>(e, N, K, pi, Mat, boundary, has_boundary, msgs, marg,
                                                                       ends_in, ends_out, changed, repair, rebase);
    }
}

template <int KP>
int launch_scans(const double* e, int64_t N, int K, const double* pi, const double* PiT, const double* Pi,
                 const double* boundary_in, int has_prev, int has_next, double* alpha, double* beta, double* marg,
                 double* endsA, double* endsB, int* changed, int* changed_host, int* rounds_host, cudaStream_t st,
                 int warm) {
    const int64_t C = (N + CHUNK - 1) / CHUNK;
    double* fa[2] = {endsA, endsA + C * K};
    double* fb[2] = {endsB, endsB + C * K};
    HmmSide* side = hmm_side();
    if (!side) { hgp_set_error("hgp_hmm_smooth: cannot create the side stream"); return HGP_E_UNSUPPORTED; }
    cudaStream_t sb = side->stream;      // backward direction
    cudaMemsetAsync(changed, 0, 2 * sizeof(int), st);
    if (!warm) {
    cudaEventRecord(side->fork, st);
    cudaStreamWaitEvent(sb, side->fork, 0);
    launch_scan<KP, false>(e, N, K, pi, PiT, boundary_in, has_prev, alpha, marg, nullptr, fa[0], changed, 0, 0, st);
    HGP_LAUNCH_CHECK("hmm forward scan");
    launch_scan<KP, true>(e, N, K, pi, Pi, boundary_in ? boundary_in + K : nullptr, has_next, beta, nullptr, nullptr, fb[0],
                          changed + 1, 0, 0, sb);
    HGP_LAUNCH_CHECK("hmm backward scan");
    cudaEventRecord(side->join, sb);
    cudaStreamWaitEvent(st, side->join, 0);
    }
    int rounds = 0;
    if (C > 1 || warm) {
        int cur = 0;
        bool need_f = true, need_b = true;
        int first_rebase = warm;     // warm start: the first repair round also re-scans the two boundary chunks
        while (need_f || need_b) {
            cudaMemsetAsync(changed, 0, 2 * sizeof(int), st);
            if (need_b) {
                cudaEventRecord(side->fork, st);
                cudaStreamWaitEvent(sb, side->fork, 0);
            }
            if (need_f) {
                launch_scan<KP, false>(e, N, K, pi, PiT, boundary_in, has_prev, alpha, marg, fa[cur], fa[cur ^ 1], changed, 1,
                                       first_rebase, st);
                HGP_LAUNCH_CHECK("hmm forward repair");
            }
            if (need_b) {
                launch_scan<KP, true>(e, N, K, pi, Pi, boundary_in ? boundary_in + K : nullptr, has_next, beta, nullptr,
                                      fb[cur], fb[cur ^ 1], changed + 1, 1, first_rebase, sb);
                HGP_LAUNCH_CHECK("hmm backward repair");
                cudaEventRecord(side->join, sb);
                cudaStreamWaitEvent(st, side->join, 0);
            }
            cudaError_t er = cudaMemcpyAsync(changed_host, changed, 2 * sizeof(int), cudaMemcpyDeviceToHost, st);
            if (er != cudaSuccess) return hgp_status(er, "hmm repair flag copy");
            er = cudaStreamSynchronize(st);
            if (er != cudaSuccess) return hgp_status(er, "hmm repair sync");
            ++rounds;
            // a direction that did not use the freshly written buffer keeps its old one
            const bool cf = need_f && changed_host[0], cb = need_b && changed_host[1];
            if (need_f && !cf) { /* converged: fa[cur^1] holds the final ends */ }
            if (need_b && !cb) { /* converged */ }
            // directions that skipped this round must see the same buffer index next round
            if (!need_f) cudaMemcpyAsync(fa[cur ^ 1], fa[cur], sizeof(double) * C * K, cudaMemcpyDeviceToDevice, st);
            if (!need_b) cudaMemcpyAsync(fb[cur ^ 1], fb[cur], sizeof(double) * C * K, cudaMemcpyDeviceToDevice, st);
            need_f = cf;
            need_b = cb;
            first_rebase = 0;
            cur ^= 1;
            if (rounds > C + 2) { hgp_set_error("hmm repair did not converge"); return HGP_E_UNSUPPORTED; }
        }
        // leave the final boundary messages in slot 0 for the caller
        if (cur != 0) {
            cudaMemcpyAsync(fa[0], fa[cur], sizeof(double) * C * K, cudaMemcpyDeviceToDevice, st);
            cudaMemcpyAsync(fb[0], fb[cur], sizeof(double) * C * K, cudaMemcpyDeviceToDevice, st);
        }
    }
    if (rounds_host) *rounds_host = rounds;
    return 0;
}

}  // namespace

extern "C" int hgp_lead_weights(const double* q, const double* snr, const double* lead_w, int64_t N, int M, int L,
                                double* qbar, double* e, double* wout, int* flags, void* stream) {
    HGP_REQUIRE(N >= 0 && M > 0 && L > 0 && L <= MAX_LEADS, "hgp_lead_weights: need 0 < L <= 16, M > 0");
    HGP_REQUIRE(snr != nullptr || lead_w != nullptr, "hgp_lead_weights: snr or lead_w required");
    cudaError_t er = cudaMemsetAsync(flags, 0, sizeof(int), (cudaStream_t)stream);
    if (er != cudaSuccess) return hgp_status(er, "hgp_lead_weights: memset");
    if (N == 0) return 0;
    int blocks = (int)hgp_min64((N + 7) / 8, 148 * 8);
    lead_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(q, snr, lead_w, N, M, L, qbar, e, wout, flags);
    HGP_LAUNCH_CHECK("hgp_lead_weights");
    return 0;
}

extern "C" int64_t hgp_hmm_workspace_bytes(int64_t N, int K) {
    const int64_t C = (N + CHUNK - 1) / CHUNK;
    // 2 directions x 2 buffers of chunk-boundary messages + flags (+ pinned-free host copy lives in the call)
    return 4 * C * K * (int64_t)sizeof(double) + 256;
}

static int hmm_smooth_impl(int warm, const double* e, int64_t N, int K, const double* pi, const double* PiT, const double* Pi,
                              const double* Pc, const double* boundary_in, int has_prev, int has_next, double* alpha,
                              double* beta, double* marg, int* z, int* zpair, double* boundary_out, void* workspace,
                              int64_t workspace_bytes, int* rounds_host, void* stream) {
    HGP_REQUIRE(N > 0 && K > 0, "hgp_hmm_smooth: need N > 0, K > 0");
    if (K > 128) { hgp_set_error("hgp_hmm_smooth: K <= 128 supported (got %d)", K); return HGP_E_UNSUPPORTED; }
    HGP_REQUIRE(!(has_prev || has_next) || boundary_in != nullptr, "hgp_hmm_smooth: boundary_in required");
    if (workspace_bytes < hgp_hmm_workspace_bytes(N, K)) { hgp_set_error("hgp_hmm_smooth: workspace too small"); return HGP_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t C = (N + CHUNK - 1) / CHUNK;
    double* endsA = reinterpret_cast<double*>(workspace);
    double* endsB = endsA + 2 * C * K;
    int* changed = reinterpret_cast<int*>(endsB + 2 * C * K);
    int changed_host[2] = {0, 0};
    int rc;
    if (K <= 8) rc = launch_scans<8>(e, N, K, pi, PiT, Pi, boundary_in, has_prev, has_next, alpha, beta, marg, endsA, endsB, changed, changed_host, rounds_host, st, warm);
    else if (K <= 16) rc = launch_scans<16>(e, N, K, pi, PiT, Pi, boundary_in, has_prev, has_next, alpha, beta, marg, endsA, endsB, changed, changed_host, rounds_host, st, warm);
    else if (K <= 32) rc = launch_scans<32>(e, N, K, pi, PiT, Pi, boundary_in, has_prev, has_next, alpha, beta, marg, endsA, endsB, changed, changed_host, rounds_host, st, warm);
    else if (K <= 64) rc = launch_scans<64>(e, N, K, pi, PiT, Pi, boundary_in, has_prev, has_next, alpha, beta, marg, endsA, endsB, changed, changed_host, rounds_host, st, warm);
    else rc = launch_scans<128>(e, N, K, pi, PiT, Pi, boundary_in, has_prev, has_next, alpha, beta, marg, endsA, endsB, changed, changed_host, rounds_host, st, warm);
    if (rc) return rc;
    const int warps = 8;
    size_t smem = sizeof(double) * warps * 2 * K;
    int blocks = (int)hgp_min64((N + warps - 1) / warps, 148 * 8);
    hard_resp_kernel<<<blocks, warps * 32, smem, st>>>(e, alpha, beta, Pc, boundary_in, has_prev, N, K, z, zpair);
    HGP_LAUNCH_CHECK("hgp_hmm_smooth: hard_resp");
    if (boundary_out) {
        // alpha of the last beat; (beta (.) e) of the first beat = backward boundary of chunk 0
        cudaError_t er = cudaMemcpyAsync(boundary_out, alpha + (N - 1) * K, sizeof(double) * K, cudaMemcpyDeviceToDevice, st);
        if (er != cudaSuccess) return hgp_status(er, "hgp_hmm_smooth: boundary copy");
        er = cudaMemcpyAsync(boundary_out + K, endsB, sizeof(double) * K, cudaMemcpyDeviceToDevice, st);
        if (er != cudaSuccess) return hgp_status(er, "hgp_hmm_smooth: boundary copy");
    }
    return 0;
}

extern "C" int hgp_hmm_smooth(const double* e, int64_t N, int K, const double* pi, const double* PiT, const double* Pi,
                              const double* Pc, const double* boundary_in, int has_prev, int has_next, double* alpha,
                              double* beta, double* marg, int* z, int* zpair, double* boundary_out, void* workspace,
                              int64_t workspace_bytes, int* rounds_host, void* stream) {
    return hmm_smooth_impl(0, e, N, K, pi, PiT, Pi, Pc, boundary_in, has_prev, has_next, alpha, beta, marg, z, zpair,
                           boundary_out, workspace, workspace_bytes, rounds_host, stream);
}

extern "C" int hgp_hmm_resmooth(const double* e, int64_t N, int K, const double* pi, const double* PiT, const double* Pi,
                                const double* Pc, const double* boundary_in, int has_prev, int has_next, double* alpha,
                                double* beta, double* marg, int* z, int* zpair, double* boundary_out, void* workspace,
                                int64_t workspace_bytes, int* rounds_host, void* stream) {
    return hmm_smooth_impl(1, e, N, K, pi, PiT, Pi, Pc, boundary_in, has_prev, has_next, alpha, beta, marg, z, zpair,
                           boundary_out, workspace, workspace_bytes, rounds_host, stream);
}

extern "C" int64_t hgp_suffstats_workspace_bytes(int64_t N, int K) {
    (void)N;
    return (int64_t)sizeof(int) * (K + (int64_t)K * K) + (int64_t)sizeof(double) * 1024 + 64;
}

extern "C" int hgp_suffstats(const int* z, const int* zpair, const double* qbar, int64_t N, int K, int is_first_slice,
                             double* Nm, double* trans, double* start, double* Qem, void* workspace,
                             int64_t workspace_bytes, void* stream) {
    HGP_REQUIRE(N >= 0 && K > 0, "hgp_suffstats: bad sizes");
    if (workspace_bytes < hgp_suffstats_workspace_bytes(N, K)) { hgp_set_error("hgp_suffstats: workspace too small"); return HGP_E_WORKSPACE; }
    cudaStream_t st = (cudaStream_t)stream;
    int* counts = reinterpret_cast<int*>(workspace);
    size_t cbytes = sizeof(int) * (K + (size_t)K * K);
    double* partial = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + ((cbytes + 63) / 64) * 64);
    cudaError_t er = cudaMemsetAsync(counts, 0, cbytes, st);
    if (er != cudaSuccess) return hgp_status(er, "hgp_suffstats: memset");
    int nblocks = (int)hgp_max64(1, hgp_min64(1024, (N + 4095) / 4096));
    if (N > 0) {
        stats_count_kernel<<<nblocks, STAT_BLOCK, 0, st>>>(z, zpair, qbar, N, K, counts, partial);
        HGP_LAUNCH_CHECK("hgp_suffstats: count");
    } else {
        nblocks = 0;
    }
    stats_finish_kernel<<<1, 256, 0, st>>>(counts, partial, nblocks, z, N, K, is_first_slice, 0, Nm, trans, start, Qem);
    HGP_LAUNCH_CHECK("hgp_suffstats: finish");
    return 0;
}
