// summaries.cuh -- filter-level read-outs evaluated on the device-resident population
// (src/filters.jl:17-42, src/particle.jl:157-172): the posterior predictive density at a batch of
// query points (a mixture over particles of mixtures of Student-t components), the weighted
// distribution of the component count, the weighted means of ncomponents / state_entropy, the
// byte-wide assignment matrix and the T x T assignment-similarity matrix.  Floating-point sums are
// per-block partials combined in a fixed order; the similarity counts are integers: every result is
// reproducible from run to run.
#pragma once
#include "model.cuh"

#define SUM_THREADS 256
#define PRED_QCHUNK 8            // query points evaluated per pass over a particle's slots

// block-wide sum of one double per thread (fixed tree), result valid in thread 0
__device__ __forceinline__ double block_sum_double(double v, double *s_red /* >= 8 */)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[wid] = v;
    __syncthreads();
    if (threadIdx.x == 0) for (int w = 1; w < (int)(blockDim.x >> 5); w++) v += s_red[w];
    return v;
}

// posterior_predictive(ps) (src/filters.jl:22-24) at Q query points: block b writes
//   part[b * (Q + 1) + q] = sum_{i in block} w_i sum_{j in 1..K_i+1} N_ij-or-alpha t_ij(x_q)      (q < Q)
//   part[b * (Q + 1) + Q] = sum_{i in block} w_i
// weights(p) = exp.(log_prior(p.stateprior)) (src/particle.jl:160, src/statepriors.jl:24-29) are N_j / (t + alpha) and
// alpha / (t + alpha) under the CRP (t = observations so far, the same for every particle: the host divides);
// t_ij = posterior_predictive of the component (src/component.jl:12-23).  The table row C(n) carries log N_j
// (log alpha for n = 0).
template <int D, typename R = double>
__global__ void __launch_bounds__(SUM_THREADS) k_predictive(const R *__restrict__ S, const double *__restrict__ logw,
                                                            const int *__restrict__ Kc, long long cap, long long n,
                                                            const NRow *__restrict__ tab, const PriorConst *__restrict__ pc,
                                                            const double *__restrict__ xs, int Q, double *__restrict__ part)
{
    using L = Layout<D>;
    __shared__ double s_red[8];
    __shared__ double s_new[PRED_QCHUNK];
    __shared__ double s_x[PRED_QCHUNK * D];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    const int K = live ? Kc[i] : 0;
    const double w = live ? exp(logw[i]) : 0.0;
    double mu0[D];
#pragma unroll
    for (int k = 0; k < D; k++) mu0[k] = pc->mu0[k];
    for (int q0 = 0; q0 < Q; q0 += PRED_QCHUNK) {
        const int nq = Q - q0 < PRED_QCHUNK ? Q - q0 : PRED_QCHUNK;
        __syncthreads();
        // the chunk's query points and the prior component's density at them (the same for every particle)
        if (threadIdx.x < nq) {
            double dx[D], z[D], dinv[D], off[L::NOFF > 0 ? L::NOFF : 1];
#pragma unroll
            for (int k = 0; k < D; k++) {
                const double x = xs[(long long)(q0 + threadIdx.x) * D + k];
                s_x[threadIdx.x * D + k] = x; dx[k] = x - mu0[k]; dinv[k] = pc->L0_dinv[k];
            }
#pragma unroll
            for (int k = 0; k < L::NOFF; k++) off[k] = pc->L0_off[k];
            const double qf = quad_form<D>(dx, dinv, off, z);
            s_new[threadIdx.x] = exp(pc->row0.C - 0.5 * pc->logdet0 - pc->row0.h * log1p(pc->row0.r * qf));
        }
        __syncthreads();
        double acc[PRED_QCHUNK];
#pragma unroll
        for (int q = 0; q < PRED_QCHUNK; q++) acc[q] = q < nq ? s_new[q] : 0.0;
        for (int j = 0; j < K; j++) {
            const R *base = S + ((long long)j * L::F) * cap + i;
            double fl[L::F];
#pragma unroll
            for (int f = 0; f < L::F; f++) fl[f] = (double)base[(long long)f * cap];
            const NRow row = tab[(int)fl[L::F_N]];
#pragma unroll
            for (int q = 0; q < PRED_QCHUNK; q++) {
                if (q < nq) {
                    double y[D], z[D], qf;
#pragma unroll
                    for (int k = 0; k < D; k++) y[k] = s_x[q * D + k];
                    acc[q] += exp(slot_logw<D>(row, fl[L::F_LOGDET], fl + L::F_MEAN, fl + L::F_DINV, fl + L::F_OFF, y, mu0, z, &qf));
                }
            }
        }
#pragma unroll
        for (int q = 0; q < PRED_QCHUNK; q++) {
            const double tot = block_sum_double(w * acc[q], s_red);
            if (threadIdx.x == 0 && q < nq) part[(long long)blockIdx.x * (Q + 1) + q0 + q] = tot;
        }
    }
    const double tw = block_sum_double(w, s_red);
    if (threadIdx.x == 0) part[(long long)blockIdx.x * (Q + 1) + Q] = tw;
}

// out[c] = sum over blocks of part[b * ncol + c], in block order (one block per column: fixed tree)
__global__ void __launch_bounds__(SUM_THREADS) k_sum_partials(const double *__restrict__ part, long long nblocks, int ncol, double *__restrict__ out)
{
    __shared__ double s_red[8];
    const int c = blockIdx.x;
    double v = 0.0;
    for (long long b = threadIdx.x; b < nblocks; b += blockDim.x) v += part[b * ncol + c];
    v = block_sum_double(v, s_red);
    if (threadIdx.x == 0) out[c] = v;
}

// state_entropy(p) = entropy(exp.(log_prior(p.stateprior))) (src/particle.jl:162, src/statepriors.jl:24-36): the
// entropy of the NORMALISED prior over the candidate states of particle i
template <int D, typename R>
__device__ __forceinline__ double particle_state_entropy(const R *__restrict__ S, long long cap, long long i, int K,
                                                         const PriorConst *__restrict__ pc, SpArrays sp, double t_obs)
{
    using L = Layout<D>;
    auto xlogx = [](double p) { return p > 0.0 ? p * log(p) : 0.0; };
    const int kind = pc->sp_kind;
    double H = 0.0;
    if (kind == SPK_CHANGEPOINT) {
        if (K == 0) return 0.0;                                      // candidates max(k,1):k+1 = 1:1
        const double ps = exp(pc->sp_lp_stay), pch = exp(pc->sp_lp_change), tot = ps + pch;
        return -(xlogx(ps / tot) + xlogx(pch / tot));
    }
    if (kind == SPK_NSTATE) {
        const double tot = pc->sp_total0 + t_obs;
        for (int j = 0; j < K; j++) H -= xlogx((pc->sp_counts0[j] + (double)S[((long long)j * L::F + L::F_N) * cap + i]) / tot);
        return H;
    }
    if (kind == SPK_STICKY) {
        if (K == 0) return 0.0;                                      // candidates 1:1
        double sumN = 0.0;
        for (int j = 0; j < K; j++) sumN += sp.N[(long long)j * cap + i];
        const double w0 = exp(pc->sp_logit_rho) * (sumN + pc->alpha), tot = w0 + sumN + pc->alpha;
        H = -xlogx(w0 / tot) - xlogx(pc->alpha / tot);
        for (int j = 0; j < K; j++) H -= xlogx(sp.N[(long long)j * cap + i] / tot);
        return H;
    }
    const double tot = t_obs + pc->alpha;                            // CRP: N_j / (t + alpha), alpha / (t + alpha)
    H = -xlogx(pc->alpha / tot);
    for (int j = 0; j < K; j++) H -= xlogx((double)S[((long long)j * L::F + L::F_N) * cap + i] / tot);
    return H;
}

// One pass over the population for the scalar summaries: block b writes a record of 3 + slots doubles
//   [0] sum w   [1] sum w K   [2] sum w H(state prior)   [3 + k] sum of w over the block's particles with K = k
// ncomponents_dist(ps) = fit(Categorical, ncomponents, weights) (src/filters.jl:36-37) is [3 + k] / [0];
// mean(ncomponents, ps), state_entropy(ps) = mean(state_entropy, ps) (src/filters.jl:17-20) are [1] / [0], [2] / [0].
template <int D, typename R = double>
__global__ void __launch_bounds__(SUM_THREADS) k_pop_summary(const R *__restrict__ S, const double *__restrict__ logw,
                                                             const int *__restrict__ Kc, long long cap, long long n, int slots,
                                                             const PriorConst *__restrict__ pc, SpArrays sp, double t_obs,
                                                             double *__restrict__ part)
{
    __shared__ double s_red[8];
    __shared__ int s_lo[8], s_hi[8];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n;
    const int K = live ? min(Kc[i], slots - 1) : 0;
    const double w = live ? exp(logw[i]) : 0.0;
    const double H = live ? particle_state_entropy<D, R>(S, cap, i, K, pc, sp, t_obs) : 0.0;
    double *rec = part + (long long)blockIdx.x * (3 + slots);
    double t = block_sum_double(w, s_red);
    if (threadIdx.x == 0) rec[0] = t;
    t = block_sum_double(w * (double)K, s_red);
    if (threadIdx.x == 0) rec[1] = t;
    t = block_sum_double(w * H, s_red);
    if (threadIdx.x == 0) rec[2] = t;
    // the block's K values span a short range: one fixed-tree reduction per value in it
    int klo = live ? K : slots, khi = live ? K : -1;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { klo = min(klo, __shfl_xor_sync(0xffffffffu, klo, d)); khi = max(khi, __shfl_xor_sync(0xffffffffu, khi, d)); }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = klo; s_hi[threadIdx.x >> 5] = khi; }
    __syncthreads();
    for (int wv = 0; wv < (int)(blockDim.x >> 5); wv++) { klo = min(klo, s_lo[wv]); khi = max(khi, s_hi[wv]); }
    for (int k = threadIdx.x; k < slots; k += blockDim.x) if (k < klo || k > khi) rec[3 + k] = 0.0;
    for (int k = klo; k <= khi; k++) {
        t = block_sum_double(live && K == k ? w : 0.0, s_red);
        if (threadIdx.x == 0) rec[3 + k] = t;
    }
}

// assignments as bytes, time-major: out8[t * nc + (i - i0)] = label (0-based) of observation t in particle i,
// for the particles [i0, i0 + nc) (assignments(p), src/particle.jl:25-32)
// off == NULL: dense log, generation t at t * cap; else generation t at off[t] (pruned arena: the stored nodes of a
// generation are the ancestors of the current population, renumbered densely)
__global__ void k_backtrack8(const unsigned int *__restrict__ gen_anc, const unsigned char *__restrict__ gen_asg, long long cap,
                             long long T, long long i0, long long nc, unsigned char *__restrict__ out8,
                             const long long *__restrict__ off = nullptr)
{
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    long long k = i0 + c;
    for (long long t = T - 1; t >= 0; t--) {
        PF_ASSERT(k >= 0 && k < cap);
        const long long o = off ? off[t] : t * cap;
        out8[t * nc + c] = gen_asg[o + k];
        k = gen_anc[o + k];
    }
}

// ------------------------------------------------------------------ genealogy pruning
// The reference's ancestor chain (src/particle.jl:25-32) is an implicit tree pruned by the garbage collector: a
// particle that leaves no descendant in the current population disappears with its history.  The device log does the
// same explicitly: from the newest generation backwards, the nodes of generation g - 1 that are still an ancestor of
// somebody are marked, renumbered densely (exclusive scan of the marks), compacted, and the parent indices of
// generation g are rewritten to the new numbering.  Coalescence makes the live history O(N log T) nodes instead of N T.
__global__ void k_gen_mark(const unsigned int *__restrict__ anc, const long long *__restrict__ n_ptr, unsigned int *__restrict__ mark)
{
    const long long n = *n_ptr;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) mark[anc[k]] = 1u;
}
__global__ void k_gen_count(const unsigned int *__restrict__ mark, const unsigned int *__restrict__ idx, long long n_old, long long *n_new)
{
    if (threadIdx.x || blockIdx.x) return;
    *n_new = n_old > 0 ? (long long)idx[n_old - 1] + mark[n_old - 1] : 0;
}
__global__ void k_gen_compact(const unsigned int *__restrict__ anc, const unsigned char *__restrict__ asg, long long n_old,
                              const unsigned int *__restrict__ mark, const unsigned int *__restrict__ idx,
                              unsigned int *__restrict__ tmp_anc, unsigned char *__restrict__ tmp_asg)
{
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n_old; j += (long long)gridDim.x * blockDim.x)
        if (mark[j]) { tmp_anc[idx[j]] = anc[j]; tmp_asg[idx[j]] = asg[j]; }
}
__global__ void k_gen_copy(const unsigned int *__restrict__ src_anc, const unsigned char *__restrict__ src_asg,
                           const long long *__restrict__ n_ptr, unsigned int *__restrict__ anc, unsigned char *__restrict__ asg)
{
    const long long n = *n_ptr;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) { anc[j] = src_anc[j]; asg[j] = src_asg[j]; }
}
__global__ void k_gen_rewrite(unsigned int *__restrict__ anc, const long long *__restrict__ n_ptr, const unsigned int *__restrict__ idx)
{
    const long long n = *n_ptr;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) anc[k] = idx[anc[k]];
}

// The same pass for a RUN of short generations (<= GEN_SMALL stored nodes each: all the old ones, since the live nodes
// only get fewer with age) in ONE launch: one block walks the generations g_hi, g_hi - 1, ..., 1 with the marks, their
// prefix and the compacted rows in shared memory.  n[] holds the old lengths on entry and the new ones on return.
#define GEN_SMALL 4096
__global__ void __launch_bounds__(1024) k_gen_prune_small(unsigned int *__restrict__ gen_anc, unsigned char *__restrict__ gen_asg,
                                                          const long long *__restrict__ off, long long *__restrict__ n, long long g_hi)
{
    __shared__ unsigned int s_idx[GEN_SMALL];          // (exclusive prefix << 1) | mark
    __shared__ unsigned int s_anc[GEN_SMALL];
    __shared__ unsigned char s_asg[GEN_SMALL];
    __shared__ unsigned int s_w[33];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int PER = GEN_SMALL / 1024;
    for (long long g = g_hi; g >= 1; g--) {
        const long long p = g - 1;
        const int ng = (int)n[g], np_old = (int)n[p];
        unsigned int *anc_g = gen_anc + off[g];
        unsigned int *anc_p = gen_anc + off[p];
        unsigned char *asg_p = gen_asg + off[p];
        for (int j = tid; j < np_old; j += 1024) s_idx[j] = 0u;
        __syncthreads();
        for (int k = tid; k < ng; k += 1024) s_idx[anc_g[k]] = 1u;
        __syncthreads();
        unsigned int m[PER], loc = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) { const int j = tid * PER + k; m[k] = j < np_old ? s_idx[j] : 0u; loc += m[k]; }
        unsigned int inc = loc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            unsigned int w = s_w[lane], wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned int o = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += o; }
            s_w[lane] = wi - w;
            if (lane == 31) s_w[32] = wi;
        }
        __syncthreads();
        unsigned int run = s_w[wid] + inc - loc;
        const int np_new = (int)s_w[32];
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int j = tid * PER + k;
            if (j < np_old) {
                s_idx[j] = (run << 1) | m[k];
                if (m[k]) { s_anc[run] = anc_p[j]; s_asg[run] = asg_p[j]; }
                run += m[k];
            }
        }
        __syncthreads();
        for (int j = tid; j < np_new; j += 1024) { anc_p[j] = s_anc[j]; asg_p[j] = s_asg[j]; }
        for (int k = tid; k < ng; k += 1024) anc_g[k] = s_idx[anc_g[k]] >> 1;
        __syncthreads();
        if (tid == 0) n[p] = np_new;
        __syncthreads();
    }
}

// assignment_similarity(ps) (src/filters.jl:39-42): sim[s, t] = 1 - Hamming(A[s, :], A[t, :]) / n, i.e. the fraction of
// particles that give observations s and t the same label.  counts[s * T + t] accumulates the number of agreeing
// particles (integer atomics: order-free).  A block owns a tile of SIM_TILE x SIM_TILE observation pairs and a slice of
// the particles; a thread keeps the tile's 64 pair counters in registers and compares 16 particles per load.
#define SIM_TILE 8
__global__ void __launch_bounds__(SUM_THREADS) k_similarity(const unsigned char *__restrict__ A8, long long T, long long nc,
                                                            unsigned long long *__restrict__ counts)
{
    const int ntile = (int)((T + SIM_TILE - 1) / SIM_TILE);
    // tile pair (ts <= tt) from the linear block index
    int p = blockIdx.x, ts = 0;
    while (p >= ntile - ts) { p -= ntile - ts; ts++; }
    const int tt = ts + p;
    const long long s0 = (long long)ts * SIM_TILE, t0 = (long long)tt * SIM_TILE;
    unsigned int cntr[SIM_TILE][SIM_TILE];
#pragma unroll
    for (int a = 0; a < SIM_TILE; a++)
#pragma unroll
        for (int b = 0; b < SIM_TILE; b++) cntr[a][b] = 0;
    const bool vec = (nc % 16 == 0) && ((size_t)A8 % 16 == 0);
    const long long n16 = vec ? nc / 16 : 0;
    for (long long k = (long long)blockIdx.y * blockDim.x + threadIdx.x; k < n16; k += (long long)gridDim.y * blockDim.x) {
        uint4 ra[SIM_TILE], rb[SIM_TILE];
#pragma unroll
        for (int a = 0; a < SIM_TILE; a++) {
            const long long s = s0 + a < T ? s0 + a : T - 1, t = t0 + a < T ? t0 + a : T - 1;
            ra[a] = reinterpret_cast<const uint4 *>(A8 + s * nc)[k];
            rb[a] = reinterpret_cast<const uint4 *>(A8 + t * nc)[k];
        }
#pragma unroll
        for (int a = 0; a < SIM_TILE; a++)
#pragma unroll
            for (int b = 0; b < SIM_TILE; b++)
                cntr[a][b] += __popc(__vcmpeq4(ra[a].x, rb[b].x)) + __popc(__vcmpeq4(ra[a].y, rb[b].y)) +
                              __popc(__vcmpeq4(ra[a].z, rb[b].z)) + __popc(__vcmpeq4(ra[a].w, rb[b].w));       // 8 bits per equal byte
    }
    // scalar tail (or the whole slice when the chunk is not 16-aligned)
    for (long long k = n16 * 16 + (long long)blockIdx.y * blockDim.x + threadIdx.x; k < nc; k += (long long)gridDim.y * blockDim.x) {
#pragma unroll
        for (int a = 0; a < SIM_TILE; a++)
#pragma unroll
            for (int b = 0; b < SIM_TILE; b++) {
                const long long s = s0 + a < T ? s0 + a : T - 1, t = t0 + b < T ? t0 + b : T - 1;
                cntr[a][b] += 8u * (A8[s * nc + k] == A8[t * nc + k]);
            }
    }
    __shared__ unsigned int s_c[SUM_THREADS / 32][SIM_TILE * SIM_TILE];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < SIM_TILE; a++)
#pragma unroll
        for (int b = 0; b < SIM_TILE; b++) {
            const unsigned int v = __reduce_add_sync(0xffffffffu, cntr[a][b] >> 3);
            if (lane == 0) s_c[wid][a * SIM_TILE + b] = v;
        }
    __syncthreads();
    if (threadIdx.x < SIM_TILE * SIM_TILE) {
        const int a = threadIdx.x / SIM_TILE, b = threadIdx.x % SIM_TILE;
        unsigned long long v = 0;
        for (int w = 0; w < SUM_THREADS / 32; w++) v += s_c[w][threadIdx.x];
        const long long s = s0 + a, t = t0 + b;
        if (s < T && t < T && v) {
            atomicAdd(&counts[s * T + t], v);
            if (ts != tt) atomicAdd(&counts[t * T + s], v);
        }
    }
}
