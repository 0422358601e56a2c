// chenliu.cuh -- fit!(::ChenLiuParticles, y) (src/chenliu.jl:54-69): per-particle
// propagation with one categorical draw (propogate_chenliu, :28-42), coefficient of
// variation (:49-52), CV-triggered stratified rejuvenation or weight normalisation.
//
// The population is either one block (single GPU) or block-sharded in global order, rank r owning
// the global particles [r q, (r + 1) q), q = ceil(N / G) (fixed: Chen-Liu keeps N particles).  What
// crosses the ranks per observation are five scalars per rank -- max weight, exact sum of weights,
// exact sum of squared deviations, low-partition mass, "prefix ready" -- pushed into every peer's
// exchange region and synchronised by sequence flags (select.cuh, "peer exchange"), and, on a
// rejuvenation step, the records of ancestors that live on another rank, which the child's owner
// PULLS from the peer's store over NVLink.  Every sum that feeds a decision is an exact integer
// (128-bit fixed point), so the G-rank run takes the decisions of the 1-rank run bit for bit.
#pragma once
#include "../../include/particles_b200.h"
#include "fearnhead.cuh"

#define CL_W_BITS 60          // rejuvenation fixed point: q = floor(W 2^(CL_W_BITS - e(max W)))
#define CL_TILE 1024          // particles per tile of the prefix scan

struct ClState {
    int cur;                  // which store buffer holds the population
    int resample;             // decision of the current step
    u64 wmax_bits;            // max W over this shard, then (after the exchange) over the population
    acc128 sumW;              // sum of W over this shard at scale SEL_TOT_BITS - e(wmax)
    acc128 sumSq;             // sum of ((W - mean) 2^-e(wmax))^2 over this shard at scale SEL_TOT_BITS - 2
    u128 sumW_g;              // ... over the population
    double mean, cv, log_total;
    u128 Qtot;                // sum of q over the population (rejuvenation scale)
    u128 Qoff;                // sum of q over the lower ranks
    u128 Qincl[PF_MAX_RANKS]; // inclusive prefix of the ranks' q sums (Qincl[G - 1] = Qtot)
    int qscale;
    int pad;
    long long n_resampled_steps;
};

// this shard's block of the population
struct ClShard {
    long long N;              // population (all shards)
    long long n_loc;          // particles of this shard
    long long goff;           // global index of the shard's first particle
    long long q;              // block length ceil(N / G)
};

// ------------------------------------------------------------------ k_cl_propagate
// Pass 1 evaluates the K + 1 putative weights (src/particle.jl:100-116) into the scratch rows of V and sums them;
// pass 2 walks them for the inverse-CDF draw; the chosen slot is updated in place.
template <int D, typename R = double>
__global__ void __launch_bounds__(256) k_cl_propagate(R *S0, R *S1, const double *__restrict__ logw0,
                                                      const double *__restrict__ logw1, int *Kc0, int *Kc1, long long cap,
                                                      ClShard sh, int k_max, const NRow *__restrict__ tab,
                                                      const PriorConst *__restrict__ pc, const StepConst *__restrict__ scp,
                                                      StepState *st, ClState *cl, const double *__restrict__ u_dev, u64 seed,
                                                      u64 step, double *__restrict__ W, unsigned char *__restrict__ asg_tmp,
                                                      double *__restrict__ V, const SelScratch *guard)
{
    using L = Layout<D>;
    if (guard->halt) return;
    __shared__ u64 s_max[8];
    __shared__ long long s_ovf[8];
    const int cur = cl->cur;
    R *S = cur ? S1 : S0;
    const double *logw = cur ? logw1 : logw0;
    int *Kc = cur ? Kc1 : Kc0;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    u64 wmax = 0;
    long long myOvf = 0;
    if (i < sh.n_loc) {
        double y[D], mu0[D];
#pragma unroll
        for (int k = 0; k < D; k++) { y[k] = scp->y[k]; mu0[k] = pc->mu0[k]; }
        const int K = Kc[i];
        const double lw0 = logw[i];
        double sum = 0.0;
        for (int j = 0; j < K; j++) {
            const R *base = S + ((long long)j * L::F) * cap + i;
            double fl[L::F], z[D], q;
#pragma unroll
            for (int f = 0; f < L::F; f++) fl[f] = (double)base[(long long)f * cap];
            const NRow row = tab[(int)fl[L::F_N]];
            const double v = exp(lw0 + slot_logw<D>(row, fl[L::F_LOGDET], fl + L::F_MEAN, fl + L::F_DINV, fl + L::F_OFF, y, mu0, z, &q));
            V[(long long)j * cap + i] = v;
            sum = j == 0 ? v : sum + v;             // weights = weight.(ps); sum(weights)  (src/chenliu.jl:34,41)
        }
        int C = K;
        if (K < k_max) {
            const double v = exp(lw0 + scp->lw_new);
            V[(long long)K * cap + i] = v;
            sum = K == 0 ? v : sum + v;
            C = K + 1;
        } else myOvf = 1;
        // wsample: StatsBase inverse-CDF draw (src/chenliu.jl:35)
        const double u = u_dev ? u_dev[i] : pf_uniform(seed, step, 1, (u64)(sh.goff + i));
        const double t = u * sum;
        int pick = 0;
        double cw = V[i];
        while (cw < t && pick < C - 1) { pick++; cw += V[(long long)pick * cap + i]; }
        PF_ASSERT(pick >= 0 && pick <= K && pick < k_max && C >= 1);
        // instantiate in place (src/particle.jl:134-149)
        R *dst = S + ((long long)pick * L::F) * cap + i;
        if (pick == K) {
#pragma unroll
            for (int f = 0; f < L::F; f++) dst[(long long)f * cap] = (R)scp->newslot[f];
            Kc[i] = K + 1;
        } else {
            double fl[L::F];
#pragma unroll
            for (int f = 0; f < L::F; f++) fl[f] = (double)dst[(long long)f * cap];
            slot_add<D>(fl, tab[(int)fl[L::F_N]], y, mu0);
#pragma unroll
            for (int f = 0; f < L::F; f++) dst[(long long)f * cap] = (R)fl[f];
        }
        W[i] = sum;
        asg_tmp[i] = (unsigned char)pick;
        wmax = (u64)__double_as_longlong(sum);
    }
    wmax = warp_max64(wmax);
    myOvf = (long long)warp_sum64((u64)myOvf);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s_max[wid] = wmax; s_ovf[wid] = myOvf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) { wmax = s_max[w] > wmax ? s_max[w] : wmax; myOvf += s_ovf[w]; }
        if (wmax) atomicMax(&cl->wmax_bits, wmax);
        if (myOvf) atomicAdd((unsigned long long *)&st->overflow, (unsigned long long)myOvf);
    }
}

// ------------------------------------------------------------------ the scalar exchanges (one thread each)
// stage bit SEL_ST_PUB: push this shard's value into every region and raise the flag; SEL_ST_RUN: wait for every
// rank's flag and combine in rank order.  x == NULL (single GPU): the shard's value is the population's.
__device__ __forceinline__ void st_acc(acc128 *dst, const acc128 &v) { for (int k = 0; k < 4; k++) dst->l[k] = v.l[k]; }
__device__ __forceinline__ u128 ld_acc_peer(const acc128 *a)
{
    acc128 t;
    for (int k = 0; k < 4; k++) t.l[k] = ld_peer(&a->l[k]);
    return acc_value(&t);
}

// 1: max weight
__global__ void k_cl_x_wmax(ClState *cl, const Xch *x, unsigned long long seq, long long step, int stage, SelScratch *guard)
{
    if (threadIdx.x || blockIdx.x || guard->halt || !x) return;
    if (stage & SEL_ST_PUB) {
        for (int p = 0; p < x->nranks; p++) x->reg[p]->clw[x->rank] = cl->wmax_bits;
        xch_signal(x, XCH_CL_W, seq);
    }
    if (stage & SEL_ST_RUN) {
        if (!xch_wait(x, XCH_CL_W, seq)) { sel_halt(guard, step, PF_HALT_TIMEOUT); return; }
        u64 m = 0;
        for (int r = 0; r < x->nranks; r++) { const u64 v = ld_peer(&x->reg[x->rank]->clw[r]); m = v > m ? v : m; }
        cl->wmax_bits = m;
    }
}

// sum of W in fixed point (this shard)
__global__ void __launch_bounds__(256) k_cl_sum(const double *__restrict__ W, long long n_loc, ClState *cl, const SelScratch *guard)
{
    if (guard->halt) return;
    __shared__ u128 s_red[8];
    const int sc = SEL_TOT_BITS - exp_of_bits(cl->wmax_bits);
    u128 acc = mk128(0, 0);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_loc; i += (long long)gridDim.x * blockDim.x)
        acc = add128(acc, fx128((u64)__double_as_longlong(W[i]), sc));
    acc = warp_sum128(acc);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s_red[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) acc = add128(acc, s_red[w]);
        atomic_add_acc(&cl->sumW, acc);
    }
}

// 2: exact sum of the weights -> mean
__global__ void k_cl_x_sum(ClState *cl, const Xch *x, unsigned long long seq, long long step, int stage, long long N, SelScratch *guard)
{
    if (threadIdx.x || blockIdx.x || guard->halt) return;
    u128 tot = acc_value(&cl->sumW);
    if (x) {
        if (stage & SEL_ST_PUB) {
            for (int p = 0; p < x->nranks; p++) st_acc(&x->reg[p]->clsum[x->rank], cl->sumW);
            xch_signal(x, XCH_CL_S, seq);
        }
        if (!(stage & SEL_ST_RUN)) return;
        if (!xch_wait(x, XCH_CL_S, seq)) { sel_halt(guard, step, PF_HALT_TIMEOUT); return; }
        tot = mk128(0, 0);
        for (int r = 0; r < x->nranks; r++) tot = add128(tot, ld_acc_peer(&x->reg[x->rank]->clsum[r]));
    }
    const int sc = SEL_TOT_BITS - exp_of_bits(cl->wmax_bits);
    cl->sumW_g = tot;
    cl->mean = ldexp(to_double128(tot), -sc) / (double)N;
}

// centred sum of squares (mean_and_std, two-pass), exact: every term ((W - mean) 2^-e)^2 < 4 is truncated to a
// multiple of 2^-(SEL_TOT_BITS - 2) and the terms are summed as integers (order-free, shard-independent)
__global__ void __launch_bounds__(256) k_cl_var(const double *__restrict__ W, long long n_loc, ClState *cl, const SelScratch *guard)
{
    if (guard->halt) return;
    __shared__ u128 s_red[8];
    const int e = exp_of_bits(cl->wmax_bits);
    const double mean = cl->mean;
    u128 acc = mk128(0, 0);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_loc; i += (long long)gridDim.x * blockDim.x) {
        const double dlt = ldexp(W[i] - mean, -e);
        acc = add128(acc, fx128((u64)__double_as_longlong(dlt * dlt), SEL_TOT_BITS - 2));
    }
    acc = warp_sum128(acc);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s_red[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) acc = add128(acc, s_red[w]);
        atomic_add_acc(&cl->sumSq, acc);
    }
}

// 3: cv = 100 std / mean (corrected), decision cv > threshold (src/chenliu.jl:57-59)
__global__ void k_cl_decide(ClState *cl, const Xch *x, unsigned long long seq, long long step, int stage, long long N, double thr,
                            SelScratch *guard)
{
    if (threadIdx.x || blockIdx.x || guard->halt) return;
    u128 ssq = acc_value(&cl->sumSq);
    if (x) {
        if (stage & SEL_ST_PUB) {
            for (int p = 0; p < x->nranks; p++) st_acc(&x->reg[p]->clsq[x->rank], cl->sumSq);
            xch_signal(x, XCH_CL_Q, seq);
        }
        if (!(stage & SEL_ST_RUN)) return;
        if (!xch_wait(x, XCH_CL_Q, seq)) { sel_halt(guard, step, PF_HALT_TIMEOUT); return; }
        ssq = mk128(0, 0);
        for (int r = 0; r < x->nranks; r++) ssq = add128(ssq, ld_acc_peer(&x->reg[x->rank]->clsq[r]));
    }
    const int e = exp_of_bits(cl->wmax_bits);
    const int sc = SEL_TOT_BITS - e;
    const double ss = ldexp(to_double128(ssq), -(SEL_TOT_BITS - 2));                 // sum of ((W - mean) 2^-e)^2
    const double sd = N > 1 ? ldexp(sqrt(ss / (double)(N - 1)), e) : 0.0;
    const double cv = N > 1 ? sd / cl->mean * 100.0 : 0.0;
    cl->cv = cv;
    cl->log_total = log(to_double128(cl->sumW_g)) - (double)sc * 0.693147180559945309417232121458;
    cl->resample = cv > thr ? 1 : 0;
    cl->qscale = CL_W_BITS - e;
}

// ------------------------------------------------------------------ prefix sums of W (rejuvenation steps only)
// tile totals -> exclusive prefix over the tiles (one block) -> inclusive prefix per particle, offset by the mass of
// the lower ranks: cum[i] = sum of q over all global particles <= goff + i
__global__ void __launch_bounds__(256) k_cl_tile_sums(const double *__restrict__ W, long long n_loc, const ClState *cl,
                                                      u128 *__restrict__ tiles, const SelScratch *guard)
{
    if (guard->halt || !cl->resample) return;
    __shared__ u128 s_red[8];
    const int sc = cl->qscale;
    const long long b = (long long)blockIdx.x * CL_TILE;
    u128 acc = mk128(0, 0);
#pragma unroll
    for (int k = 0; k < CL_TILE / 256; k++) {
        const long long i = b + k * 256 + threadIdx.x;
        if (i < n_loc) acc = add128(acc, fx128((u64)__double_as_longlong(W[i]), sc));
    }
    acc = warp_sum128(acc);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s_red[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) acc = add128(acc, s_red[w]);
        tiles[blockIdx.x] = acc;
    }
}

// 4: exclusive prefix over this shard's tiles, the shard's total to every rank, offsets of the ranks
__global__ void __launch_bounds__(1024) k_cl_tile_prefix(u128 *__restrict__ tiles, long long ntiles, ClState *cl, const Xch *x,
                                                         unsigned long long seq, long long step, int stage, SelScratch *guard)
{
    if (guard->halt || !cl->resample) return;
    __shared__ u128 s_warp[33];
    if (stage & SEL_ST_PUB) {
        const long long per = (ntiles + blockDim.x - 1) / blockDim.x;
        const long long b = (long long)threadIdx.x * per, e = b + per < ntiles ? b + per : ntiles;
        u128 loc = mk128(0, 0);
        for (long long t = b; t < e; t++) loc = add128(loc, tiles[t]);
        u128 tot;
        u128 run = block_excl_scan128(loc, s_warp, &tot);
        for (long long t = b; t < e; t++) { const u128 v = tiles[t]; tiles[t] = run; run = add128(run, v); }
        if (threadIdx.x == 0) {
            if (x) {
                for (int p = 0; p < x->nranks; p++) x->reg[p]->clq[x->rank] = tot;
                xch_signal(x, XCH_CL_T, seq);
            } else {
                cl->Qoff = mk128(0, 0); cl->Qtot = tot; cl->Qincl[0] = tot;
            }
        }
    }
    if ((stage & SEL_ST_RUN) && x && threadIdx.x == 0) {
        if (!xch_wait(x, XCH_CL_T, seq)) { sel_halt(guard, step, PF_HALT_TIMEOUT); return; }
        u128 run = mk128(0, 0);
        for (int r = 0; r < x->nranks; r++) {
            if (r == x->rank) cl->Qoff = run;
            u128 v;
            v.lo = ld_peer(reinterpret_cast<const unsigned long long *>(&x->reg[x->rank]->clq[r].lo));
            v.hi = ld_peer(reinterpret_cast<const unsigned long long *>(&x->reg[x->rank]->clq[r].hi));
            run = add128(run, v);
            cl->Qincl[r] = run;
        }
        cl->Qtot = run;
    }
}

__global__ void __launch_bounds__(256) k_cl_tile_apply(const double *__restrict__ W, long long n_loc, const ClState *cl,
                                                       const u128 *__restrict__ tiles, u128 *__restrict__ cum, const SelScratch *guard)
{
    if (guard->halt || !cl->resample) return;
    __shared__ u128 s_warp[33];
    const int sc = cl->qscale;
    const long long b = (long long)blockIdx.x * CL_TILE + (long long)threadIdx.x * (CL_TILE / 256);
    u128 q[CL_TILE / 256], loc = mk128(0, 0);
#pragma unroll
    for (int k = 0; k < CL_TILE / 256; k++) {
        q[k] = (b + k < n_loc) ? fx128((u64)__double_as_longlong(W[b + k]), sc) : mk128(0, 0);
        loc = add128(loc, q[k]);
    }
    u128 tot;
    u128 run = add128(add128(cl->Qoff, tiles[blockIdx.x]), block_excl_scan128(loc, s_warp, &tot));
#pragma unroll
    for (int k = 0; k < CL_TILE / 256; k++) {
        run = add128(run, q[k]);
        if (b + k < n_loc) cum[b + k] = run;
    }
}

// 5: every rank's prefix array is complete (peers binary-search it)
__global__ void k_cl_x_ready(const ClState *cl, const Xch *x, unsigned long long seq, long long step, int stage, SelScratch *guard)
{
    if (threadIdx.x || blockIdx.x || guard->halt || !x || !cl->resample) return;
    if (stage & SEL_ST_PUB) xch_signal(x, XCH_CL_R, seq);
    if ((stage & SEL_ST_RUN) && !xch_wait(x, XCH_CL_R, seq)) sel_halt(guard, step, PF_HALT_TIMEOUT);
}

// arrays of every rank a rejuvenating child reads its ancestor from (single GPU: entry 0 = the local arrays)
struct ClPeers {
    const void *S[2][PF_MAX_RANKS];          // (elements of the handle's storage type)
    const int *Kc[2][PF_MAX_RANKS];
    const u128 *cum[PF_MAX_RANKS];
    const unsigned char *asg_tmp[PF_MAX_RANKS];
};

// normalise (src/chenliu.jl:65-66) or rejuvenate (src/chenliu.jl:61, stratified variant): child gi takes the ancestor
//   a = min{k : cum_k N 2^53 > (gi 2^53 + floor(u 2^53)) Qtot} = min{k : cum_k > floor((gi 2^53 + mu) Qtot / (N 2^53))}
// found by a search over the ranks' totals and a binary search in the owner's prefix array
template <int D, typename R = double>
__global__ void __launch_bounds__(256) k_cl_finish(R *S0, R *S1, int *Kc0, int *Kc1, double *logw0, double *logw1,
                                                   long long cap, ClShard sh, int nranks, const ClState *__restrict__ cl,
                                                   const double *__restrict__ W, const ClPeers *__restrict__ peers,
                                                   const double *__restrict__ u_dev, u64 seed, u64 step,
                                                   const unsigned char *__restrict__ asg_tmp, unsigned int *__restrict__ gen_anc,
                                                   unsigned char *__restrict__ gen_asg, const SelScratch *guard)
{
    using L = Layout<D>;
    if (guard->halt) return;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= sh.n_loc) return;
    const long long gi = sh.goff + i;
    const int cur = cl->cur;
    if (!cl->resample) {
        double *logw = cur ? logw1 : logw0;
        logw[i] = log(W[i]) - cl->log_total;
        if (gen_anc) { gen_anc[i] = (unsigned int)gi; gen_asg[i] = asg_tmp[i]; }
        return;
    }
    R *S2 = cur ? S0 : S1;
    int *Kc2 = cur ? Kc0 : Kc1;
    double *logw2 = cur ? logw0 : logw1;
    const double u = u_dev ? u_dev[i] : pf_uniform(seed, step, 2, (u64)gi);
    const u64 mu = (u64)(u * 9007199254740992.0);
    const u128 Rpos = add128(shl128(mk128((u64)gi, 0), 53), mk128(mu, 0));        // gi 2^53 + mu  (< 2^81)
    u64 X[4];
    mul128x128(Rpos, cl->Qtot, X);                                                     // < 2^168
    const u128 Xs = mk128((X[0] >> 53) | (X[1] << 11), (X[1] >> 53) | (X[2] << 11));   // X >> 53  (< 2^115)
    unsigned int rem;
    const u128 thr = divmod128_32(Xs, (unsigned int)sh.N, &rem);
    int r = 0;
    while (r < nranks - 1 && !lt128(thr, cl->Qincl[r])) r++;
    const u128 *cum = peers->cum[r];
    const long long n_r = (r + 1) * sh.q < sh.N ? sh.q : sh.N - (long long)r * sh.q;
    long long lo = 0, hi = n_r - 1;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        u128 c;
        c.lo = __ldcg(reinterpret_cast<const unsigned long long *>(&cum[mid].lo));
        c.hi = __ldcg(reinterpret_cast<const unsigned long long *>(&cum[mid].hi));
        if (lt128(thr, c)) hi = mid; else lo = mid + 1;
    }
    const long long a = lo;
    PF_ASSERT(r >= 0 && r < nranks && n_r > 0 && a >= 0 && a < n_r);
    const R *S = static_cast<const R *>(peers->S[cur][r]);
    const int K = __ldcg(peers->Kc[cur][r] + a);
    Kc2[i] = K;
    logw2[i] = -log((double)sh.N);                     // weights reset to 1/N
    if (gen_anc) { gen_anc[i] = (unsigned int)((long long)r * sh.q + a); gen_asg[i] = __ldcg(peers->asg_tmp[r] + a); }
    const R *src = S + a;
    R *dst = S2 + i;
    const int nrows = K * L::F;
    int row = 0;
    for (; row + 8 <= nrows; row += 8) {
        R tmp[8];
#pragma unroll
        for (int k = 0; k < 8; k++) tmp[k] = __ldcg(src + (long long)(row + k) * cap);
#pragma unroll
        for (int k = 0; k < 8; k++) dst[(long long)(row + k) * cap] = tmp[k];
    }
    for (; row < nrows; row++) dst[(long long)row * cap] = __ldcg(src + (long long)row * cap);
}

// step record; sharded: "this rank no longer reads the peers' stores" (exchange kind D)
__global__ void k_cl_steplog(StepState *st, ClState *cl, StepLog *lg, long long N, long long n_loc, const Xch *x, unsigned long long seq,
                             SelScratch *guard)
{
    if (threadIdx.x || blockIdx.x || guard->halt) return;
    StepLog r;
    r.n_in = N; r.M = 0; r.n_kept = 0; r.n_req = 0; r.n_got = 0; r.n_out = N; r.overflow = st->overflow; r.n_cand = 0; r.n_plateau = 0; r.status = 0; r.pad = 0;
    r.log_total = cl->log_total; r.lw_resamp = -log((double)N); r.thr = 0.0; r.cv = cl->cv;
    r.resampled = cl->resample; r.rounds = 0;
    for (int k = 0; k < 4; k++) r.sel_us[k] = 0.0;
    *lg = r;
    if (cl->resample) { cl->cur ^= 1; cl->n_resampled_steps++; }
    cl->resample = 0; cl->wmax_bits = 0;
    for (int k = 0; k < 4; k++) { cl->sumW.l[k] = 0; cl->sumSq.l[k] = 0; }
    st->n_live = n_loc; st->n_next = n_loc;
    if (x) xch_signal(x, XCH_D, seq);
}
