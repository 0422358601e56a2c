// chenliu.cuh -- fit!(::ChenLiuParticles, y) (src/chenliu.jl:54-69): per-particle
// propagation with one categorical draw (propogate_chenliu, :28-42), coefficient of
// variation (:49-52), CV-triggered stratified rejuvenation or weight normalisation.
#pragma once
#include "../../include/particles_b200.h"
#include "fearnhead.cuh"

#define CL_W_BITS 60          // rejuvenation fixed point: q = floor(W 2^(CL_W_BITS - e(max W)))

struct ClState {
    int cur;                  // which store buffer holds the population
    int resample;             // decision of the current step
    u64 wmax_bits;
    acc128 sumW;              // sum of W at scale SEL_TOT_BITS - e(wmax)
    double mean, cv, log_total;
    u128 Qtot;                // sum of q over the population (rejuvenation scale)
    int qscale;
    long long n_resampled_steps;
};

// ------------------------------------------------------------------ k_cl_propagate
template <int D>
__global__ void __launch_bounds__(256) k_cl_propagate(double *S0, double *S1, const double *__restrict__ logw0,
                                                      const double *__restrict__ logw1, int *Kc0, int *Kc1, long long cap,
                                                      long long N, int k_max, const NRow *__restrict__ tab,
                                                      const PriorConst *__restrict__ pc, const StepConst *__restrict__ scp,
                                                      StepState *st, ClState *cl, const double *__restrict__ u_dev, u64 seed,
                                                      u64 step, double *__restrict__ W, unsigned char *__restrict__ asg_tmp)
{
    using L = Layout<D>;
    __shared__ u64 s_max[8];
    __shared__ long long s_ovf[8];
    const int cur = cl->cur;
    double *S = cur ? S1 : S0;
    const double *logw = cur ? logw1 : logw0;
    int *Kc = cur ? Kc1 : Kc0;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    u64 wmax = 0;
    long long myOvf = 0;
    if (i < N) {
        double y[D], mu0[D];
#pragma unroll
        for (int k = 0; k < D; k++) { y[k] = scp->y[k]; mu0[k] = pc->mu0[k]; }
        const int K = Kc[i];
        const double lw0 = logw[i];
        double v[PF_MAX_SLOTS];
        for (int j = 0; j < K; j++) {
            const double *base = S + ((long long)j * L::F) * cap + i;
            double n = base[(long long)L::F_N * cap];
            double logdet = base[(long long)L::F_LOGDET * cap];
            double m[D], dinv[D], off[L::NOFF > 0 ? L::NOFF : 1], z[D], q;
#pragma unroll
            for (int k = 0; k < D; k++) { m[k] = base[(long long)(L::F_MEAN + k) * cap]; dinv[k] = base[(long long)(L::F_DINV + k) * cap]; }
#pragma unroll
            for (int k = 0; k < L::NOFF; k++) off[k] = base[(long long)(L::F_OFF + k) * cap];
            const NRow row = tab[(int)n];
            v[j] = exp(lw0 + slot_logw<D>(row, logdet, m, dinv, off, y, mu0, z, &q));
        }
        int C = K;
        if (K < k_max) { v[K] = exp(lw0 + scp->lw_new); C = K + 1; } else myOvf = 1;
        // weights = weight.(ps); sum(weights)  (src/chenliu.jl:34,41)
        double sum = v[0];
        for (int j = 1; j < C; j++) sum += v[j];
        // wsample: StatsBase inverse-CDF draw (src/chenliu.jl:35)
        double u = u_dev ? u_dev[i] : pf_uniform(seed, step, 1, (u64)i);
        double t = u * sum;
        int pick = 0;
        double cw = v[0];
        while (cw < t && pick < C - 1) { pick++; cw += v[pick]; }
        // instantiate in place (src/particle.jl:134-149)
        double *dst = S + ((long long)pick * L::F) * cap + i;
        if (pick == K) {
#pragma unroll
            for (int f = 0; f < L::F; f++) dst[(long long)f * cap] = scp->newslot[f];
            Kc[i] = K + 1;
        } else {
            double n = dst[(long long)L::F_N * cap];
            double logdet = dst[(long long)L::F_LOGDET * cap];
            double m[D], dinv[D], off[L::NOFF > 0 ? L::NOFF : 1], dx[D], z[D];
#pragma unroll
            for (int k = 0; k < D; k++) { m[k] = dst[(long long)(L::F_MEAN + k) * cap]; dinv[k] = dst[(long long)(L::F_DINV + k) * cap]; }
#pragma unroll
            for (int k = 0; k < L::NOFF; k++) off[k] = dst[(long long)(L::F_OFF + k) * cap];
            const NRow row = tab[(int)n];
#pragma unroll
            for (int k = 0; k < D; k++) dx[k] = y[k] - (mu0[k] + row.g * (m[k] - mu0[k]));
            double q = quad_form<D>(dx, dinv, off, z);
            double sr = sqrt(row.r);
#pragma unroll
            for (int k = 0; k < D; k++) dx[k] *= sr;
            chol_rank1<D>(dinv, off, dx);
            double tw = n + 1.0;
            dst[(long long)L::F_N * cap] = tw;
            dst[(long long)L::F_LOGDET * cap] = logdet + log1p(row.r * q);
#pragma unroll
            for (int k = 0; k < D; k++) {
                double delta = y[k] - m[k];
                dst[(long long)(L::F_MEAN + k) * cap] = m[k] + delta / tw;
                dst[(long long)(L::F_DINV + k) * cap] = dinv[k];
            }
#pragma unroll
            for (int k = 0; k < L::NOFF; k++) dst[(long long)(L::F_OFF + k) * cap] = off[k];
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

// sum of W in fixed point
__global__ void __launch_bounds__(256) k_cl_sum(const double *__restrict__ W, long long N, ClState *cl)
{
    __shared__ u128 s_red[8];
    const int sc = SEL_TOT_BITS - exp_of_bits(cl->wmax_bits);
    u128 acc = mk128(0, 0);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
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

// centred sum of squares, per-block partials in a fixed order (mean_and_std, two-pass)
__global__ void __launch_bounds__(256) k_cl_var(const double *__restrict__ W, long long N, const ClState *cl, double *part)
{
    __shared__ double s_red[256];
    const int sc = SEL_TOT_BITS - exp_of_bits(cl->wmax_bits);
    const double mean = ldexp(to_double128(acc_value(&cl->sumW)), -sc) / (double)N;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        double dlt = W[i] - mean;
        acc += dlt * dlt;
    }
    s_red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) s_red[threadIdx.x] += s_red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = s_red[0];
}

// cv = 100 std / mean (corrected), decision cv > threshold (src/chenliu.jl:57-59)
__global__ void k_cl_decide(ClState *cl, const double *part, int nparts, long long N, double thr)
{
    if (threadIdx.x || blockIdx.x) return;
    const int sc = SEL_TOT_BITS - exp_of_bits(cl->wmax_bits);
    u128 Q = acc_value(&cl->sumW);
    double tot = ldexp(to_double128(Q), -sc);
    double mean = tot / (double)N;
    double ss = 0.0;
    for (int k = 0; k < nparts; k++) ss += part[k];
    double sd = N > 1 ? sqrt(ss / (double)(N - 1)) : 0.0;
    double cv = N > 1 ? sd / mean * 100.0 : 0.0;
    cl->mean = mean; cl->cv = cv;
    cl->log_total = log(to_double128(Q)) - (double)sc * 0.693147180559945309417232121458;
    cl->resample = cv > thr ? 1 : 0;
    cl->qscale = CL_W_BITS - exp_of_bits(cl->wmax_bits);
}

// inclusive fixed-point prefix sums of W (only when rejuvenating); one block
__global__ void __launch_bounds__(1024) k_cl_scan(const double *__restrict__ W, long long N, ClState *cl, u128 *__restrict__ cum)
{
    if (!cl->resample) return;
    __shared__ u128 s_warp[33];
    const int sc = cl->qscale;
    const long long per = (N + blockDim.x - 1) / blockDim.x;
    const long long b = (long long)threadIdx.x * per, e = b + per < N ? b + per : N;
    u128 loc = mk128(0, 0);
    for (long long i = b; i < e; i++) loc = add128(loc, fx128((u64)__double_as_longlong(W[i]), sc));
    u128 tot;
    u128 run = block_excl_scan128(loc, s_warp, &tot);
    for (long long i = b; i < e; i++) { run = add128(run, fx128((u64)__double_as_longlong(W[i]), sc)); cum[i] = run; }
    if (threadIdx.x == 0) cl->Qtot = tot;
}

// normalise (src/chenliu.jl:65-66) or rejuvenate (src/chenliu.jl:61, stratified variant)
template <int D>
__global__ void __launch_bounds__(256) k_cl_finish(double *S0, double *S1, int *Kc0, int *Kc1, double *logw0, double *logw1,
                                                   long long cap, long long N, const ClState *__restrict__ cl,
                                                   const double *__restrict__ W, const u128 *__restrict__ cum,
                                                   const double *__restrict__ u_dev, u64 seed, u64 step,
                                                   const unsigned char *__restrict__ asg_tmp, unsigned int *__restrict__ gen_anc,
                                                   unsigned char *__restrict__ gen_asg)
{
    using L = Layout<D>;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int cur = cl->cur;
    if (!cl->resample) {
        double *logw = cur ? logw1 : logw0;
        logw[i] = log(W[i]) - cl->log_total;
        if (gen_anc) { gen_anc[i] = (unsigned int)i; gen_asg[i] = asg_tmp[i]; }
        return;
    }
    const double *S = cur ? S1 : S0;
    double *S2 = cur ? S0 : S1;
    const int *Kc = cur ? Kc1 : Kc0;
    int *Kc2 = cur ? Kc0 : Kc1;
    double *logw2 = cur ? logw0 : logw1;
    // ancestor = min{k : cum_k * N * 2^53 > (i 2^53 + floor(u 2^53)) * Qtot}
    double u = u_dev ? u_dev[i] : pf_uniform(seed, step, 2, (u64)i);
    u64 mu = (u64)(u * 9007199254740992.0);
    // R = i * 2^53 + mu  (< 2^81)
    u128 R = add128(shl128(mk128((u64)i, 0), 53), mk128(mu, 0));
    u64 rhs[4];
    mul128x128(R, cl->Qtot, rhs);
    u128 NS = shl128(mk128((u64)N, 0), 53);
    long long lo = 0, hi = N - 1;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        u64 lhs[4];
        mul128x128(cum[mid], NS, lhs);
        if (gt256(lhs, rhs)) hi = mid; else lo = mid + 1;
    }
    const long long a = lo;
    const int K = Kc[a];
    Kc2[i] = K;
    logw2[i] = -log((double)N);                        // weights reset to 1/N
    if (gen_anc) { gen_anc[i] = (unsigned int)a; gen_asg[i] = asg_tmp[a]; }
    for (int jj = 0; jj < K; jj++) {
        const double *src = S + ((long long)jj * L::F) * cap + a;
        double *dst = S2 + ((long long)jj * L::F) * cap + i;
#pragma unroll
        for (int f = 0; f < L::F; f++) dst[(long long)f * cap] = src[(long long)f * cap];
    }
}

__global__ void k_cl_steplog(StepState *st, ClState *cl, StepLog *lg, long long N)
{
    if (threadIdx.x || blockIdx.x) return;
    StepLog r;
    r.n_in = N; r.M = 0; r.n_kept = 0; r.n_req = 0; r.n_got = 0; r.n_out = N; r.overflow = st->overflow; r.n_cand = 0; r.n_plateau = 0; r.status = 0; r.pad = 0;
    r.log_total = cl->log_total; r.lw_resamp = -log((double)N); r.thr = 0.0; r.cv = cl->cv;
    r.resampled = cl->resample; r.rounds = 0;
    for (int k = 0; k < 4; k++) r.sel_us[k] = 0.0;
    *lg = r;
    if (cl->resample) { cl->cur ^= 1; cl->n_resampled_steps++; }
    cl->resample = 0; cl->wmax_bits = 0;
    for (int k = 0; k < 4; k++) cl->sumW.l[k] = 0;
    st->n_live = N; st->n_next = N;
}
