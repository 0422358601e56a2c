// fearnhead.cuh -- the per-observation kernels of fit!(::FearnheadParticles, y)
// (src/fearnhead.jl:130-184): expansion, survivor scan (systematic comb), instantiate.
#pragma once
#include "model.cuh"
#include "select.cuh"

// Device-resident control block of one filter (no host round trip inside a step).
struct StepState {
    long long n_live;       // population entering the step
    long long n_next;       // population leaving the step (written by k_scan)
    long long M;            // putatives of this step
    long long overflow;     // cumulative dropped new-cluster candidates
    long long step;         // observations consumed so far
    u64 vmax_bits;          // max putative weight (bit pattern)
    unsigned int ticket;    // tile ticket of k_scan
    unsigned int ticket2;   // tile ticket of the second k_scan of a sorted-order step
    long long n_picked;     // resampled survivors
    long long n_kept;       // kept survivors
};

struct StepLog {            // one record per step, read back after pf_filter
    long long n_in, M, n_kept, n_req, n_got, n_out, overflow, n_cand;
    double log_total, lw_resamp, thr, cv;
    int resampled, rounds;
};

// ------------------------------------------------------------------ k_prestep
// One thread: roll the control block over and derive the per-observation constants:
// the new-cluster putative (src/particle.jl:108-112 with comp = p.prior) and the slot a
// new cluster starts with (add(prior, y), src/component.jl:85).
template <int D>
__global__ void k_prestep(StepState *st, const PriorConst *pc, const double *y, StepConst *sc_out, int first)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    using L = Layout<D>;
    if (!first) st->n_live = st->n_next;
    st->n_next = 0;
    st->M = 0; st->vmax_bits = 0; st->ticket = 0; st->ticket2 = 0; st->n_picked = 0; st->n_kept = 0;
    StepConst c;
    double dx[D], z[D], dinv[D], off[L::NOFF > 0 ? L::NOFF : 1];
#pragma unroll
    for (int i = 0; i < D; i++) { c.y[i] = y[i]; dx[i] = y[i] - pc->mu0[i]; dinv[i] = pc->L0_dinv[i]; }
#pragma unroll
    for (int k = 0; k < L::NOFF; k++) off[k] = pc->L0_off[k];
    double q = quad_form<D>(dx, dinv, off, z);
    const NRow r0 = pc->row0;
    double l1p = log1p(r0.r * q);
    c.lw_new = r0.C - 0.5 * pc->logdet0 - r0.h * l1p;
    // new slot: n = 1, mean = y, L = cholupdate(L0, sqrt(r0) (y - mu0)), logdet = logdet0 + log1p(r0 q)
    double x[D];
    double sr = sqrt(r0.r);
#pragma unroll
    for (int i = 0; i < D; i++) x[i] = sr * dx[i];
    chol_rank1<D>(dinv, off, x);
    c.newslot[L::F_N] = 1.0;
    c.newslot[L::F_LOGDET] = pc->logdet0 + l1p;
#pragma unroll
    for (int i = 0; i < D; i++) { c.newslot[L::F_MEAN + i] = y[i]; c.newslot[L::F_DINV + i] = dinv[i]; }
#pragma unroll
    for (int k = 0; k < L::NOFF; k++) c.newslot[L::F_OFF + k] = off[k];
    *sc_out = c;
}

// ------------------------------------------------------------------ k_expand
// putatives(p, y) for every particle (src/fearnhead.jl:132, src/particle.jl:100-116):
// thread i evaluates the K_i + 1 candidate assignments of particle i and writes their
// LINEAR weights exp(logweight) slot-major into V.  One thread per particle keeps all 32
// lanes on the FP64 pipe and makes every (slot, field) load a 256-byte coalesced read.
template <int D>
__global__ void __launch_bounds__(256) k_expand(const double *__restrict__ S, const double *__restrict__ logw,
                                                const int *__restrict__ Kc, long long cap, int k_max,
                                                const NRow *__restrict__ tab, const PriorConst *__restrict__ pc,
                                                const StepConst *__restrict__ scp, StepState *st,
                                                double *__restrict__ V, unsigned char *__restrict__ cnt)
{
    using L = Layout<D>;
    __shared__ u64 s_max[8];
    __shared__ long long s_M[8], s_ovf[8];
    const long long n_live = st->n_live;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    u64 vmax = 0;
    long long myM = 0, myOvf = 0;
    if (i < n_live) {
        double y[D], mu0[D];
#pragma unroll
        for (int k = 0; k < D; k++) { y[k] = scp->y[k]; mu0[k] = pc->mu0[k]; }
        const int K = Kc[i];
        const double lw0 = logw[i];
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
            double lw = lw0 + slot_logw<D>(row, logdet, m, dinv, off, y, mu0, z, &q);
            double v = exp(lw);
            V[(long long)j * cap + i] = v;
            u64 b = (u64)__double_as_longlong(v);
            vmax = b > vmax ? b : vmax;
        }
        int c = K;
        if (K < k_max) {
            double v = exp(lw0 + scp->lw_new);
            V[(long long)K * cap + i] = v;
            u64 b = (u64)__double_as_longlong(v);
            vmax = b > vmax ? b : vmax;
            c = K + 1;
        } else {
            myOvf = 1;
        }
        cnt[i] = (unsigned char)c;
        myM = c;
    }
    vmax = warp_max64(vmax);
    myM = (long long)warp_sum64((u64)myM);
    myOvf = (long long)warp_sum64((u64)myOvf);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s_max[wid] = vmax; s_M[wid] = myM; s_ovf[wid] = myOvf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int nw = blockDim.x >> 5;
        for (int w = 1; w < nw; w++) { vmax = s_max[w] > vmax ? s_max[w] : vmax; myM += s_M[w]; myOvf += s_ovf[w]; }
        if (vmax) atomicMax(&st->vmax_bits, vmax);
        if (myM) atomicAdd((unsigned long long *)&st->M, (unsigned long long)myM);
        if (myOvf) atomicAdd((unsigned long long *)&st->overflow, (unsigned long long)myOvf);
    }
}

// ------------------------------------------------------------------ k_scan
// Single pass over the putatives in (particle, slot) order with two decoupled look-back
// chains: (1) the running fixed-point mass X = n * sum(q) of the low partition plus the
// number kept, (2) the number picked.  Item picked <=> the number of comb teeth
// {Uoff + k T} strictly below X grows (sample_stratified!, src/fearnhead.jl:97-114, in
// exact integer arithmetic).  Survivors are written in index order.
#define SCAN_THREADS 256
struct TileDesc {            // 32 bytes
    u128 X;
    unsigned int kept;
    unsigned int picked;
    unsigned int flag1;   // 2*epoch: aggregate of chain 1 published, 2*epoch+1: inclusive prefix published
    unsigned int flag2;   // same for chain 2 (picked); smaller values belong to earlier steps
};
struct TileIncl {
    u128 X;
    unsigned int kept;
    unsigned int picked;
    unsigned int pad[2];
};

__device__ __forceinline__ unsigned int ld_flag(const unsigned int *p) { return *(const volatile unsigned int *)p; }
__device__ __forceinline__ u128 ld_v128(const u128 *p)
{
    const volatile u64 *q = (const volatile u64 *)p;
    u128 r; r.lo = q[0]; r.hi = q[1]; return r;
}
__device__ __forceinline__ unsigned int ld_vu32(const unsigned int *p) { return *(const volatile unsigned int *)p; }

__global__ void __launch_bounds__(SCAN_THREADS) k_scan(const double *__restrict__ V, const unsigned char *__restrict__ cnt,
                                                       long long cap, StepState *st, const SelResult *__restrict__ res,
                                                       TileDesc *tiles, TileIncl *incl, unsigned int *__restrict__ surv_parent,
                                                       unsigned char *__restrict__ surv_slot, unsigned int epoch,
                                                       const long long *n_items_ptr, unsigned int *ticket, int mode)
{
    // mode 0: index-order filter step (always runs); mode 1: flat sorted list, runs only when
    // the step resamples, writes the PICKED positions only (kept = suffix of the sorted list);
    // mode 2: index-order pass of a sorted-order filter, runs only when the step keeps all.
    if (mode == 1 && res->keep_all) return;
    if (mode == 2 && !res->keep_all) return;
    __shared__ u128 s_warp[33];
    __shared__ unsigned int s_wc[33];
    __shared__ unsigned int s_tile;
    __shared__ u128 s_preX;
    __shared__ unsigned int s_preK, s_preP;
    const long long n_live = *n_items_ptr;
    const long long ntiles = (n_live + SCAN_THREADS - 1) / SCAN_THREADS;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    if (tile >= ntiles) return;
    const unsigned int F_AGG = 2u * epoch, F_INC = 2u * epoch + 1u;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long i = (long long)tile * SCAN_THREADS + tid;
    const u64 thr = res->thr_bits;
    const int scale = res->scale;
    const u64 n = (u64)res->n;
    const u128 T = res->T, U = res->Uoff;
    const bool comb = n > 0 && !isz128(T);
    const int c = (i < n_live) ? (cnt ? cnt[i] : 1) : 0;

    // pass A: thread totals
    u128 myX = mk128(0, 0);
    unsigned int myKept = 0;
    for (int j = 0; j < c; j++) {
        u64 bits = (u64)__double_as_longlong(V[(long long)j * cap + i]);
        if (bits >= thr) myKept++;
        else if (comb) myX = add128(myX, mul128_64(fx128(bits, scale), n));
    }
    u128 totX;
    u128 exX = block_excl_scan128(myX, s_warp, &totX);
    unsigned int incK = myKept;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { unsigned int o = __shfl_up_sync(0xffffffffu, incK, d); if (lane >= d) incK += o; }
    if (lane == 31) s_wc[wid] = incK;
    __syncthreads();
    if (wid == 0) {
        unsigned int w = lane < (SCAN_THREADS / 32) ? s_wc[lane] : 0, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned int o = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += o; }
        s_wc[lane] = wi - w;
        if (lane == 31) s_wc[32] = wi;
    }
    __syncthreads();
    unsigned int exK = s_wc[wid] + incK - myKept;
    unsigned int totK = s_wc[32];
    __syncthreads();

    // chain 1: publish the tile aggregate, look back for the exclusive prefix
    if (tid == 0) {
        TileDesc *td = &tiles[tile];
        u128 pX = mk128(0, 0); unsigned int pK = 0;
        if (tile > 0) {
            td->X = totX; td->kept = totK;
            __threadfence();
            *(volatile unsigned int *)&td->flag1 = F_AGG;
            long long t = (long long)tile - 1;
            while (true) {
                unsigned int f;
                while ((f = ld_flag(&tiles[t].flag1)) < F_AGG) { }
                __threadfence();
                if (f == F_INC) { pX = add128(pX, ld_v128(&incl[t].X)); pK += ld_vu32(&incl[t].kept); break; }
                pX = add128(pX, ld_v128(&tiles[t].X)); pK += ld_vu32(&tiles[t].kept);
                t--;
            }
        }
        incl[tile].X = add128(pX, totX); incl[tile].kept = pK + totK;
        __threadfence();
        *(volatile unsigned int *)&td->flag1 = F_INC;
        s_preX = pX; s_preK = pK;
    }
    __syncthreads();
    const u128 X0 = add128(s_preX, exX);

    // pass B: decide the picks among this thread's items
    unsigned long long pickmask = 0;
    unsigned int myPicked = 0;
    if (comb && c > 0) {
        u128 X = X0;
        u64 t = teeth_below(X, U, T);
        u128 P = add128(U, mul128_64(T, t));          // position of the next tooth not yet passed
        for (int j = 0; j < c; j++) {
            u64 bits = (u64)__double_as_longlong(V[(long long)j * cap + i]);
            if (bits < thr) {
                X = add128(X, mul128_64(fx128(bits, scale), n));
                if (lt128(P, X)) {
                    pickmask |= 1ull << j; myPicked++;
                    do { P = add128(P, T); } while (lt128(P, X));
                }
            }
        }
    }
    unsigned int incP = myPicked;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { unsigned int o = __shfl_up_sync(0xffffffffu, incP, d); if (lane >= d) incP += o; }
    if (lane == 31) s_wc[wid] = incP;
    __syncthreads();
    if (wid == 0) {
        unsigned int w = lane < (SCAN_THREADS / 32) ? s_wc[lane] : 0, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned int o = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += o; }
        s_wc[lane] = wi - w;
        if (lane == 31) s_wc[32] = wi;
    }
    __syncthreads();
    unsigned int exP = s_wc[wid] + incP - myPicked;
    unsigned int totP = s_wc[32];

    // chain 2: picked counts
    if (tid == 0) {
        TileDesc *td = &tiles[tile];
        unsigned int pP = 0;
        if (tile > 0) {
            td->picked = totP;
            __threadfence();
            *(volatile unsigned int *)&td->flag2 = F_AGG;
            long long t = (long long)tile - 1;
            while (true) {
                unsigned int f;
                while ((f = ld_flag(&tiles[t].flag2)) < F_AGG) { }
                __threadfence();
                if (f == F_INC) { pP += ld_vu32(&incl[t].picked); break; }
                pP += ld_vu32(&tiles[t].picked);
                t--;
            }
        }
        incl[tile].picked = pP + totP;
        __threadfence();
        *(volatile unsigned int *)&td->flag2 = F_INC;
        s_preP = pP;
        if (tile == ntiles - 1) {
            st->n_kept = (long long)(s_preK + totK);
            st->n_picked = (long long)(pP + totP);
            st->n_next = (long long)(s_preK + totK) + (long long)(pP + totP);
        }
    }
    __syncthreads();
    // write survivors in index order
    if (c > 0) {
        long long pos = (mode == 1) ? (long long)s_preP + exP : (long long)s_preK + exK + (long long)s_preP + exP;
        for (int j = 0; j < c; j++) {
            u64 bits = (u64)__double_as_longlong(V[(long long)j * cap + i]);
            bool kept = (mode == 1) ? false : bits >= thr;
            bool picked = (pickmask >> j) & 1ull;
            if (kept || picked) {
                surv_parent[pos] = (unsigned int)i;
                surv_slot[pos] = (unsigned char)(j | (picked ? 0x80 : 0));
                pos++;
            }
        }
    }
}

// ------------------------------------------------------------------ k_instantiate
// instantiate.(survivors) (src/particle.jl:134-149): thread t builds particle t of the new
// population from its parent's slots (copy), applies add(comp, y) to the chosen slot as a
// Welford mean update + rank-1 Cholesky update, appends (ancestor, assignment) to the
// genealogy and sets the normalised log-weight (src/fearnhead.jl:172-179).
template <int D>
__global__ void __launch_bounds__(256) k_instantiate(const double *__restrict__ S, const int *__restrict__ Kc,
                                                     double *__restrict__ S2, int *__restrict__ Kc2,
                                                     double *__restrict__ logw2, long long cap,
                                                     const NRow *__restrict__ tab, const PriorConst *__restrict__ pc,
                                                     const StepConst *__restrict__ scp, const StepState *__restrict__ st,
                                                     const SelResult *__restrict__ res, const double *__restrict__ V,
                                                     const unsigned int *__restrict__ surv_parent,
                                                     const unsigned char *__restrict__ surv_slot,
                                                     unsigned int *__restrict__ gen_anc, unsigned char *__restrict__ gen_asg)
{
    using L = Layout<D>;
    const long long n_next = st->n_next;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_next) return;
    const unsigned int par = surv_parent[t];
    const unsigned char sl = surv_slot[t];
    const int j = sl & 0x7F;
    const bool picked = sl & 0x80;
    const int K = Kc[par];
    const double v = V[(long long)j * cap + par];
    logw2[t] = picked ? res->lw_resamp : (log(v) - res->log_total);
    Kc2[t] = K + (j == K ? 1 : 0);
    if (gen_anc) { gen_anc[t] = par; gen_asg[t] = (unsigned char)j; }
    for (int jj = 0; jj < K; jj++) {
        const double *src = S + ((long long)jj * L::F) * cap + par;
        double *dst = S2 + ((long long)jj * L::F) * cap + t;
        if (jj != j) {
#pragma unroll
            for (int f = 0; f < L::F; f++) dst[(long long)f * cap] = src[(long long)f * cap];
        } else {
            double n = src[(long long)L::F_N * cap];
            double logdet = src[(long long)L::F_LOGDET * cap];
            double m[D], dinv[D], off[L::NOFF > 0 ? L::NOFF : 1], dx[D], z[D];
#pragma unroll
            for (int k = 0; k < D; k++) { m[k] = src[(long long)(L::F_MEAN + k) * cap]; dinv[k] = src[(long long)(L::F_DINV + k) * cap]; }
#pragma unroll
            for (int k = 0; k < L::NOFF; k++) off[k] = src[(long long)(L::F_OFF + k) * cap];
            const NRow row = tab[(int)n];
#pragma unroll
            for (int k = 0; k < D; k++) dx[k] = scp->y[k] - (pc->mu0[k] + row.g * (m[k] - pc->mu0[k]));
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
                double delta = scp->y[k] - m[k];                       // Welford, src/component.jl:46-47
                dst[(long long)(L::F_MEAN + k) * cap] = m[k] + delta / tw;
                dst[(long long)(L::F_DINV + k) * cap] = dinv[k];
            }
#pragma unroll
            for (int k = 0; k < L::NOFF; k++) dst[(long long)(L::F_OFF + k) * cap] = off[k];
        }
    }
    if (j == K) {
        double *dst = S2 + ((long long)K * L::F) * cap + t;
#pragma unroll
        for (int f = 0; f < L::F; f++) dst[(long long)f * cap] = scp->newslot[f];
    }
}

// ------------------------------------------------------------------ reference-literal order
// flat (particle-major, candidate-minor) key list; dead slots and dead particles get the
// all-ones pattern so that a stable ascending sort leaves the M live putatives first, ties
// in putative-index order (the declared tie-break of sort!, src/fearnhead.jl:148).
__global__ void k_flat_keys(const double *__restrict__ V, const unsigned char *__restrict__ cnt, long long cap,
                            const StepState *__restrict__ st, int slots, long long n_pad, u64 *__restrict__ keys)
{
    long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pad) return;
    long long i = k / slots;
    int j = (int)(k - i * slots);
    u64 key = ~0ull;
    if (i < st->n_live && j < cnt[i]) key = (u64)__double_as_longlong(V[(long long)j * cap + i]);
    keys[k] = key;
}

// survivors of a sorted-order step in the reference's output order
// [resampled (ascending weight) ; kept (ascending weight)] (src/fearnhead.jl:172-179)
__global__ void k_sorted_finalize(const StepState *__restrict__ st, const SelResult *__restrict__ res,
                                  const unsigned int *__restrict__ pick_pos, const unsigned int *__restrict__ vals, int slots,
                                  unsigned int *__restrict__ surv_parent, unsigned char *__restrict__ surv_slot)
{
    if (res->keep_all) return;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= st->n_next) return;
    const long long np = st->n_picked;
    const bool picked = t < np;
    const long long flat = picked ? (long long)pick_pos[t] : (st->M - res->A) + (t - np);
    const unsigned int code = vals[flat];
    surv_parent[t] = code / (unsigned int)slots;
    surv_slot[t] = (unsigned char)((code % (unsigned int)slots) | (picked ? 0x80 : 0));
}

// fill the per-step log record (one thread)
__global__ void k_steplog(const StepState *st, const SelResult *res, StepLog *lg, long long N)
{
    if (threadIdx.x || blockIdx.x) return;
    StepLog r;
    r.n_in = st->n_live; r.M = st->M; r.n_kept = st->n_kept; r.n_req = res->keep_all ? 0 : res->n;
    r.n_got = st->n_picked; r.n_out = st->n_next; r.overflow = st->overflow; r.n_cand = res->n_cand_first;
    r.log_total = res->log_total; r.lw_resamp = res->lw_resamp; r.thr = res->thr_value; r.cv = 0.0;
    r.resampled = res->keep_all ? 0 : 1; r.rounds = res->rounds;
    (void)N;
    *lg = r;
}
