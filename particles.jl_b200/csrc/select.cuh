// select.cuh -- Fearnhead-Clifford threshold search (cutoff_ascending, src/fearnhead.jl:59-77)
// as ONE cooperative persistent kernel: sample -> bracket -> classify pass -> radix descent
// on the bracket's candidates -> exact solve.  All sums that feed a decision are 128-bit
// fixed-point integers (order-free, hence deterministic and shard-independent).
//
// Items are the putative weights of the current step, addressed as (particle i, slot j):
//     value = V[j * cap + i],   j < cnt[i]   (cnt == nullptr: one item per index)
//
// Result (SelResult): kept set = {v >= thr}; everything else is the low partition with
// fixed-point total T at scale 2^scale; n = N - A particles are to be drawn from it by the
// systematic comb of k_scan (sample_stratified!, src/fearnhead.jl:97-114).
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
namespace cg = cooperative_groups;

#define SEL_THREADS 512
#define SEL_SAMPLE_CAP 8192
#define SEL_FINAL_CAP 2048
#define SEL_BINS 1024
#define SEL_CHUNK 256
#define SEL_ITEM_BITS 69            // items q = floor(v 2^scale) < 2^(SEL_ITEM_BITS+1)
#define SEL_TOT_BITS 89             // total-weight accumulator scale relative to vmax

struct SelResult {
    u64 thr_bits;                   // kept = {bits(v) >= thr_bits}; 0 = keep everything
    int scale;                      // q = floor(v * 2^scale) for low items
    int keep_all;                   // M <= N (src/fearnhead.jl:135-138)
    u128 T;                         // sum of q over the low partition
    u128 Uoff;                      // floor(u * T): comb offset
    long long A;                    // number kept
    long long n;                    // number to draw (N - A)
    double log_total;               // log(sum of all putative weights)
    double lw_resamp;               // log(c / total), c = T / n      (src/fearnhead.jl:173)
    double log_c;                   // log(c)
    double thr_value;
    int rounds;
    long long n_cand_first;
    int status;                     // 0 ok
};

struct SelScratch {
    // bracket state (written by block 0, read by all after grid.sync)
    u64 lo, hi;
    int scale;
    int phase;                      // control word for the next loop iteration
    int mode;                       // 0 bracket round, 1 keep everything, 2 comb-rescale round (lo = hi = thr)
    long long count_in;             // candidates inside [lo, hi)
    // accumulators of the classify pass
    unsigned long long A_hi;        // #items >= hi
    unsigned long long n_cand;      // #items in [lo, hi)
    acc128 B_lo;                    // sum q of items < lo
    acc128 S_cand;                  // sum q of items in [lo, hi)
    acc128 Tot;                     // sum of all items at scale SEL_TOT_BITS - e(vmax)
    // running totals of the descent
    u128 B_run;                     // sum q of items < lo (after narrowing)
    long long A_run;                // #items >= hi (after narrowing)
    // candidate list
    unsigned long long cand_chunks; // chunks handed out
    // final list
    unsigned int final_n;
    double final_vals[SEL_FINAL_CAP];
    // histogram of one descent level
    unsigned int h_cnt[SEL_BINS];
    u64 h_sum[3][SEL_BINS];
    int h_shift;
};

struct SelArgs {
    const double *V;
    const unsigned char *cnt;
    long long cap;
    const long long *n_live;        // device
    const long long *M;             // device: total number of items
    const u64 *vmax_bits;           // device
    long long N;
    const double *u;                // device: the uniform of this step
    double *cand;                   // candidate list, chunked
    long long cand_cap;             // in doubles
    SelScratch *sc;
    SelResult *res;
};

__device__ __forceinline__ bool pred128(u128 B, long long N, long long A_geq, u128 qx)
{
    long long dd = N - A_geq;
    if (dd < 0) return false;
    return le128(B, mul128_64(qx, (u64)dd));
}

// exclusive scan of one u128 per thread over the block; returns exclusive prefix, *total = block sum
__device__ inline u128 block_excl_scan128(u128 v, u128 *smem_warp /* >= 32 */, u128 *total)
{
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    u128 inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u128 o = shfl_up128(inc, d);
        if (lane >= d) inc = add128(inc, o);
    }
    if (lane == 31) smem_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        u128 w = lane < nw ? smem_warp[lane] : mk128(0, 0);
        u128 winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u128 o = shfl_up128(winc, d);
            if (lane >= d) winc = add128(winc, o);
        }
        smem_warp[lane] = sub128(winc, w);          // exclusive warp offsets
        if (lane == 31) smem_warp[32] = winc;        // total
    }
    __syncthreads();
    u128 off = smem_warp[wid];
    *total = smem_warp[32];
    u128 r = add128(off, sub128(inc, v));
    __syncthreads();
    return r;
}

__device__ inline void bitonic_sort_smem(u64 *a, int n /* power of two */)
{
    for (int k = 2; k <= n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < n; t += blockDim.x) {
                int ixj = t ^ j;
                if (ixj > t) {
                    u64 x = a[t], y = a[ixj];
                    bool up = (t & k) == 0;
                    if ((x > y) == up) { a[t] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
}

// The comb needs the low partition's mass T at a scale fine relative to c = T / n (every
// low weight is < c, so c bounds the items).  The threshold round's scale follows the
// threshold instead, which can sit many binades above c when a gap separates the kept
// group from the rest.  Returns true (and programs a mode-2 round: lo = hi = thr at a
// finer scale) when T has fewer than 52 bits per drawn particle.
__device__ inline bool comb_rescale(SelScratch *sc, const SelResult *res, long long N, u64 vmax)
{
    const long long n = N - res->A;
    if (n <= 0) return false;
    const u128 T = res->T;
    const int scale = res->scale;
    const int max_scale = 1074 + SEL_ITEM_BITS;
    int nscale;
    if (isz128(T)) {
        if (scale >= max_scale) return false;               // the low partition really is all zero
        nscale = scale + 64;
    } else {
        if (!lt128(T, shl128(mk128((u64)n, 0), 52))) return false;
        int blT = T.hi ? 128 - __clzll((long long)T.hi) : 64 - __clzll((long long)T.lo);
        int bln = 64 - __clzll(n);
        int e_c = blT - bln + 1 - scale;                    // c < 2^(e_c + 1)
        nscale = SEL_ITEM_BITS - 3 - e_c;
        if (nscale <= scale) return false;
    }
    if (nscale > max_scale) nscale = max_scale;
    (void)vmax;
    sc->lo = res->thr_bits; sc->hi = res->thr_bits; sc->scale = nscale; sc->mode = 2;
    sc->A_hi = 0; sc->n_cand = 0; sc->cand_chunks = 0; sc->final_n = 0;
    for (int k = 0; k < 4; k++) { sc->B_lo.l[k] = 0; sc->S_cand.l[k] = 0; }
    return true;
}

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SEL_THREADS, 2) k_select(SelArgs a)
{
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u64 *s_u64 = reinterpret_cast<u64 *>(smem_raw);               // SEL_SAMPLE_CAP u64 (sample / final sort)
    __shared__ u128 s_warp[33];
    __shared__ unsigned int s_cnt;
    __shared__ long long s_ll[4];
    __shared__ u128 s_red[3][SEL_THREADS / 32];
    __shared__ unsigned long long s_redc[2][SEL_THREADS / 32];

    SelScratch *sc = a.sc;
    SelResult *res = a.res;
    const long long n_live = *a.n_live;
    const long long M = *a.M;
    const u64 vmax = *a.vmax_bits;
    const long long N = a.N;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long gwarp = (long long)blockIdx.x * (SEL_THREADS / 32) + wid;
    const long long nwarps = (long long)gridDim.x * (SEL_THREADS / 32);
    const int e_vmax = exp_of_bits(vmax);
    const int tot_scale = SEL_TOT_BITS - e_vmax;

    // ---------------- phase 0: first bracket (block 0)
    if (blockIdx.x == 0) {
        if (tid == 0) {
            sc->A_hi = 0; sc->n_cand = 0; sc->cand_chunks = 0; sc->final_n = 0;
            for (int k = 0; k < 4; k++) { sc->B_lo.l[k] = 0; sc->S_cand.l[k] = 0; sc->Tot.l[k] = 0; }
            res->rounds = 0; res->status = 0; res->n_cand_first = 0;
        }
        if (M <= N || vmax == 0) {
            // keep everything (src/fearnhead.jl:135-138); an all-zero step keeps nothing
            if (tid == 0) { sc->lo = PF_INF_BITS; sc->hi = PF_INF_BITS; sc->scale = SEL_ITEM_BITS - e_vmax; sc->phase = 1; sc->mode = 1; }
        } else if (M <= SEL_FINAL_CAP) {
            if (tid == 0) { sc->lo = 1; sc->hi = PF_INF_BITS; sc->scale = SEL_ITEM_BITS - e_vmax; sc->phase = 1; sc->mode = 0; }
        } else {
            // strided sample of whole particles
            if (tid == 0) s_cnt = 0;
            __syncthreads();
            long long stride = (M + (SEL_SAMPLE_CAP * 7 / 8) - 1) / (SEL_SAMPLE_CAP * 7 / 8);
            if (stride < 1) stride = 1;
            for (long long p = tid; p * stride < n_live; p += SEL_THREADS) {
                long long i = p * stride;
                int c = a.cnt ? a.cnt[i] : 1;
                for (int j = 0; j < c; j++) {
                    unsigned int pos = atomicAdd(&s_cnt, 1u);
                    if (pos < SEL_SAMPLE_CAP) s_u64[pos] = __double_as_longlong(a.V[(long long)j * a.cap + i]);
                }
            }
            __syncthreads();
            int m = min(s_cnt, (unsigned)SEL_SAMPLE_CAP);
            for (int t = m + tid; t < SEL_SAMPLE_CAP; t += SEL_THREADS) s_u64[t] = ~0ull;
            __syncthreads();
            bitonic_sort_smem(s_u64, SEL_SAMPLE_CAP);
            // approximate solve on the sample (FP64): rho = M / m
            double rho = (double)M / (double)m;
            const int per = SEL_SAMPLE_CAP / SEL_THREADS;
            double loc = 0.0;
            for (int k = 0; k < per; k++) { int t = tid * per + k; if (t < m) loc += __longlong_as_double(s_u64[t]); }
            // block exclusive scan of doubles through the u128 helper is overkill: do it in smem
            double *s_d = reinterpret_cast<double *>(s_red);
            __shared__ double s_pref[SEL_THREADS];
            s_pref[tid] = loc;
            __syncthreads();
            if (tid == 0) { double run = 0.0; for (int t = 0; t < SEL_THREADS; t++) { double x = s_pref[t]; s_pref[t] = run; run += x; } }
            __syncthreads();
            (void)s_d;
            double run = s_pref[tid];
            int first = m;
            for (int k = 0; k < per; k++) {
                int t = tid * per + k;
                if (t < m) {
                    double w = __longlong_as_double(s_u64[t]);
                    if (w != 0.0 && first == m) {
                        double ngeq = rho * (double)(m - t);
                        if (rho * run <= ((double)N - ngeq) * w) first = t;
                    }
                    run += w;
                }
            }
            if (tid == 0) s_ll[0] = m;
            __syncthreads();
            atomicMin(&s_ll[0], (long long)first);
            __syncthreads();
            if (tid == 0) {
                int istar = (int)s_ll[0];
                int delta = (int)(3.0 * sqrt((double)m)) + 8;
                int il = istar - delta, ih = istar + delta;
                u64 lo = il <= 0 ? 1ull : s_u64[il];
                u64 hi = ih >= m ? PF_INF_BITS : s_u64[ih];
                if (lo < 1) lo = 1;
                if (hi <= lo) hi = lo + 1;
                if (hi > PF_INF_BITS) hi = PF_INF_BITS;
                sc->lo = lo; sc->hi = hi;
                sc->scale = SEL_ITEM_BITS - exp_of_bits(hi >= PF_INF_BITS ? vmax : hi);
                sc->phase = 1; sc->mode = 0;
            }
        }
    }
    grid.sync();

    // ---------------- rounds
    for (int round = 0; round < 40; round++) {
        const u64 lo = sc->lo, hi = sc->hi;
        const int scale = sc->scale;
        const int mode = sc->mode;
        const bool keep_all = (mode == 1);
        // ---- classify pass over all items
        {
            unsigned long long A_hi = 0, n_cand = 0;
            u128 B_lo = mk128(0, 0), S_c = mk128(0, 0), Tot = mk128(0, 0);
            long long chunk_base = -1;
            int chunk_fill = SEL_CHUNK;
            for (long long base = gwarp * 32; base < n_live; base += nwarps * 32) {
                long long i = base + lane;
                int c = (i < n_live) ? (a.cnt ? a.cnt[i] : 1) : 0;
                int maxc = c;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, d));
                for (int j = 0; j < maxc; j++) {
                    bool have = j < c;
                    u64 bits = have ? (u64)__double_as_longlong(a.V[(long long)j * a.cap + i]) : 0ull;
                    bool is_cand = false;
                    if (have) {
                        if (round == 0) Tot = add128(Tot, fx128(bits, tot_scale));
                        if (bits >= hi) A_hi++;
                        else if (bits < lo) { if (!keep_all) B_lo = add128(B_lo, fx128(bits, scale)); }
                        else { is_cand = true; n_cand++; S_c = add128(S_c, fx128(bits, scale)); }
                    }
                    unsigned int bal = __ballot_sync(0xffffffffu, is_cand);
                    if (bal) {
                        int nb = __popc(bal);
                        if (chunk_fill + nb > SEL_CHUNK) {
                            // pad the rest of the old chunk, take a new one
                            if (chunk_base >= 0)
                                for (int t = chunk_fill + lane; t < SEL_CHUNK; t += 32) a.cand[chunk_base + t] = __longlong_as_double(~0ll);
                            long long nb_ = 0;
                            if (lane == 0) nb_ = (long long)atomicAdd(&sc->cand_chunks, 1ull) * SEL_CHUNK;
                            chunk_base = __shfl_sync(0xffffffffu, nb_, 0);
                            chunk_fill = 0;
                        }
                        if (is_cand) {
                            int r = __popc(bal & ((1u << lane) - 1));
                            a.cand[chunk_base + chunk_fill + r] = __longlong_as_double((long long)bits);
                        }
                        chunk_fill += nb;
                    }
                }
            }
            if (chunk_base >= 0)
                for (int t = chunk_fill + lane; t < SEL_CHUNK; t += 32) a.cand[chunk_base + t] = __longlong_as_double(~0ll);
            // block reduce, then one set of atomics per block
            A_hi = warp_sum64(A_hi); n_cand = warp_sum64(n_cand);
            B_lo = warp_sum128(B_lo); S_c = warp_sum128(S_c); Tot = warp_sum128(Tot);
            if (lane == 0) { s_redc[0][wid] = A_hi; s_redc[1][wid] = n_cand; s_red[0][wid] = B_lo; s_red[1][wid] = S_c; s_red[2][wid] = Tot; }
            __syncthreads();
            if (wid == 0) {
                const int nw = SEL_THREADS / 32;
                unsigned long long x0 = lane < nw ? s_redc[0][lane] : 0, x1 = lane < nw ? s_redc[1][lane] : 0;
                u128 y0 = lane < nw ? s_red[0][lane] : mk128(0, 0), y1 = lane < nw ? s_red[1][lane] : mk128(0, 0),
                     y2 = lane < nw ? s_red[2][lane] : mk128(0, 0);
                x0 = warp_sum64(x0); x1 = warp_sum64(x1);
                y0 = warp_sum128(y0); y1 = warp_sum128(y1); y2 = warp_sum128(y2);
                if (lane == 0) {
                    if (x0) atomicAdd(&sc->A_hi, x0);
                    if (x1) atomicAdd(&sc->n_cand, x1);
                    atomic_add_acc(&sc->B_lo, y0);
                    atomic_add_acc(&sc->S_cand, y1);
                    if (round == 0) atomic_add_acc(&sc->Tot, y2);
                }
            }
            __syncthreads();
        }
        grid.sync();

        // ---- block 0: verify the bracket, start the descent
        if (blockIdx.x == 0 && tid == 0) {
            res->rounds = round + 1;
            if (round == 0) res->n_cand_first = (long long)sc->n_cand;
            int phase = 2;       // 2 = descend, 3 = redo classify with a new bracket, 4 = done
            if (keep_all) phase = 4;
            else if (mode == 2) {
                // comb-rescale round: T at the new scale; lo = hi = thr so A_hi = A again
                res->T = acc_value(&sc->B_lo); res->scale = scale;
                phase = comb_rescale(sc, res, N, vmax) ? 3 : 4;
            } else {
                u128 B = acc_value(&sc->B_lo), Sc = acc_value(&sc->S_cand);
                long long Ahi = (long long)sc->A_hi, nc = (long long)sc->n_cand;
                bool fail_lo = (lo > 1) && pred128(B, N, Ahi + nc, fx128(lo, scale));
                bool fail_hi = (hi < PF_INF_BITS) && !pred128(add128(B, Sc), N, Ahi, fx128(hi, scale));
                if (fail_lo || fail_hi) {
                    u64 nlo = fail_lo ? 1ull : lo, nhi = fail_hi ? PF_INF_BITS : hi;
                    sc->lo = nlo; sc->hi = nhi; sc->mode = 0;
                    sc->scale = SEL_ITEM_BITS - exp_of_bits(nhi >= PF_INF_BITS ? vmax : nhi);
                    phase = 3;
                } else {
                    sc->B_run = B; sc->A_run = Ahi; sc->count_in = nc;
                }
            }
            if (phase == 3) {
                sc->A_hi = 0; sc->n_cand = 0; sc->cand_chunks = 0;
                for (int k = 0; k < 4; k++) { sc->B_lo.l[k] = 0; sc->S_cand.l[k] = 0; }
            }
            sc->phase = phase;
        }
        grid.sync();
        int phase = sc->phase;
        if (phase == 3) continue;
        if (phase == 4) break;

        // ---- radix descent over the candidate list
        const long long list_n = (long long)sc->cand_chunks * SEL_CHUNK;
        bool refine = false;
        for (int level = 0; level < 8; level++) {
            const u64 blo = sc->lo, bhi = sc->hi;
            const long long count_in = sc->count_in;
            const bool to_final = count_in <= SEL_FINAL_CAP;
            const bool single = (bhi - blo) == 1;
            if (to_final) {
                // compact the bracket's items into the final list
                for (long long t = (long long)blockIdx.x * SEL_THREADS + tid; t < list_n; t += (long long)gridDim.x * SEL_THREADS) {
                    u64 bits = (u64)__double_as_longlong(a.cand[t]);
                    if (bits >= blo && bits < bhi) {
                        unsigned int pos = atomicAdd(&sc->final_n, 1u);
                        if (pos < SEL_FINAL_CAP) sc->final_vals[pos] = __longlong_as_double((long long)bits);
                    }
                }
                grid.sync();
                break;
            }
            if (single) break;             // one value with > SEL_FINAL_CAP ties: solved in closed form below
            // histogram level
            u64 width = bhi - blo;
            int shift = 0;
            while (((width - 1) >> shift) >= SEL_BINS) shift++;
            {
                unsigned int *h_cnt = reinterpret_cast<unsigned int *>(smem_raw);
                u64 *h_s0 = reinterpret_cast<u64 *>(smem_raw + SEL_BINS * 4);
                u64 *h_s1 = h_s0 + SEL_BINS, *h_s2 = h_s1 + SEL_BINS;
                for (int t = tid; t < SEL_BINS; t += SEL_THREADS) { h_cnt[t] = 0; h_s0[t] = 0; h_s1[t] = 0; h_s2[t] = 0; }
                // global histogram is zeroed by block 0 before the level starts (see below)
                __syncthreads();
                for (long long t = (long long)blockIdx.x * SEL_THREADS + tid; t < list_n; t += (long long)gridDim.x * SEL_THREADS) {
                    u64 bits = (u64)__double_as_longlong(a.cand[t]);
                    if (bits >= blo && bits < bhi) {
                        int b = (int)((bits - blo) >> shift);
                        u128 q = fx128(bits, scale);
                        atomicAdd(&h_cnt[b], 1u);
                        atomicAdd(&h_s0[b], q.lo & 0xffffffffull);
                        atomicAdd(&h_s1[b], q.lo >> 32);
                        atomicAdd(&h_s2[b], q.hi);
                    }
                }
                __syncthreads();
                for (int t = tid; t < SEL_BINS; t += SEL_THREADS)
                    if (h_cnt[t]) {
                        atomicAdd(&sc->h_cnt[t], h_cnt[t]);
                        atomicAdd(&sc->h_sum[0][t], h_s0[t]);
                        atomicAdd(&sc->h_sum[1][t], h_s1[t]);
                        atomicAdd(&sc->h_sum[2][t], h_s2[t]);
                    }
                __syncthreads();
            }
            grid.sync();
            if (blockIdx.x == 0 && wid == 0) {
                // lane l owns bins [32 l, 32 l + 32)
                const int per = SEL_BINS / 32;
                u128 ls = mk128(0, 0);
                long long lc = 0;
                for (int k = 0; k < per; k++) {
                    int b = lane * per + k;
                    u128 s = add128(add128(mk128(sc->h_sum[0][b], 0), shl128(mk128(sc->h_sum[1][b], 0), 32)), mk128(0, sc->h_sum[2][b]));
                    ls = add128(ls, s);
                    lc += sc->h_cnt[b];
                }
                u128 inc = ls; long long incc = lc;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    u128 o = shfl_up128(inc, d);
                    long long oc = __shfl_up_sync(0xffffffffu, incc, d);
                    if (lane >= d) { inc = add128(inc, o); incc += oc; }
                }
                u128 pre = sub128(inc, ls);            // sum of bins before this lane's range
                long long prec = incc - lc;
                u128 B0 = sc->B_run; long long A0 = sc->A_run;
                int found = SEL_BINS;                   // first bin whose UPPER boundary satisfies the predicate
                for (int k = 0; k < per; k++) {
                    int b = lane * per + k;
                    u128 s = add128(add128(mk128(sc->h_sum[0][b], 0), shl128(mk128(sc->h_sum[1][b], 0), 32)), mk128(0, sc->h_sum[2][b]));
                    pre = add128(pre, s);
                    prec += sc->h_cnt[b];
                    // boundary x = blo + (b+1) << shift (clamped to bhi)
                    u64 off = ((u64)(b + 1)) << shift;
                    u64 x = (off >= width) ? bhi : blo + off;
                    bool ok;
                    if (x >= PF_INF_BITS) ok = true;
                    else ok = pred128(add128(B0, pre), N, A0 + (count_in - prec), fx128(x, scale));
                    if (ok && found == SEL_BINS) found = b;
                    if (off >= width) break;
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) found = min(found, __shfl_xor_sync(0xffffffffu, found, d));
                if (found >= SEL_BINS) found = (int)((width - 1) >> shift);      // cannot happen (bhi verified)
                // totals strictly before / after the chosen bin
                u128 before = mk128(0, 0); long long cbefore = 0, cin = 0;
                for (int k = 0; k < per; k++) {
                    int b = lane * per + k;
                    u128 s = add128(add128(mk128(sc->h_sum[0][b], 0), shl128(mk128(sc->h_sum[1][b], 0), 32)), mk128(0, sc->h_sum[2][b]));
                    if (b < found) { before = add128(before, s); cbefore += sc->h_cnt[b]; }
                    if (b == found) cin = sc->h_cnt[b];
                }
                before = warp_sum128(before);
                cbefore = (long long)warp_sum64((u64)cbefore);
                cin = (long long)warp_sum64((u64)cin);
                if (lane == 0) {
                    u64 nlo = blo + (((u64)found) << shift);
                    u64 off = ((u64)(found + 1)) << shift;
                    u64 nhi = (off >= width) ? bhi : blo + off;
                    sc->B_run = add128(B0, before);
                    sc->A_run = A0 + (count_in - cbefore - cin);
                    sc->lo = nlo; sc->hi = nhi; sc->count_in = cin;
                }
                __syncwarp();
                for (int t = lane; t < SEL_BINS; t += 32) { sc->h_cnt[t] = 0; sc->h_sum[0][t] = 0; sc->h_sum[1][t] = 0; sc->h_sum[2][t] = 0; }
            }
            grid.sync();
        }

        // ---- block 0: exact solve on the final list
        if (blockIdx.x == 0) {
            const u64 blo = sc->lo, bhi = sc->hi;
            const long long count_in = sc->count_in;
            const u128 B0 = sc->B_run;
            const long long A0 = sc->A_run;
            u64 thr; u128 T; long long A;
            if (count_in > SEL_FINAL_CAP) {
                // (bhi - blo) == 1: count_in copies of the single value blo
                if (tid == 0) {
                    u128 q = fx128(blo, scale);
                    bool ok = pred128(B0, N, A0 + count_in, q);
                    s_ll[1] = ok ? (long long)blo : (long long)bhi;
                    s_ll[2] = ok ? A0 + count_in : A0;
                    s_warp[0] = ok ? B0 : add128(B0, mul128_64(q, (u64)count_in));
                }
                __syncthreads();
            } else {
                int nf = (int)count_in;
                int np2 = 1; while (np2 < nf) np2 <<= 1;
                if (np2 < 2) np2 = 2;
                for (int t = tid; t < np2; t += SEL_THREADS) s_u64[t] = t < nf ? (u64)__double_as_longlong(sc->final_vals[t]) : ~0ull;
                __syncthreads();
                bitonic_sort_smem(s_u64, np2);
                const int per = SEL_FINAL_CAP / SEL_THREADS;
                u128 loc = mk128(0, 0);
                for (int k = 0; k < per; k++) { int t = tid * per + k; if (t < nf) loc = add128(loc, fx128(s_u64[t], scale)); }
                u128 tot;
                u128 pre = block_excl_scan128(loc, s_warp, &tot);
                int first = nf;
                for (int k = 0; k < per; k++) {
                    int t = tid * per + k;
                    if (t < nf) {
                        u64 bits = s_u64[t];
                        bool head = (t == 0) || (s_u64[t - 1] != bits);
                        if (head && first == nf && pred128(add128(B0, pre), N, A0 + (nf - t), fx128(bits, scale))) first = t;
                        pre = add128(pre, fx128(bits, scale));
                    }
                }
                if (tid == 0) s_ll[0] = nf;
                __syncthreads();
                atomicMin(&s_ll[0], (long long)first);
                __syncthreads();
                int p = (int)s_ll[0];
                // T = B0 + sum of final items before p
                u128 part = mk128(0, 0);
                for (int k = 0; k < per; k++) { int t = tid * per + k; if (t < nf && t < p) part = add128(part, fx128(s_u64[t], scale)); }
                u128 tot2;
                block_excl_scan128(part, s_warp, &tot2);
                if (tid == 0) {
                    s_ll[1] = p < nf ? (long long)s_u64[p] : (long long)bhi;
                    s_ll[2] = A0 + (nf - p);
                    s_warp[0] = add128(B0, tot2);
                }
                __syncthreads();
            }
            if (tid == 0) {
                thr = (u64)s_ll[1]; A = s_ll[2]; T = s_warp[0];
                // precision: the scale must sit within 8 binades of the threshold
                bool precise = (thr >= PF_INF_BITS) ? (scale == SEL_ITEM_BITS - e_vmax)
                                                   : ((SEL_ITEM_BITS - scale) - exp_of_bits(thr) <= 8);
                if (!precise && round < 30) {
                    int e = (thr >= PF_INF_BITS) ? e_vmax : exp_of_bits(thr);
                    // new bracket: [2^(e-1), 2^(e+2)) clipped; scale from its upper end
                    long long elo = e - 1 + 1023, ehi = e + 2 + 1023;
                    u64 nlo = elo <= 0 ? 1ull : ((u64)elo << 52);
                    u64 nhi = ehi >= 2047 ? PF_INF_BITS : ((u64)ehi << 52);
                    if (thr >= PF_INF_BITS) { nlo = 1; nhi = PF_INF_BITS; }
                    sc->lo = nlo; sc->hi = nhi; sc->mode = 0;
                    sc->scale = SEL_ITEM_BITS - exp_of_bits(nhi >= PF_INF_BITS ? vmax : nhi);
                    sc->A_hi = 0; sc->n_cand = 0; sc->cand_chunks = 0; sc->final_n = 0;
                    for (int k = 0; k < 4; k++) { sc->B_lo.l[k] = 0; sc->S_cand.l[k] = 0; }
                    sc->phase = 3;
                } else {
                    res->thr_bits = thr; res->A = A; res->T = T; res->scale = scale;
                    sc->phase = comb_rescale(sc, res, N, vmax) ? 3 : 4;
                }
            }
        }
        (void)refine;
        grid.sync();
        if (sc->phase == 4) break;
    }

    // ---------------- outputs (block 0, thread 0)
    if (blockIdx.x == 0 && tid == 0) {
        u128 Tot = acc_value(&sc->Tot);
        double total = ldexp(to_double128(Tot), -tot_scale);
        res->log_total = log(to_double128(Tot)) - (double)tot_scale * 0.693147180559945309417232121458;
        (void)total;
        if (M <= N || vmax == 0) {
            res->keep_all = 1; res->thr_bits = (vmax == 0 && M > N) ? PF_INF_BITS : 0ull; res->A = (vmax == 0 && M > N) ? 0 : M; res->n = 0;
            res->T = mk128(0, 0); res->Uoff = mk128(0, 0); res->scale = sc->scale;
            res->lw_resamp = __longlong_as_double(0xFFF0000000000000ll); res->log_c = res->lw_resamp;
            res->thr_value = 0.0;
        } else {
            res->keep_all = 0;
            long long n = N - res->A;
            res->n = n;
            u128 T = res->T;
            // Uoff = floor(mu * T / 2^53), mu = floor(u 2^53)
            double u = *a.u;
            u64 mu = (u64)(u * 9007199254740992.0);
            // 192-bit product (T.hi:T.lo) * mu, then >> 53
            u64 p0 = T.lo * mu, p0h = mulhi64(T.lo, mu);
            u64 p1 = T.hi * mu, p1h = mulhi64(T.hi, mu);
            u64 w0 = p0, w1 = p0h + p1, w2 = p1h + (w1 < p0h ? 1ull : 0ull);
            u128 U;
            U.lo = (w0 >> 53) | (w1 << 11);
            U.hi = (w1 >> 53) | (w2 << 11);
            res->Uoff = U;
            double c = n > 0 ? ldexp(to_double128(T) / (double)n, -res->scale) : __longlong_as_double(0x7FF8000000000000ll);
            res->log_c = log(c);
            res->lw_resamp = res->log_c - res->log_total;
            res->thr_value = __longlong_as_double((long long)res->thr_bits);
        }
    }
}
