// select.cuh -- Fearnhead-Clifford threshold search (cutoff_ascending, src/fearnhead.jl:59-77).
//
//   bracket   a strided sample of the putative weights (about 2^20 of them) is solved
//             approximately by the same radix descent that solves the real problem, and widened
//             by a sampling-error margin into [lo, hi);                     k_sel_bracket
//   classify  one pass over all putatives: exact fixed-point mass below lo, count above hi, the
//             few in between become the candidate list (the fused expansion kernel does this on
//             the fly for the first round);                                  k_sel_classify
//   solve     verify the bracket (predicate false at lo and true at hi, else move / widen it
//             and ask for another classify round), resolve the plateau tie, 1024-bin radix
//             histogram levels over the candidate list until <= 2048 values remain, sort them
//             and evaluate the predicate of src/fearnhead.jl:69 exactly; if the comb needs a
//             finer scale than the threshold did, ask for one more classify round.  k_sel_solve
//
// k_sel_bracket and k_sel_solve work on short lists (~10^6 values); they are cooperative full-grid
// kernels: a 16-CTA cluster with DSMEM histograms was built and measured 2-4x slower (the list passes
// are instruction-bound on 16 SMs, profiles/r2/cluster_search_experiment.md).  Every CTA keeps the
// same search state and selects the next bin redundantly from the level's global histogram, so a
// radix level costs ONE grid barrier.  k_sel_classify streams the whole V and is a plain full-grid
// kernel.  A round that asks for another one (phase 3) is followed by queued retry launches of
// k_sel_classify / k_sel_solve, which are no-ops otherwise: no host round trip inside a step.
//
// Sharded runs (particles block-sharded over GPUs): k_sel_bracket / k_sel_solve begin by
// writing the local sample / candidate list and accumulator block into EVERY peer's exchange
// region (stores into peer memory over NVLink) and wait on per-rank sequence flags, so the
// all-gather is part of the kernel; every rank then solves the same gathered problem in
// integer arithmetic and reaches the same threshold, bit-identical to the single-GPU result.
//
// All sums that feed a decision are 128-bit fixed-point integers (order-free, hence
// deterministic and shard-independent).  Items are addressed as (particle i, slot j):
//     value = V[j * cap + i],   j < cnt[i]   (cnt == nullptr: one item per index)
// Result (SelResult): kept set = {v >= thr}; everything else is the low partition with
// fixed-point total T at scale 2^scale; n = N - A particles are drawn from it by the
// systematic comb of the survivor scan (sample_stratified!, src/fearnhead.jl:97-114).
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
namespace cg = cooperative_groups;

#define SEL_THREADS 512
#define SEL_WARPS (SEL_THREADS / 32)
#define SEL_FINAL_CAP 2048
#define SEL_BINS 1024
#define SEL_CHUNK 32
#ifndef SEL_SAMPLE_TARGET
#define SEL_SAMPLE_TARGET (1 << 20)
#endif
#define SEL_ITEM_BITS 69            // items q = floor(v 2^scale) < 2^(SEL_ITEM_BITS+1)
#define SEL_TOT_BITS 89             // total-weight accumulator scale relative to vmax
#define SEL_SMEM_BYTES (SEL_BINS * 4 + 2 * SEL_BINS * 8)   // one CTA's level histogram; also the scratch of select_bin / the final sort
#define SEL_RETRIES 2               // queued extra rounds per step (no-ops unless the previous round asked for one)
#define PF_MAX_RANKS 16

struct SelResult {
    u64 thr_bits;                   // kept = {bits(v) >= thr_bits}; 0 = keep everything
    int scale;                      // q = floor(v * 2^scale) for low items
    int keep_all;                   // M <= N (src/fearnhead.jl:135-138)
    u128 T;                         // sum of q over the low partition
    u128 Uoff;                      // floor(u * T): comb offset
    u128 Tq;                        // floor(T / n): tooth spacing in units of the running mass
    unsigned int Tr, pad0;          // T mod n
    long long A;                    // number kept
    long long n;                    // number to draw (N - A)
    double log_total;               // log(sum of all putative weights)
    double lw_resamp;               // log(c / total), c = T / n      (src/fearnhead.jl:173)
    double log_c;                   // log(c)
    double thr_value;
    int rounds;
    long long n_cand_first;
    long long n_plateau_first;      // in-bracket copies of the plateau value in the first round (counted, not listed)
    int status;                     // bit set, see pf_step_stats.select_status
    unsigned long long t_ns[5];     // globaltimer at: start, bracket ready, first classify done, solve done, end
};

struct SelScratch {
    // bracket state (written by CTA 0 of the search kernels, read by all after a barrier)
    u64 lo, hi;
    int scale;
    int phase;                      // 2 descend, 3 another classify round is needed, 4 done, 6 sample solve finished, 8 plateau
    int mode;                       // 0 bracket round, 1 keep everything, 2 comb-rescale round (lo = hi = thr)
    long long count_in;             // list items inside [lo, hi)
    long long mult;                 // sample stride while solving the sample, 1 otherwise
    long long halt;                 // 1 + index of the step that could not be completed (all later kernels are no-ops)
    int halt_reason;                // PF_HALT_*
    int pad0;
    // fused path: the expansion kernel classifies against a bracket found BEFORE it
    long long list_n;               // exact number of entries of the sample / candidate list (list_exact)
    int list_exact;                 // 1: the list is dense (block-reserved ranges), 0: chunked with padding
    int fuse_valid;                 // 1: k_expand classified against [lo, hi) and filled the accumulators / tile records
    int fuse_scale;                 // item scale the expansion used (= scale of the bracket)
    int fuse_tot_scale;             // scale of the kept-side total the expansion used
    int tot_scale_used;             // scale of the kept-side total of the current classify round
    int n_fail;                     // bracket verifications that failed in this step (geometric widening, then open)
    int fuse_overflow;              // a block had more candidates than its staging buffer
    int need_tot;                   // the total weight of this step has not been computed yet
    int tiles_ok;                   // the survivor-scan tile records written by the expansion are usable
    int low_max_valid;              // the round was classified by k_sel_classify (which tracks low_max)
    // accumulators of the classify pass (SelAcc: the block a sharded run exchanges)
    unsigned long long A_hi;        // #items >= hi
    unsigned long long n_cand;      // #items in [lo, hi)
    acc128 B_lo;                    // sum q of items < lo
    acc128 S_cand;                  // sum q of items in [lo, hi)
    acc128 Tot;                     // sum of the items >= hi at scale kept_tot_scale(hi, vmax)
    unsigned long long cand_chunks; // chunks handed out of the candidate / sample list
    unsigned long long overflow;    // sharded run: the list outgrew its exchange segment
    unsigned long long M_local;     // sharded run: this shard's putative count and max weight bits
    unsigned long long vmax_local;
    unsigned long long vs_bits;     // max weight of the strided sample (its exponent fixes the sample solve's scale)
    unsigned long long n_plateau;   // in-bracket items equal to the plateau value (counted, not listed)
    unsigned long long low_max;     // largest item below lo seen by k_sel_classify (bit pattern; valid if low_max_valid)
    // plateau resolution: candidates below the plateau value (count, fixed-point sum)
    unsigned long long pl_cnt;
    acc128 pl_sum;
    // running totals of the descent
    u128 B_run;                     // sum q of list items < lo (after narrowing)
    long long A_run;                // #list items >= hi (after narrowing)
    // the global problem of the current launch (sharded: summed over the ranks)
    long long gM;
    u64 gvmax;
    // final list
    unsigned int final_n;
    double final_vals[SEL_FINAL_CAP];
    // histograms of the descent levels: three rotating buffers, all zero between launches
    unsigned int h_cnt[3][SEL_BINS];
    u64 h_sum[3][2][SEL_BINS];      // bin sum = h_sum[.][0] + h_sum[.][1] 2^32
};
enum { PF_HALT_SEARCH = 1,          // the threshold search needed more rounds than were queued
       PF_HALT_TIMEOUT = 2,         // a peer's flag did not arrive (sharded run)
       PF_HALT_XCH_OVERFLOW = 3,    // a sample / candidate list outgrew its exchange segment (sharded run)
       PF_HALT_SURVIVORS = 4,       // a shard produced more survivors than its survivor list holds
       PF_HALT_LABEL = 5 };         // a labeled observation named a state that is not a candidate (ArgumentError, src/particle.jl:137-139)

// the accumulator block of SelScratch as a stand-alone record (sharded runs exchange one per rank)
struct SelAcc {
    unsigned long long A_hi, n_cand;
    acc128 B_lo, S_cand, Tot;
    unsigned long long cand_chunks, overflow, M_local, vmax_local, vs_bits, n_plateau, low_max;
};
#define SEL_ACC_OFFSET offsetof(SelScratch, A_hi)

// ------------------------------------------------------------------ peer exchange (sharded runs)
// Every rank owns one exchange region in device memory, mapped into every other rank (CUDA IPC between
// processes, plain pointers inside one process).  Rank r PUSHES its contribution into slot r of every
// region (stores into peer memory over NVLink), makes the stores visible (__threadfence_system) and
// raises flag[kind][r] to a sequence number; a consumer spins on the G flags of its OWN region, so all
// reads are local.  Sequence numbers only grow, nothing is ever reset.
//   kind A  strided sample            (header XchHdrA + list)        seq = step + 1
//   kind B  candidates of a round     (header SelAcc  + list)        seq = (step + 1) * 64 + round, two buffers by round parity
//   kind C  survivor-scan total       (XchTot)                       seq = step + 1
//   kind D  next generation complete  (flag only)                    seq = step + 1
// A rank publishes step t + 1 only after it has seen every peer's D flag of step t, which a peer raises
// after its last kernel of step t: one buffer per kind is enough across steps; within a step the rounds
// of kind B alternate between two buffers because a fast rank may publish round k + 1 while a slow one
// still reads round k.
// Chen-Liu (chenliu.cuh): kinds CL_W max weight, CL_S sum of weights, CL_Q sum of squared deviations, CL_T mass of the
// shard (rejuvenation), CL_R prefix array ready -- one scalar per rank each, seq = step + 1
enum { XCH_A = 0, XCH_B = 1, XCH_C = 2, XCH_D = 3, XCH_CL_W = 4, XCH_CL_S = 5, XCH_CL_Q = 6, XCH_CL_T = 7, XCH_CL_R = 8, XCH_KINDS = 9 };
struct XchHdrA { unsigned long long n, vmax, M_pred, overflow; };
struct XchTot { u128 X; unsigned int kept; unsigned int pad[3]; };         // = TileSum (fearnhead.cuh)
struct XchRegion {
    unsigned long long flag[XCH_KINDS][PF_MAX_RANKS];
    XchHdrA a[PF_MAX_RANKS];
    SelAcc b[2][PF_MAX_RANKS];
    XchTot c[PF_MAX_RANKS];
    unsigned long long clw[PF_MAX_RANKS];
    acc128 clsum[PF_MAX_RANKS], clsq[PF_MAX_RANKS];
    u128 clq[PF_MAX_RANKS];
    // followed (256-byte aligned) by listA[nranks][capA] and listB[2][nranks][capB] (doubles)
};
struct Xch {
    int rank, nranks;
    long long capA, capB;
    unsigned long long timeout_ns;
    XchRegion *reg[PF_MAX_RANKS];   // region of every rank as mapped in this process (reg[rank]: own)
};
__host__ __device__ __forceinline__ size_t xch_header_bytes() { return (sizeof(XchRegion) + 255) / 256 * 256; }
__host__ __device__ __forceinline__ double *xch_listA(XchRegion *reg, long long capA, int src)
{
    return reinterpret_cast<double *>(reinterpret_cast<char *>(reg) + xch_header_bytes()) + (long long)src * capA;
}
__host__ __device__ __forceinline__ double *xch_listB(XchRegion *reg, int nranks, long long capA, long long capB, int parity, int src)
{
    return reinterpret_cast<double *>(reinterpret_cast<char *>(reg) + xch_header_bytes()) + (long long)nranks * capA +
           ((long long)parity * nranks + src) * capB;
}
__host__ __device__ __forceinline__ size_t xch_region_bytes(int nranks, long long capA, long long capB)
{
    return xch_header_bytes() + sizeof(double) * ((size_t)nranks * capA + 2 * (size_t)nranks * capB);
}

__device__ __forceinline__ unsigned long long gtimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// one thread: wait until every rank's flag of `kind` in the OWN region reached `seq`; false on time-out
__device__ inline bool xch_wait(const Xch *x, int kind, unsigned long long seq)
{
    const unsigned long long *f = x->reg[x->rank]->flag[kind];
    const unsigned long long t0 = gtimer();
    for (int r = 0; r < x->nranks; r++) {
        unsigned int spins = 0;
        while (ld_acquire_sys(f + r) < seq) {
            if ((++spins & 1023u) == 0 && gtimer() - t0 > x->timeout_ns) return false;
            __nanosleep(40);
        }
    }
    return true;
}
// one thread: raise this rank's flag of `kind` to `seq` in every region (own included)
__device__ inline void xch_signal(const Xch *x, int kind, unsigned long long seq)
{
    __threadfence_system();
    for (int p = 0; p < x->nranks; p++) st_release_sys(&x->reg[p]->flag[kind][x->rank], seq);
}
// a u64 that a peer wrote into local memory: read it from L2 (never from a stale L1 line)
__device__ __forceinline__ unsigned long long ld_peer(const unsigned long long *p) { return __ldcg(p); }

// stages of the search kernels (k_sel_bracket, k_sel_solve)
enum { SEL_ST_PUB = 1,        // sharded: write the local list / header into the peers' regions and raise the flag
       SEL_ST_RUN = 2 };      // do the stage's work (sharded: after waiting for every rank's flag)

struct SelArgs {
    const double *V;
    const unsigned char *cnt;
    long long cap;
    const long long *n_live;        // device
    const long long *M;             // device: number of items (sharded: of this shard)
    const u64 *vmax_bits;           // device: maximum item (sharded: of this shard)
    long long N;
    const double *u;                // device: the uniform of this step
    double *cand;                   // local sample / candidate list
    long long cand_cap;             // in doubles
    SelScratch *sc;
    SelResult *res;
    int stage;                      // SEL_ST_* bit mask
    int margin;                     // half-width of the sample bracket in sigmas of the sampling error (8; tests shrink it)
    int retry;                      // 1: a queued extra round -- runs only if the previous round asked for it (phase 3)
    int external_sample;            // 1: the sample came from k_expand (sample mode): M is the predicted count, vmax the sample's
    int preclassified;              // 1: round 0 was classified by the expansion kernel (if sc->fuse_valid)
    int pad0;
    long long fixed_stride;         // sample stride (the same strided sample on every path)
    const unsigned long long *plateau_ptr;   // device: value of the step's mass tie (StepConst.plateau_bits); NULL = none
    const Xch *x;                   // sharded run: exchange descriptor (device); NULL = single GPU
    unsigned long long seq;         // sharded run: step + 1
    long long step;                 // index of this step (halt mark)
};

// the list a search kernel works on: one segment (single GPU) or one per rank (own exchange region)
struct ListView {
    const double *seg[PF_MAX_RANKS];
    long long off[PF_MAX_RANKS + 1];    // exclusive prefix of the segment lengths
    int nseg;
};
__device__ __forceinline__ u64 lv_load(const ListView &lv, long long t)
{
    int s = 0;
    while (s + 1 < lv.nseg && t >= lv.off[s + 1]) s++;
    return (u64)__double_as_longlong(__ldcg(lv.seg[s] + (t - lv.off[s])));      // L2: peers may have written this memory
}

// Every thread of the grid visits its share of the list, SEL_BATCH independent loads in flight per thread.
// Padding entries (all-ones pattern) are skipped.
#define SEL_BATCH 4
template <class F>
__device__ __forceinline__ void lv_for_each(const ListView &lv, long long start, long long step, F f)
{
    const long long n = lv.off[lv.nseg];
    for (long long t0 = start; t0 < n; t0 += step * SEL_BATCH) {
        u64 b[SEL_BATCH];
#pragma unroll
        for (int k = 0; k < SEL_BATCH; k++) { const long long t = t0 + k * step; b[k] = t < n ? lv_load(lv, t) : ~0ull; }
#pragma unroll
        for (int k = 0; k < SEL_BATCH; k++) if (b[k] != ~0ull) f(b[k]);
    }
}

// Scale of the kept-side total (items >= hi).  52 - e(hi) represents every such double exactly
// (no truncation), so the total weight does not depend on which kernel accumulated it; the
// sum stays below 2^127 while the maximum is within 46 binades of hi.
__host__ __device__ __forceinline__ int kept_tot_scale(u64 hi, u64 vmax)
{
    const int e_v = exp_of_bits(vmax);
    if (hi < PF_INF_BITS && e_v - exp_of_bits(hi) <= 46) return 52 - exp_of_bits(hi);
    return 99 - e_v;
}

// log of the step's total weight = (kept side: integer Tot at scale ts) + (low side: integer low at scale `scale`).
// Where the scales are close (the usual case: scale - ts = 17) the two integers are added EXACTLY (192 bits)
// and rounded to a double ONCE, so the result depends on the sum only, not on where the bracket split it --
// runs whose brackets differ (another sample stride, a retried round, another sharding) agree bit for bit.
__host__ __device__ inline double log_total_exact(u128 Tot, int ts, u128 low, int scale)
{
    const double LN2 = 0.693147180559945309417232121458;
    const int d = scale - ts;                           // the low side's scale minus the kept side's
    if (!isz128(Tot) && !isz128(low) && (d > 60 || d < -60)) {
        // scales far apart (kept side far above the bracket: its sum was truncated anyway): plain floating point
        const double tsum = to_double128(Tot) + ldexp(to_double128(low), ts - scale);
        return log(tsum) - (double)ts * LN2;
    }
    // total = A * 2^sh + B at scale sc (the finer of the two scales)
    u128 A = mk128(0, 0), B = low;
    int sh = 0, sc = scale;
    if (isz128(Tot)) { }
    else if (isz128(low)) { B = Tot; sc = ts; }
    else if (d >= 0) { A = Tot; sh = d; }
    else { A = low; B = Tot; sh = -d; sc = ts; }
    // w = A * 2^sh + B, little-endian 64-bit limbs
    u64 w[3];
    w[0] = A.lo << sh; w[1] = (A.hi << sh) | (sh ? A.lo >> (64 - sh) : 0ull); w[2] = sh ? A.hi >> (64 - sh) : 0ull;
    u64 c = 0;
    w[0] += B.lo; c = w[0] < B.lo ? 1ull : 0ull;
    u64 t1 = w[1] + B.hi; u64 c1 = t1 < B.hi ? 1ull : 0ull;
    w[1] = t1 + c; c1 += w[1] < t1 ? 1ull : 0ull;
    w[2] += c1;
    // top 64 significant bits + sticky, converted with one round-to-nearest-even
    int top = w[2] ? 2 : (w[1] ? 1 : 0);
    u64 hi64 = w[top];
#ifdef __CUDA_ARCH__
    const int lz = hi64 ? __clzll((long long)hi64) : 64;
#else
    const int lz = hi64 ? __builtin_clzll(hi64) : 64;
#endif
    if (lz == 64) return log(0.0);
    u64 m = hi64 << lz;
    bool sticky = false;
    if (top > 0) {
        const u64 nxt = w[top - 1];
        if (lz) { m |= nxt >> (64 - lz); sticky = (nxt << lz) != 0; } else sticky = nxt != 0;
        if (top > 1) sticky = sticky || w[top - 2] != 0;
    }
    if (sticky) m |= 1ull;
    const double v = (double)m;                         // w = m * 2^(64 top - lz) up to the one rounding of this conversion
    const int e = 64 * top - lz - sc;
    if (e > -1000 && e < 900) return log(ldexp(v, e));  // the total as ONE correctly rounded double, then its log
    return log(v) + (double)e * LN2;
}

// predicate of cutoff_ascending (src/fearnhead.jl:69): tot <= (N - n_geq) * w
__device__ __forceinline__ bool pred128(u128 B, long long N, long long A_geq, u128 qx)
{
    long long dd = N - A_geq;
    if (dd < 0) return false;
    return le128(B, mul128_64(qx, (u64)dd));
}

// exclusive scan of one u128 per thread over the block; returns exclusive prefix, *total = block sum
__device__ inline u128 block_excl_scan128(u128 v, u128 *smem_warp /* >= 33 */, u128 *total)
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

// warp-aggregated append of the lanes with `flag` to a chunked list (chunks of SEL_CHUNK
// doubles handed out by one atomic each; unused tails are padded with NaN-pattern ~0).
struct ChunkWriter {
    long long base;
    int fill;
    __device__ __forceinline__ void init() { base = -1; fill = SEL_CHUNK; }
    __device__ __forceinline__ void push(bool flag, u64 bits, double *list, unsigned long long *chunks, int lane)
    {
        unsigned int bal = __ballot_sync(0xffffffffu, flag);
        if (!bal) return;
        int nb = __popc(bal);
        if (fill + nb > SEL_CHUNK) {
            if (base >= 0)
                for (int t = fill + lane; t < SEL_CHUNK; t += 32) list[base + t] = __longlong_as_double(~0ll);
            long long nb_ = 0;
            if (lane == 0) nb_ = (long long)atomicAdd(chunks, 1ull) * SEL_CHUNK;
            base = __shfl_sync(0xffffffffu, nb_, 0);
            fill = 0;
        }
        if (flag) list[base + fill + __popc(bal & ((1u << lane) - 1))] = __longlong_as_double((long long)bits);
        fill += nb;
    }
    __device__ __forceinline__ void finish(double *list, int lane)
    {
        if (base >= 0)
            for (int t = fill + lane; t < SEL_CHUNK; t += 32) list[base + t] = __longlong_as_double(~0ll);
    }
};

// search state of the radix descent: identical in every CTA (each computes it from the same global histogram),
// so a level needs only ONE grid barrier (after the histogram is complete); CTA 0 mirrors it to SelScratch
struct SrchSt {
    u64 lo, hi;
    long long count_in, A_run, mult;
    u128 B_run;
    int phase;
};

// One radix level: every CTA histograms (count, fixed-point sum as two 64-bit words) its share of the list items
// inside [blo, bhi) in shared memory (1024 bins of width 2^shift bit patterns) and adds its non-empty bins to the
// level's global buffer `buf`.  Lanes of a warp that hit the same bin are grouped (match.any), their items summed
// with three 32-bit warp reductions (24-bit limbs) and the group's first lane does the shared-memory atomics:
// the first level's bins are whole binades, a handful of them hold most items, and 64-bit shared-memory atomics
// are compare-and-swap loops.  CTA 0 also clears the buffer the NEXT level will use (three buffers rotate: the
// one being cleared was last read two levels ago, before the previous barrier).
__device__ inline void hist_level(const ListView &lv, u64 blo, u64 bhi, int shift, int scale, SelScratch *sc, int buf,
                                  unsigned char *smem_raw)
{
    unsigned int *h_cnt = reinterpret_cast<unsigned int *>(smem_raw);
    u64 *h_s0 = reinterpret_cast<u64 *>(smem_raw + SEL_BINS * 4);
    u64 *h_s1 = h_s0 + SEL_BINS;
    const int tid = threadIdx.x, lane = tid & 31;
    for (int t = tid; t < SEL_BINS; t += SEL_THREADS) { h_cnt[t] = 0; h_s0[t] = 0; h_s1[t] = 0; }
    if (blockIdx.x == 0) {
        const int nb = (buf + 1) % 3;
        for (int t = tid; t < SEL_BINS; t += SEL_THREADS) { sc->h_cnt[nb][t] = 0; sc->h_sum[nb][0][t] = 0; sc->h_sum[nb][1][t] = 0; }
    }
    __syncthreads();
    const long long n = lv.off[lv.nseg];
    const long long step = (long long)gridDim.x * SEL_THREADS;
    for (long long t0 = (long long)blockIdx.x * SEL_THREADS + (tid & ~31); t0 < n; t0 += step * SEL_BATCH) {      // (warp-uniform trip count)
        u64 bb[SEL_BATCH];
#pragma unroll
        for (int k = 0; k < SEL_BATCH; k++) { const long long t = t0 + lane + k * step; bb[k] = t < n ? lv_load(lv, t) : ~0ull; }
#pragma unroll
        for (int k = 0; k < SEL_BATCH; k++) {
            const u64 bits = bb[k];
            const bool in = bits >= blo && bits < bhi;
            if (!__any_sync(0xffffffffu, in)) continue;
            const int b = in ? (int)((bits - blo) >> shift) : -1 - lane;
            const unsigned int peers = __match_any_sync(0xffffffffu, b);
            if (in) {
                const u128 q = fx70(bits, scale);                                              // < 2^71
                const int np = __popc(peers);
                if (np == 1) {
                    atomicAdd(&h_cnt[b], 1u);
                    atomicAdd(&h_s0[b], q.lo & 0xffffffffull);
                    atomicAdd(&h_s1[b], (q.lo >> 32) | (q.hi << 32));
                } else {
                    const unsigned int s0 = __reduce_add_sync(peers, (unsigned int)(q.lo & 0xffffffu));
                    const unsigned int s1 = __reduce_add_sync(peers, (unsigned int)((q.lo >> 24) & 0xffffffu));
                    const unsigned int s2 = __reduce_add_sync(peers, (unsigned int)((q.lo >> 48) | (q.hi << 16)));
                    if (lane == __ffs(peers) - 1) {
                        const u128 tot = add128(add128(mk128(s0, 0), mk128((u64)s1 << 24, 0)), shl128(mk128(s2, 0), 48));
                        atomicAdd(&h_cnt[b], (unsigned int)np);
                        atomicAdd(&h_s0[b], tot.lo & 0xffffffffull);
                        atomicAdd(&h_s1[b], (tot.lo >> 32) | (tot.hi << 32));
                    }
                }
            }
        }
    }
    __syncthreads();
    for (int t = tid; t < SEL_BINS; t += SEL_THREADS)
        if (h_cnt[t]) {
            atomicAdd(&sc->h_cnt[buf][t], h_cnt[t]);
            atomicAdd(&sc->h_sum[buf][0][t], h_s0[t]);
            atomicAdd(&sc->h_sum[buf][1][t], h_s1[t]);           // bin sum = h_sum[0] + h_sum[1] 2^32
        }
    __threadfence();
}

__device__ __forceinline__ unsigned int block_excl_scan_u32(unsigned int v, unsigned int *smem_warp /* >= 33 */)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    unsigned int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { unsigned int o = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) smem_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned int w = lane < nw ? smem_warp[lane] : 0, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned int o = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += o; }
        smem_warp[lane] = wi - w;
    }
    __syncthreads();
    const unsigned int r = smem_warp[wid] + inc - v;
    __syncthreads();
    return r;
}

// After a histogram level (and the grid barrier behind it), in EVERY CTA, all threads: find the first bin whose
// UPPER boundary satisfies the predicate and narrow the bracket to it.  With sampling != 0 the list is a strided
// sample (sums and counts are multiplied by st->mult): the level either narrows to the bin and its two neighbours,
// or -- once the bin is finer than the sampling error -- ends the sample solve by widening the bin to `delta`
// sample items on each side (st->phase = 6).  Two bins per thread, block-wide prefix sums.
__device__ inline void select_bin(SrchSt *st, const SelScratch *sc, int buf, long long N, u64 blo, u64 bhi, int shift, int scale,
                                  int sampling, unsigned char *smem_raw, u128 *s_warp, unsigned int *s_wc)
{
    unsigned int *pc = reinterpret_cast<unsigned int *>(smem_raw);            // inclusive count prefix
    u128 *ps = reinterpret_cast<u128 *>(smem_raw + SEL_BINS * 4);            // inclusive sum prefix
    __shared__ int s_found;
    const int tid = threadIdx.x;
    constexpr int PER = SEL_BINS / SEL_THREADS;
    u128 v[PER]; unsigned int c[PER];
    u128 loc = mk128(0, 0); unsigned int lc = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int b = tid * PER + k;
        c[k] = __ldcg(&sc->h_cnt[buf][b]);
        v[k] = add128(mk128(__ldcg(&sc->h_sum[buf][0][b]), 0), shl128(mk128(__ldcg(&sc->h_sum[buf][1][b]), 0), 32));
        loc = add128(loc, v[k]); lc += c[k];
    }
    if (tid == 0) s_found = SEL_BINS;
    u128 tot;
    u128 pre = block_excl_scan128(loc, s_warp, &tot);
    unsigned int prec = block_excl_scan_u32(lc, s_wc);
    const u64 width = bhi - blo;
    const long long count_in = st->count_in;
    const long long mult = st->mult;
    const u128 B0 = st->B_run; const long long A0 = st->A_run;
    const int last_bin = (int)((width - 1) >> shift);
    int found = SEL_BINS;
#pragma unroll
    for (int k = 0; k < PER; k++) {
        const int b = tid * PER + k;
        pre = add128(pre, v[k]); prec += c[k];
        ps[b] = pre; pc[b] = prec;
        if (b <= last_bin) {
            u64 off = ((u64)(b + 1)) << shift;
            u64 x = (off >= width) ? bhi : blo + off;
            bool ok = (x >= PF_INF_BITS) ? true
                                         : pred128(mul128_64(add128(B0, pre), (u64)mult), N, (A0 + (count_in - (long long)prec)) * mult,
                                                   fx70(x, scale));
            if (ok && found == SEL_BINS) found = b;
        }
    }
    if (found < SEL_BINS) atomicMin(&s_found, found);
    __syncthreads();
    if (tid == 0) {
        int f = s_found;
        if (f > last_bin) f = last_bin;
        int b_lo = f, b_hi = f;
        int next_phase = 2;
        const long long cnt_found = (long long)pc[f] - (f > 0 ? (long long)pc[f - 1] : 0);
        if (sampling) {
            const long long above = A0 + (count_in - (long long)pc[f]) + cnt_found;
            const long long delta = (long long)((double)sampling * sqrt((double)(above > 1 ? above : 1))) + 4 * sampling;   // `sampling` = margin in sigmas (8)
            if (cnt_found > 2 * delta && shift > 0) {
                b_lo = f > 0 ? f - 1 : 0;
                b_hi = f < last_bin ? f + 1 : last_bin;
            } else {
                const long long below_found = f > 0 ? (long long)pc[f - 1] : 0;
                while (b_lo > 0 && below_found - (b_lo > 0 ? (long long)pc[b_lo - 1] : 0) < delta) b_lo--;
                while (b_hi < last_bin && (long long)pc[b_hi] - (long long)pc[f] < delta) b_hi++;
                next_phase = 6;
            }
        }
        const u128 before = b_lo > 0 ? ps[b_lo - 1] : mk128(0, 0);
        const long long cbefore = b_lo > 0 ? (long long)pc[b_lo - 1] : 0;
        const long long cin = (long long)pc[b_hi] - cbefore;
        const u64 nlo = blo + (((u64)b_lo) << shift);
        const u64 off = ((u64)(b_hi + 1)) << shift;
        const u64 nhi = (off >= width) ? bhi : blo + off;
        st->B_run = add128(B0, before);
        st->A_run = A0 + (count_in - cbefore - cin);
        st->lo = nlo; st->hi = nhi; st->count_in = cin;
        st->phase = next_phase;
    }
    __syncthreads();
}

// all CTAs: clear the three histogram buffers (after a grid barrier that follows the last bin selection)
__device__ __forceinline__ void hist_clear(SelScratch *sc)
{
    unsigned int *c = &sc->h_cnt[0][0];
    u64 *s = &sc->h_sum[0][0][0];
    for (long long t = (long long)blockIdx.x * SEL_THREADS + threadIdx.x; t < 3 * SEL_BINS; t += (long long)gridDim.x * SEL_THREADS) c[t] = 0;
    for (long long t = (long long)blockIdx.x * SEL_THREADS + threadIdx.x; t < 6 * SEL_BINS; t += (long long)gridDim.x * SEL_THREADS) s[t] = 0;
}

__device__ inline void reset_round(SelScratch *sc)
{
    sc->A_hi = 0; sc->n_cand = 0; sc->cand_chunks = 0; sc->final_n = 0; sc->overflow = 0; sc->list_n = 0; sc->list_exact = 0;
    sc->n_plateau = 0; sc->pl_cnt = 0; for (int k = 0; k < 4; k++) sc->pl_sum.l[k] = 0;
    sc->low_max = 0; sc->low_max_valid = 0;
    for (int k = 0; k < 4; k++) { sc->B_lo.l[k] = 0; sc->S_cand.l[k] = 0; }
}

// The comb needs the low partition's mass T at a scale fine relative to c = T / n (every
// low weight is < c, so c bounds the items).  The threshold round's scale follows the
// threshold instead, which can sit many binades above c when a gap separates the kept
// group from the rest.  Returns true (and programs a mode-2 round: lo = hi = thr at a
// finer scale) when T has fewer than 52 bits per drawn particle.
__device__ inline bool comb_rescale(SelScratch *sc, const SelResult *res, long long N)
{
    const long long n = N - res->A;
    if (n <= 0) return false;
    const u128 T = res->T;
    const int scale = res->scale;
    const int max_scale = 1074 + SEL_ITEM_BITS;
    int nscale;
    if (isz128(T)) {
        if (scale >= max_scale) return false;               // the low partition really is all zero
        if (sc->low_max_valid) {
            // the classify pass saw the largest low item: go straight to the scale that resolves it
            if (sc->low_max == 0) return false;             // nothing but exact zeros below the threshold
            nscale = SEL_ITEM_BITS - 3 - exp_of_bits(sc->low_max);
            if (nscale <= scale) nscale = scale + 64;
        } else nscale = scale + 64;                         // (fused round: the next, classic, round will know)
    } else {
        if (!lt128(T, shl128(mk128((u64)n, 0), 52))) return false;
        int blT = T.hi ? 128 - __clzll((long long)T.hi) : 64 - __clzll((long long)T.lo);
        int bln = 64 - __clzll(n);
        int e_c = blT - bln + 1 - scale;                    // c < 2^e_c
        nscale = SEL_ITEM_BITS - 3 - e_c;
        if (nscale <= scale) return false;
    }
    if (nscale > max_scale) nscale = max_scale;
    sc->lo = res->thr_bits; sc->hi = res->thr_bits; sc->scale = nscale; sc->mode = 2; sc->mult = 1;
    reset_round(sc);
    return true;
}

// mark the step as not completable: every later kernel of the handle is a no-op
__device__ __forceinline__ void sel_halt(SelScratch *sc, long long step, int reason)
{
    if (!sc->halt) { sc->halt = step + 1; sc->halt_reason = reason; }
}

// sharded: all threads of the grid copy `n` doubles of the local list into segment `rank` of every region
__device__ inline void xch_copy_list(const Xch *x, const double *src, long long n, int kind, int parity)
{
    const long long nth = (long long)gridDim.x * SEL_THREADS;
    const long long t0 = (long long)blockIdx.x * SEL_THREADS + threadIdx.x;
    for (int p = 0; p < x->nranks; p++) {
        double *dst = kind == XCH_A ? xch_listA(x->reg[p], x->capA, x->rank)
                                    : xch_listB(x->reg[p], x->nranks, x->capA, x->capB, parity, x->rank);
        for (long long t = t0; t < n; t += nth) dst[t] = src[t];
    }
    __threadfence_system();
}

// ------------------------------------------------------------------ classic path: reset + strided sample
// (pf_select, the sorted traversal order and PF_NO_FUSE; the fused path samples inside k_expand)
__global__ void k_sel_reset(SelArgs a)
{
    if (threadIdx.x || blockIdx.x) return;
    SelScratch *sc = a.sc; SelResult *res = a.res;
    if (sc->halt) return;
    reset_round(sc);
    for (int k = 0; k < 4; k++) sc->Tot.l[k] = 0;
    res->rounds = 0; res->status = 0; res->n_cand_first = 0; res->n_plateau_first = 0; sc->n_fail = 0;
    res->t_ns[0] = gtimer();
    sc->mult = 1; sc->need_tot = 1; sc->fuse_valid = 0; sc->fuse_overflow = 0; sc->tiles_ok = 0; sc->vs_bits = 0; sc->phase = 0;
}

__global__ void __launch_bounds__(SEL_THREADS) k_sel_sample(SelArgs a)
{
    SelScratch *sc = a.sc;
    if (sc->halt) return;
    const long long n_live = *a.n_live, M = *a.M;
    const u64 vmax = *a.vmax_bits;
    if (M <= a.N || vmax == 0 || M <= SEL_FINAL_CAP) return;           // solved directly
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long gwarp = (long long)blockIdx.x * SEL_WARPS + wid, nwarps = (long long)gridDim.x * SEL_WARPS;
    const long long stride = a.fixed_stride;
    ChunkWriter cw; cw.init();
    unsigned long long cnt_s = 0;
    u64 vs_max = 0;
    // particle g is in the sample iff (g / 4) % stride == 0 (aligned groups of 4; the same
    // rule as the sample expansion, so every path brackets the threshold on the same sample)
    for (long long base = gwarp * 32; (base >> 2) * (4 * stride) < n_live; base += nwarps * 32) {
        const long long g = base + lane;
        const long long i = (g >> 2) * (4 * stride) + (g & 3);
        int c = (i < n_live) ? (a.cnt ? a.cnt[i] : 1) : 0;
        int maxc = c;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, d));
        for (int j = 0; j < maxc; j++) {
            bool have = j < c;
            u64 bits = have ? (u64)__double_as_longlong(a.V[(long long)j * a.cap + i]) : 0ull;
            have = have && bits != 0;             // zeros never decide anything
            cnt_s += have;
            vs_max = bits > vs_max ? bits : vs_max;
            cw.push(have, bits, a.cand, &sc->cand_chunks, lane);
        }
    }
    cw.finish(a.cand, lane);
    cnt_s = warp_sum64(cnt_s);
    vs_max = warp_max64(vs_max);
    if (lane == 0 && cnt_s) { atomicAdd(&sc->n_cand, cnt_s); atomicMax(&sc->vs_bits, vs_max); }
}

// ------------------------------------------------------------------ k_sel_bracket (cooperative grid)
__global__ void __launch_bounds__(SEL_THREADS, 1) k_sel_bracket(SelArgs a)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ __align__(16) unsigned char smem_raw[SEL_SMEM_BYTES];
    __shared__ ListView s_lv;
    __shared__ SrchSt s_st;
    __shared__ u128 s_warp[33];
    __shared__ unsigned int s_wc[33];
    SelScratch *sc = a.sc;
    SelResult *res = a.res;
    const Xch *x = a.x;
    if (sc->halt) return;
    const int tid = threadIdx.x;
    const bool lead = blockIdx.x == 0 && tid == 0;
    const long long N = a.N;

    // ---- the sample: local list, or (sharded) every rank's list gathered through the exchange regions
    if (x) {
        if (a.stage & SEL_ST_PUB) {
            const long long ns = sc->list_n;
            const long long n_pub = ns <= x->capA ? ns : x->capA;
            xch_copy_list(x, a.cand, n_pub, XCH_A, 0);
            grid.sync();
            if (lead) {
                XchHdrA hd; hd.n = (unsigned long long)n_pub; hd.vmax = *a.vmax_bits; hd.M_pred = (unsigned long long)*a.M;
                hd.overflow = ns > x->capA ? 1ull : 0ull;
                for (int p = 0; p < x->nranks; p++) x->reg[p]->a[x->rank] = hd;
                xch_signal(x, XCH_A, a.seq);
            }
        }
        if (!(a.stage & SEL_ST_RUN)) return;
        if (lead) {
            const bool ok = xch_wait(x, XCH_A, a.seq);
            unsigned long long ovf = 0;
            long long gM = 0; u64 gv = 0;
            if (ok) {
                const XchHdrA *ha = x->reg[x->rank]->a;
                for (int r = 0; r < x->nranks; r++) {
                    gM += (long long)ld_peer(&ha[r].M_pred);
                    const u64 v = ld_peer(&ha[r].vmax); gv = v > gv ? v : gv;
                    ovf |= ld_peer(&ha[r].overflow);
                }
            }
            if (!ok) sel_halt(sc, a.step, PF_HALT_TIMEOUT); else if (ovf) sel_halt(sc, a.step, PF_HALT_XCH_OVERFLOW);
            sc->gM = gM; sc->gvmax = gv;
            __threadfence();
        }
        grid.sync();
        if (sc->halt) return;
        if (tid == 0) {
            const XchHdrA *ha = x->reg[x->rank]->a;
            s_lv.nseg = x->nranks; s_lv.off[0] = 0;
            for (int r = 0; r < x->nranks; r++) {
                s_lv.seg[r] = xch_listA(x->reg[x->rank], x->capA, r);
                s_lv.off[r + 1] = s_lv.off[r] + (long long)ld_peer(&ha[r].n);
            }
        }
    } else {
        if (tid == 0) {
            s_lv.nseg = 1; s_lv.seg[0] = a.cand; s_lv.off[0] = 0;
            s_lv.off[1] = a.external_sample ? sc->list_n : (long long)sc->cand_chunks * SEL_CHUNK;
        }
    }
    __syncthreads();
    const long long M = x ? sc->gM : *a.M;
    const u64 vmax = x ? sc->gvmax : *a.vmax_bits;      // (external sample: the sample's maximum)
    const int vmax_scale = SEL_ITEM_BITS - exp_of_bits(vmax);
    const bool trivial = a.external_sample ? (M <= N) : (M <= N || vmax == 0);
    const bool sampled = a.external_sample ? !trivial : (!trivial && M > SEL_FINAL_CAP);
    const long long ns_total = s_lv.off[s_lv.nseg];
    const unsigned long long n_cand0 = a.external_sample ? (unsigned long long)ns_total : sc->n_cand;   // sample items (classic: counted by k_sel_sample)
    const u64 vs_bits = a.external_sample ? vmax : sc->vs_bits;
    grid.sync();                                        // every CTA has read the sample's size / counters before the lead resets them

    if (lead) {
        if (a.external_sample) {
            // fused path: k_expand (sample mode) filled the list; this is also the step's reset
            reset_round(sc);
            for (int k = 0; k < 4; k++) sc->Tot.l[k] = 0;
            res->rounds = 0; res->status = 0; res->n_cand_first = 0; res->n_plateau_first = 0; sc->n_fail = 0;
            res->t_ns[0] = gtimer();
            sc->need_tot = 1; sc->fuse_valid = 0; sc->fuse_overflow = 0; sc->tiles_ok = 0;
            sc->vs_bits = vmax;
        }
        if (!x) { sc->gM = M; sc->gvmax = vmax; }
        sc->mult = 1;
        // keep everything (src/fearnhead.jl:135-138): lo = hi = 0 puts every putative on the kept side, whose sum is the
        // (practically) exact one -- the step's total weight does not depend on a bracket scale
        if (trivial) { sc->lo = 0; sc->hi = 0; sc->scale = vmax_scale; sc->mode = 1; }
        else { sc->lo = 1; sc->hi = PF_INF_BITS; sc->scale = vmax_scale; sc->mode = 0; }                     // direct solve, or the sample solve's start
    }
    if (sampled) {
        // the sample solve: every CTA keeps the same search state; one grid barrier per level
        if (tid == 0) { s_st.lo = 1; s_st.hi = PF_INF_BITS; s_st.count_in = (long long)n_cand0; s_st.A_run = 0; s_st.mult = a.fixed_stride; s_st.B_run = mk128(0, 0); s_st.phase = 2; }
        __syncthreads();
        const int ss = SEL_ITEM_BITS - exp_of_bits(vs_bits);         // the sample solve runs at the sample's own scale
        for (int level = 0; level < 8; level++) {
            const u64 blo = s_st.lo, bhi = s_st.hi;
            if (s_st.count_in == 0 || (bhi - blo) <= 1) break;
            const u64 width = bhi - blo;
            int shift = 0;
            while (((width - 1) >> shift) >= SEL_BINS) shift++;
            __syncthreads();
            hist_level(s_lv, blo, bhi, shift, ss, sc, level % 3, smem_raw);
            grid.sync();
            select_bin(&s_st, sc, level % 3, N, blo, bhi, shift, ss, a.margin > 0 ? a.margin : 8, smem_raw, s_warp, s_wc);
            if (s_st.phase == 6) break;
        }
        grid.sync();                                    // every CTA is done with the histogram buffers
        hist_clear(sc);
        if (lead) {
            u64 lo = s_st.lo, hi = s_st.hi;
            if (lo < 1) lo = 1;
            if (hi <= lo) hi = lo + 1;
            sc->lo = lo; sc->hi = hi; sc->mode = 0; sc->mult = 1; sc->n_fail = 0;
            sc->scale = SEL_ITEM_BITS - exp_of_bits(hi >= PF_INF_BITS ? vmax : hi);
            reset_round(sc);
            for (int k = 0; k < 4; k++) sc->Tot.l[k] = 0;
            if (a.external_sample && hi < PF_INF_BITS) {
                // the expansion kernel may classify against this bracket: dense candidate list from 0,
                // kept-side total at the exact scale 52 - e(hi) (valid while the maximum is within 46 binades of hi)
                sc->fuse_valid = 1; sc->fuse_scale = sc->scale; sc->fuse_tot_scale = 52 - exp_of_bits(hi); sc->list_exact = 1;
            }
        }
    } else if (lead && a.external_sample) {
        reset_round(sc);                                 // the classic rounds start from a clean slate
    }
    if (lead) { sc->phase = 0; res->t_ns[1] = gtimer(); }
}

// ------------------------------------------------------------------ k_sel_classify (full grid)
// One pass over all putatives against the current bracket.  Round 0 of a fused step was classified by the
// expansion kernel (no-op here); a queued retry round runs only if the previous solve asked for it.
__global__ void __launch_bounds__(SEL_THREADS, 2) k_sel_classify(SelArgs a)
{
    SelScratch *sc = a.sc;
    if (sc->halt) return;
    if (a.retry ? sc->phase != 3 : (a.preclassified && sc->fuse_valid)) return;
    __shared__ u128 s_red[3][SEL_WARPS];
    __shared__ unsigned long long s_redc[2][SEL_WARPS];
    const long long n_live = *a.n_live;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long gwarp = (long long)blockIdx.x * SEL_WARPS + wid;
    const long long nwarps = (long long)gridDim.x * SEL_WARPS;
    const u64 lo = sc->lo, hi = sc->hi;
    const int scale = sc->scale;
    const bool need_tot = sc->need_tot != 0;
    // kept-side scale from the maximum known here: the exact one on a single GPU; in a sharded run the global
    // one of the previous exchange (round 0: the SAMPLE's maximum -- the solve checks the choice against the
    // true maximum and redoes the round if it was not the exact scale)
    const int ts_round = kept_tot_scale(hi, a.x ? sc->gvmax : *a.vmax_bits);
    const u64 plateau = a.plateau_ptr ? *a.plateau_ptr : 0ull;
    if (blockIdx.x == 0 && tid == 0) { sc->tot_scale_used = ts_round; sc->low_max_valid = 1; }
    unsigned long long A_hi = 0, n_cand = 0, n_plat = 0;
    u64 low_max = 0;
    u128 B_lo = mk128(0, 0), S_c = mk128(0, 0), Tot = mk128(0, 0);
    ChunkWriter cw; cw.init();
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
                if (bits >= hi) { A_hi++; if (need_tot) Tot = add128(Tot, fx128(bits, ts_round)); }
                else if (bits < lo) { B_lo = add128(B_lo, fx70(bits, scale)); low_max = bits > low_max ? bits : low_max; }
                else if (bits == plateau) n_plat++;                  // the mass tie: counted, not listed
                else { is_cand = true; n_cand++; S_c = add128(S_c, fx70(bits, scale)); }
            }
            cw.push(is_cand, bits, a.cand, &sc->cand_chunks, lane);
        }
    }
    cw.finish(a.cand, lane);
    // block reduce, then one set of atomics per block
    A_hi = warp_sum64(A_hi); n_cand = warp_sum64(n_cand); n_plat = warp_sum64(n_plat); low_max = warp_max64(low_max);
    if (lane == 0 && n_plat) atomicAdd(&sc->n_plateau, n_plat);
    if (lane == 0 && low_max) atomicMax(&sc->low_max, low_max);
    B_lo = warp_sum128(B_lo); S_c = warp_sum128(S_c); Tot = warp_sum128(Tot);
    if (lane == 0) { s_redc[0][wid] = A_hi; s_redc[1][wid] = n_cand; s_red[0][wid] = B_lo; s_red[1][wid] = S_c; s_red[2][wid] = Tot; }
    __syncthreads();
    if (wid == 0) {
        unsigned long long x0 = lane < SEL_WARPS ? s_redc[0][lane] : 0, x1 = lane < SEL_WARPS ? s_redc[1][lane] : 0;
        u128 y0 = lane < SEL_WARPS ? s_red[0][lane] : mk128(0, 0), y1 = lane < SEL_WARPS ? s_red[1][lane] : mk128(0, 0),
             y2 = lane < SEL_WARPS ? s_red[2][lane] : mk128(0, 0);
        x0 = warp_sum64(x0); x1 = warp_sum64(x1);
        y0 = warp_sum128(y0); y1 = warp_sum128(y1); y2 = warp_sum128(y2);
        if (lane == 0) {
            if (x0) atomicAdd(&sc->A_hi, x0);
            if (x1) atomicAdd(&sc->n_cand, x1);
            atomic_add_acc(&sc->B_lo, y0);
            atomic_add_acc(&sc->S_cand, y1);
            if (need_tot) atomic_add_acc(&sc->Tot, y2);
        }
    }
}

// ------------------------------------------------------------------ k_sel_solve (cooperative grid)
__global__ void __launch_bounds__(SEL_THREADS, 1) k_sel_solve(SelArgs a)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ __align__(16) unsigned char smem_raw[SEL_SMEM_BYTES];
    u64 *s_u64 = reinterpret_cast<u64 *>(smem_raw);               // SEL_FINAL_CAP u64 for the final sort
    __shared__ ListView s_lv;
    __shared__ SrchSt s_st;
    __shared__ u128 s_warp[33];
    __shared__ unsigned int s_wc[33];
    __shared__ long long s_ll[4];
    __shared__ u128 s_red[SEL_WARPS];
    __shared__ unsigned long long s_redc[SEL_WARPS];

    SelScratch *sc = a.sc;
    SelResult *res = a.res;
    const Xch *x = a.x;
    if (sc->halt) return;
    if (a.retry && sc->phase != 3) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const bool lead = blockIdx.x == 0 && tid == 0;
    const long long N = a.N;
    const int round = a.retry ? res->rounds : 0;
    const bool pre_round = !a.retry && a.preclassified && sc->fuse_valid;      // the expansion kernel classified this round
    const u64 plateau = a.plateau_ptr ? *a.plateau_ptr : 0ull;

    // ---- the candidate list and the accumulators: local, or (sharded) gathered through the exchange regions
    if (x) {
        const int parity = round & 1;
        const unsigned long long seq = a.seq * 64 + (unsigned long long)round;
        if (a.stage & SEL_ST_PUB) {
            const long long nl = sc->list_exact ? sc->list_n : (long long)sc->cand_chunks * SEL_CHUNK;
            const long long n_pub = nl <= x->capB ? nl : x->capB;
            xch_copy_list(x, a.cand, n_pub, XCH_B, parity);
            grid.sync();
            if (lead) {
                SelAcc hd = *reinterpret_cast<const SelAcc *>(reinterpret_cast<const char *>(sc) + SEL_ACC_OFFSET);
                hd.cand_chunks = (unsigned long long)n_pub;            // entries of this rank's segment
                hd.overflow = nl > x->capB ? 1ull : 0ull;
                hd.M_local = (unsigned long long)*a.M; hd.vmax_local = *a.vmax_bits;
                for (int p = 0; p < x->nranks; p++) x->reg[p]->b[parity][x->rank] = hd;
                xch_signal(x, XCH_B, seq);
            }
        }
        if (!(a.stage & SEL_ST_RUN)) return;
        if (lead) {
            const bool ok = xch_wait(x, XCH_B, seq);
            if (ok) {
                const SelAcc *g = x->reg[x->rank]->b[parity];
                unsigned long long A = 0, nc = 0, npl = 0, ovf = 0, Mg = 0, vm = 0, lm = 0;
                acc128 b, c, t;
                for (int k = 0; k < 4; k++) { b.l[k] = 0; c.l[k] = 0; t.l[k] = 0; }
                for (int r = 0; r < x->nranks; r++) {
                    A += ld_peer(&g[r].A_hi); nc += ld_peer(&g[r].n_cand); npl += ld_peer(&g[r].n_plateau); ovf |= ld_peer(&g[r].overflow);
                    Mg += ld_peer(&g[r].M_local);
                    const u64 v = ld_peer(&g[r].vmax_local); vm = v > vm ? v : vm;
                    const u64 l = ld_peer(&g[r].low_max); lm = l > lm ? l : lm;
                    for (int k = 0; k < 4; k++) { b.l[k] += ld_peer(&g[r].B_lo.l[k]); c.l[k] += ld_peer(&g[r].S_cand.l[k]); t.l[k] += ld_peer(&g[r].Tot.l[k]); }
                }
                sc->A_hi = A; sc->n_cand = nc; sc->B_lo = b; sc->S_cand = c; sc->Tot = t; sc->n_plateau = npl; sc->low_max = lm;
                sc->gM = (long long)Mg; sc->gvmax = vm;
                if (ovf) sel_halt(sc, a.step, PF_HALT_XCH_OVERFLOW);
            } else sel_halt(sc, a.step, PF_HALT_TIMEOUT);
            __threadfence();
        }
        grid.sync();
        if (sc->halt) return;
        if (tid == 0) {
            const SelAcc *g = x->reg[x->rank]->b[parity];
            s_lv.nseg = x->nranks; s_lv.off[0] = 0;
            for (int r = 0; r < x->nranks; r++) {
                s_lv.seg[r] = xch_listB(x->reg[x->rank], x->nranks, x->capA, x->capB, parity, r);
                s_lv.off[r + 1] = s_lv.off[r] + (long long)ld_peer(&g[r].cand_chunks);
            }
        }
    } else {
        if (tid == 0) {
            s_lv.nseg = 1; s_lv.seg[0] = a.cand; s_lv.off[0] = 0;
            s_lv.off[1] = sc->list_exact ? sc->list_n : (long long)sc->cand_chunks * SEL_CHUNK;
        }
    }
    __syncthreads();
    const long long M = x ? sc->gM : *a.M;
    const u64 vmax = x ? sc->gvmax : *a.vmax_bits;
    const int vmax_scale = SEL_ITEM_BITS - exp_of_bits(vmax);
    const bool trivial = (M <= N || vmax == 0);

    const u64 lo = sc->lo, hi = sc->hi;
    const int scale = sc->scale;
    const int mode = sc->mode;
    const bool keep_all = (mode == 1);
    const bool need_tot = sc->need_tot != 0;
    grid.sync();                             // everyone has read the round's bracket before CTA 0 may change it

    // ---- CTA 0: verify the bracket, start the descent
    if (lead) {
        if (!x) { sc->gM = M; sc->gvmax = vmax; }
        res->rounds = round + 1;
        // fused path: the expansion's classification is unusable if a block overflowed its candidate
        // buffer or the true maximum lies more than 46 binades above hi (kept-side scale not exact)
        const bool pre_bad = pre_round && (sc->fuse_overflow || kept_tot_scale(hi, vmax) != sc->fuse_tot_scale);
        // sharded classic round: k_sel_classify chose the kept-side scale from the maximum it knew; redo the
        // round (the true maximum is known from now on) if that was not the scale the true maximum asks for
        bool ts_bad = !pre_round && x && need_tot && kept_tot_scale(hi, vmax) != sc->tot_scale_used;
        // a bracket that is open at the top was scaled by the maximum known when it was set (the SAMPLE's maximum
        // behind a sample expansion); the true maximum is known now: redo the round at its scale if they differ
        if (!pre_round && !keep_all && mode == 0 && hi >= PF_INF_BITS && scale != vmax_scale && vmax != 0) { ts_bad = true; sc->scale = vmax_scale; }
        if (need_tot && !pre_bad && !ts_bad) {
            if (round == 0) { res->n_cand_first = (long long)sc->n_cand; res->n_plateau_first = (long long)sc->n_plateau; res->t_ns[2] = gtimer(); }
            // total weight = kept side (>= hi) + low side (< hi, bracket scale)
            const int ts = pre_round ? sc->fuse_tot_scale : sc->tot_scale_used;
            u128 low = add128(acc_value(&sc->B_lo), acc_value(&sc->S_cand));
            if (sc->n_plateau) low = add128(low, mul128_64(fx70(plateau, scale), (u64)sc->n_plateau));
            res->log_total = log_total_exact(acc_value(&sc->Tot), ts, low, scale);
            sc->need_tot = 0;
        }
        int phase = 2;       // 2 = descend, 3 = another classify round with a new bracket / scale, 4 = done
        if (pre_bad || ts_bad) {
            reset_round(sc); for (int k = 0; k < 4; k++) sc->Tot.l[k] = 0; sc->fuse_valid = 0; phase = 3;
            res->status |= (pre_round && sc->fuse_overflow) ? 2 : 4;
        }
        else if (keep_all) phase = 4;
        else if (mode == 2) {
            // comb-rescale round: T at the new scale; lo = hi = thr so A_hi = A again
            res->T = acc_value(&sc->B_lo); res->scale = scale;
            phase = comb_rescale(sc, res, N) ? 3 : 4;
        } else {
            if (pre_round) res->status |= 1;
            u128 B = acc_value(&sc->B_lo), Sc = acc_value(&sc->S_cand);
            long long Ahi = (long long)sc->A_hi, nc = (long long)sc->n_cand;
            const long long np = (long long)sc->n_plateau;            // in-bracket copies of the plateau value (not in the list)
            const u128 Pm = np ? mul128_64(fx70(plateau, scale), (u64)np) : mk128(0, 0);
            bool fail_lo = (lo > 1) && pred128(B, N, Ahi + nc + np, fx70(lo, scale));
            bool fail_hi = (hi < PF_INF_BITS) && !pred128(add128(add128(B, Sc), Pm), N, Ahi, fx70(hi, scale));
            if (fail_lo || fail_hi) {
                res->status |= fail_lo ? 8 : 16;
                // the sample missed by more than its margin: move the bracket to the failing side and
                // make it 8x wider (64x the second time); the third failure opens it completely
                const int nf = ++sc->n_fail;
                const u64 w = hi - lo;
                const bool bounded = nf <= 2 && hi < PF_INF_BITS && lo > 1 && w < (1ull << 56);
                const u64 grow = bounded ? w << 3 : 0;
                u64 nlo, nhi;
                if (fail_hi) { nlo = hi; nhi = (bounded && hi + grow < PF_INF_BITS) ? hi + grow : PF_INF_BITS; }
                else { nhi = lo; nlo = (bounded && lo > grow + 1) ? lo - grow : 1ull; }
                sc->lo = nlo; sc->hi = nhi; sc->mode = 0;
                sc->scale = SEL_ITEM_BITS - exp_of_bits(nhi >= PF_INF_BITS ? vmax : nhi);
                reset_round(sc);
                phase = 3;
            } else {
                sc->B_run = B; sc->A_run = Ahi; sc->count_in = nc;
                if (np) phase = 8;           // first decide on which side of the plateau the threshold lies
            }
        }
        sc->phase = phase;
        __threadfence();
    }
    grid.sync();
    if (sc->phase == 8) {
        // ---- plateau resolution: mass and count of the candidates below the plateau value P, then the
        // predicate at P itself; the bracket shrinks to [lo, P) or (P, hi) and never contains the tie
        u128 sb = mk128(0, 0);
        unsigned long long cb = 0;
        lv_for_each(s_lv, (long long)blockIdx.x * SEL_THREADS + tid, (long long)gridDim.x * SEL_THREADS, [&](u64 bits) {
            if (bits >= lo && bits < plateau) { cb++; sb = add128(sb, fx70(bits, scale)); }
        });
        sb = warp_sum128(sb); cb = warp_sum64(cb);
        if (lane == 0) { s_red[wid] = sb; s_redc[wid] = cb; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < SEL_WARPS; w++) { sb = add128(sb, s_red[w]); cb += s_redc[w]; }
            if (cb) atomicAdd(&sc->pl_cnt, cb);
            atomic_add_acc(&sc->pl_sum, sb);
            __threadfence();
        }
        grid.sync();
        if (lead) {
            const long long np = (long long)sc->n_plateau, cbt = (long long)sc->pl_cnt;
            const u128 qP = fx70(plateau, scale);
            const u128 Bp = add128(sc->B_run, acc_value(&sc->pl_sum));
            const long long A_geqP = sc->A_run + (sc->count_in - cbt) + np;
            if (pred128(Bp, N, A_geqP, qP)) { sc->hi = plateau; sc->A_run = A_geqP; sc->count_in = cbt; }
            else { sc->lo = plateau + 1; sc->B_run = add128(Bp, mul128_64(qP, (u64)np)); sc->count_in = sc->count_in - cbt; }
            sc->phase = 2;
            __threadfence();
        }
        grid.sync();
    }
    if (sc->phase == 3) return;              // a queued retry round (k_sel_classify + k_sel_solve) takes over

    if (sc->phase == 2) {
        // ---- radix descent over the candidate list: every CTA keeps the same search state, one grid barrier per level
        if (tid == 0) { s_st.lo = sc->lo; s_st.hi = sc->hi; s_st.count_in = sc->count_in; s_st.A_run = sc->A_run; s_st.mult = 1; s_st.B_run = sc->B_run; s_st.phase = 2; }
        __syncthreads();
        int levels = 0;
        for (int level = 0; level < 8; level++) {
            const u64 blo = s_st.lo, bhi = s_st.hi;
            if (s_st.count_in <= SEL_FINAL_CAP) {
                // compact the bracket's items into the final list
                lv_for_each(s_lv, (long long)blockIdx.x * SEL_THREADS + tid, (long long)gridDim.x * SEL_THREADS, [&](u64 bits) {
                    if (bits >= blo && bits < bhi) {
                        unsigned int pos = atomicAdd(&sc->final_n, 1u);
                        if (pos < SEL_FINAL_CAP) sc->final_vals[pos] = __longlong_as_double((long long)bits);
                    }
                });
                __threadfence();
                break;
            }
            if ((bhi - blo) == 1) break;       // one value with > SEL_FINAL_CAP ties: closed form below
            const u64 width = bhi - blo;
            int shift = 0;
            while (((width - 1) >> shift) >= SEL_BINS) shift++;
            __syncthreads();
            hist_level(s_lv, blo, bhi, shift, scale, sc, level % 3, smem_raw);
            grid.sync();
            select_bin(&s_st, sc, level % 3, N, blo, bhi, shift, scale, 0, smem_raw, s_warp, s_wc);
            levels++;
        }
        grid.sync();                          // final list complete; every CTA is done with the histogram buffers
        if (levels) hist_clear(sc);

        // ---- CTA 0: exact solve on the final list
        if (blockIdx.x == 0) {
            const u64 blo = s_st.lo, bhi = s_st.hi;
            const long long count_in = s_st.count_in;
            const u128 B0 = s_st.B_run;
            const long long A0 = s_st.A_run;
            if (count_in > SEL_FINAL_CAP) {
                // (bhi - blo) == 1: count_in copies of the single value blo
                if (tid == 0) {
                    u128 q = fx70(blo, scale);
                    bool ok = pred128(B0, N, A0 + count_in, q);
                    s_ll[1] = ok ? (long long)blo : (long long)bhi;
                    s_ll[2] = ok ? A0 + count_in : A0;
                    s_warp[0] = ok ? B0 : add128(B0, mul128_64(q, (u64)count_in));
                }
                __syncthreads();
            } else {
                int nf = (int)count_in;
                int np2 = 2; while (np2 < nf) np2 <<= 1;
                for (int t = tid; t < np2; t += SEL_THREADS) s_u64[t] = t < nf ? (u64)__double_as_longlong(__ldcg(&sc->final_vals[t])) : ~0ull;
                __syncthreads();
                bitonic_sort_smem(s_u64, np2);
                const int per = SEL_FINAL_CAP / SEL_THREADS;
                u128 loc = mk128(0, 0);
                for (int k = 0; k < per; k++) { int t = tid * per + k; if (t < nf) loc = add128(loc, fx70(s_u64[t], scale)); }
                u128 tot;
                u128 pre = block_excl_scan128(loc, s_warp, &tot);
                int first = nf;
                for (int k = 0; k < per; k++) {
                    int t = tid * per + k;
                    if (t < nf) {
                        u64 bits = s_u64[t];
                        bool head = (t == 0) || (s_u64[t - 1] != bits);
                        if (head && first == nf && pred128(add128(B0, pre), N, A0 + (nf - t), fx70(bits, scale))) first = t;
                        pre = add128(pre, fx70(bits, scale));
                    }
                }
                if (tid == 0) s_ll[0] = nf;
                __syncthreads();
                atomicMin(&s_ll[0], (long long)first);
                __syncthreads();
                int p = (int)s_ll[0];
                u128 part = mk128(0, 0);
                for (int k = 0; k < per; k++) { int t = tid * per + k; if (t < nf && t < p) part = add128(part, fx70(s_u64[t], scale)); }
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
                const u64 thr = (u64)s_ll[1];
                const long long A = s_ll[2];
                const u128 T = s_warp[0];
                sc->lo = blo; sc->hi = bhi; sc->count_in = count_in; sc->B_run = B0; sc->A_run = A0;      // (mirror of the search state)
                // precision: the scale must sit within 8 binades of the threshold
                bool precise = (thr >= PF_INF_BITS) ? (scale == vmax_scale) : ((SEL_ITEM_BITS - scale) - exp_of_bits(thr) <= 8);
                if (!precise && round < 30) {
                    res->status |= 32;
                    int e = exp_of_bits(thr);
                    long long elo = e - 1 + 1023, ehi = e + 2 + 1023;
                    u64 nlo = elo <= 0 ? 1ull : ((u64)elo << 52);
                    u64 nhi = ehi >= 2047 ? PF_INF_BITS : ((u64)ehi << 52);
                    if (thr >= PF_INF_BITS) { nlo = 1; nhi = PF_INF_BITS; }
                    sc->lo = nlo; sc->hi = nhi; sc->mode = 0;
                    sc->scale = SEL_ITEM_BITS - exp_of_bits(nhi >= PF_INF_BITS ? vmax : nhi);
                    reset_round(sc);
                    sc->phase = 3;
                } else {
                    res->thr_bits = thr; res->A = A; res->T = T; res->scale = scale;
                    sc->phase = comb_rescale(sc, res, N) ? 3 : 4;
                    if (sc->phase == 3) res->status |= 64;
                }
                __threadfence();
            }
        }
        grid.sync();
        if (sc->phase != 4) return;
    }

    // ---------------- outputs (CTA 0, thread 0)
    if (lead) {
        res->t_ns[3] = gtimer();
        if (trivial) {
            const bool none = (vmax == 0 && M > N);
            res->keep_all = 1; res->thr_bits = none ? PF_INF_BITS : 0ull; res->A = none ? 0 : M; res->n = 0;
            res->T = mk128(0, 0); res->Uoff = mk128(0, 0); res->Tq = mk128(0, 0); res->Tr = 0; res->scale = sc->scale;
            res->lw_resamp = __longlong_as_double(0xFFF0000000000000ll); res->log_c = res->lw_resamp;
            res->thr_value = 0.0;
            sc->tiles_ok = 0;
        } else {
            res->keep_all = 0;
            long long n = N - res->A;
            if (n < 0) { n = 0; res->status |= 128; }          // cannot happen with consistent inputs
            res->n = n;
            const u128 T = res->T;
            // Uoff = floor(mu * T / 2^53), mu = floor(u 2^53): 192-bit product, then >> 53
            const double u = *a.u;
            const u64 mu = (u64)(u * 9007199254740992.0);
            u64 p0 = T.lo * mu, p0h = mulhi64(T.lo, mu);
            u64 p1 = T.hi * mu, p1h = mulhi64(T.hi, mu);
            u64 w0 = p0, w1 = p0h + p1, w2 = p1h + (w1 < p0h ? 1ull : 0ull);
            u128 U;
            U.lo = (w0 >> 53) | (w1 << 11);
            U.hi = (w1 >> 53) | (w2 << 11);
            res->Uoff = U;
            if (n > 0) { unsigned int tr; res->Tq = divmod128_32(T, (unsigned int)n, &tr); res->Tr = tr; }
            else { res->Tq = mk128(0, 0); res->Tr = 0; }
            const double c = n > 0 ? ldexp(to_double128(T) / (double)n, -res->scale) : __longlong_as_double(0x7FF8000000000000ll);
            res->log_c = log(c);
            res->lw_resamp = res->log_c - res->log_total;
            res->thr_value = __longlong_as_double((long long)res->thr_bits);
            // the tile records the expansion wrote stay usable if no classic round overwrote the list or the scale
            sc->tiles_ok = (a.preclassified && sc->fuse_valid && res->rounds == 1 && res->scale == sc->fuse_scale) ? 1 : 0;
        }
        res->t_ns[4] = gtimer();
    }
}
