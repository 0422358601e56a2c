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
    unsigned int ticket;    // (unused)
    unsigned int ticket2;
    long long n_picked;     // resampled survivors
    long long n_kept;       // kept survivors
    long long M_pred;       // putatives of this step, known before the expansion (counted by the previous instantiate)
    long long M_next;       // putatives of the next step (accumulated by k_instantiate)
    u64 vmax_sample;        // max putative weight of the sample expansion
    long long goff;         // sharded run: global index of this shard's first particle
    long long n_global;     // sharded run: population over all shards after the step
};

struct StepLog {            // one record per step, read back after pf_filter
    long long n_in, M, n_kept, n_req, n_got, n_out, overflow, n_cand, n_plateau;
    double log_total, lw_resamp, thr, cv;
    int resampled, rounds, status, pad;
    double sel_us[4];
};

// ------------------------------------------------------------------ k_prestep
// One thread: roll the control block over and derive the per-observation constants:
// the new-cluster putative (src/particle.jl:108-112 with comp = p.prior) and the slot a
// new cluster starts with (add(prior, y), src/component.jl:85).
template <int D>
__global__ void k_prestep(StepState *st, const PriorConst *pc, const double *y, StepConst *sc_out, int first,
                          SelScratch *guard, const SelResult *prev, const Xch *xch, unsigned long long seq, long long step, int forced = -1)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (guard && guard->halt) return;
    // sharded run: the previous generation is complete on every rank (peers stored children into this shard)
    if (xch && !xch_wait(xch, XCH_D, seq - 1)) { sel_halt(guard, step, PF_HALT_TIMEOUT); return; }
    if (guard) guard->list_n = 0;                      // the sample expansion appends from 0
    using L = Layout<D>;
    // first: 1 = the filter's first observation, 2 = a halted step is run again (the roll-over already happened),
    // 3 = pipelined step: the previous step's children are instantiated WHILE they are expanded for this observation,
    // so the putative count is not known yet (k_count_next accumulates it)
    if (first == 0 || first == 3) st->n_live = st->n_next;
    if (first == 3) st->M_pred = 0;
    else if (first != 2) st->M_pred = first ? st->n_live : st->M_next;      // (first step: empty particles, one putative each)
    st->M_next = 0; st->vmax_sample = 0;
    st->n_next = 0;
    st->M = 0; st->vmax_bits = 0; st->ticket = 0; st->ticket2 = 0; st->n_picked = 0; st->n_kept = 0;
    StepConst c;
    c.forced = forced; c.pad = 0;
    c.prev_lw_resamp = prev ? prev->lw_resamp : 0.0; c.prev_log_total = prev ? prev->log_total : 0.0;
    double dx[D], z[D], dinv[D], off[L::NOFF > 0 ? L::NOFF : 1];
#pragma unroll
    for (int i = 0; i < D; i++) { c.y[i] = y[i]; dx[i] = y[i] - pc->mu0[i]; dinv[i] = pc->L0_dinv[i]; }
#pragma unroll
    for (int k = 0; k < L::NOFF; k++) off[k] = pc->L0_off[k];
    double q = quad_form<D>(dx, dinv, off, z);
    const NRow r0 = pc->row0;
    double l1p = log1p(r0.r * q);
    c.lw_new = r0.C - 0.5 * pc->logdet0 - r0.h * l1p;
    // exactly what k_expand evaluates for the new-cluster putative of a particle whose log-weight is lw_resamp
    c.plateau_bits = (prev && first != 1) ? (unsigned long long)__double_as_longlong(exp(prev->lw_resamp + c.lw_new)) : 0ull;
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

#define SCAN_THREADS 256
struct TileSum {           // survivor-scan record of one tile of SCAN_THREADS particles
    u128 X;
    unsigned int kept;
    unsigned int pad[3];
};

// ------------------------------------------------------------------ k_expand
// putatives(p, y) for every particle (src/fearnhead.jl:132, src/particle.jl:100-116):
// thread i evaluates the K_i + 1 candidate assignments of particle i and writes their
// LINEAR weights exp(logweight) slot-major into V.  One thread per particle keeps all 32
// lanes on the FP64 pipe and makes every (slot, field) load a 256-byte coalesced read.
//
// MODE 0  plain expansion.
// MODE 1  sample expansion: only every `stride`-th particle, nothing is stored except the
//         putative weights themselves, appended densely to the sample list (the threshold
//         search brackets the threshold on them BEFORE the full expansion runs).
// MODE 2  fused expansion: as MODE 0, and every putative is classified on the fly against the
//         bracket [lo, hi) (exact fixed-point mass below lo, count above hi, candidates in
//         between appended densely), so the separate classify pass over V disappears; the
//         block's sums are also the survivor scan's tile record (one block = one tile).
// Block size of the full expansions (MODE 0 / 2).  The survivor scan's tile (SCAN_THREADS particles) is EXP_SUB such
// blocks: a block leaves a SUB-tile record, k_tile_scan1 adds them up.  Smaller blocks shorten the tail in which a
// block's finished warps wait for its slowest one (17 % of the stall samples at 256 threads), but every block ends with
// ~15 atomics on the step's accumulators, and those serialise per address.  Measured at 2^24 particles, fused kernel:
// register-prefetch form 3.83 ms with 256-thread blocks, 3.81 with 128, 5.50 with 64, 8.59 with 32; asynchronous-copy
// form (below) 3.48 with 256, 3.40 with 128 (profiles/r2/README.md, ab_cpa_4.jsonl).
#ifndef EXP_THREADS
#define EXP_THREADS 128
#endif
#define EXP_SUB (256 / EXP_THREADS)
#define EXP_CAND_CAP (4 * EXP_THREADS)   // MODE 2: candidates one block may stage
#define EXP_SAMPLE_THREADS 32      // MODE 1 runs one warp per block: at most 32 * PF_MAX_SLOTS sample values per block
struct TileFuse {
    unsigned int cand_off, cand_cnt, n_plateau, pad;
};
// Per-particle record of the fused expansion for the survivor scan: the fixed-point mass of the particle's
// putatives below the bracket (< 2^77: 120 bits are plenty) with the number of putatives above the bracket
// in the top byte, and a 64-bit mask of the slots that fell INSIDE the bracket (candidates and plateau
// copies).  k_scan_apply then needs V only for the in-bracket slots (~0.5 % of the putatives) when it
// forms the particle's exact (mass, kept) pair.
struct ScanRec {
    ulonglong2 *x;
    unsigned long long *m;
    unsigned long long *h;     // mask of the slots ABOVE the bracket (kept whatever the threshold turns out to be)
};

#ifndef PF_CPASYNC_DEPTH
#define PF_CPASYNC_DEPTH 1                 // slots in flight as asynchronous global -> shared copies (0: register prefetch)
#endif
#ifndef PF_CPASYNC_STAGES
#define PF_CPASYNC_STAGES (PF_CPASYNC_DEPTH + 1)   // staging slots; 1 (with depth 1): the slot is refilled right after it has been read
#endif
#ifndef PF_CPASYNC_SMEM_MAX
#define PF_CPASYNC_SMEM_MAX (16 * 1024)    // staging buffer per block (14 KB for the 2-D FP64 record); larger records keep the register
                                           // prefetch: shared memory is taken from the L1 the parent gather lives on (deeper pipelines lost to it)
#endif
// dynamic shared memory of the fused kernels k_expand<D, 2, true, R, ...> (single GPU and sharded): staging slots x F fields x block
template <int D, typename R>
__host__ __device__ constexpr size_t instexp_staging_bytes()
{
    return PF_CPASYNC_DEPTH > 0 && (size_t)PF_CPASYNC_STAGES * Layout<D>::F * EXP_THREADS * sizeof(R) <= PF_CPASYNC_SMEM_MAX
               ? (size_t)PF_CPASYNC_STAGES * Layout<D>::F * EXP_THREADS * sizeof(R) : 0;
}
template <int BYTES>
__device__ __forceinline__ void cp_async_elem(void *smem_dst, const void *gsrc)
{
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(d), "l"(gsrc), "n"(BYTES));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
#ifndef PF_UNROLL2
#define PF_UNROLL2 1                       // slot loop of the INST kernels on two alternating register sets
#endif
#ifndef PF_EXPAND_MINBLOCKS
#define PF_EXPAND_MINBLOCKS 4
#endif
#ifndef PF_INST_MINBLOCKS
#define PF_INST_MINBLOCKS 1
#endif
// Sharded run (particles block-sharded over ranks, global order = rank-major): where this
// rank's survivors sit in the global survivor order.  The new population is re-partitioned into
// equal blocks of q = ceil(n / G) particles at EVERY step (rank r owns global [r q, (r+1) q)):
// k_instantiate stores a child that changes owner straight into its new owner's arrays.
struct RankOff {
    u128 S_off;                    // fixed-point low mass of all lower ranks
    long long kept_off;            // survivors kept on all lower ranks
    long long g_first;             // global index of this rank's first survivor
    long long n_local;             // survivors produced by this rank
    long long n_mine;              // particles this rank owns after re-partitioning
    long long n_global, q;
    long long goff_prev;           // global index of this rank's first PARENT (genealogy)
    long long goff_next;           // global index of this rank's first particle after the step
    int rank, nranks;
};

// Every rank's next-generation arrays mapped into this process (cudaIpc, or plain pointers when
// the ranks live in one process): k_instantiate stores a child that changes owner straight into
// its new owner's store over NVLink -- no packing, no send/recv buffers, no host-known sizes.
struct PeerTable {
    void *S[PF_MAX_RANKS];                 // (elements of the handle's storage type)
    int *Kc[PF_MAX_RANKS];
    double *logw[PF_MAX_RANKS];
    unsigned int *gen_anc[PF_MAX_RANKS];
    unsigned char *gen_asg[PF_MAX_RANKS];
    // pipelined sharded steps: the parent's owner also expands the child for the next observation and delivers the
    // putative weights (two alternating buffers), their count and the per-particle scan record with the child
    double *V[2][PF_MAX_RANKS];
    unsigned char *cnt[PF_MAX_RANKS];
    ulonglong2 *recx[PF_MAX_RANKS];
    unsigned long long *recm[PF_MAX_RANKS], *rech[PF_MAX_RANKS];
};

// INST (pipelined step): particle i is child i of the PREVIOUS step's survivor list.  Its record does not exist yet:
// the thread gathers the parent's slots, applies add(comp, y_prev) to the chosen one in registers
// (instantiate, src/particle.jl:131-149), stores the child (MODE 2) and evaluates the child's putatives for THIS
// observation from the registers -- the separate read of the new population by the expansion disappears.
struct InstArgs {
    const unsigned int *surv_parent;       // survivor list of the previous step
    const unsigned char *surv_slot;
    const double *V_prev;                  // the previous step's putative weights
    const StepConst *scp_prev;             // ... observation and new-cluster slot
    void *S2; int *Kc2; double *logw2;     // the children's store (MODE 2; elements of the kernel's storage type)
    unsigned int *gen_anc; unsigned char *gen_asg;
    // sharded run: child t of this rank's survivor list is global particle ro->g_first + t of the next generation and is
    // delivered to its new owner (peer memory over NVLink) together with its putatives for the next observation
    const RankOff *ro;
    const PeerTable *peers;                // tables of the store buffer the children are written to
    int vnext;                             // which of the two V buffers receives the next observation's putatives
    long long gen_off;                     // offset of this generation's row in the genealogy arrays
    long long vb0;                         // STRIDED launch: first tile (= number of tiles the main launch covered)
    // the two observations and the prior mean BY VALUE: kernel parameters live in the constant bank and are used as
    // operands directly, which frees the registers the fused kernel would otherwise spend on copies of them
    double yc[PF_MAX_D], y_prev[PF_MAX_D], mu0c[PF_MAX_D];
};
#ifndef PF_INSTEXP_MINBLOCKS
#define PF_INSTEXP_MINBLOCKS 3
#endif
#ifndef PF_INSTEXP_REMOTE_MINBLOCKS
#define PF_INSTEXP_REMOTE_MINBLOCKS 3      // (the sharded variant carries the destination pointers: 208 bytes of spills at 80 registers)
#endif
// R: storage type of the particle store (double, or float under PF_F32: the arithmetic stays in double, every field
// is rounded to R when it is stored, so the store traffic halves)
// REMOTE (sharded INST): the child and its putatives are delivered to the child's new owner (InstArgs.ro / peers).
// STRIDED (a second, small launch of the sharded variant): the tiles from InstArgs.vb0 on, blocks striding over them -- a
// rank may produce more survivors than the main launch's grid covers; keeping the loop out of the main launch is worth
// 11 % of its time.
// (records larger than the 2-D one: the fused kernels trade occupancy for registers -- at 80 registers the 8-D kernel
// spills 2.6 KB per thread; PF_INSTEXP_BY_DIM=0 keeps the 2-D setting for every d)
#ifndef PF_INSTEXP_BY_DIM
#define PF_INSTEXP_BY_DIM 1
#endif
template <int D, bool REMOTE>
__host__ __device__ constexpr int instexp_minblocks()
{
    const int base = REMOTE ? PF_INSTEXP_REMOTE_MINBLOCKS : PF_INSTEXP_MINBLOCKS;
    return !PF_INSTEXP_BY_DIM || D <= 2 ? base : (D <= 4 ? 2 : 1);
}
template <int D>
__host__ __device__ constexpr int expand_minblocks()          // the expansions from the store (64 registers at d <= 2)
{
    return !PF_INSTEXP_BY_DIM || D <= 2 ? PF_EXPAND_MINBLOCKS : (D <= 4 ? 2 : 1);
}
template <int D, int MODE, bool INST = false, typename R = double, bool REMOTE = false, bool STRIDED = false>
__global__ void __launch_bounds__(EXP_THREADS, (INST ? instexp_minblocks<D, REMOTE>() : expand_minblocks<D>()) * EXP_SUB) k_expand(const R *__restrict__ S, const double *__restrict__ logw,
                                                const int *__restrict__ Kc, long long cap, int k_max,
                                                const NRow *__restrict__ tab, const PriorConst *__restrict__ pc,
                                                const StepConst *__restrict__ scp, StepState *st,
                                                double *__restrict__ V, unsigned char *__restrict__ cnt,
                                                SelScratch *sc, long long stride, double *__restrict__ list,
                                                TileSum *__restrict__ tiles, TileFuse *__restrict__ tfuse,
                                                ScanRec rec = ScanRec{nullptr, nullptr, nullptr}, InstArgs ia = InstArgs{})
{
    using L = Layout<D>;
    if (sc && sc->halt) return;
    __shared__ u64 s_max[8];
    __shared__ long long s_M[8], s_ovf[8];
    constexpr int CAND_CAP = (MODE == 1) ? EXP_SAMPLE_THREADS * PF_MAX_SLOTS : (MODE == 2 ? EXP_CAND_CAP : 1);
    __shared__ u64 s_cand[CAND_CAP];
    __shared__ unsigned int s_ncand, s_goff;
    __shared__ u128 s_tot[8];                    // kept-side total per warp
    __shared__ u128 s_red[8];
    __shared__ unsigned int s_redk[8], s_redp[8];
    // asynchronous-copy form of the slot loop (fused kernels, records small enough for the staging buffer)
    constexpr bool CPA = MODE == 2 && INST && instexp_staging_bytes<D, R>() > 0;
    extern __shared__ __align__(16) unsigned char s_dyn[];      // (the launch passes instexp_staging_bytes<D, R>())
    R *const s_pre = reinterpret_cast<R *>(s_dyn);
    constexpr bool remote = INST && REMOTE;
    const long long n_live = remote ? ia.ro->n_local : st->n_live;
    constexpr bool strided = STRIDED;
    const long long vb_end = strided ? (n_live + blockDim.x - 1) / blockDim.x : 0;
    long long vb = (strided ? ia.vb0 : 0) + blockIdx.x;
    if (strided && vb >= vb_end) return;
  do {
    const long long gid = vb * blockDim.x + threadIdx.x;
    // MODE 1 samples aligned groups of 4 consecutive particles (one 32-byte sector per field row):
    // GLOBAL particle g is in the sample iff (g / 4) % stride == 0; a shard starts at global index st->goff
    long long i = gid;
    if (MODE == 1) {
        const long long goff = remote ? ia.ro->g_first : st->goff;
        i = (goff / (4 * stride) + (gid >> 2)) * (4 * stride) + (gid & 3) - goff;
    }
    // where particle i's outputs go (sharded INST: the arrays of the child's new owner `dest`; -1 = this rank's own)
    long long di = i;
    double *Vd = V;
    R *S2d = static_cast<R *>(ia.S2);
    int dest = -1;
    const bool fused = (MODE == 2) && sc->fuse_valid;
    u64 lo = 0, hi = 0; int scale = 0, tscale = 0;
    if (fused) { lo = sc->lo; hi = sc->hi; scale = sc->fuse_scale; tscale = sc->fuse_tot_scale; }
    if (MODE != 0) {
        if (threadIdx.x == 0) s_ncand = 0;
        __syncthreads();
    }
    u64 vmax = 0;
    long long myM = 0, myOvf = 0;
    u128 B_lo = mk128(0, 0), Tot = mk128(0, 0);
    unsigned int kept_hi = 0, n_plat = 0;
    unsigned long long in_mask = 0, hi_mask = 0;
    const u64 plateau = scp->plateau_bits;
    // what happens to one putative weight
    auto emit = [&](int j, double v) {
        const u64 b = (u64)__double_as_longlong(v);
        vmax = b > vmax ? b : vmax;
        if (MODE == 1) {
            if (b != 0) { unsigned int pos = atomicAdd(&s_ncand, 1u); if (pos < CAND_CAP) s_cand[pos] = b; }
            return;
        }
        Vd[(long long)j * cap + di] = v;
        if (MODE == 2 && fused) {
            if (b >= hi) {
                hi_mask |= 1ull << j;
                Tot = add128(Tot, fx128(b, tscale));              // < 2^99 each by the 28-binade guard of the search
            } else if (b < lo) {
                B_lo = add128(B_lo, fx70(b, scale));
            } else if (b == plateau) {
                n_plat++;                                          // the mass tie: counted, not listed
                in_mask |= 1ull << j;
            } else {
                unsigned int pos = atomicAdd(&s_ncand, 1u);
                if (pos < CAND_CAP) s_cand[pos] = b;
                in_mask |= 1ull << j;
            }
        }
    };
    if (i >= 0 && i < n_live) {
        double y[D], mu0[D];
#pragma unroll
        for (int k = 0; k < D; k++) { y[k] = INST ? ia.yc[k] : scp->y[k]; mu0[k] = INST ? ia.mu0c[k] : pc->mu0[k]; }
        long long src = i;                   // the particle whose slots are read
        int jsel = -1;                       // INST: the slot that receives the previous observation
        int K; double lw0;
        if (INST) {
            const unsigned int par = ia.surv_parent[i];
            const unsigned char sl = ia.surv_slot[i];
            jsel = sl & 0x7F;
            K = Kc[par];
            PF_ASSERT((long long)par < cap && K >= 0 && K <= k_max && jsel <= K && jsel < k_max && i < cap);
            const double v = ia.V_prev[(long long)jsel * cap + par];
            lw0 = (sl & 0x80) ? scp->prev_lw_resamp : (log(v) - scp->prev_log_total);      // src/fearnhead.jl:172-179
            src = par;
            int *Kc2d = ia.Kc2; double *logw2d = ia.logw2;
            unsigned int *gen_ancd = ia.gen_anc; unsigned char *gen_asgd = ia.gen_asg;
            unsigned int anc_off = 0;
            if (remote) {
                const long long g = ia.ro->g_first + i;
                const int dr = (int)(g / ia.ro->q);
                PF_ASSERT(dr >= 0 && dr < ia.ro->nranks);
                di = g - (long long)dr * ia.ro->q;
                anc_off = (unsigned int)ia.ro->goff_prev;
                if (MODE == 2 && dr != ia.ro->rank) {
                    dest = dr;
                    S2d = static_cast<R *>(ia.peers->S[dr]); Kc2d = ia.peers->Kc[dr]; logw2d = ia.peers->logw[dr];
                    gen_ancd = ia.peers->gen_anc[dr] ? ia.peers->gen_anc[dr] + ia.gen_off : nullptr;
                    gen_asgd = ia.peers->gen_asg[dr] ? ia.peers->gen_asg[dr] + ia.gen_off : nullptr;
                    Vd = ia.peers->V[ia.vnext][dr];
                }
            }
            if (MODE == 2) {
                logw2d[di] = lw0; Kc2d[di] = K + (jsel == K ? 1 : 0);
                if (gen_ancd) { gen_ancd[di] = par + anc_off; gen_asgd[di] = (unsigned char)jsel; }
            }
        } else {
            K = Kc[i]; lw0 = logw[i];
            PF_ASSERT(K >= 0 && K <= k_max);
        }
        // software pipeline: the fields of slot j+1 are in flight while slot j is evaluated
        double fcur[L::F], fnxt[L::F];
        if (K > 0 && !CPA) {
            const R *base = S + src;
#pragma unroll
            for (int f = 0; f < L::F; f++) fcur[f] = (double)base[(long long)f * cap];
        }
        auto eval = [&](const double *fl, int j) {
            double z[D], q;
            const NRow row = tab[(int)fl[L::F_N]];
            double lw = lw0 + slot_logw<D>(row, fl[L::F_LOGDET], fl + L::F_MEAN, fl + L::F_DINV, fl + L::F_OFF, y, mu0, z, &q);
            emit(j, exp(lw));
        };
        auto store_child = [&](const double *fl, int j) {
            R *dst = S2d + ((long long)j * L::F) * cap + di;
#pragma unroll
            for (int f = 0; f < L::F; f++) dst[(long long)f * cap] = (R)fl[f];
        };
      if constexpr (CPA) {
        // the fields of the next slots travel global -> shared memory as asynchronous copies (LDGSTS, per-thread commit
        // groups): nothing of a slot that is still in flight occupies a register, and the copies run PF_CPASYNC_DEPTH
        // slots ahead of the arithmetic.  A thread reads back only what it copied itself: no barrier.
        constexpr int NS = PF_CPASYNC_STAGES;
        auto prefetch = [&](int j) {
            const R *base = S + ((long long)j * L::F) * cap + src;
            R *dst = s_pre + ((j % NS) * L::F) * EXP_THREADS + threadIdx.x;
#pragma unroll
            for (int f = 0; f < L::F; f++) cp_async_elem<sizeof(R)>(dst + f * EXP_THREADS, base + (long long)f * cap);
        };
#pragma unroll
        for (int p = 0; p < PF_CPASYNC_DEPTH; p++) { if (p < K) prefetch(p); cp_async_commit(); }
        for (int j = 0; j < K; j++) {
            if (NS > 1) {
                if (j + PF_CPASYNC_DEPTH < K) prefetch(j + PF_CPASYNC_DEPTH);
                cp_async_commit();
                cp_async_wait<PF_CPASYNC_DEPTH>();              // all but the newest DEPTH groups have landed: slot j is here
            } else {
                cp_async_wait<0>();
            }
            const R *srcs = s_pre + ((j % NS) * L::F) * EXP_THREADS + threadIdx.x;
#pragma unroll
            for (int f = 0; f < L::F; f++) fcur[f] = (double)srcs[f * EXP_THREADS];
            if (NS == 1) {
                // one staging slot: it is refilled as soon as its content sits in registers (the empty asm makes the
                // shared-memory loads complete before the copies that overwrite their source are issued)
#pragma unroll
                for (int f = 0; f < L::F; f++) asm volatile("" ::"d"(fcur[f]) : "memory");
                if (j + 1 < K) prefetch(j + 1);
                cp_async_commit();
            }
            if (j == jsel) {
                double yp[D];
#pragma unroll
                for (int k = 0; k < D; k++) yp[k] = ia.y_prev[k];
                slot_add<D>(fcur, tab[(int)fcur[L::F_N]], yp, mu0);
#pragma unroll
                for (int f = 0; f < L::F; f++) fcur[f] = (double)(R)fcur[f];          // what the stored child holds (no-op for R = double)
            }
            store_child(fcur, j);
            eval(fcur, j);
        }
        cp_async_wait<0>();
      } else
#if PF_UNROLL2
      if constexpr (INST) {
        // two register sets take turns (slots j, j + 1): the fields that arrive for the next slot stay where they were
        // loaded -- no copy from "next" to "current" (14 moves per slot) and fewer values live across the slot's
        // arithmetic (spills of the fused kernel: 108 -> 24 bytes stored per thread).  Measured at 2^24: fused kernel
        // 3.84 -> 3.74 ms; the plain expansions (64 registers, 4 blocks per SM) lose 11 % with it and keep the copy.
        auto load_slot = [&](double *fl, int j) {
            const R *base = S + ((long long)j * L::F) * cap + src;
#pragma unroll
            for (int f = 0; f < L::F; f++) fl[f] = (double)base[(long long)f * cap];
        };
        auto process = [&](double *fl, int j) {
            if (j == jsel) {
                double yp[D];
#pragma unroll
                for (int k = 0; k < D; k++) yp[k] = ia.y_prev[k];
                slot_add<D>(fl, tab[(int)fl[L::F_N]], yp, mu0);
#pragma unroll
                for (int f = 0; f < L::F; f++) fl[f] = (double)(R)fl[f];          // what the stored child holds (no-op for R = double)
            }
            if (MODE == 2) store_child(fl, j);
            eval(fl, j);
        };
        for (int j = 0; j < K; j += 2) {
            if (j + 1 < K) load_slot(fnxt, j + 1);
            process(fcur, j);
            if (j + 1 < K) {
                if (j + 2 < K) load_slot(fcur, j + 2);
                process(fnxt, j + 1);
            }
        }
      } else
#endif
      {
        for (int j = 0; j < K; j++) {
            if (j + 1 < K) {
                const R *base = S + ((long long)(j + 1) * L::F) * cap + src;
#pragma unroll
                for (int f = 0; f < L::F; f++) fnxt[f] = (double)base[(long long)f * cap];
            }
            if (INST) {
                if (j == jsel) {
                    double yp[D];
#pragma unroll
                    for (int k = 0; k < D; k++) yp[k] = ia.y_prev[k];
                    slot_add<D>(fcur, tab[(int)fcur[L::F_N]], yp, mu0);
#pragma unroll
                    for (int f = 0; f < L::F; f++) fcur[f] = (double)(R)fcur[f];      // what the stored child holds (no-op for R = double)
                }
                if (MODE == 2) store_child(fcur, j);
            }
            eval(fcur, j);
#pragma unroll
            for (int f = 0; f < L::F; f++) fcur[f] = fnxt[f];
        }
      }
        int Kn = K;
        if (INST && jsel == K) {             // the previous observation opened a new cluster (src/particle.jl:141-146)
#pragma unroll
            for (int f = 0; f < L::F; f++) fcur[f] = (double)(R)ia.scp_prev->newslot[f];
            if (MODE == 2) store_child(fcur, K);
            eval(fcur, K);
            Kn = K + 1;
        }
        int c = Kn;
        if (Kn < k_max) { emit(Kn, exp(lw0 + scp->lw_new)); c = Kn + 1; }
        else myOvf = 1;
        if (MODE != 1) ((remote && dest >= 0) ? ia.peers->cnt[dest] : cnt)[di] = (unsigned char)c;
        myM = c;
        kept_hi = (unsigned int)__popcll(hi_mask);
        if (MODE == 2 && fused && rec.x) {
            const bool far = remote && dest >= 0;
            (far ? ia.peers->recx[dest] : rec.x)[di] = make_ulonglong2(B_lo.lo, B_lo.hi | ((u64)kept_hi << 56));
            (far ? ia.peers->recm[dest] : rec.m)[di] = in_mask;
            (far ? ia.peers->rech[dest] : rec.h)[di] = hi_mask;
        }
    }
    // ---- block epilogue
    vmax = warp_max64(vmax);
    myM = (long long)warp_sum64((u64)myM);
    myOvf = (long long)warp_sum64((u64)myOvf);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (fused) { B_lo = warp_sum128(B_lo); Tot = warp_sum128(Tot); kept_hi = (unsigned int)warp_sum64(kept_hi); n_plat = (unsigned int)warp_sum64(n_plat); }
    if (lane == 0) { s_max[wid] = vmax; s_M[wid] = myM; s_ovf[wid] = myOvf; if (fused) { s_red[wid] = B_lo; s_tot[wid] = Tot; s_redk[wid] = kept_hi; s_redp[wid] = n_plat; } }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int w = 1; w < nw; w++) { vmax = s_max[w] > vmax ? s_max[w] : vmax; myM += s_M[w]; myOvf += s_ovf[w]; }
        if (MODE == 1) {
            if (vmax) atomicMax(&st->vmax_sample, vmax);
        } else {
            if (vmax) atomicMax(&st->vmax_bits, vmax);
            if (myM) atomicAdd((unsigned long long *)&st->M, (unsigned long long)myM);
            if (myOvf) atomicAdd((unsigned long long *)&st->overflow, (unsigned long long)myOvf);
        }
        if (MODE != 0 && (MODE == 1 || fused)) {
            unsigned int n = s_ncand;
            if (n > CAND_CAP) { n = CAND_CAP; if (MODE == 2) sc->fuse_overflow = 1; }
            s_goff = n ? (unsigned int)atomicAdd((unsigned long long *)&sc->list_n, (unsigned long long)n) : 0u;
            PF_ASSERT((long long)s_goff + n <= (long long)cap * (k_max + 1));
            s_ncand = n;
        }
    }
    if (MODE != 0 && (MODE == 1 || fused)) {
        __shared__ u128 s_sc[8];
        __syncthreads();
        const unsigned int n = s_ncand, goff = s_goff;
        u128 Sc = mk128(0, 0);
        for (unsigned int t = threadIdx.x; t < n; t += blockDim.x) {
            const u64 b = s_cand[t];
            list[(long long)goff + t] = __longlong_as_double((long long)b);
            if (MODE == 2) Sc = add128(Sc, fx70(b, scale));
        }
        if (MODE == 2) {
            Sc = warp_sum128(Sc);
            if (lane == 0) s_sc[wid] = Sc;
            __syncthreads();
            if (threadIdx.x == 0) {
                // block totals: the per-warp B_lo / kept_hi were parked in s_red / s_redk above
                u128 B = s_red[0], C2 = s_sc[0], tot = s_tot[0];
                unsigned int kh = s_redk[0], np = s_redp[0];
                for (int w = 1; w < (int)(blockDim.x >> 5); w++) { B = add128(B, s_red[w]); C2 = add128(C2, s_sc[w]); tot = add128(tot, s_tot[w]); kh += s_redk[w]; np += s_redp[w]; }
                if (tiles) {           // (sharded INST expansion: the children's new owners form their own tile records)
                    tiles[vb].X = B; tiles[vb].kept = kh;
                    tfuse[vb].cand_off = goff; tfuse[vb].cand_cnt = n; tfuse[vb].n_plateau = np;
                }
                if (np) atomicAdd(&sc->n_plateau, (unsigned long long)np);
                if (kh) atomicAdd(&sc->A_hi, (unsigned long long)kh);
                if (n) atomicAdd(&sc->n_cand, (unsigned long long)n);
                atomic_add_acc(&sc->B_lo, B);
                atomic_add_acc(&sc->S_cand, C2);
                atomic_add_acc(&sc->Tot, tot);
            }
        }
    }
    if (!strided) break;                               // (a compile-time constant: no loop in the main launches)
    vb += gridDim.x;
    if (vb >= vb_end) break;
    __syncthreads();                                   // (the next tile re-uses the block's shared scratch)
  } while (true);
}

// ------------------------------------------------------------------ k_expand_sp
// putatives(p, y) under the other state priors (src/statepriors.jl:93-221) and for Labeled observations
// (src/labeled.jl:10-11): the same predictive arithmetic as k_expand, the candidates and their log-priors per
// prior.  Putative index jj of a particle with K components (row jj of V):
//   ChineseRestaurantProcess   jj = slot (jj = K: new)                                  candidates 1:K+1
//   StickyCRP                  K = 0: jj = 0 new; else jj = 0 "stick" (slot last), 1 + slot, K + 1 new   0:K+1
//   ChangePoint                K = 0: jj = 0 new; else jj = 0 stay (slot K-1), jj = 1 change (slot K)    max(k,1):k+1
//   NStatePrior                jj = slot                                                 1:n
//   labeled                    jj = 0 only: the labelled slot
template <int D, typename R = double>
__global__ void __launch_bounds__(256) k_expand_sp(const R *__restrict__ S, const double *__restrict__ logw,
                                                   const int *__restrict__ Kc, long long cap, int k_max,
                                                   const NRow *__restrict__ tab, const PriorConst *__restrict__ pc,
                                                   const StepConst *__restrict__ scp, StepState *st, double *__restrict__ V,
                                                   unsigned char *__restrict__ cnt, SelScratch *sc, SpArrays sp, long long step)
{
    using L = Layout<D>;
    if (sc->halt) return;
    __shared__ u64 s_max[8];
    __shared__ long long s_M[8], s_ovf[8];
    const long long n_live = st->n_live;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    u64 vmax = 0;
    long long myM = 0, myOvf = 0;
    bool bad_label = false;
    if (i < n_live) {
        double y[D], mu0[D];
#pragma unroll
        for (int k = 0; k < D; k++) { y[k] = scp->y[k]; mu0[k] = pc->mu0[k]; }
        const int K = Kc[i];
        const double lw0 = logw[i];
        const int kind = pc->sp_kind, forced = scp->forced;
        const double lp_newc = kind == SPK_STICKY ? pc->sp_log_alpha : 0.0;      // (CRP: log alpha is inside lw_new)
        double n_of = 0.0;
        auto pred = [&](int j) {       // log predictive of y under slot j (+ log N_j under the CRP: folded into the table)
            const R *base = S + ((long long)j * L::F) * cap + i;
            double fl[L::F], z[D], q;
#pragma unroll
            for (int f = 0; f < L::F; f++) fl[f] = (double)base[(long long)f * cap];
            n_of = fl[L::F_N];
            const NRow row = tab[(int)fl[L::F_N]];
            return slot_logw<D>(row, fl[L::F_LOGDET], fl + L::F_MEAN, fl + L::F_DINV, fl + L::F_OFF, y, mu0, z, &q);
        };
        auto put = [&](int jj, double lw) {
            const double v = exp(lw);
            const u64 b = (u64)__double_as_longlong(v);
            vmax = b > vmax ? b : vmax;
            V[(long long)jj * cap + i] = v;
        };
        auto lp_slot = [&](int j, double nj) {     // log-prior of the non-sticky candidate "existing slot j"
            if (kind == SPK_STICKY) return log(sp.N[(long long)j * cap + i]);
            if (kind == SPK_CHANGEPOINT) return pc->sp_lp_stay;
            if (kind == SPK_NSTATE) return log(pc->sp_counts0[j] + nj) - log(pc->sp_total0 + (double)step);
            return 0.0;                              // CRP: in the table
        };
        int c = 0;
        if (forced >= 0) {
            // fit(p, y, label): one putative.  Valid states: CRP / StickyCRP 1..K+1, ChangePoint k or k+1, NStatePrior 1..n
            const int j = forced;
            bool ok = kind == SPK_NSTATE ? j < pc->sp_n : (kind == SPK_CHANGEPOINT ? (K == 0 ? j == 0 : (j == K - 1 || j == K)) : j <= K);
            if (j >= k_max) ok = false;
            if (!ok) { bad_label = true; put(0, -INFINITY); }
            else if (j < K) { const double p = pred(j); put(0, lw0 + lp_slot(j, n_of) + p); }
            else put(0, lw0 + (kind == SPK_CHANGEPOINT ? (K == 0 ? 0.0 : pc->sp_lp_change) : lp_newc) + scp->lw_new);
            c = 1;
        } else if (kind == SPK_CHANGEPOINT) {
            if (K == 0) { put(0, lw0 + scp->lw_new); c = 1; }
            else {
                const double p = pred(K - 1);
                put(0, lw0 + pc->sp_lp_stay + p); c = 1;
                if (K < k_max) { put(1, lw0 + pc->sp_lp_change + scp->lw_new); c = 2; } else myOvf = 1;
            }
        } else if (kind == SPK_NSTATE) {
            for (int j = 0; j < K; j++) { const double p = pred(j); put(j, lw0 + lp_slot(j, n_of) + p); }
            c = K;
        } else if (kind == SPK_STICKY) {
            if (K == 0) { put(0, lw0 + lp_newc + scp->lw_new); c = 1; }
            else {
                const int last = sp.last[i];
                double sumN = 0.0, p_last = 0.0;
                for (int j = 0; j < K; j++) {
                    const double Nj = sp.N[(long long)j * cap + i];
                    sumN += Nj;
                    const double p = pred(j);
                    if (j == last) p_last = p;
                    put(1 + j, lw0 + log(Nj) + p);
                }
                put(0, lw0 + pc->sp_logit_rho + log(sumN + pc->alpha) + p_last);
                c = K + 1;
                if (K < k_max) { put(K + 1, lw0 + lp_newc + scp->lw_new); c = K + 2; } else myOvf = 1;
            }
        } else {                                      // CRP (labeled streams with missing labels)
            for (int j = 0; j < K; j++) { const double p = pred(j); put(j, lw0 + p); }
            c = K;
            if (K < k_max) { put(K, lw0 + scp->lw_new); c = K + 1; } else myOvf = 1;
        }
        cnt[i] = (unsigned char)c;
        myM = c;
    }
    if (bad_label) sel_halt(sc, step, PF_HALT_LABEL);
    vmax = warp_max64(vmax);
    myM = (long long)warp_sum64((u64)myM);
    myOvf = (long long)warp_sum64((u64)myOvf);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s_max[wid] = vmax; s_M[wid] = myM; s_ovf[wid] = myOvf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); w++) { vmax = s_max[w] > vmax ? s_max[w] : vmax; myM += s_M[w]; myOvf += s_ovf[w]; }
        if (vmax) atomicMax(&st->vmax_bits, vmax);
        if (myM) atomicAdd((unsigned long long *)&st->M, (unsigned long long)myM);
        if (myOvf) atomicAdd((unsigned long long *)&st->overflow, (unsigned long long)myOvf);
    }
}

// Survivor scan, level 1: one block owns a "supertile" of SUPER_TILES consecutive tile records.  On the fused
// path each record first receives its tile's in-bracket candidates (their side is known now that the
// threshold is); then the block turns its records into exclusive prefixes WITHIN the supertile and
// leaves the supertile's total in supers[] (level 2, k_tile_prefix on supers[], scans those).
#define SUPER_TILES 256
// ------------------------------------------------------------------ survivor scan
// The systematic comb (sample_stratified!, src/fearnhead.jl:97-114) over the putatives in
// (particle, slot) order, in exact integer arithmetic.  With X = n * (running fixed-point
// mass of the low partition) the comb teeth sit at Uoff + k T; a low item receives as many
// copies as teeth fall strictly below its running X but not below its predecessor's (one
// at most, because every low weight is < c = T / n), so the number of resampled survivors
// before an item is teeth_below(X_before) and the output position of every survivor is
//     (#kept before it) + (#teeth before it)
// -- one exclusive scan of (X, kept), no second dependent scan.  Three launches, no
// spinning: per-tile sums, one-block prefix over the tiles, apply.


// one thread: wait for every rank's survivor-scan total (exchange kind C), turn the totals (mass of the
// low partition, kept count) into this rank's offsets and the new ownership
__global__ void k_rank_offsets(const Xch *x, unsigned long long seq, long long step, const SelResult *__restrict__ res,
                               StepState *st, RankOff *ro, SelScratch *guard, long long surv_cap)
{
    if (threadIdx.x || blockIdx.x) return;
    if (guard->halt) { ro->n_local = 0; return; }
    if (guard->phase != 4) { sel_halt(guard, step, PF_HALT_SEARCH); ro->n_local = 0; return; }
    if (!xch_wait(x, XCH_C, seq)) { sel_halt(guard, step, PF_HALT_TIMEOUT); ro->n_local = 0; return; }
    const XchTot *gathered = x->reg[x->rank]->c;
    const int rank = x->rank, nranks = x->nranks;
    const bool comb = res->n > 0 && !isz128(res->T);
    const u64 n = (u64)res->n;
    u128 Spre = mk128(0, 0);
    long long kpre = 0, gpre = 0, tpre = 0;
    long long my_kept = 0, my_picked = 0, my_first = 0, my_n = 0;
    for (int r = 0; r < nranks; r++) {
        u128 X;
        X.lo = ld_peer(reinterpret_cast<const unsigned long long *>(&gathered[r].X.lo));
        X.hi = ld_peer(reinterpret_cast<const unsigned long long *>(&gathered[r].X.hi));
        const long long kept = (long long)(unsigned int)ld_peer(reinterpret_cast<const unsigned long long *>(&gathered[r].kept));
        u128 Snext = add128(Spre, X);
        long long tnext = comb ? (long long)teeth_below(mul128_64(Snext, n), res->Uoff, res->T) : 0;
        long long picked = tnext - tpre;
        if (r == rank) { ro->S_off = Spre; ro->kept_off = kpre; my_kept = kept; my_picked = picked; my_first = gpre; my_n = kept + picked; }
        Spre = Snext; kpre += kept; tpre = tnext; gpre += kept + picked;
    }
    const long long n_global = gpre;
    long long q = (n_global + nranks - 1) / nranks;
    if (q < 1) q = 1;
    const long long b0 = (long long)rank * q < n_global ? (long long)rank * q : n_global;
    const long long b1 = (long long)(rank + 1) * q < n_global ? (long long)(rank + 1) * q : n_global;
    ro->g_first = my_first; ro->n_local = my_n; ro->n_global = n_global; ro->q = q;
    ro->n_mine = b1 - b0; ro->goff_next = b0;
    ro->rank = rank; ro->nranks = nranks; ro->goff_prev = st->goff;
    st->n_kept = my_kept; st->n_picked = my_picked; st->n_next = ro->n_mine; st->n_global = n_global;
    if (my_n > surv_cap) { sel_halt(guard, step, PF_HALT_SURVIVORS); ro->n_local = 0; }
}

// pipelined sharded steps: "this rank has delivered every child of the generation" (exchange kind D) is raised after the
// fused instantiate + expansion, and awaited before the next step reads what the peers delivered
__global__ void k_mg_signal_generation(const Xch *x, unsigned long long seq, const SelScratch *guard)
{
    if (threadIdx.x || blockIdx.x || guard->halt) return;
    xch_signal(x, XCH_D, seq);
}
__global__ void k_mg_await_generation(const Xch *x, unsigned long long seq, long long step, SelScratch *guard)
{
    if (threadIdx.x || blockIdx.x || guard->halt) return;
    if (!xch_wait(x, XCH_D, seq)) sel_halt(guard, step, PF_HALT_TIMEOUT);
}

// mode 0: index-order step (always runs); mode 1: flat sorted list, runs only when the step
// resamples and emits the PICKED positions only (kept = suffix of the sorted list);
// mode 2: index-order pass of a sorted-order filter, runs only when the step keeps all.
__device__ __forceinline__ bool scan_mode_skips(int mode, const SelResult *res)
{
    return (mode == 1 && res->keep_all) || (mode == 2 && !res->keep_all);
}

__device__ __forceinline__ void thread_sums(const double *__restrict__ V, long long cap, long long i, int c, u64 thr, int scale,
                                            bool comb, u128 &S, unsigned int &kept)
{
    S = mk128(0, 0); kept = 0;
    for (int j = 0; j < c; j++) {
        u64 bits = (u64)__double_as_longlong(V[(long long)j * cap + i]);
        if (bits >= thr) kept++;
        else if (comb) S = add128(S, fx70(bits, scale));
    }
}

__global__ void __launch_bounds__(SUPER_TILES) k_tile_scan1(TileSum *__restrict__ tiles, const TileSum *__restrict__ subs, const TileFuse *__restrict__ tfuse,
                                                           const double *__restrict__ list, const SelResult *__restrict__ res,
                                                           const SelScratch *__restrict__ sc, const long long *__restrict__ n_items_ptr,
                                                           const StepConst *__restrict__ scp, TileSum *__restrict__ supers, int mode,
                                                           const SelScratch *guard)
{
    if (guard && (guard->phase != 4 || guard->halt)) return;
    if (scan_mode_skips(mode, res)) return;
    __shared__ u128 s_warp[33];
    __shared__ unsigned int s_wc[33];
    const long long ntiles = (*n_items_ptr + SCAN_THREADS - 1) / SCAN_THREADS;
    if ((long long)blockIdx.x * SUPER_TILES >= ntiles) return;
    const long long t = (long long)blockIdx.x * SUPER_TILES + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool comb = res->n > 0 && !isz128(res->T);
    u128 X = mk128(0, 0);
    unsigned int kept = 0;
    if (t < ntiles) {
        if (tfuse && sc && sc->tiles_ok) {
            // the fused expansion left EXP_SUB sub-tile records holding the putatives outside the bracket: add them up
            // and add each one's candidates
            const u64 thr = res->thr_bits;
            const int scale = res->scale;
            for (int sb = 0; sb < EXP_SUB; sb++) {
                const long long r = t * EXP_SUB + sb;
                if (r * EXP_THREADS >= *n_items_ptr) break;
                X = add128(X, subs[r].X); kept += subs[r].kept;
                const unsigned int off = tfuse[r].cand_off, n = tfuse[r].cand_cnt;
                PF_ASSERT((long long)off + n <= sc->list_n);
                for (unsigned int k = 0; k < n; k++) {
                    const u64 b = (u64)__double_as_longlong(list[(long long)off + k]);
                    if (b >= thr) kept++;
                    else if (comb) X = add128(X, fx70(b, scale));
                }
                const unsigned int np = tfuse[r].n_plateau;
                if (np) {
                    if (scp->plateau_bits >= thr) kept += np;
                    else if (comb) X = add128(X, mul128_64(fx70(scp->plateau_bits, scale), (u64)np));
                }
            }
            if (!comb) X = mk128(0, 0);
        } else {
            X = tiles[t].X; kept = tiles[t].kept;
        }
    }
    u128 totX;
    const u128 exX = block_excl_scan128(X, s_warp, &totX);
    unsigned int incK = kept;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { unsigned int o = __shfl_up_sync(0xffffffffu, incK, d); if (lane >= d) incK += o; }
    if (lane == 31) s_wc[wid] = incK;
    __syncthreads();
    if (wid == 0) {
        unsigned int w = lane < (SUPER_TILES / 32) ? s_wc[lane] : 0, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { unsigned int o = __shfl_up_sync(0xffffffffu, wi, d); if (lane >= d) wi += o; }
        s_wc[lane] = wi - w;
        if (lane == 31) s_wc[32] = wi;
    }
    __syncthreads();
    if (t < ntiles) { tiles[t].X = exX; tiles[t].kept = s_wc[wid] + incK - kept; }
    if (threadIdx.x == 0) { supers[blockIdx.x].X = totX; supers[blockIdx.x].kept = s_wc[32]; }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_tile_sums(const double *__restrict__ V, const unsigned char *__restrict__ cnt,
                                                            long long cap, const SelResult *__restrict__ res,
                                                            const long long *__restrict__ n_items_ptr, TileSum *__restrict__ tiles,
                                                            int mode, const SelScratch *guard, const SelScratch *fusedsc,
                                                            ScanRec rec = ScanRec{nullptr, nullptr, nullptr}, const SelScratch *recsc = nullptr)
{
    if (guard && (guard->phase != 4 || guard->halt)) return;
    if (fusedsc && fusedsc->tiles_ok) return;              // the fused expansion already wrote the tile records
    // (pipelined sharded steps: the particles were expanded by their parents' owners, which delivered the per-particle
    // scan records but could not form this shard's tile records: sum the records instead of the putatives)
    const bool use_rec = rec.x && recsc && recsc->tiles_ok;
    if (scan_mode_skips(mode, res)) return;
    __shared__ u128 s_x[SCAN_THREADS / 32];
    __shared__ unsigned int s_k[SCAN_THREADS / 32];
    const long long n_items = *n_items_ptr;
    const long long ntiles = (n_items + SCAN_THREADS - 1) / SCAN_THREADS;
    const u64 n = (u64)res->n;
    const bool comb = n > 0 && !isz128(res->T);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {         // (a small grid: the kernel is a no-op on the fused path)
        const long long i = tile * SCAN_THREADS + threadIdx.x;
        const int c = (i < n_items) ? (cnt ? cnt[i] : 1) : 0;
        u128 X; unsigned int kept;
        if (use_rec) {
            X = mk128(0, 0); kept = 0;
            if (c) {
                const ulonglong2 x = rec.x[i];
                kept = (unsigned int)(x.y >> 56);
                if (comb) X = mk128(x.x, x.y & 0x00FFFFFFFFFFFFFFull);
                unsigned long long m = rec.m[i];
                while (m) {
                    const int j = __ffsll((long long)m) - 1; m &= m - 1;
                    const u64 bits = (u64)__double_as_longlong(V[(long long)j * cap + i]);
                    if (bits >= res->thr_bits) kept++;
                    else if (comb) X = add128(X, fx70(bits, res->scale));
                }
            }
        } else {
            thread_sums(V, cap, i, c, res->thr_bits, res->scale, comb, X, kept);
        }
        X = warp_sum128(X);
        kept = (unsigned int)warp_sum64(kept);
        if (lane == 0) { s_x[wid] = X; s_k[wid] = kept; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < SCAN_THREADS / 32; w++) { X = add128(X, s_x[w]); kept += s_k[w]; }
            tiles[tile].X = X; tiles[tile].kept = kept;
        }
        __syncthreads();
    }
}

// exclusive prefix over the tiles (in place) and the step's survivor counts; one block of 32
// warps, warp w owns a contiguous run of tiles and walks it 32 tiles at a time (coalesced
// 1 KB loads, shuffle scan, carry), then the warps' totals are scanned
__global__ void __launch_bounds__(1024) k_tile_prefix(TileSum *tiles, const SelResult *__restrict__ res,
                                                      const long long *__restrict__ n_items_ptr, StepState *st, int mode,
                                                      const Xch *xch /* sharded run: the total goes to every rank (exchange kind C) */,
                                                      unsigned long long seq, const SelScratch *guard, int super_level)
{
    if (guard && (guard->phase != 4 || guard->halt)) return;
    if (scan_mode_skips(mode, res)) return;
    __shared__ u128 s_wx[32];
    __shared__ unsigned int s_wk[32];
    const long long n_items = *n_items_ptr;
    long long ntiles = (n_items + SCAN_THREADS - 1) / SCAN_THREADS;
    if (super_level) ntiles = (ntiles + SUPER_TILES - 1) / SUPER_TILES;           // `tiles` holds supertile totals
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long per = ((ntiles + 31) / 32 + 31) / 32 * 32;          // tiles per warp, multiple of 32
    const long long b = (long long)wid * per, e = b + per < ntiles ? b + per : ntiles;
    // pass 1: warp totals
    u128 wx = mk128(0, 0); unsigned int wk = 0;
    for (long long t0 = b; t0 < e; t0 += 32) {
        const long long t = t0 + lane;
        if (t < e) { wx = add128(wx, tiles[t].X); wk += tiles[t].kept; }
    }
    wx = warp_sum128(wx); wk = (unsigned int)warp_sum64(wk);
    if (lane == 0) { s_wx[wid] = wx; s_wk[wid] = wk; }
    __syncthreads();
    if (wid == 0) {
        u128 x = s_wx[lane], ix = x; unsigned int k = s_wk[lane], ik = k;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u128 ox = shfl_up128(ix, d); unsigned int ok = __shfl_up_sync(0xffffffffu, ik, d);
            if (lane >= d) { ix = add128(ix, ox); ik += ok; }
        }
        s_wx[lane] = sub128(ix, x); s_wk[lane] = ik - k;                 // exclusive warp offsets
        if (lane == 31) {
            if (xch) {
                XchTot tot; tot.X = ix; tot.kept = ik; tot.pad[0] = tot.pad[1] = tot.pad[2] = 0;
                for (int p = 0; p < xch->nranks; p++) xch->reg[p]->c[xch->rank] = tot;
                xch_signal(xch, XCH_C, seq);
            } else {
                const bool comb = res->n > 0 && !isz128(res->T);
                long long picked = comb ? (long long)teeth_below(mul128_64(ix, (u64)res->n), res->Uoff, res->T) : 0;
                st->n_kept = ik; st->n_picked = picked; st->n_next = (long long)ik + picked; st->n_global = st->n_next;
            }
        }
    }
    __syncthreads();
    // pass 2: exclusive prefixes in place
    u128 runX = s_wx[wid]; unsigned int runK = s_wk[wid];
    for (long long t0 = b; t0 < e; t0 += 32) {
        const long long t = t0 + lane;
        u128 x = mk128(0, 0); unsigned int k = 0;
        if (t < e) { x = tiles[t].X; k = tiles[t].kept; }
        u128 ix = x; unsigned int ik = k;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            u128 ox = shfl_up128(ix, d); unsigned int ok = __shfl_up_sync(0xffffffffu, ik, d);
            if (lane >= d) { ix = add128(ix, ox); ik += ok; }
        }
        if (t < e) { tiles[t].X = add128(runX, sub128(ix, x)); tiles[t].kept = runK + ik - k; }
        runX = add128(runX, shfl_xor_bcast128(ix, 31)); runK += __shfl_sync(0xffffffffu, ik, 31);
    }
}

// Every tile's starting point of the comb, one THREAD per tile: running mass and kept count before the tile, and its first
// tooth (the one 128-bit division of tooth_base), so that thread 0 of a k_scan_apply block does not run that serial
// chain of ~600 instructions while the other 255 wait at the scan's barrier.
struct __align__(16) TileTooth {
    u128 S0, Q0;
    long long k0;
    u64 t0;
    double inv;
    unsigned int R0, pad;
};
__global__ void __launch_bounds__(256) k_tooth_bases(const TileSum *__restrict__ tiles, const TileSum *__restrict__ supers,
                                                     const SelResult *__restrict__ res, const long long *__restrict__ n_items_ptr,
                                                     const RankOff *__restrict__ ro, TileTooth *__restrict__ out, int mode,
                                                     const SelScratch *guard)
{
    if (guard && (guard->phase != 4 || guard->halt)) return;
    if (scan_mode_skips(mode, res)) return;
    const long long ntiles = (*n_items_ptr + SCAN_THREADS - 1) / SCAN_THREADS;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const u64 n = (u64)res->n;
    const u128 T = res->T;
    TileTooth tt;
    tt.S0 = add128(tiles[t].X, supers[t / SUPER_TILES].X);          // prefix within the supertile + prefix of the supertiles
    tt.k0 = (long long)tiles[t].kept + (long long)supers[t / SUPER_TILES].kept;
    if (ro) { tt.S0 = add128(tt.S0, ro->S_off); tt.k0 += ro->kept_off - ro->g_first; }   // positions relative to this rank's first survivor
    tt.Q0 = mk128(0, 0); tt.t0 = 0; tt.inv = 0.0; tt.R0 = 0; tt.pad = 0;
    if (n > 0 && !isz128(T)) {
        const ToothBase b = tooth_base(tt.S0, n, res->Uoff, T);
        tt.Q0 = b.Q0; tt.R0 = b.R0; tt.t0 = b.t0; tt.inv = b.inv;
    }
    out[t] = tt;
}

// The walk over one particle's putatives (one lane per particle; all 32 lanes of the warp call it, idle lanes with
// c = 0): kept ones are listed; for the low ones the question "has the running mass passed the next tooth?" is
// Q - S_particle < (sum of the q_j so far).  It is asked in double precision first -- the items' values scaled by
// 2^scale, summed with DADD, against D = (double)(Q - S_particle) -- and the answer is taken only when it clears the
// rounding error by a wide margin (the integer q_j = floor(v_j 2^scale) differ from the scaled doubles by < 1 each, the
// doubles by 2^-53 relative); a lane that comes closer than that (or meets a subnormal weight) redoes its particle in
// exact 128-bit arithmetic.  Decisions are those of the exact walk; the common case costs one DADD per putative instead
// of a 128-bit conversion, add and compare.  The putatives are fetched eight at a time (independent loads in flight).
__device__ __forceinline__ void walk_particle(const double *__restrict__ V, long long cap, long long i, int c, u64 thr, int scale,
                                              int mode, u128 S, long long kcount, u64 t, u128 Q, unsigned int R, bool tooth_inside,
                                              u128 Tq, unsigned int Tr, unsigned int n32, unsigned int *__restrict__ surv_parent,
                                              unsigned char *__restrict__ surv_slot, long long surv_cap)
{
    const long long kcount0 = kcount;
    const u64 t_first = t;
    const u128 Q_first = Q; const unsigned int R_first = R;
    bool amb = false;
    {
        double cum = 0.0;
        // distance to the next tooth; a lane without a tooth inside its particle gets one it cannot reach (finite, so that
        // the comparisons below stay ordinary: no per-slot test of `tooth_inside`)
        double Dd = tooth_inside ? to_double128(sub128(Q, S)) : 1e300;
        // below Dlo the running mass has certainly not reached the tooth and is not within the rounding margin of it
        // (margin = 64 + (cum + Dd) 2^-45 < 64 + Dd 2^-44 while cum < Dd): one comparison per putative in the common case
        double Dlo = walk_reject_below(Dd);
        const long long sadd = (long long)scale << 52;
        int cmax = c;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cmax = max(cmax, __shfl_xor_sync(0xffffffffu, cmax, d));
        const double *__restrict__ pv = V + i;
        for (int j0 = 0; j0 < cmax; j0 += 8) {
            u64 vb[8];
#pragma unroll
            for (int k = 0; k < 8; k++) { vb[k] = (j0 + k < c) ? (u64)__double_as_longlong(*pv) : 0ull; pv += cap; }
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int j = j0 + k;
                const u64 bits = vb[k];
                if (j >= c || amb) continue;
                if (bits >= thr) {
                    if (mode != 1) {
                        long long pos = kcount + (long long)t;
                        PF_ASSERT(pos >= 0 && pos < surv_cap);
                        if (pos >= 0 && pos < surv_cap) { surv_parent[pos] = (unsigned int)i; surv_slot[pos] = (unsigned char)j; }
                    }
                    kcount++;
                } else if (bits != 0) {
                    const int eb = (int)(bits >> 52) + scale;                       // exponent field of v 2^scale
                    if ((bits >> 52) == 0 || (unsigned int)(eb - 1) >= 2045u) { amb = tooth_inside; continue; }      // subnormal, or v 2^scale not a normal double (matters to a lane with a tooth only)
                    cum += __longlong_as_double((long long)bits + sadd);              // v_j 2^scale, exact
                    if (cum < Dlo) continue;
                    for (;;) {
                        const int passed = walk_compare(cum, Dd);
                        if (passed > 0) {
                            long long pos = (mode == 1 ? 0 : kcount) + (long long)t;
                            PF_ASSERT(pos >= 0 && pos < surv_cap);
                            if (pos >= 0 && pos < surv_cap) { surv_parent[pos] = (unsigned int)i; surv_slot[pos] = (unsigned char)(j | 0x80); }
                            t++;
                            tooth_step(Q, R, Tq, Tr, n32);
                            Dd = to_double128(sub128(Q, S));
                            Dlo = walk_reject_below(Dd);
                            continue;
                        }
                        if (passed == 0) amb = true;
                        break;
                    }
                }
            }
        }
    }
    if (amb) {
        // exact walk of this particle from its start (rewrites the survivors the fast walk listed)
        kcount = kcount0; t = t_first; Q = Q_first; R = R_first;
        for (int j = 0; j < c; j++) {
            u64 bits = (u64)__double_as_longlong(V[(long long)j * cap + i]);
            if (bits >= thr) {
                if (mode != 1) {
                    long long pos = kcount + (long long)t;
                    if (pos >= 0 && pos < surv_cap) { surv_parent[pos] = (unsigned int)i; surv_slot[pos] = (unsigned char)j; }
                }
                kcount++;
            } else {
                S = add128(S, fx70(bits, scale));
                while (lt128(Q, S)) {
                    long long pos = (mode == 1 ? 0 : kcount) + (long long)t;
                    if (pos >= 0 && pos < surv_cap) { surv_parent[pos] = (unsigned int)i; surv_slot[pos] = (unsigned char)(j | 0x80); }
                    t++;
                    tooth_step(Q, R, Tq, Tr, n32);
                }
            }
        }
    }
}

// a particle that has a comb tooth inside its low mass, parked for the compacted walk
struct __align__(16) WalkItem {
    u128 S, Q;
    long long kcount;
    u64 t;
    unsigned int R;
    int tid, c, pad;
};

#ifndef PF_SCAN_MINBLOCKS
#define PF_SCAN_MINBLOCKS 4       // (64 registers; measured at 2^24: 4 -> 1.05 ms, 5 -> 1.06, 6 -> 1.19, unconstrained 1 -> 1.81)
#endif
__global__ void __launch_bounds__(SCAN_THREADS, PF_SCAN_MINBLOCKS) k_scan_apply(const double *__restrict__ V, const unsigned char *__restrict__ cnt,
                                                             long long cap, const SelResult *__restrict__ res,
                                                             const long long *__restrict__ n_items_ptr,
                                                             const TileSum *__restrict__ tiles, const TileSum *__restrict__ supers,
                                                             unsigned int *__restrict__ surv_parent,
                                                             unsigned char *__restrict__ surv_slot, int mode,
                                                             const RankOff *__restrict__ ro, const SelScratch *guard,
                                                             long long surv_cap, ScanRec rec = ScanRec{nullptr, nullptr, nullptr},
                                                             const SelScratch *recsc = nullptr, const TileTooth *__restrict__ tooth = nullptr)
{
    if (guard && (guard->phase != 4 || guard->halt)) return;
    if (scan_mode_skips(mode, res)) return;
    __shared__ u128 s_warp[33];
    const long long n_items = *n_items_ptr;
    const long long ntiles = (n_items + SCAN_THREADS - 1) / SCAN_THREADS;
    if (blockIdx.x >= ntiles) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const long long i = (long long)blockIdx.x * SCAN_THREADS + tid;
    const u64 thr = res->thr_bits;
    const int scale = res->scale;
    const u64 n = (u64)res->n;
    const u128 T = res->T, U = res->Uoff;
    const bool comb = n > 0 && !isz128(T);
    const int c = (i < n_items) ? (cnt ? cnt[i] : 1) : 0;
    __shared__ u128 s_S0, s_Q0;
    __shared__ unsigned int s_R0;
    __shared__ u64 s_t0;
    __shared__ long long s_k0;
    __shared__ double s_inv;
    __shared__ unsigned int s_nq;
    if (tid == 0 && tooth) {
        const TileTooth tt = tooth[blockIdx.x];                  // (k_tooth_bases)
        s_S0 = tt.S0; s_k0 = tt.k0; s_nq = 0;
        s_Q0 = tt.Q0; s_R0 = tt.R0; s_t0 = tt.t0; s_inv = tt.inv;
    } else if (tid == 0) {
        u128 S0 = add128(tiles[blockIdx.x].X, supers[blockIdx.x / SUPER_TILES].X);      // prefix within the supertile + prefix of the supertiles
        long long k0 = (long long)tiles[blockIdx.x].kept + (long long)supers[blockIdx.x / SUPER_TILES].kept;
        if (ro) { S0 = add128(S0, ro->S_off); k0 += ro->kept_off - ro->g_first; }   // positions relative to this rank's first survivor
        s_S0 = S0; s_k0 = k0; s_nq = 0;
        if (comb) {
            const ToothBase b = tooth_base(S0, n, U, T);
            s_Q0 = b.Q0; s_R0 = b.R0; s_t0 = b.t0; s_inv = b.inv;
        }
    }
    // the fused expansion left, per particle: mass below the bracket, the slots inside the bracket (the only ones whose
    // side of the threshold was open: ~0.5 % of the putatives) and the slots above it
    const bool use_rec = rec.x && recsc->tiles_ok;
    u128 myX; unsigned int myKept;
    unsigned long long in_mask = 0, hi_mask = 0;
    if (use_rec) {
        myX = mk128(0, 0); myKept = 0;
        if (c) {
            const ulonglong2 x = rec.x[i];
            in_mask = rec.m[i]; hi_mask = rec.h[i];
            myKept = (unsigned int)(x.y >> 56);
            if (comb) myX = mk128(x.x, x.y & 0x00FFFFFFFFFFFFFFull);
            unsigned long long m = in_mask, kept_in = 0;
            while (m) {
                const int j = __ffsll((long long)m) - 1; m &= m - 1;
                const u64 bits = (u64)__double_as_longlong(V[(long long)j * cap + i]);
                if (bits >= thr) { myKept++; kept_in |= 1ull << j; }
                else if (comb) myX = add128(myX, fx70(bits, scale));
            }
            hi_mask |= kept_in;                    // from here: the particle's kept slots
        }
    } else {
        thread_sums(V, cap, i, c, thr, scale, comb, myX, myKept);
    }
    // ONE exclusive scan over the block for both quantities: the kept count rides in bits 104.. of the 128-bit mass (a
    // tile's mass is < 2^86, its kept count < 2^15), and with 8 warps every warp adds up the totals of the warps before
    // it by itself -- one block barrier instead of the five of two separate two-level scans
    PF_ASSERT(myX.hi < (1ull << 22) && myKept <= PF_MAX_SLOTS + 1);
    u128 pk = mk128(myX.lo, myX.hi | ((u64)myKept << 40));
    u128 inc = pk;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { u128 o = shfl_up128(inc, d); if (lane >= d) inc = add128(inc, o); }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    u128 ex = sub128(inc, pk);
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32 - 1; w++) if (w < wid) ex = add128(ex, s_warp[w]);
    const u128 exX = mk128(ex.lo, ex.hi & ((1ull << 40) - 1));
    const unsigned int exK = (unsigned int)(ex.hi >> 40);
    // running low-partition mass S (fixed point) and the next tooth k: tooth k is passed by S
    // <=> Uoff + k T < n S <=> floor((Uoff + k T) / n) < S; (Q, R) = divmod(Uoff + k T, n) steps
    // by (Tq, Tr) = divmod(T, n) from tooth to tooth.  The tile's first tooth (t0, Q0, R0) is
    // found once per block (the one 128-bit division); a thread reaches its own tooth
    // t0 + k with Q = Q0 + k Tq + floor((R0 + k Tr) / n), k from an FP estimate corrected exactly.
    const u128 S = add128(s_S0, exX);
    const long long kcount = s_k0 + exK;
    u64 t = 0;
    u128 Q = mk128(0, 0);
    unsigned int R = 0;
    const unsigned int n32 = (unsigned int)n, Tr = res->Tr;
    const u128 Tq = res->Tq;
    bool tooth_inside = false;
    if (comb) {
        t = s_t0 + tooth_seek(s_Q0, s_R0, s_inv, S, Tq, Tr, n32, Q, R);
        tooth_inside = lt128(Q, add128(S, myX));
    }
    if (!use_rec) {
        if (!__any_sync(0xffffffffu, tooth_inside) && (mode == 1 || !__any_sync(0xffffffffu, myKept != 0))) return;      // nothing to list in this warp
        walk_particle(V, cap, i, c, thr, scale, mode, S, kcount, t, Q, R, tooth_inside, Tq, Tr, n32, surv_parent, surv_slot, surv_cap);
        return;
    }
    // With the records: a particle without a tooth only lists its kept slots (known as a mask: no pass over its
    // putatives); the particles with a tooth (about a third) are parked and walked by densely packed warps, so a warp
    // does not run the full walk for a few of its lanes.
    __shared__ WalkItem s_q[SCAN_THREADS];
    if (tooth_inside) {
        const unsigned int slot = atomicAdd(&s_nq, 1u);
        PF_ASSERT(slot < SCAN_THREADS);
        WalkItem it;
        it.S = S; it.Q = Q; it.kcount = kcount; it.t = t; it.R = R; it.tid = tid; it.c = c; it.pad = 0;
        s_q[slot] = it;
    } else {
        unsigned long long m = hi_mask;
        long long pos = kcount + (long long)t;
        while (m) {
            const int j = __ffsll((long long)m) - 1; m &= m - 1;
            PF_ASSERT(pos >= 0 && pos < surv_cap && j < c);
            if (pos >= 0 && pos < surv_cap) { surv_parent[pos] = (unsigned int)i; surv_slot[pos] = (unsigned char)j; }
            pos++;
        }
    }
    __syncthreads();
    const unsigned int nq = s_nq;
    for (unsigned int base = (unsigned int)wid * 32; base < nq; base += SCAN_THREADS) {
        const unsigned int q = base + lane;
        WalkItem it;
        it.c = 0; it.tid = 0; it.S = mk128(0, 0); it.Q = mk128(0, 0); it.kcount = 0; it.t = 0; it.R = 0;
        if (q < nq) it = s_q[q];
        walk_particle(V, cap, (long long)blockIdx.x * SCAN_THREADS + it.tid, it.c, thr, scale, mode, it.S, it.kcount, it.t, it.Q, it.R, q < nq,
                      Tq, Tr, n32, surv_parent, surv_slot, surv_cap);
    }
}

// ------------------------------------------------------------------ k_instantiate
// instantiate.(survivors) (src/particle.jl:134-149): thread t builds particle t of the new
// population from its parent's slots (copy), applies add(comp, y) to the chosen slot as a
// Welford mean update + rank-1 Cholesky update, appends (ancestor, assignment) to the
// genealogy and sets the normalised log-weight (src/fearnhead.jl:172-179).
template <int D, typename R>
__device__ __forceinline__ void instantiate_one(const long long t, const R *__restrict__ S, const int *__restrict__ Kc,
                                                     R *S2, int *Kc2,
                                                     double *logw2, long long cap,
                                                     const NRow *__restrict__ tab, const PriorConst *__restrict__ pc,
                                                     const StepConst *__restrict__ scp, const StepState *__restrict__ st,
                                                     const SelResult *__restrict__ res, const double *__restrict__ V,
                                                     const unsigned int *__restrict__ surv_parent,
                                                     const unsigned char *__restrict__ surv_slot,
                                                     unsigned int *gen_anc, unsigned char *gen_asg,
                                                     const RankOff *__restrict__ ro, int k_max,
                                                     const SelScratch *guard, long long *m_next,
                                                     const PeerTable *__restrict__ peers, long long gen_off, SpArrays sp_cur, SpArrays sp_nxt)
{
    using L = Layout<D>;
    // destination of child t: a local index, or (sharded run) an index in the arrays of its new owner
    R *dstb = S2 + t;
    long long idx = t;
    unsigned int anc_off = 0;
    if (ro) {
        const long long g = ro->g_first + t;
        const int dest = (int)(g / ro->q);
        PF_ASSERT(dest >= 0 && dest < ro->nranks);
        anc_off = (unsigned int)ro->goff_prev;
        idx = g - (long long)dest * ro->q;
        if (dest != ro->rank) {
            // the child is written into its new owner's arrays (peer memory over NVLink)
            S2 = static_cast<R *>(peers->S[dest]); Kc2 = peers->Kc[dest]; logw2 = peers->logw[dest];
            gen_anc = peers->gen_anc[dest] ? peers->gen_anc[dest] + gen_off : nullptr;
            gen_asg = peers->gen_asg[dest] ? peers->gen_asg[dest] + gen_off : nullptr;
        }
        dstb = S2 + idx;
    }
    const long long stride = cap;
    const unsigned int par = surv_parent[t];
    const unsigned char sl = surv_slot[t];
    const int jj = sl & 0x7F;                    // putative index = row of V
    const bool picked = sl & 0x80;
    const int K = Kc[par];
    PF_ASSERT((long long)par < cap && idx >= 0 && idx < cap && K >= 0 && K <= k_max && jj <= K + 1);
    // the component the putative assigns y to (state_to_index, src/statepriors.jl:22,103; k_expand_sp's table)
    int j = jj;
    bool sticky = false;
    const int sp_kind = pc->sp_kind;
    if (scp->forced >= 0) j = scp->forced;
    else if (sp_kind == SPK_CHANGEPOINT) j = K == 0 ? 0 : K - 1 + jj;
    else if (sp_kind == SPK_STICKY && K > 0) { if (jj == 0) { j = sp_cur.last[par]; sticky = true; } else j = jj - 1; }
    const double v = V[(long long)jj * cap + par];
    const double lw = picked ? res->lw_resamp : (log(v) - res->log_total);
    const int Knew = K + (j == K ? 1 : 0);
    if (m_next) {        // putatives this child will have at the next observation
        const unsigned int am = __activemask();
        const unsigned int c2 = __reduce_add_sync(am, (unsigned int)(Knew + (Knew < k_max ? 1 : 0)));
        if ((threadIdx.x & 31) == (unsigned)(__ffs(am) - 1)) atomicAdd((unsigned long long *)m_next, (unsigned long long)c2);
    }
    logw2[idx] = lw; Kc2[idx] = Knew;
    if (gen_anc) { gen_anc[idx] = par + anc_off; gen_asg[idx] = (unsigned char)j; }
    if (sp_kind == SPK_STICKY) {
        // add(crp::StickyCRP, x, sticky), src/statepriors.jl:106-128
        const int last = sp_cur.last[par];
        for (int k = 0; k < K; k++) {
            sp_nxt.N[(long long)k * cap + idx] = sp_cur.N[(long long)k * cap + par];
            sp_nxt.Nsame[(long long)k * cap + idx] = sp_cur.Nsame[(long long)k * cap + par];
        }
        double Nj = j < K ? sp_cur.N[(long long)j * cap + par] : 0.0, Sj = j < K ? sp_cur.Nsame[(long long)j * cap + par] : 0.0;
        if (j == last) Sj += 1.0;
        if (!sticky) Nj += 1.0;
        sp_nxt.N[(long long)j * cap + idx] = Nj; sp_nxt.Nsame[(long long)j * cap + idx] = Sj;
        sp_nxt.last[idx] = j;
    }
    // copy the parent's slots (the chosen one is rewritten below)
    {
        const R *src = S + par;
        const int nrows = K * L::F;
        int r = 0;
        for (; r + 8 <= nrows; r += 8) {
            R tmp[8];
#pragma unroll
            for (int k = 0; k < 8; k++) tmp[k] = src[(long long)(r + k) * cap];
#pragma unroll
            for (int k = 0; k < 8; k++) dstb[(long long)(r + k) * stride] = tmp[k];
        }
        for (; r < nrows; r++) dstb[(long long)r * stride] = src[(long long)r * cap];
    }
    if (j < K) {
        const R *src = S + ((long long)j * L::F) * cap + par;
        R *dst = dstb + ((long long)j * L::F) * stride;
        double fl[L::F], y[D], mu0[D];
#pragma unroll
        for (int f = 0; f < L::F; f++) fl[f] = (double)src[(long long)f * cap];
#pragma unroll
        for (int k = 0; k < D; k++) { y[k] = scp->y[k]; mu0[k] = pc->mu0[k]; }
        slot_add<D>(fl, tab[(int)fl[L::F_N]], y, mu0);
#pragma unroll
        for (int f = 0; f < L::F; f++) dst[(long long)f * stride] = (R)fl[f];
    } else {
        R *dst = dstb + ((long long)K * L::F) * stride;
#pragma unroll
        for (int f = 0; f < L::F; f++) dst[(long long)f * stride] = (R)scp->newslot[f];
    }
}

template <int D, typename R = double>
__global__ void __launch_bounds__(256, PF_INST_MINBLOCKS) k_instantiate(const R *__restrict__ S, const int *__restrict__ Kc,
                                                     R *S2, int *Kc2,
                                                     double *logw2, long long cap,
                                                     const NRow *__restrict__ tab, const PriorConst *__restrict__ pc,
                                                     const StepConst *__restrict__ scp, const StepState *__restrict__ st,
                                                     const SelResult *__restrict__ res, const double *__restrict__ V,
                                                     const unsigned int *__restrict__ surv_parent,
                                                     const unsigned char *__restrict__ surv_slot,
                                                     unsigned int *gen_anc, unsigned char *gen_asg,
                                                     const RankOff *__restrict__ ro, int k_max,
                                                     const SelScratch *guard, long long *m_next,
                                                     const PeerTable *__restrict__ peers, long long gen_off,
                                                     SpArrays sp_cur = SpArrays{nullptr, nullptr, nullptr}, SpArrays sp_nxt = SpArrays{nullptr, nullptr, nullptr})
{
    if (guard && (guard->phase != 4 || guard->halt)) return;
    const long long n_out = ro ? ro->n_local : st->n_next;
    // (grid-stride: a sharded rank sizes its grid for a balanced share and loops when it produced more)
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n_out; t += (long long)gridDim.x * blockDim.x)
        instantiate_one<D, R>(t, S, Kc, S2, Kc2, logw2, cap, tab, pc, scp, st, res, V, surv_parent, surv_slot, gen_anc, gen_asg, ro, k_max, guard, m_next, peers, gen_off, sp_cur, sp_nxt);
}

// pipelined step: the number of putatives the children will have (M of the next observation, exact: it decides
// "keep everything" and scales the sample bracket), from the survivor list alone
__global__ void __launch_bounds__(256) k_count_next(const StepState *__restrict__ st, const int *__restrict__ Kc,
                                                    const unsigned int *__restrict__ surv_parent, const unsigned char *__restrict__ surv_slot,
                                                    int k_max, long long *m_pred, const SelScratch *guard, const RankOff *ro = nullptr)
{
    if (guard->halt) return;
    const long long n = ro ? ro->n_local : st->n_live;      // (sharded: the survivors this rank produced)
    unsigned int c = 0;
    // four independent (list, gather) chains per thread in flight
    const long long nth = (long long)gridDim.x * blockDim.x;
    for (long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; t0 < n; t0 += 4 * nth) {
        unsigned int par[4]; unsigned char sl[4]; int K[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { const long long t = t0 + k * nth; par[k] = t < n ? surv_parent[t] : 0u; sl[k] = t < n ? surv_slot[t] : 0; }
#pragma unroll
        for (int k = 0; k < 4; k++) K[k] = (t0 + k * nth < n) ? Kc[par[k]] : 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (t0 + k * nth < n) {
                const int Kn = K[k] + (((int)(sl[k] & 0x7F) == K[k]) ? 1 : 0);
                c += (unsigned int)(Kn + (Kn < k_max ? 1 : 0));
            }
        }
    }
    c = __reduce_add_sync(0xffffffffu, c);
    __shared__ unsigned int s_c[8];
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tot = 0;
        for (int w = 0; w < 8; w++) tot += s_c[w];
        if (tot) atomicAdd((unsigned long long *)m_pred, tot);
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
                                  unsigned int *__restrict__ surv_parent, unsigned char *__restrict__ surv_slot,
                                  const SelScratch *guard)
{
    if (guard && (guard->phase != 4 || guard->halt)) return;
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

// fill the per-step log record (one thread) and close the step: a search that did not finish in the
// queued rounds halts the handle (every later kernel is a no-op, the host reports it); a sharded run
// moves the shard's global offset and tells every rank that this rank's part of the next generation
// is complete (exchange kind D)
__global__ void k_steplog(StepState *st, const SelResult *res, StepLog *lg, SelScratch *sc, long long step, int dbg_status,
                          const RankOff *ro, const Xch *x, unsigned long long seq)
{
    if (threadIdx.x || blockIdx.x) return;
    if (sc->halt) return;
    if (sc->phase != 4) { sel_halt(sc, step, PF_HALT_SEARCH); return; }
    StepLog r;
    r.n_in = st->n_live; r.M = st->M; r.n_kept = st->n_kept; r.n_req = res->keep_all ? 0 : res->n;
    r.n_got = st->n_picked; r.n_out = ro ? ro->n_global : st->n_next; r.overflow = st->overflow; r.n_cand = res->n_cand_first; r.n_plateau = res->n_plateau_first;
    r.log_total = res->log_total; r.lw_resamp = res->lw_resamp; r.thr = res->thr_value; r.cv = 0.0;
    r.resampled = res->keep_all ? 0 : 1; r.rounds = res->rounds + (dbg_status ? 1000 * res->status : 0); r.status = res->status; r.pad = 0;
    r.sel_us[0] = (res->t_ns[1] - res->t_ns[0]) * 1e-3; r.sel_us[1] = (res->t_ns[2] - res->t_ns[1]) * 1e-3;
    r.sel_us[2] = (res->t_ns[3] - res->t_ns[2]) * 1e-3; r.sel_us[3] = (res->t_ns[4] - res->t_ns[0]) * 1e-3;
    *lg = r;
    if (ro) st->goff = ro->goff_next;
    if (x) xch_signal(x, XCH_D, seq);
}
