/*
 * particles_b200.h -- C ABI of libparticles_b200.so
 *
 * B200-native (sm_100a) implementation of ONE hot path of kleinschmidt/Particles.jl:
 * the online particle-filter step for Dirichlet-process mixture clustering
 * (fit!/filter! on FearnheadParticles and ChenLiuParticles with NormalInverseChisq and
 * NormalInverseWishart priors under a ChineseRestaurantProcess state prior).
 *
 * The reference is pure Julia and has no FFI of its own; the boundary it offers is its
 * generic-function surface.  Each entry point below names the reference method it stands
 * behind (file:line in the reference repository).  The Julia shim that binds these with
 * `ccall` is particles.jl_b200/julia/Particles.jl; the Python mirror used by the tests is
 * particles.jl_b200/particles_b200/.
 *
 * Conventions
 *   - every function returns an int32 status (PF_OK = 0); pf_last_error(h) gives the text;
 *   - the caller owns all host buffers and keeps them alive for the duration of the call;
 *     the library owns all device memory;
 *   - arrays are column-major Float64; indices cross the ABI 0-based (the shim adds 1);
 *   - one handle is driven from one host thread at a time;
 *   - there is NO CPU fallback: every entry point that computes needs a CUDA device and
 *     fails with PF_ERR_CUDA if none is usable.
 */
#ifndef PARTICLES_B200_H
#define PARTICLES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pf_filter_s *pf_handle;

enum { PF_OK = 0, PF_ERR_ARG = 1, PF_ERR_CUDA = 2, PF_ERR_STATE = 3, PF_ERR_HISTORY = 4,
       PF_ERR_OVERFLOW = 5, PF_ERR_NCCL = 6 };

enum { PF_FEARNHEAD = 0,   /* FearnheadParticles, src/fearnhead.jl:13-30 */
       PF_CHENLIU = 1 };   /* ChenLiuParticles,  src/chenliu.jl:7-26    */
enum { PF_NIX2 = 0,        /* NormalInverseChisq(mu, sigma2, kappa, nu)  (ConjugatePriors) */
       PF_NIW = 1 };       /* NormalInverseWishart(mu, kappa, Lambda, nu)                 */
enum { PF_F64 = 0, PF_F32 = 1 };
/* Traversal order of the low-weight partition in Fearnhead's resampling step:
 * PF_ORDER_SORTED reproduces the reference literally (ascending weight after sort!,
 * src/fearnhead.jl:148, ties broken by putative index; output [resampled ; kept]);
 * PF_ORDER_INDEX is the sort-free variant the reference's TODO proposes
 * (src/fearnhead.jl:140-145): same kept set, same sampler, putative-index order. */
enum { PF_ORDER_INDEX = 0, PF_ORDER_SORTED = 1 };

typedef struct {
    int32_t filter_kind;      /* PF_FEARNHEAD | PF_CHENLIU */
    int32_t prior_kind;       /* PF_NIX2 | PF_NIW */
    int32_t d;                /* observation dimension (NIX2: 1) */
    int32_t precision;        /* PF_F64 (PF_F32: reserved) */
    int64_t n_particles;      /* N: particle budget (src/fearnhead.jl:15, src/chenliu.jl:9) */
    int32_t k_max;            /* cluster slots per particle (reference: unbounded, src/particle.jl:142-143);
                                 at K == k_max the new-cluster candidate is dropped and counted */
    int32_t order;            /* PF_ORDER_INDEX | PF_ORDER_SORTED (Fearnhead only) */
    int64_t history;          /* generations of (ancestor, assignment) kept for assignments(); 0 = none */
    int32_t device;           /* CUDA device ordinal */
    int32_t reserved0;
    int32_t rank, nranks;     /* sharded run: this handle owns shard `rank` of `nranks` (0, 0 or 0, 1: unsharded) */
    double alpha;             /* ChineseRestaurantProcess(alpha), src/statepriors.jl:47-52 */
    double rejuv_threshold;   /* ChenLiuParticles rejuvination_threshold (percent CV), src/chenliu.jl:25 */
    double kappa0, nu0;
    double sigma2_0;          /* NIX2 only */
    const double *mu0;        /* [d] */
    const double *lambda0;    /* NIW only: [d*d] column-major, symmetric positive definite */
    uint64_t seed;            /* device uniform stream, used when a step is given uniforms == NULL */
} pf_config;

typedef struct {
    int64_t step;             /* observations filtered so far */
    int64_t n_in;             /* population before the step */
    int64_t n_putative;       /* M, src/fearnhead.jl:135 */
    int64_t n_kept;           /* src/fearnhead.jl:154 */
    int64_t n_resamp_req;     /* src/fearnhead.jl:157 */
    int64_t n_resamp_got;     /* src/fearnhead.jl:165-167 */
    int64_t n_out;            /* population after the step */
    int64_t overflow;         /* new-cluster candidates dropped because K == k_max (cumulative) */
    double log_total_w;       /* log(total_w): the log-evidence increment of this observation */
    double log_w_resamp;      /* log(w_resamp / total_w), src/fearnhead.jl:173 */
    double threshold;         /* kept set = {w >= threshold} (linear putative weight) */
    double cv;                /* Chen-Liu coefficient of variation, src/chenliu.jl:49-52 */
    int32_t resampled;        /* Fearnhead: M > N ; Chen-Liu: cv > threshold */
    int32_t select_rounds;    /* bracket rounds the threshold search needed */
    int64_t n_candidates;     /* putatives inside the first bracket */
    double select_phase_us[4]; /* threshold search: bracket, classify pass(es), descent + solve, total */
} pf_step_stats;

/* ---- life cycle ---------------------------------------------------------------------- */
/* FearnheadParticles(n, prior, stateprior) src/fearnhead.jl:29-30 (ONE empty particle);
 * ChenLiuParticles(n, prior, stateprior; rejuv) src/chenliu.jl:25-26 (N empty particles). */
int32_t pf_create(const pf_config *cfg, pf_handle *out);
int32_t pf_destroy(pf_handle h);
/* Text of the last error on this handle (h == NULL: last pf_create failure). */
const char *pf_last_error(pf_handle h);

/* ---- the hot path -------------------------------------------------------------------- */
/* fit!(ps, y): src/fearnhead.jl:130-184, src/chenliu.jl:54-69.
 * y[d].  uniforms: Fearnhead consumes 1 (the rand() of src/fearnhead.jl:99); Chen-Liu
 * consumes n_particles (the categorical draw of src/chenliu.jl:35) followed by n_particles
 * more (stratified rejuvenation, read only when cv > threshold).  uniforms == NULL draws
 * them from the device stream pf_uniform(seed, step, stream, i) defined below. */
int32_t pf_step(pf_handle h, const double *y, const double *uniforms, int64_t n_uniforms);
/* filter!(ps, ys, false): src/filters.jl:6-13 for the no-callback case.  ys[d x T]
 * column-major; uniforms[uniforms_per_step x T] or NULL; log_evidence[T] or NULL receives
 * log(total_w) per observation.  Runs asynchronously on the handle's stream and
 * synchronises before returning. */
int32_t pf_filter(pf_handle h, const double *ys, int64_t T, const double *uniforms,
                  int64_t uniforms_per_step, double *log_evidence);

/* ---- read-out (each synchronises) ---------------------------------------------------- */
int64_t pf_n_particles(pf_handle h);          /* length(particles(ps)), src/filters.jl:15 */
int64_t pf_n_steps(pf_handle h);
int32_t pf_get_logweights(pf_handle h, double *out);      /* log(weight(p)), src/particle.jl:19 */
int32_t pf_get_ncomponents(pf_handle h, int32_t *out);    /* ncomponents(p), src/particle.jl:157-158 */
/* One particle, for materialising particles(ps)[i] lazily (src/filters.jl:15):
 * counts[K] = CRP N (src/statepriors.jl:49) = nobs per component; means[d x K] = sample
 * means (NormalStats.m / MvNormalStats.m); chol[d x d x K] = lower Cholesky factor of the
 * posterior Lambda_n (NIX2: sqrt(nu_n sigma2_n)); logdet[K] = log det Lambda_n.
 * Buffers sized for k_max components. */
int32_t pf_get_particle(pf_handle h, int64_t i, int32_t *K, double *counts, double *means,
                        double *chol, double *logdet);
/* assignments(ps), src/filters.jl:26-34: out[T x n] column-major (column = particle),
 * 0-based component labels.  Needs history >= steps. */
int32_t pf_get_assignments(pf_handle h, int32_t *out);
int32_t pf_get_last_step(pf_handle h, pf_step_stats *out);

/* ---- debugging taps used by the parity tests ------------------------------------------ */
/* Putative weights exp(logweight) of the LAST step (src/particle.jl:114) in
 * particle-major / candidate-minor order (src/fearnhead.jl:132); *M receives their count. */
int32_t pf_get_putative_weights(pf_handle h, double *out, int64_t cap, int64_t *M);
int32_t pf_get_last_ancestors(pf_handle h, int32_t *out);    /* p.ancestor as 0-based index */
int32_t pf_get_last_assignments(pf_handle h, int32_t *out);  /* p.assignment - 1 */

/* The selection step of fit!(::FearnheadParticles) (cutoff_ascending + sample_stratified!,
 * src/fearnhead.jl:59-77, 97-114, 146-179) on caller-supplied weights w[M]: the same device
 * kernels the filter runs, exposed so that "identical weights and uniforms => identical
 * survivors" can be tested directly.  survivors[<= N] receives 0-based indices into w in
 * output order, resampled[<= N] is 1 where the survivor came from the resampled partition.
 * stats: n_kept, n_resamp_req, n_resamp_got, n_out, threshold, log_w_resamp (= log c,
 * unnormalised), log_total_w. */
int32_t pf_select(int32_t device, const double *w, int64_t M, int64_t N, double u, int32_t order,
                  int64_t *survivors, uint8_t *resampled, pf_step_stats *stats);

/* ---- multi-GPU: particles block-sharded, one process (rank) per GPU ------------------
 * The population of a Fearnhead filter is block-sharded in global (rank-major) order; every
 * rank runs the same kernels on its shard and the host program supplies the collectives
 * between the stages below ("bring your own communicator": torch.distributed / NCCL in
 * particles_b200/distributed.py, MPI.jl or NCCL.jl under the Julia shim).  What crosses the
 * ranks per observation: the strided sample and the bracket's candidates with their
 * accumulator blocks (all-gather of a few MB), one 32-byte total per rank (all-gather) and,
 * only when a share outgrows its shard, the records of the particles that change owner in
 * the re-partitioning (all-to-all-v).  All decisions are integer arithmetic on gathered data, so every rank
 * reaches the same threshold and an N-rank run reproduces the 1-rank run bit for bit.
 *
 *   pf_mg_begin(y, u)                     expansion of the local shard
 *   pf_mg_select(PF_MG_SAMPLE, pad_to = sample_entries)        local strided sample -> list, acc
 *   all_gather(acc) ; all_gather(list[0 : sample_entries])
 *   pf_mg_select(PF_MG_BRACKET | PF_MG_CLASSIFY, gathered acc, gathered list, sample_entries, pad_to = cand_entries)
 *                                         bracket from the gathered sample, classify the shard
 *   all_gather(acc) ; all_gather(list[0 : cand_entries])
 *   pf_mg_select(PF_MG_SOLVE, gathered acc, gathered list, cand_entries)
 *                                         verify / descend / solve on the gathered candidates
 *   pf_mg_tiles()                         local survivor-scan totals -> local_total
 *   all_gather(local_total)
 *   pf_mg_finish(gathered totals)         offsets, survivor scan, instantiate (local + send buffer)
 *   pf_mg_plan(...)                       THE host synchronisation of the step: search outcome
 *                                         (phase 4 = done; 3 / 7 = another PF_MG_CLASSIFY + exact-size
 *                                         gathers + PF_MG_SOLVE, then tiles/finish again) and the
 *                                         exchange plan; the kernels after PF_MG_SOLVE are no-ops
 *                                         until the search is done, so they may be queued behind it
 *   [exchange != 0] all_to_all_v(send buffer -> recv buffer, split sizes = counts * rows)
 *   pf_mg_unpack(recv buffer)             received particles -> local store; step done          */
enum { PF_MG_SAMPLE = 1, PF_MG_BRACKET = 2, PF_MG_CLASSIFY = 4, PF_MG_SOLVE = 8 };
typedef struct {
    void *scalars;          /* int64[2] device: {M (sum), max weight bit pattern (max)} */
    void *acc;              /* device: accumulator block of the threshold search, acc_bytes bytes */
    int64_t acc_bytes;
    void *list;             /* device: local sample / candidate list (doubles), list_capacity entries */
    int64_t list_capacity;
    void *local_total;      /* device: 32-byte survivor-scan total of this shard */
    int64_t total_bytes;
    void *send_buffer;      /* device: packed records of emigrating particles (doubles) */
    void *recv_buffer;      /* device: packed records of immigrating particles (doubles) */
    int64_t buffer_capacity;   /* doubles in each of send_buffer / recv_buffer */
    int64_t rows_per_particle; /* doubles per migrating particle */
    int64_t local_capacity;    /* particles this shard can hold */
    int64_t sample_entries;    /* fixed gather size (doubles) of the sample list */
    int64_t cand_entries;      /* fixed gather size (doubles) of the candidate list */
    int64_t window;            /* neighbour exchange: particles per message slot (send/recv buffers hold 2 slots) */
} pf_mg_ptrs;
int32_t pf_set_stream(pf_handle h, void *cuda_stream);    /* run on the caller's stream (e.g. torch's current stream) */
int32_t pf_mg_info(pf_handle h, pf_mg_ptrs *out);
int32_t pf_mg_begin(pf_handle h, const double *y, double u);
int32_t pf_mg_select(pf_handle h, int32_t stage, int32_t rounds_done, const void *gathered_acc, const void *gathered_list,
                     int64_t list_each, int64_t pad_to);
/* state[0] = phase (4 done, 3 classify again), [1] = local list entries in use (multiple of 256),
 * [2] = 1 if this step needs the sample, [3] = classify rounds done.  Synchronises the stream. */
int32_t pf_mg_state(pf_handle h, int64_t state[4]);
int32_t pf_mg_tiles(pf_handle h);
int32_t pf_mg_finish(pf_handle h, const void *gathered_totals);
/* Exchange plan of the re-partitioning, in particles: send_counts[nranks] by destination,
 * recv_counts[nranks] by source.  Synchronises the stream. */
int32_t pf_mg_plan(pf_handle h, int64_t *send_counts, int64_t *recv_counts, int64_t *n_local, int64_t *n_global,
                   int64_t *phase, int64_t *exchange);
int32_t pf_mg_unpack(pf_handle h, const void *recv_buffer);
/* Direct exchange: when every rank's next-generation arrays are mapped into this process
 * (pf_mg_ipc_handles on the owner -> pf_mg_set_peer here; or plain pointers from pf_mg_local_arrays when
 * the ranks live in one process), k_instantiate stores a child that changes owner straight into its
 * new owner's store over NVLink and the population is re-partitioned into equal blocks at EVERY
 * step: no send/recv buffers, no all-to-all, no host-known sizes.  The host program must place a
 * collective (e.g. the next step's first all-gather, or a barrier) between pf_mg_finish and the
 * next pf_mg_begin so that remote stores are complete before their owner reads them. */
int32_t pf_mg_local_arrays(pf_handle h, void **out /* [8] */);
int32_t pf_mg_ipc_handles(pf_handle h, void *out /* 8 x 64 bytes */);
int32_t pf_mg_set_peer(pf_handle h, int32_t peer, void *const *ptrs, const void *ipc_handles);
int32_t pf_mg_enable_direct(pf_handle h);
/* Look-ahead: pf_mg_close() after pf_mg_finish marks the step as "needs the host" on the device when the search
 * did not finish or particles must change owner; every kernel of later steps is then a no-op, so a host program
 * may queue several observations ahead and call pf_mg_drain() (synchronises; out = {halt (1 + step index, or 0),
 * phase, local population, global population}) only now and then.  After completing a halted step through the
 * synchronous stages, pf_mg_resume(step) rewinds the host mirrors to `step` and clears the mark. */
int32_t pf_mg_close(pf_handle h);
int32_t pf_mg_set_obs(pf_handle h, const double *y, double u);   /* re-upload (y, u) of a halted step before completing it */
int32_t pf_mg_drain(pf_handle h, int64_t out[4]);
int32_t pf_mg_resume(pf_handle h, int64_t step);
/* Raw control-block words of the last threshold search / survivor scan (debugging aid). */
int32_t pf_debug_dump(pf_handle h, int64_t *out /* [24] */);

/* ---- instrumentation ------------------------------------------------------------------ */
/* Device milliseconds (CUDA events on the handle's stream) and kernel launches of the last
 * pf_filter / pf_step call. */
int32_t pf_last_timing(pf_handle h, double *device_ms, int64_t *kernel_launches);
/* Per-kernel-class device times: when profiling is on, every launch of the step sequence is
 * bracketed by CUDA events on the handle's stream; times accumulate until the next
 * pf_set_profiling call.  Classes: 0 prestep, 1 expand, 2 select, 3 scan, 4 instantiate,
 * 5 steplog, 6 sort, 7 cl_propagate, 8 cl_stats, 9 cl_finish. */
#define PF_KERNEL_CLASSES 10
int32_t pf_set_profiling(pf_handle h, int32_t on);
int32_t pf_get_kernel_times(pf_handle h, double *ms /* [PF_KERNEL_CLASSES] */, int64_t *launches /* [PF_KERNEL_CLASSES] */);
const char *pf_kernel_class_name(int32_t k);
/* Raw stream pointer (cudaStream_t) so that a host program can order its own work. */
void *pf_stream(pf_handle h);
const char *pf_version(void);

/* Device uniform stream (SplitMix64 finaliser), reproducible on the host:
 *   u = (mix(mix(seed + step * 0xD1342543DE82EF95 + stream) + i) >> 11) * 2^-53
 * stream 0: Fearnhead comb offset (i = 0); 1: Chen-Liu categorical draw; 2: rejuvenation. */
#ifdef __CUDACC__
#define PF_HD __host__ __device__
#else
#define PF_HD
#endif
PF_HD static inline uint64_t pf_mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
PF_HD static inline double pf_uniform(uint64_t seed, uint64_t step, uint64_t stream, uint64_t i)
{
    uint64_t k = pf_mix64(seed + step * 0xD1342543DE82EF95ull + stream);
    return (double)(pf_mix64(k + i) >> 11) * (1.0 / 9007199254740992.0);
}

#ifdef __cplusplus
}
#endif
#endif /* PARTICLES_B200_H */
