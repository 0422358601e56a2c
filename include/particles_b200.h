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
 * `ccall` is particles.jl_b200/julia/device.jl; the Python mirror used by the tests is
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
/* State priors, src/statepriors.jl: ChineseRestaurantProcess(alpha) :47-69; StickyCRP(alpha, rho) :93-145;
 * ChangePoint(p) :162-190; NStatePrior(counts) :203-221.  Priors other than the CRP run the Fearnhead filter on a
 * single GPU (the classic expansion -> search sequence). */
enum { PF_SP_CRP = 0, PF_SP_STICKY = 1, PF_SP_CHANGEPOINT = 2, PF_SP_NSTATE = 3 };

typedef struct {
    int32_t filter_kind;      /* PF_FEARNHEAD | PF_CHENLIU */
    int32_t prior_kind;       /* PF_NIX2 | PF_NIW */
    int32_t d;                /* observation dimension (NIX2: 1) */
    int32_t precision;        /* storage type of the particle store: PF_F64, or PF_F32 (every cluster field is kept as a
                                 float: half the store traffic; the predictive arithmetic, the putative weights and the
                                 selection stay as in PF_F64.  Log-weights agree with the FP64 filter to ~1e-6 relative) */
    int64_t n_particles;      /* N: particle budget (src/fearnhead.jl:15, src/chenliu.jl:9) */
    int32_t k_max;            /* cluster slots per particle (reference: unbounded, src/particle.jl:142-143);
                                 at K == k_max the new-cluster candidate is dropped and counted */
    int32_t order;            /* PF_ORDER_INDEX | PF_ORDER_SORTED (Fearnhead only) */
    int64_t history;          /* generations of (ancestor, assignment) kept for assignments(): > 0 fixed capacity (dense), 0 none,
                                 < 0 grows on demand (the reference's ancestor chain, src/particle.jl:25-32, is unbounded): on
                                 a single GPU an arena pruned of dead lineages (pf_genealogy_info), sharded a dense log */
    int32_t device;           /* CUDA device ordinal */
    int32_t stateprior_kind;  /* PF_SP_* */
    int32_t rank, nranks;     /* sharded run: this handle owns shard `rank` of `nranks` (0, 0 or 0, 1: unsharded) */
    double alpha;             /* ChineseRestaurantProcess(alpha), src/statepriors.jl:47-52 */
    double rejuv_threshold;   /* ChenLiuParticles rejuvination_threshold (percent CV), src/chenliu.jl:25 */
    double kappa0, nu0;
    double sigma2_0;          /* NIX2 only */
    const double *mu0;        /* [d] */
    const double *lambda0;    /* NIW only: [d*d] column-major, symmetric positive definite */
    uint64_t seed;            /* device uniform stream, used when a step is given uniforms == NULL */
    double sp_param;          /* StickyCRP: rho ; ChangePoint: probability of a change */
    const double *sp_counts;  /* NStatePrior: counts[sp_n] (the particles start with sp_n empty components) */
    int32_t sp_n;
    int32_t reserved2;
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
    int64_t n_plateau;        /* in-bracket copies of the step's exactly tied value (counted, not listed) */
    int32_t select_status;    /* bit set of the threshold search: 1 fused classification used, 2 candidate staging overflowed,
                                 4 kept-side scale unusable, 8 / 16 bracket failed low / high, 32 rescaled, 64 comb rescaled */
    int32_t reserved1;
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
/* filter!(ps, Labeled.(ys, labels)): src/labeled.jl:1-11.  labels[T]: 0-based component of observation t, or -1
 * (missing: the ordinary step).  A labeled observation gives every particle the single putative fit(p, y, label)
 * (src/particle.jl:154); a label beyond K (K + 1 under NStatePrior ... any invalid state) is the ArgumentError of
 * src/particle.jl:137-139: PF_ERR_ARG.  Fearnhead filter, single GPU. */
int32_t pf_filter_labeled(pf_handle h, const double *ys, const int32_t *labels, int64_t T, const double *uniforms,
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
/* p.stateprior of particle i: StickyCRP N[K], Nsame[K], last (0-based) (src/statepriors.jl:95-100); the other
 * priors' counts follow from pf_get_particle (CRP N = NStatePrior counts - initial counts = nobs per component). */
int32_t pf_get_stateprior(pf_handle h, int64_t i, double *N, double *Nsame, int32_t *last);
/* assignments(ps), src/filters.jl:26-34: out[T x n] column-major (column = particle),
 * 0-based component labels.  Needs history >= steps. */
int32_t pf_get_assignments(pf_handle h, int32_t *out);
/* assignments(p) of ONE particle (src/particle.jl:25-32): out[T], 0-based.  Stays cheap when the T x n matrix of a long
 * run over a large population is too big to bring to the host (2^24 particles x 10^4 observations = 1.7e11 labels). */
int32_t pf_get_particle_assignments(pf_handle h, int64_t i, int32_t *out);
/* The genealogy log: out[0] generations stored, out[1] nodes stored, out[2] capacity in nodes, out[3] pruning passes so
 * far.  With history < 0 on a single GPU the log is an arena that is pruned of dead lineages when it fills up (the
 * reference's ancestor chain is pruned by the garbage collector, src/particle.jl:25-32): a generation then keeps only
 * the nodes that are still an ancestor of somebody in the current population. */
int32_t pf_genealogy_info(pf_handle h, int64_t *out /* [4] */);
int32_t pf_get_last_step(pf_handle h, pf_step_stats *out);

/* ---- filter-level summaries, evaluated on the device-resident population ------------------ */
/* posterior_predictive(ps), src/filters.jl:22-24 with src/particle.jl:156-172 and src/component.jl:12-23: the density
 * of the next observation, sum_i w_i sum_j pi_ij t_ij(x), at the Q query points xs[d x Q] (column-major).  out[Q].
 * ChineseRestaurantProcess state prior. */
int32_t pf_posterior_predictive(pf_handle h, const double *xs, int64_t Q, double *out);
/* mean(f, ps), src/filters.jl:17-18, for the two per-particle functions the reference ships:
 * ncomponents (src/particle.jl:157-158) and state_entropy (src/particle.jl:162, src/filters.jl:20). */
enum { PF_MEAN_NCOMPONENTS = 0, PF_MEAN_STATE_ENTROPY = 1 };
int32_t pf_weighted_mean(pf_handle h, int32_t what, double *out);
/* ncomponents_dist(ps), src/filters.jl:36-37: out[k] = weighted share of the particles with k components, k = 0..k_max. */
int32_t pf_ncomponents_dist(pf_handle h, double *out /* [k_max + 1] */);
/* assignment_similarity(ps), src/filters.jl:39-42: out[T x T] = 1 - Hamming(assignments of s, of t) / n.  Needs history. */
int32_t pf_assignment_similarity(pf_handle h, double *out);
/* randindex(ps, truth, type), src/filters.jl:53-80: truth[T] labels; adjusted != 0 -> :adjusted.  T != observations
 * filtered is the DimensionMismatch of :57-58: PF_ERR_ARG. */
int32_t pf_randindex(pf_handle h, const int32_t *truth, int64_t T, int32_t adjusted, double *out);

/* ---- checkpoint / restore -------------------------------------------------- */
/* The whole state of a single-GPU handle (population, control blocks, genealogy, step counter) in one file;
 * a loaded handle continues bit for bit where the saved one stopped. */
int32_t pf_save(pf_handle h, const char *path);
int32_t pf_load(const char *path, int32_t device, pf_handle *out);
/* The configuration a handle was created (or loaded) with; the array members of *out are NULL, pf_get_prior copies
 * them: mu0[d], lambda0[d x d] (NIX2: nu0 sigma2_0), sp_counts[sp_n]. */
int32_t pf_get_config(pf_handle h, pf_config *out);
int32_t pf_get_prior(pf_handle h, double *mu0, double *lambda0, double *sp_counts);

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

/* ---- multi-GPU: particles block-sharded over the GPUs of one node ------------------------
 * The population of a Fearnhead filter is block-sharded in global (rank-major) order, one handle
 * (shard) per GPU: pf_config.rank / nranks.  Every rank runs the same kernels on its shard; what
 * crosses the ranks per observation -- the strided sample, the bracket's candidates with their
 * accumulator blocks (a few MB), one 32-byte survivor-scan total per rank, and the children that
 * change owner when the new population is re-partitioned into equal blocks -- is written by the
 * kernels themselves into the peers' device memory over NVLink (CUDA IPC between processes, peer
 * access inside one process) and synchronised by per-rank sequence flags the kernels spin on:
 * there is no collective library and no host round trip inside pf_mg_filter.  All decisions are
 * integer arithmetic on the gathered data, so every rank reaches the same threshold and a G-rank
 * run reproduces the 1-rank run bit for bit (src/fearnhead.jl:130-184 on the whole population).
 *
 * One process per GPU (e.g. under torchrun / MPI; any side channel can carry the 1088-byte blobs):
 *     pf_create(cfg with rank, nranks)            on every rank
 *     pf_mg_ipc_handles(h, blob)                  -> all-gather the blobs (host side, once)
 *     pf_mg_set_peer(h, r, NULL, blob_r)          for every other rank r
 *     pf_mg_connect(&h, 1)
 *     pf_mg_filter(&h, 1, ys, T, uniforms, 1, log_evidence)      the same call on every rank
 * One process driving several GPUs (one Julia / C host thread, no launcher):
 *     pf_mg_create_local(cfg, devices, n, handles) ;  pf_mg_filter(handles, n, ...)
 * Read-outs (pf_get_logweights, pf_get_ncomponents, pf_get_last_ancestors, ...) return each shard's
 * part; concatenated in rank order they are the global arrays (ancestors are global indices).     */
#define PF_MG_NARRAYS 17        /* S[0], S[1], K[0], K[1], logw[0], logw[1], gen_anc, gen_asg, exchange region, Chen-Liu prefix array and
                                   assignments, V[0], V[1], putative counts, scan records (x, m, h) */
#define PF_MG_IPC_BYTES 64      /* sizeof(cudaIpcMemHandle_t) */
int32_t pf_set_stream(pf_handle h, void *cuda_stream);    /* run on the caller's stream */
int32_t pf_mg_local_arrays(pf_handle h, void **out /* [PF_MG_NARRAYS] */);
int32_t pf_mg_ipc_handles(pf_handle h, void *out /* PF_MG_NARRAYS x PF_MG_IPC_BYTES bytes */);
/* Arrays of rank `peer`: plain device pointers (ptrs, from pf_mg_local_arrays of a handle in this
 * process) or the IPC blob pf_mg_ipc_handles produced in the peer's process. */
int32_t pf_mg_set_peer(pf_handle h, int32_t peer, void *const *ptrs, const void *ipc_handles);
/* All peers set: upload the peer tables of the n_local handles this process drives. */
int32_t pf_mg_connect(pf_handle *hs, int32_t n_local);
/* filter!(ps, ys, false) (src/filters.jl:6-13) on the sharded population; hs = the shards of this
 * process.  Every process of the group makes the same call.  uniforms[uniforms_per_step x T] (the
 * first of each column is the step's rand(), src/fearnhead.jl:99) or NULL for the device stream.
 * PF_ERR_STATE / PF_ERR_OVERFLOW: a step could not be completed on the device (pf_last_error). */
int32_t pf_mg_filter(pf_handle *hs, int32_t n_local, const double *ys, int64_t T, const double *uniforms,
                     int64_t uniforms_per_step, double *log_evidence);
int64_t pf_mg_n_global(pf_handle h);          /* population over all shards */
/* assignments(ps) (src/filters.jl:26-34) of this shard's particles: out[T x pf_n_particles(h)] column-major, 0-based.
 * A particle's ancestor chain is followed through the genealogy arrays of whichever rank owned each ancestor (peer
 * memory); the shards' blocks concatenated in rank order are the unsharded filter's matrix.  Needs history. */
int32_t pf_mg_get_assignments(pf_handle h, int32_t *out);
/* Filter-level summaries of a sharded population (src/filters.jl:17-42, 53-80): this shard's UNNORMALISED sums; added
 * over the shards they give pred[q] / pred[Q] / (steps + alpha) = pdf of posterior_predictive(ps) at xs[:, q];
 * pop[1] / pop[0] = mean(ncomponents, ps), pop[2] / pop[0] = state_entropy(ps), pop[3 + k] / pop[0] = ncomponents_dist(ps);
 * sim[s * T + t] / n = assignment_similarity(ps) (counts of particles that agree; chains cross the shards through peer
 * memory).  pred [Q + 1] (with xs [d x Q]), pop [3 + k_max + 1], sim [T x T]: any of them may be NULL. */
int32_t pf_mg_summary_partials(pf_handle h, const double *xs, int64_t Q, double *pred, double *pop, uint64_t *sim);
int32_t pf_mg_create_local(const pf_config *cfg, const int32_t *devices, int32_t n, pf_handle *out /* [n] */);
/* Raw control-block words of the last threshold search / survivor scan (debugging aid). */
int32_t pf_debug_dump(pf_handle h, int64_t *out /* [24] */);

/* ---- instrumentation ------------------------------------------------------------------ */
/* Device milliseconds (CUDA events on the handle's stream) and kernel launches of the last
 * pf_filter / pf_step call. */
int32_t pf_last_timing(pf_handle h, double *device_ms, int64_t *kernel_launches);
/* Per-kernel-class device times: when profiling is on, every launch of the step sequence is
 * bracketed by CUDA events on the handle's stream; times accumulate until the next
 * pf_set_profiling call.  Classes: 0 prestep, 1 expand, 2 select, 3 scan, 4 instantiate,
 * 5 steplog, 6 sort, 7 cl_propagate, 8 cl_stats, 9 cl_finish, 10 inst_expand (instantiate fused with the next
 * observation's expansion: the pipelined steps of pf_filter). */
#define PF_KERNEL_CLASSES 11
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
