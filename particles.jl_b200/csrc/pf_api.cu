// pf_api.cu -- C ABI of libparticles_b200.so (see include/particles_b200.h) and the host
// orchestration of the per-observation kernel sequence.  No CPU fallback: every compute
// entry point needs a CUDA device.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/particles_b200.h"
#include "fearnhead.cuh"
#include "chenliu.cuh"
#include "sort.cuh"
#include "summaries.cuh"

static thread_local std::string g_create_error;

struct pf_filter_s {
    pf_config cfg;
    int d, F, k_max;
    long long N, cap;
    std::vector<double> mu0, lam0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // particle store (ping-pong)
    void *S[2] = {nullptr, nullptr};      // elements of the storage type: double, or float under PF_F32
    bool f32 = false;
    size_t esz = sizeof(double);          // bytes per store element
    double *logw[2] = {nullptr, nullptr};
    int *Kc[2] = {nullptr, nullptr};
    int cur = 0;
    // step scratch
    double *V = nullptr;                  // putative weights of the current step
    double *V_alt = nullptr;              // pipelined steps: the buffer the NEXT step's putatives are written to (swapped with V)
    StepConst *stc_alt = nullptr;         // ... and its per-observation constants (swapped with stc)
    double *Vbuf[2] = {nullptr, nullptr}; // the two V buffers by identity (peers address them by index); V == Vbuf[vcur]
    int vcur = 0;
    bool pipeline_mg = false;             // ... the same fusion on the sharded path (children and their putatives go to the new owner)
    bool instexp_attr = false;            // the fused kernel's dynamic shared memory limit has been raised
    bool pipeline = false;                // instantiate of step t fused with the expansion of step t + 1 inside pf_filter
    bool pre_expanded = false;            // the step about to run was expanded by the previous step's pipelined tail
    unsigned char *cnt = nullptr;
    unsigned int *surv_parent = nullptr;
    unsigned char *surv_slot = nullptr;
    TileSum *tiles = nullptr;
    TileSum *supers = nullptr;            // totals of SUPER_TILES consecutive tile records (level 2 of the survivor scan)
    TileFuse *tfuse = nullptr;            // per expansion block (EXP_SUB per tile)
    TileTooth *ttooth = nullptr;          // per tile of the index-order survivor scan: offsets and first comb tooth (k_tooth_bases)
    TileSum *tsub = nullptr;              // sub-tile records of the fused expansion (EXP_SUB per tile)
    ScanRec srec = ScanRec{nullptr, nullptr, nullptr};   // per-particle scan records of the fused expansion (PF_SCAN_REC)
    long long fuse_stride = 1;
    bool fused = true;                    // single-GPU index-order Fearnhead: classify inside the expansion
    int sel_margin = 8;                   // half-width of the sample bracket in sigmas (PF_SEL_MARGIN; tests shrink it)
    int dbg_status = 0;                   // PF_DEBUG_STATUS (read once at pf_create): selection status bits in the step log
    double *cand = nullptr;
    long long cand_cap = 0;
    SelScratch *sc = nullptr;
    SelResult *res = nullptr;
    StepState *st = nullptr;
    PriorConst *pc = nullptr;
    StepConst *stc = nullptr;
    NRow *tab = nullptr;
    long long tab_cap = 0;
    NRow row0;
    StepLog *slog = nullptr;
    long long slog_cap = 0;
    double *ys_dev = nullptr, *u_dev = nullptr;
    long long ys_cap = 0, u_cap = 0;
    // state prior (src/statepriors.jl) and labeled observations (src/labeled.jl)
    int sp_kind = PF_SP_CRP;
    std::vector<double> sp_counts;
    SpArrays sp[2] = {SpArrays{nullptr, nullptr, nullptr}, SpArrays{nullptr, nullptr, nullptr}};
    int cur_label = -1;                   // label of the observation being queued (-1: missing)
    // reference-literal order scratch
    u64 *keys = nullptr, *keys2 = nullptr;
    unsigned int *vals = nullptr, *vals2 = nullptr, *hist = nullptr, *pick_pos = nullptr;
    // Chen-Liu scratch
    ClState *cl = nullptr;
    double *W = nullptr;
    u128 *cum = nullptr;
    unsigned char *asg_tmp = nullptr;
    u128 *cl_tiles = nullptr;             // tile totals / prefixes of the rejuvenation scan
    ClPeers *clp = nullptr;               // every rank's store / prefix arrays (single GPU: entry 0 = the local arrays)
    ClPeers clp_host;
    ClShard shard{0, 0, 0, 0};            // this shard's block of the Chen-Liu population
    // sharded run
    int rank = 0, nranks = 1;
    RankOff *ro = nullptr;
    long long surv_cap = 0;
    bool own_stream = true;
    std::shared_ptr<void> shared_stream;  // lock-step group: one stream for all shards of a device, destroyed with the last of them
    XchRegion *xreg = nullptr;            // this rank's exchange region (peers store into it)
    size_t xreg_bytes = 0;
    long long capA = 0, capB = 0;
    Xch xch_host;                         // regions of all ranks as mapped in this process
    Xch *xch = nullptr;                   // device copy (null until pf_mg_connect)
    PeerTable *peers[2] = {nullptr, nullptr};   // peer arrays of the store buffer written at even / odd steps
    PeerTable peers_host[2];
    bool connected = false, lockstep = false;
    std::vector<void *> ipc_opened;
    // genealogy
    unsigned int *gen_anc = nullptr;
    unsigned char *gen_asg = nullptr;
    long long hist_cap = 0;
    bool hist_auto = false;               // pf_config.history < 0: the genealogy grows on demand
    // history < 0 on a single GPU: the log is an ARENA of generations of varying length (gen_off / gen_n), appended
    // densely and pruned of dead lineages when it fills up (prune_genealogy)
    bool arena = false;
    long long arena_cap = 0, arena_used = 0, arena_budget = 0;     // entries
    std::vector<long long> gen_off, gen_n;
    long long *gen_off_dev = nullptr;     // device copy of gen_off (uploaded before a backtrack)
    long long gen_off_dev_cap = 0;
    long long n_prunes = 0, last_prune_step = 0;
    std::vector<long long> gen_q;         // sharded run: block length q = ceil(n / G) of the population after each step
    // host mirrors
    long long steps = 0;
    long long n_host = 0;                 // population size after the last synchronised step
    long long ub_live = 1;                // host upper bound on n_live (grid sizing)
    std::vector<StepLog> logs;            // records of the last call
    StepLog last{};
    bool have_last = false, last_stale = false;
    int sel_grid = 0;
    int coop_grid = 0;                    // CTAs of the cooperative bracket / solve kernels (one per SM)
    int retries = SEL_RETRIES;            // queued extra search rounds per step
    int redo = 0;                         // the next step re-runs a halted one
    int num_sms = 0;
    double last_ms = 0.0;
    long long last_launches = 0;
    long long launches = 0;
    unsigned int epoch = 0;
    // per-kernel-class CUDA-event timing (pf_set_profiling)
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_cls;
    size_t ev_used = 0;
    double cls_ms[PF_KERNEL_CLASSES] = {0};
    long long cls_launches[PF_KERNEL_CLASSES] = {0};
    std::string err;
};

static int32_t set_error(pf_handle h, int32_t code, const std::string &msg)
{
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}
static int32_t set_cuda_error(pf_handle h, cudaError_t e, const char *what, int line)
{
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %s at pf_api.cu:%d: %s", cudaGetErrorName(e), line, what);
    cudaGetLastError();
    return set_error(h, PF_ERR_CUDA, buf);
}

// ------------------------------------------------------------------ per-kernel timing
static const char *k_class_names[PF_KERNEL_CLASSES] = {"prestep", "expand", "select", "scan", "instantiate", "steplog",
                                                        "sort", "cl_propagate", "cl_stats", "cl_finish", "inst_expand"};
enum { KC_PRESTEP = 0, KC_EXPAND, KC_SELECT, KC_SCAN, KC_INSTANTIATE, KC_STEPLOG, KC_SORT, KC_CL_PROP, KC_CL_STATS, KC_CL_FINISH, KC_INST_EXPAND };

struct LaunchTimer {            // records an event pair around one launch (or launch group) when profiling
    pf_handle h; int cls; size_t slot = 0; bool on;
    LaunchTimer(pf_handle h_, int cls_, int nlaunch = 1) : h(h_), cls(cls_), on(h_->profiling)
    {
        h->launches += nlaunch;
        h->cls_launches[cls] += nlaunch;
        if (!on) return;
        if (h->ev_used + 2 > h->ev_pool.size()) {
            for (int k = 0; k < 64; k++) { cudaEvent_t e; cudaEventCreate(&e); h->ev_pool.push_back(e); }
        }
        slot = h->ev_used; h->ev_used += 2;
        h->ev_cls.push_back(cls);
        cudaEventRecord(h->ev_pool[slot], h->stream);
    }
    ~LaunchTimer() { if (on) cudaEventRecord(h->ev_pool[slot + 1], h->stream); }
};

static void collect_profile(pf_handle h)
{
    if (!h->profiling) return;
    for (size_t k = 0; k < h->ev_cls.size(); k++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev_pool[2 * k], h->ev_pool[2 * k + 1]) == cudaSuccess) h->cls_ms[h->ev_cls[k]] += ms;
    }
    h->ev_cls.clear(); h->ev_used = 0;
}

// ------------------------------------------------------------------ host-side model setup
static NRow make_row(const pf_filter_s *f, long long n)
{
    return make_nrow(f->cfg.kappa0, f->cfg.nu0, f->cfg.alpha, f->d, n, f->sp_kind == PF_SP_CRP);
}

static int32_t ensure_table(pf_handle h, long long need)
{
    if (need <= h->tab_cap) return PF_OK;
    long long ncap = h->tab_cap ? h->tab_cap : 4096;
    while (ncap < need) ncap *= 2;
    std::vector<NRow> rows(ncap);
    for (long long n = 0; n < ncap; n++) rows[n] = make_row(h, n);
    NRow *nt = nullptr;
    CUDA_TRY(cudaMalloc(&nt, sizeof(NRow) * ncap));
    CUDA_TRY(cudaMemcpyAsync(nt, rows.data(), sizeof(NRow) * ncap, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (h->tab) cudaFree(h->tab);
    h->tab = nt; h->tab_cap = ncap;
    return PF_OK;
}

static bool chol_lower(int d, const double *A /*col-major*/, double *Lo /*col-major*/)
{
    std::fill(Lo, Lo + d * d, 0.0);
    for (int j = 0; j < d; j++)
        for (int i = j; i < d; i++) {
            double acc = A[i + j * d];
            for (int k = 0; k < j; k++) acc -= Lo[i + k * d] * Lo[j + k * d];
            if (i == j) { if (!(acc > 0.0)) return false; Lo[i + j * d] = std::sqrt(acc); }
            else Lo[i + j * d] = acc / Lo[j + j * d];
        }
    return true;
}

template <typename T>
static cudaError_t dmalloc(T **p, size_t n) { return cudaMalloc((void **)p, n * sizeof(T)); }

// CALL sees `D` (observation dimension) and `R` (storage type of the particle store of handle `h`)
#define DISPATCH_P(...) do { if (h->f32) { using R = float; __VA_ARGS__; } else { using R = double; __VA_ARGS__; } } while (0)
#define DISPATCH_D(dd, ...)                                          \
    switch (dd) {                                                    \
    case 1: { constexpr int D = 1; DISPATCH_P(__VA_ARGS__); } break;        \
    case 2: { constexpr int D = 2; DISPATCH_P(__VA_ARGS__); } break;        \
    case 3: { constexpr int D = 3; DISPATCH_P(__VA_ARGS__); } break;        \
    case 4: { constexpr int D = 4; DISPATCH_P(__VA_ARGS__); } break;        \
    case 8: { constexpr int D = 8; DISPATCH_P(__VA_ARGS__); } break;        \
    default: break;                                                  \
    }

static bool supported_d(int d) { return d == 1 || d == 2 || d == 3 || d == 4 || d == 8; }

extern "C" const char *pf_version(void) { return "particles_b200 0.1 (sm_100a)"; }

extern "C" const char *pf_last_error(pf_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int32_t pf_destroy(pf_handle h)
{
    if (!h) return PF_OK;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    void *ptrs[] = {h->S[0], h->S[1], h->logw[0], h->logw[1], h->Kc[0], h->Kc[1], h->V, h->V_alt, h->stc_alt, h->cnt, h->surv_parent,
                    h->surv_slot, h->tiles, h->supers, h->tfuse, h->tsub, h->ttooth, h->cand, h->sc, h->res, h->st, h->pc, h->stc, h->tab, h->slog,
                    h->ys_dev, h->u_dev, h->gen_anc, h->gen_asg, h->gen_off_dev, h->cl, h->W, h->cum, h->asg_tmp, h->cl_tiles, h->clp, h->keys, h->keys2, h->vals, h->vals2, h->hist, h->pick_pos, h->ro, h->xreg, h->xch,
                    h->peers[0], h->peers[1], h->srec.x, h->srec.m, h->srec.h, h->sp[0].N, h->sp[0].Nsame, h->sp[0].last, h->sp[1].N,
                    h->sp[1].Nsame, h->sp[1].last};
    for (void *p : ptrs) if (p) cudaFree(p);
    for (void *p : h->ipc_opened) cudaIpcCloseMemHandle(p);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream && h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
    return PF_OK;
}

static int32_t setup_search_kernels(pf_handle h);

extern "C" int32_t pf_create(const pf_config *cfg, pf_handle *out)
{
    if (!cfg || !out) return set_error(nullptr, PF_ERR_ARG, "pf_create: null argument");
    *out = nullptr;
    if (cfg->filter_kind != PF_FEARNHEAD && cfg->filter_kind != PF_CHENLIU)
        return set_error(nullptr, PF_ERR_ARG, "pf_create: unknown filter_kind");
    if (cfg->prior_kind != PF_NIX2 && cfg->prior_kind != PF_NIW)
        return set_error(nullptr, PF_ERR_ARG, "pf_create: unknown prior_kind");
    int d = cfg->prior_kind == PF_NIX2 ? 1 : cfg->d;
    if (!supported_d(d)) return set_error(nullptr, PF_ERR_ARG, "pf_create: d must be one of 1, 2, 3, 4, 8");
    if (cfg->precision != PF_F64 && cfg->precision != PF_F32) return set_error(nullptr, PF_ERR_ARG, "pf_create: unknown precision");
    if (cfg->n_particles < 1 || cfg->n_particles > (1ll << 27)) return set_error(nullptr, PF_ERR_ARG, "pf_create: n_particles out of range [1, 2^27]");
    if (cfg->k_max < 1 || cfg->k_max + 1 > PF_MAX_SLOTS) return set_error(nullptr, PF_ERR_ARG, "pf_create: k_max out of range [1, 63]");
    if (!(cfg->alpha > 0.0) || !(cfg->kappa0 > 0.0)) return set_error(nullptr, PF_ERR_ARG, "pf_create: alpha and kappa0 must be positive");
    if (!(cfg->nu0 > d - 1.0)) return set_error(nullptr, PF_ERR_ARG, "pf_create: nu0 must exceed d - 1");
    if (!cfg->mu0) return set_error(nullptr, PF_ERR_ARG, "pf_create: mu0 is null");
    if (cfg->order != PF_ORDER_INDEX && cfg->order != PF_ORDER_SORTED) return set_error(nullptr, PF_ERR_ARG, "pf_create: unknown order");
    if (cfg->history < 0 && cfg->nranks > 1) return set_error(nullptr, PF_ERR_ARG, "pf_create: a sharded handle needs a fixed history capacity (>= 0)");
    const int spk = cfg->stateprior_kind;
    if (spk < PF_SP_CRP || spk > PF_SP_NSTATE) return set_error(nullptr, PF_ERR_ARG, "pf_create: unknown stateprior_kind");
    if (spk != PF_SP_CRP && (cfg->filter_kind != PF_FEARNHEAD || cfg->nranks > 1))
        return set_error(nullptr, PF_ERR_ARG, "pf_create: state priors other than the CRP run the single-GPU Fearnhead filter");
    if (spk == PF_SP_STICKY && !(cfg->sp_param >= 0.0 && cfg->sp_param < 1.0)) return set_error(nullptr, PF_ERR_ARG, "pf_create: StickyCRP rho must be in [0, 1)");
    if (spk == PF_SP_STICKY && cfg->k_max + 2 > PF_MAX_SLOTS) return set_error(nullptr, PF_ERR_ARG, "pf_create: StickyCRP needs k_max <= 62");
    if (spk == PF_SP_CHANGEPOINT && !(cfg->sp_param >= 0.0 && cfg->sp_param <= 1.0))
        return set_error(nullptr, PF_ERR_ARG, "pf_create: ChangePoint probability p must be between 0 and 1");      // src/statepriors.jl:169
    if (spk == PF_SP_NSTATE && (!cfg->sp_counts || cfg->sp_n < 1 || cfg->sp_n > cfg->k_max)) return set_error(nullptr, PF_ERR_ARG, "pf_create: NStatePrior needs 1 <= sp_n <= k_max counts");
    if (cfg->nranks > 1) {
        if (cfg->nranks > PF_MAX_RANKS || cfg->rank < 0 || cfg->rank >= cfg->nranks) return set_error(nullptr, PF_ERR_ARG, "pf_create: bad rank / nranks");
        if (cfg->filter_kind == PF_FEARNHEAD && cfg->order != PF_ORDER_INDEX) return set_error(nullptr, PF_ERR_ARG, "pf_create: sharded Fearnhead runs use the index order");
        if (cfg->filter_kind == PF_CHENLIU && cfg->n_particles < cfg->nranks) return set_error(nullptr, PF_ERR_ARG, "pf_create: fewer particles than ranks");
    }

    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0) { cudaGetLastError(); return set_error(nullptr, PF_ERR_CUDA, "pf_create: no CUDA device (this library has no CPU fallback)"); }
    if (cfg->device < 0 || cfg->device >= ndev) return set_error(nullptr, PF_ERR_ARG, "pf_create: bad device ordinal");

    pf_filter_s *h = new pf_filter_s();
    h->cfg = *cfg;
    h->d = d; h->k_max = cfg->k_max; h->N = cfg->n_particles;
    h->nranks = cfg->nranks > 1 ? cfg->nranks : 1;
    h->rank = h->nranks > 1 ? cfg->rank : 0;
    h->cap = (cfg->n_particles + h->nranks - 1) / h->nranks;
    if (h->nranks > 1) h->cap += 64;                      // equal blocks of ceil(n / G) at every step
    h->F = 2 + d + d * (d + 1) / 2;
    h->f32 = cfg->precision == PF_F32;
    h->esz = h->f32 ? sizeof(float) : sizeof(double);
    h->sp_kind = spk;
    if (spk == PF_SP_NSTATE) h->sp_counts.assign(cfg->sp_counts, cfg->sp_counts + cfg->sp_n);
    h->mu0.assign(cfg->mu0, cfg->mu0 + d);
    h->lam0.assign(d * d, 0.0);
    if (cfg->prior_kind == PF_NIX2) {
        if (!(cfg->sigma2_0 > 0.0)) { delete h; return set_error(nullptr, PF_ERR_ARG, "pf_create: sigma2_0 must be positive"); }
        h->lam0[0] = cfg->nu0 * cfg->sigma2_0;        // NIW(d=1) == NIX2, test/niw.jl:5-6
    } else {
        if (!cfg->lambda0) { delete h; return set_error(nullptr, PF_ERR_ARG, "pf_create: lambda0 is null"); }
        h->lam0.assign(cfg->lambda0, cfg->lambda0 + d * d);
    }
    h->cfg.mu0 = nullptr; h->cfg.lambda0 = nullptr; h->cfg.sp_counts = nullptr;
    std::vector<double> L0(d * d);
    if (!chol_lower(d, h->lam0.data(), L0.data())) { delete h; return set_error(nullptr, PF_ERR_ARG, "pf_create: lambda0 is not positive definite"); }

#define CT(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { int32_t rc_ = set_cuda_error(nullptr, e_, #x, __LINE__); g_create_error = h->err.empty() ? g_create_error : h->err; pf_destroy(h); return rc_; } } while (0)
    CT(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CT(cudaGetDeviceProperties(&prop, cfg->device));
    h->num_sms = prop.multiProcessorCount;
    CT(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CT(cudaEventCreate(&h->ev0));
    CT(cudaEventCreate(&h->ev1));
    const long long cap = h->cap;
    const size_t store = (size_t)h->k_max * h->F * cap;
    for (int b = 0; b < 2; b++) {
        CT(cudaMalloc(&h->S[b], store * h->esz));
        CT(dmalloc(&h->logw[b], cap));
        CT(dmalloc(&h->Kc[b], cap));
    }
    CT(dmalloc(&h->V, (size_t)(h->k_max + 2) * cap));
    // single-GPU index-order CRP Fearnhead filter: instantiate of one observation is fused with the expansion of the next
    h->pipeline = cfg->filter_kind == PF_FEARNHEAD && cfg->order == PF_ORDER_INDEX && h->nranks == 1 && spk == PF_SP_CRP &&
                  !getenv("PF_NO_PIPELINE") && !getenv("PF_NO_FUSE");
    h->pipeline_mg = cfg->filter_kind == PF_FEARNHEAD && cfg->order == PF_ORDER_INDEX && h->nranks > 1 && spk == PF_SP_CRP &&
                     !getenv("PF_NO_PIPELINE") && !getenv("PF_NO_FUSE") && !getenv("PF_NO_SCAN_REC");
    if (h->pipeline || h->pipeline_mg) { CT(dmalloc(&h->V_alt, (size_t)(h->k_max + 2) * cap)); CT(dmalloc(&h->stc_alt, 1)); }
    h->Vbuf[0] = h->V; h->Vbuf[1] = h->V_alt; h->vcur = 0;
    CT(dmalloc(&h->cnt, cap));
    h->surv_cap = h->nranks > 1 ? std::min<long long>(h->N, cap * (h->k_max + 1)) + 1024 : cap;
    CT(dmalloc(&h->surv_parent, h->surv_cap));
    CT(dmalloc(&h->surv_slot, h->surv_cap));
    if (h->nranks > 1) {
        CT(dmalloc(&h->ro, 1));
        CT(cudaMemsetAsync(h->ro, 0, sizeof(RankOff), h->stream));
    }
    long long ntiles = (cap + SCAN_THREADS - 1) / SCAN_THREADS;
    if (cfg->filter_kind == PF_FEARNHEAD && cfg->order == PF_ORDER_SORTED) {
        const long long n_pad = cap * (h->k_max + 1);
        ntiles = (n_pad + SCAN_THREADS - 1) / SCAN_THREADS;
        CT(dmalloc(&h->keys, (size_t)n_pad)); CT(dmalloc(&h->keys2, (size_t)n_pad));
        CT(dmalloc(&h->vals, (size_t)n_pad)); CT(dmalloc(&h->vals2, (size_t)n_pad));
        CT(dmalloc(&h->hist, (size_t)(256 + 1) * ((n_pad + RS_TILE - 1) / RS_TILE) + 1024));
        CT(dmalloc(&h->pick_pos, (size_t)cap));
    }
    CT(dmalloc(&h->tiles, ntiles));
    CT(dmalloc(&h->supers, ntiles / SUPER_TILES + 2));
    CT(dmalloc(&h->tfuse, ntiles * EXP_SUB));
    CT(dmalloc(&h->tsub, ntiles * EXP_SUB));
    if (cfg->filter_kind == PF_FEARNHEAD && !getenv("PF_NO_TOOTH_BASES")) CT(dmalloc(&h->ttooth, (cap + SCAN_THREADS - 1) / SCAN_THREADS));
    // per-particle scan records of the fused expansion (24 bytes per particle; PF_NO_SCAN_REC=1 disables them)
    if (cfg->filter_kind == PF_FEARNHEAD && cfg->order == PF_ORDER_INDEX && !getenv("PF_NO_SCAN_REC")) {
        CT(dmalloc(&h->srec.x, (size_t)h->cap));
        CT(dmalloc(&h->srec.m, (size_t)h->cap));
        CT(dmalloc(&h->srec.h, (size_t)h->cap));
    }
    h->fuse_stride = std::max<long long>(1, (h->N * (h->k_max + 1) + 2 * SEL_SAMPLE_TARGET - 1) / (2 * SEL_SAMPLE_TARGET));
    h->fused = getenv("PF_NO_FUSE") == nullptr;
    h->sel_margin = getenv("PF_SEL_MARGIN") ? std::max(1, atoi(getenv("PF_SEL_MARGIN"))) : 8;
    h->dbg_status = getenv("PF_DEBUG_STATUS") ? 1 : 0;
    CT(dmalloc(&h->sc, 1));
    CT(dmalloc(&h->res, 1));
    CT(dmalloc(&h->st, 1));
    CT(dmalloc(&h->pc, 1));
    CT(dmalloc(&h->stc, 1));
    CT(cudaMemsetAsync(h->sc, 0, sizeof(SelScratch), h->stream));
    CT(cudaMemsetAsync(h->res, 0, sizeof(SelResult), h->stream));
    // grids of the classify pass and of the cooperative bracket / solve kernels
    {
        int occ = 0;
        CT(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sel_classify, SEL_THREADS, 0));
        h->sel_grid = h->num_sms * (occ > 2 ? 2 : (occ < 1 ? 1 : occ));
        int32_t rc = setup_search_kernels(h);
        if (rc != PF_OK) { g_create_error = h->err; pf_destroy(h); return rc; }
    }
    h->cand_cap = ((long long)(h->k_max + 1) * cap + (long long)h->sel_grid * (SEL_THREADS / 32) * SEL_CHUNK * 2 + SEL_CHUNK + 2) & ~1ll;
    if (h->nranks > 1) {
        // exchange region: sample segments (a shard's sample is at most its sampled particles x slots) and
        // candidate segments (two round buffers; a widened bracket may hold a few per cent of the shard's putatives)
        h->capA = (((cap + h->fuse_stride - 1) / h->fuse_stride + 8) * (h->k_max + 1) + 1024 + 1) & ~1ll;       // even: segments stay 16-byte aligned
        h->capB = std::max<long long>(1 << 20, 2 * cap) & ~1ll;
        if (getenv("PF_MG_CAPB")) h->capB = std::max<long long>(64, atoll(getenv("PF_MG_CAPB"))) & ~1ll;      // tests: force the overflow path
        if (cfg->filter_kind == PF_CHENLIU) h->capA = h->capB = 2;          // Chen-Liu exchanges scalars only
        h->xreg_bytes = xch_region_bytes(h->nranks, h->capA, h->capB);
        CT(cudaMalloc((void **)&h->xreg, h->xreg_bytes));
        CT(cudaMemsetAsync(h->xreg, 0, xch_header_bytes(), h->stream));
        memset(&h->xch_host, 0, sizeof h->xch_host);
        h->xch_host.rank = h->rank; h->xch_host.nranks = h->nranks; h->xch_host.capA = h->capA; h->xch_host.capB = h->capB;
        h->xch_host.timeout_ns = 1000000ull * (unsigned long long)(getenv("PF_MG_TIMEOUT_MS") ? atoll(getenv("PF_MG_TIMEOUT_MS")) : 20000);
        h->xch_host.reg[h->rank] = h->xreg;
        memset(h->peers_host, 0, sizeof h->peers_host);
    }
    CT(dmalloc(&h->cand, (size_t)h->cand_cap));
    h->hist_auto = cfg->history < 0;
    if (h->hist_auto && h->nranks == 1) {
        // arena mode: budget = a share of what is free now (the filter's own arrays are allocated), PF_GENEALOGY_BYTES overrides
        size_t free_b = 0, total_b = 0;
        CT(cudaMemGetInfo(&free_b, &total_b));
        double bytes = std::min(0.45 * (double)free_b, 64e9);
        if (getenv("PF_GENEALOGY_BYTES")) bytes = atof(getenv("PF_GENEALOGY_BYTES"));
        h->arena = true;
        h->arena_budget = std::max<long long>((long long)(bytes / 5.0), 4 * cap);
        h->hist_cap = 1ll << 40;                          // (rows are not the limit in this mode)
    }
    if (cfg->history > 0) {
        h->hist_cap = cfg->history;
        CT(dmalloc(&h->gen_anc, (size_t)h->hist_cap * cap));
        CT(dmalloc(&h->gen_asg, (size_t)h->hist_cap * cap));
    }
    // prior constants
    PriorConst pc;
    memset(&pc, 0, sizeof pc);
    for (int i = 0; i < d; i++) { pc.mu0[i] = h->mu0[i]; pc.L0_dinv[i] = 1.0 / L0[i + i * d]; }
    for (int i = 1, o = 0; i < d; i++) for (int k = 0; k < i; k++) pc.L0_off[o++] = L0[i + k * d];
    pc.kappa0 = cfg->kappa0; pc.nu0 = cfg->nu0; pc.alpha = cfg->alpha;
    double ld = 0.0;
    for (int i = 0; i < d; i++) ld += std::log(L0[i + i * d]);
    pc.logdet0 = 2.0 * ld;
    h->row0 = make_row(h, 0);
    pc.row0 = h->row0;
    pc.sp_kind = spk; pc.sp_n = spk == PF_SP_NSTATE ? cfg->sp_n : 0;
    pc.sp_log_alpha = std::log(cfg->alpha);
    if (spk == PF_SP_STICKY) pc.sp_logit_rho = std::log(cfg->sp_param / (1.0 - cfg->sp_param));
    if (spk == PF_SP_CHANGEPOINT) {
        const double logp = std::log(cfg->sp_param);
        pc.sp_lp_change = logp;
        pc.sp_lp_stay = logp < -0.6931471805599453 ? std::log1p(-std::exp(logp)) : std::log(-std::expm1(logp));     // StatsFuns.log1mexp
    }
    if (spk == PF_SP_NSTATE) for (int k = 0; k < cfg->sp_n; k++) { pc.sp_counts0[k] = h->sp_counts[k]; pc.sp_total0 += h->sp_counts[k]; }
    CT(cudaMemcpyAsync(h->pc, &pc, sizeof pc, cudaMemcpyHostToDevice, h->stream));
    if (spk == PF_SP_STICKY)
        for (int b = 0; b < 2; b++) {
            CT(dmalloc(&h->sp[b].N, (size_t)h->k_max * cap)); CT(dmalloc(&h->sp[b].Nsame, (size_t)h->k_max * cap)); CT(dmalloc(&h->sp[b].last, (size_t)cap));
            CT(cudaMemsetAsync(h->sp[b].last, 0, sizeof(int) * cap, h->stream));        // StickyCRP(alpha, rho): last = 1 (src/statepriors.jl:102)
        }
    {
        int32_t rc = ensure_table(h, 4096);
        if (rc != PF_OK) { g_create_error = h->err; pf_destroy(h); return rc; }
    }
    // initial population
    StepState st0;
    memset(&st0, 0, sizeof st0);
    if (cfg->filter_kind == PF_FEARNHEAD) {
        st0.n_live = h->rank == 0 ? 1 : 0; st0.n_next = st0.n_live;     // src/fearnhead.jl:29-30 (the particle lives on rank 0)
        h->n_host = 1; h->ub_live = 1;
    } else {
        // src/chenliu.jl:25-26: N empty particles; rank r owns the global particles [r q, (r + 1) q)
        h->shard.N = h->N; h->shard.q = (h->N + h->nranks - 1) / h->nranks; h->shard.goff = (long long)h->rank * h->shard.q;
        h->shard.n_loc = std::max<long long>(0, std::min<long long>(h->N, h->shard.goff + h->shard.q) - h->shard.goff);
        st0.n_live = h->shard.n_loc; st0.n_next = h->shard.n_loc;
        h->n_host = h->shard.n_loc; h->ub_live = h->shard.n_loc;
        CT(dmalloc(&h->cl, 1));
        CT(cudaMemsetAsync(h->cl, 0, sizeof(ClState), h->stream));
        CT(dmalloc(&h->W, cap));
        CT(dmalloc(&h->cum, cap));
        CT(dmalloc(&h->asg_tmp, cap));
        CT(dmalloc(&h->cl_tiles, (size_t)(cap + CL_TILE - 1) / CL_TILE + 1));
        CT(dmalloc(&h->clp, 1));
        memset(&h->clp_host, 0, sizeof h->clp_host);
        if (h->nranks == 1) {
            for (int b = 0; b < 2; b++) { h->clp_host.S[b][0] = h->S[b]; h->clp_host.Kc[b][0] = h->Kc[b]; }
            h->clp_host.cum[0] = h->cum; h->clp_host.asg_tmp[0] = h->asg_tmp;
            CT(cudaMemcpyAsync(h->clp, &h->clp_host, sizeof(ClPeers), cudaMemcpyHostToDevice, h->stream));
        }
    }
    CT(cudaMemcpyAsync(h->st, &st0, sizeof st0, cudaMemcpyHostToDevice, h->stream));
    CT(cudaMemsetAsync(h->logw[0], 0, sizeof(double) * cap, h->stream));   // weight 1.0, src/particle.jl:62
    CT(cudaMemsetAsync(h->Kc[0], 0, sizeof(int) * cap, h->stream));
    if (spk == PF_SP_NSTATE) {
        // the particle starts with sp_n EMPTY components (test/labeled.jl:13-18): n = 0, Lambda_n = Lambda_0
        std::vector<double> slot(h->F, 0.0);
        slot[1] = pc.logdet0;
        for (int k = 0; k < d; k++) slot[2 + d + k] = pc.L0_dinv[k];
        for (int k = 0; k < d * (d - 1) / 2; k++) slot[2 + 2 * d + k] = pc.L0_off[k];
        for (int j = 0; j < cfg->sp_n; j++)
            for (int f = 0; f < h->F; f++)
                {
                    const float sf = (float)slot[f];
                    CT(cudaMemcpyAsync((char *)h->S[0] + ((size_t)j * h->F + f) * cap * h->esz, h->f32 ? (const void *)&sf : (const void *)&slot[f], h->esz,
                                       cudaMemcpyHostToDevice, h->stream));
                    CT(cudaStreamSynchronize(h->stream));
                }
        CT(cudaStreamSynchronize(h->stream));
        const int Kn = cfg->sp_n;
        CT(cudaMemcpyAsync(h->Kc[0], &Kn, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    }
    CT(cudaStreamSynchronize(h->stream));
#undef CT
    *out = h;
    return PF_OK;
}

// ------------------------------------------------------------------ the step sequences
// offset of generation t in the genealogy arrays
static inline size_t gen_row(const pf_filter_s *h, long long t) { return h->arena ? (size_t)h->gen_off[t] : (size_t)t * h->cap; }

static inline int grid_for(long long n, int bs) { long long g = (n + bs - 1) / bs; return (int)(g < 1 ? 1 : g); }

// ---- the threshold search: cooperative launches, argument block, launch sequences
static cudaError_t launch_coop(void (*kern)(SelArgs), int grid, cudaStream_t stream, SelArgs &a)
{
    void *args[] = {&a};
    return cudaLaunchCooperativeKernel((void *)kern, dim3(grid), dim3(SEL_THREADS), args, 0, stream);
}

// the bracket / solve kernels synchronise grid-wide: one resident CTA per SM
static int32_t setup_search_kernels(pf_handle h)
{
    void (*kerns[2])(SelArgs) = {k_sel_bracket, k_sel_solve};
    for (int k = 0; k < 2; k++) {
        int occ = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kerns[k], SEL_THREADS, 0));
        if (occ < 1) return set_error(h, PF_ERR_CUDA, "a threshold-search kernel does not fit on an SM");
    }
    h->coop_grid = h->num_sms;
    return PF_OK;
}

static SelArgs sel_args(pf_handle h, const double *V, const unsigned char *cnt, long long cap, const long long *n_live,
                        const long long *M, const u64 *vmax, long long N, const double *u_dev, double *cand, long long cand_cap,
                        SelScratch *sc, SelResult *res)
{
    SelArgs a;
    memset(&a, 0, sizeof a);
    a.V = V; a.cnt = cnt; a.cap = cap; a.n_live = n_live; a.M = M; a.vmax_bits = vmax; a.N = N; a.u = u_dev;
    a.cand = cand; a.cand_cap = cand_cap; a.sc = sc; a.res = res;
    a.stage = SEL_ST_PUB | SEL_ST_RUN; a.margin = h->sel_margin; a.fixed_stride = h->fuse_stride;
    a.plateau_ptr = (sc == h->sc && h->stc) ? &h->stc->plateau_bits : nullptr;
    a.step = h->steps; a.seq = (unsigned long long)h->steps + 1;
    return a;
}

// [classify] + solve of one round (retry = 0: the round behind the expansion; 1: a queued extra round)
static int32_t launch_round(pf_handle h, SelArgs a, int retry, int solve_stage = SEL_ST_PUB | SEL_ST_RUN)
{
    a.retry = retry;
    k_sel_classify<<<h->sel_grid, SEL_THREADS, 0, h->stream>>>(a);
    a.stage = solve_stage;
    CUDA_TRY(launch_coop(k_sel_solve, h->coop_grid, h->stream, a));
    h->launches += 2;
    return PF_OK;
}

// classic search on an existing V: reset, strided sample, bracket, rounds
static int32_t launch_classic_search(pf_handle h, SelArgs a)
{
    a.external_sample = 0; a.preclassified = 0;
    k_sel_reset<<<1, 32, 0, h->stream>>>(a);
    k_sel_sample<<<h->sel_grid, SEL_THREADS, 0, h->stream>>>(a);
    CUDA_TRY(launch_coop(k_sel_bracket, h->coop_grid, h->stream, a));
    h->launches += 3;
    for (int r = 0; r <= h->retries; r++) {
        int32_t rc = launch_round(h, a, r > 0);
        if (rc != PF_OK) return rc;
    }
    return PF_OK;
}

// y_next / u_next: the next observation of the same call when the step may run its tail pipelined (instantiate fused
// with the next expansion), else null
template <int D, typename R>
static int32_t fearnhead_step(pf_handle h, const double *y_dev, const double *u_dev, StepLog *log_dev,
                              const double *y_next = nullptr, const double *u_next = nullptr, const double *y_host = nullptr)
{
    const long long cap = h->cap;
    const int cur = h->cur, nxt = cur ^ 1;
    const long long ub = h->ub_live;                                     // n_live <= ub
    long long ub_next = ub * (long long)(std::max<long long>(std::min<long long>(h->k_max, h->steps + 1), h->sp_kind == PF_SP_NSTATE ? h->cfg.sp_n : 0) +
                                         (h->sp_kind == PF_SP_STICKY ? 2 : 1));
    if (ub_next > h->N || ub_next < 0) ub_next = h->N;
    unsigned int *ga = nullptr; unsigned char *gs = nullptr;
    if (h->hist_cap) {
        if (h->steps >= h->hist_cap) return set_error(h, PF_ERR_HISTORY, "history capacity exhausted (pf_config.history)");
        ga = h->gen_anc + gen_row(h, h->steps); gs = h->gen_asg + gen_row(h, h->steps);
    }
    h->epoch++;
    const int forced = h->cur_label;
    // other state priors and labeled observations: their own expansion kernel, then the classic search
    const bool sp_path = h->sp_kind != PF_SP_CRP || forced >= 0;
    const bool pre = h->pre_expanded;          // prestep, sample, bracket and expansion ran in the previous step's tail
    h->pre_expanded = false;
    if (!pre) {
        LaunchTimer lt(h, KC_PRESTEP);
        k_prestep<D><<<1, 32, 0, h->stream>>>(h->st, h->pc, y_dev, h->stc, h->redo ? 2 : (h->steps == 0 ? 1 : 0), h->sc, h->res, nullptr, 0, h->steps, forced);
    }
    const bool fused = h->fused && h->cfg.order == PF_ORDER_INDEX && !sp_path;
    if (fused && pre) {
        LaunchTimer lt(h, KC_SELECT, 2 * (h->retries + 1));
        SelArgs a = sel_args(h, h->V, h->cnt, cap, &h->st->n_live, &h->st->M, &h->st->vmax_bits, h->N, u_dev, h->cand, h->cand_cap, h->sc, h->res);
        a.external_sample = 1; a.preclassified = 1;
        const long long l0 = h->launches;
        for (int r = 0; r <= h->retries; r++) {
            int32_t rc = launch_round(h, a, r > 0);
            if (rc != PF_OK) return rc;
        }
        h->launches = l0;
    } else if (fused) {
        // bracket the threshold on a sample expansion, then classify inside the full expansion
        {
            LaunchTimer lt(h, KC_SELECT, 2);
            const long long ns = 4 * ((ub + 4 * h->fuse_stride - 1) / (4 * h->fuse_stride));     // groups of 4 particles
            k_expand<D, 1, false, R><<<grid_for(ns, EXP_SAMPLE_THREADS), EXP_SAMPLE_THREADS, 0, h->stream>>>((const R *)h->S[cur], h->logw[cur], h->Kc[cur], cap, h->k_max, h->tab, h->pc,
                                                                       h->stc, h->st, h->V, h->cnt, h->sc, h->fuse_stride, h->cand, nullptr,
                                                                       nullptr);
            SelArgs a = sel_args(h, h->V, h->cnt, cap, &h->st->n_live, &h->st->M_pred, &h->st->vmax_sample, h->N, u_dev, h->cand, h->cand_cap, h->sc, h->res);
            a.external_sample = 1;
            CUDA_TRY(launch_coop(k_sel_bracket, h->coop_grid, h->stream, a));
        }
        {
            LaunchTimer lt(h, KC_EXPAND);
            k_expand<D, 2, false, R><<<grid_for(ub, EXP_THREADS), EXP_THREADS, 0, h->stream>>>((const R *)h->S[cur], h->logw[cur], h->Kc[cur], cap, h->k_max, h->tab, h->pc,
                                                                       h->stc, h->st, h->V, h->cnt, h->sc, 1, h->cand, h->tsub, h->tfuse, h->srec);
        }
        {
            LaunchTimer lt(h, KC_SELECT, 2 * (h->retries + 1));
            SelArgs a = sel_args(h, h->V, h->cnt, cap, &h->st->n_live, &h->st->M, &h->st->vmax_bits, h->N, u_dev, h->cand, h->cand_cap, h->sc, h->res);
            a.external_sample = 1; a.preclassified = 1;
            const long long l0 = h->launches;
            for (int r = 0; r <= h->retries; r++) {
                int32_t rc = launch_round(h, a, r > 0);
                if (rc != PF_OK) return rc;
            }
            h->launches = l0;                   // (counted by the LaunchTimer)
        }
    } else {
        {
            LaunchTimer lt(h, KC_EXPAND);
            if (sp_path)
                k_expand_sp<D, R><<<grid_for(ub, 256), 256, 0, h->stream>>>((const R *)h->S[cur], h->logw[cur], h->Kc[cur], cap, h->k_max, h->tab, h->pc, h->stc,
                                                                          h->st, h->V, h->cnt, h->sc, h->sp[cur], h->steps);
            else
                k_expand<D, 0, false, R><<<grid_for(ub, EXP_THREADS), EXP_THREADS, 0, h->stream>>>((const R *)h->S[cur], h->logw[cur], h->Kc[cur], cap, h->k_max, h->tab, h->pc,
                                                                           h->stc, h->st, h->V, h->cnt, h->sc, 1, nullptr, nullptr, nullptr);
        }
        {
            LaunchTimer lt(h, KC_SELECT, 3 + 2 * (h->retries + 1));
            SelArgs a = sel_args(h, h->V, h->cnt, cap, &h->st->n_live, &h->st->M, &h->st->vmax_bits, h->N, u_dev, h->cand, h->cand_cap, h->sc, h->res);
            const long long l0 = h->launches;
            int32_t rc = launch_classic_search(h, a);
            if (rc != PF_OK) return rc;
            h->launches = l0;
        }
    }
    if (h->cfg.order == PF_ORDER_SORTED) {
        const int slots = h->k_max + 1;
        const long long n_pad = ub * slots;
        {
            LaunchTimer lt(h, KC_SORT, 1 + 1 + 5 * 8);
            k_flat_keys<<<grid_for(n_pad, 256), 256, 0, h->stream>>>(h->V, h->cnt, cap, h->st, slots, n_pad, h->keys2);
            if (radix_sort_pairs(h->stream, h->keys2, h->keys, h->keys2, h->vals, h->vals2, h->hist, n_pad, nullptr))
                return set_error(h, PF_ERR_CUDA, "radix sort launch failed");
        }
        // resampling step: comb over the sorted list; exhaustive step: index order
        LaunchTimer lt(h, KC_SCAN, 9);
        const int gflat = grid_for(n_pad, SCAN_THREADS), gidx = grid_for(ub, SCAN_THREADS);
        k_tile_sums<<<std::min(gflat, h->num_sms * 8), SCAN_THREADS, 0, h->stream>>>((const double *)h->keys, nullptr, n_pad, h->res, &h->st->M, h->tiles, 1, h->sc, nullptr);
        k_tile_scan1<<<grid_for(gflat, SUPER_TILES), SUPER_TILES, 0, h->stream>>>(h->tiles, nullptr, nullptr, nullptr, h->res, nullptr, &h->st->M, h->stc, h->supers, 1, h->sc);
        k_tile_prefix<<<1, 1024, 0, h->stream>>>(h->supers, h->res, &h->st->M, h->st, 1, nullptr, 0, h->sc, 1);
        k_scan_apply<<<gflat, SCAN_THREADS, 0, h->stream>>>((const double *)h->keys, nullptr, n_pad, h->res, &h->st->M, h->tiles, h->supers,
                                                             h->pick_pos, h->surv_slot, 1, nullptr, h->sc, cap);
        k_sorted_finalize<<<grid_for(ub_next, 256), 256, 0, h->stream>>>(h->st, h->res, h->pick_pos, h->vals, slots, h->surv_parent,
                                                                           h->surv_slot, h->sc);
        k_tile_sums<<<std::min(gidx, h->num_sms * 8), SCAN_THREADS, 0, h->stream>>>(h->V, h->cnt, cap, h->res, &h->st->n_live, h->tiles, 2, h->sc, nullptr);
        k_tile_scan1<<<grid_for(gidx, SUPER_TILES), SUPER_TILES, 0, h->stream>>>(h->tiles, nullptr, nullptr, nullptr, h->res, nullptr, &h->st->n_live, h->stc, h->supers, 2, h->sc);
        k_tile_prefix<<<1, 1024, 0, h->stream>>>(h->supers, h->res, &h->st->n_live, h->st, 2, nullptr, 0, h->sc, 1);
        k_scan_apply<<<gidx, SCAN_THREADS, 0, h->stream>>>(h->V, h->cnt, cap, h->res, &h->st->n_live, h->tiles, h->supers, h->surv_parent,
                                                            h->surv_slot, 2, nullptr, h->sc, cap);
    } else {
        LaunchTimer lt(h, KC_SCAN, h->ttooth ? 5 : 4);
        const int gidx = grid_for(ub, SCAN_THREADS);
        k_tile_sums<<<std::min(gidx, h->num_sms * 8), SCAN_THREADS, 0, h->stream>>>(h->V, h->cnt, cap, h->res, &h->st->n_live, h->tiles, 0, h->sc, fused ? h->sc : nullptr);
        k_tile_scan1<<<grid_for(gidx, SUPER_TILES), SUPER_TILES, 0, h->stream>>>(h->tiles, h->tsub, fused ? h->tfuse : nullptr, h->cand, h->res, h->sc, &h->st->n_live, h->stc,
                                                                                     h->supers, 0, h->sc);
        k_tile_prefix<<<1, 1024, 0, h->stream>>>(h->supers, h->res, &h->st->n_live, h->st, 0, nullptr, 0, h->sc, 1);
        if (h->ttooth) k_tooth_bases<<<grid_for(gidx, 256), 256, 0, h->stream>>>(h->tiles, h->supers, h->res, &h->st->n_live, nullptr, h->ttooth, 0, h->sc);
        k_scan_apply<<<gidx, SCAN_THREADS, 0, h->stream>>>(h->V, h->cnt, cap, h->res, &h->st->n_live, h->tiles, h->supers, h->surv_parent,
                                                            h->surv_slot, 0, nullptr, h->sc, cap, fused ? h->srec : ScanRec{nullptr, nullptr, nullptr}, h->sc, h->ttooth);
    }
    if (fused && y_next && y_host && h->pipeline) {
        // pipelined tail: close this step's record, then build the children WHILE expanding them for the next observation
        {
            LaunchTimer lt(h, KC_STEPLOG);
            k_steplog<<<1, 32, 0, h->stream>>>(h->st, h->res, log_dev, h->sc, h->steps, h->dbg_status, nullptr, nullptr, 0);
        }
        h->steps++;                                   // the host mirrors now describe the next step
        h->ub_live = ub_next;
        h->epoch++;
        {
            LaunchTimer lt(h, KC_PRESTEP, 2);
            k_prestep<D><<<1, 32, 0, h->stream>>>(h->st, h->pc, y_next, h->stc_alt, 3, h->sc, h->res, nullptr, 0, h->steps, -1);
            k_count_next<<<std::min(grid_for(ub_next, 1024), h->num_sms * 16), 256, 0, h->stream>>>(h->st, h->Kc[cur], h->surv_parent, h->surv_slot, h->k_max,
                                                                                                 &h->st->M_pred, h->sc);
        }
        InstArgs ia{};
        ia.surv_parent = h->surv_parent; ia.surv_slot = h->surv_slot; ia.V_prev = h->V; ia.scp_prev = h->stc;
        ia.S2 = h->S[nxt]; ia.Kc2 = h->Kc[nxt]; ia.logw2 = h->logw[nxt]; ia.gen_anc = ga; ia.gen_asg = gs;
        for (int k = 0; k < h->d; k++) { ia.y_prev[k] = y_host[k]; ia.yc[k] = y_host[h->d + k]; ia.mu0c[k] = h->mu0[k]; }
        {
            LaunchTimer lt(h, KC_SELECT, 2);
            const long long ns = 4 * ((ub_next + 4 * h->fuse_stride - 1) / (4 * h->fuse_stride));
            k_expand<D, 1, true, R><<<grid_for(ns, EXP_SAMPLE_THREADS), EXP_SAMPLE_THREADS, 0, h->stream>>>((const R *)h->S[cur], nullptr, h->Kc[cur], cap, h->k_max, h->tab, h->pc,
                                                                       h->stc_alt, h->st, h->V_alt, h->cnt, h->sc, h->fuse_stride, h->cand, nullptr,
                                                                       nullptr, ScanRec{nullptr, nullptr, nullptr}, ia);
            std::swap(h->stc, h->stc_alt);            // (sel_args points the search at the next step's plateau value)
            SelArgs a = sel_args(h, h->V_alt, h->cnt, cap, &h->st->n_live, &h->st->M_pred, &h->st->vmax_sample, h->N, u_next, h->cand, h->cand_cap, h->sc, h->res);
            std::swap(h->stc, h->stc_alt);
            a.external_sample = 1;
            CUDA_TRY(launch_coop(k_sel_bracket, h->coop_grid, h->stream, a));
        }
        {
            LaunchTimer lt(h, KC_INST_EXPAND);
            constexpr size_t stage_bytes = instexp_staging_bytes<D, R>();      // asynchronous-copy staging of the slot loop
            if (stage_bytes && !h->instexp_attr) {
                CUDA_TRY(cudaFuncSetAttribute(k_expand<D, 2, true, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
                h->instexp_attr = true;
            }
            k_expand<D, 2, true, R><<<grid_for(ub_next, EXP_THREADS), EXP_THREADS, stage_bytes, h->stream>>>((const R *)h->S[cur], nullptr, h->Kc[cur], cap, h->k_max, h->tab, h->pc,
                                                                       h->stc_alt, h->st, h->V_alt, h->cnt, h->sc, 1, h->cand, h->tsub, h->tfuse, h->srec, ia);
        }
        CUDA_TRY(cudaGetLastError());
        std::swap(h->V, h->V_alt); h->vcur ^= 1;
        std::swap(h->stc, h->stc_alt);
        h->cur = nxt;
        h->pre_expanded = true;
        return PF_OK;
    }
    {
        LaunchTimer lt(h, KC_INSTANTIATE);
        k_instantiate<D, R><<<grid_for(ub_next, 256), 256, 0, h->stream>>>((const R *)h->S[cur], h->Kc[cur], (R *)h->S[nxt], h->Kc[nxt], h->logw[nxt], cap,
                                                                          h->tab, h->pc, h->stc, h->st, h->res, h->V, h->surv_parent,
                                                                          h->surv_slot, ga, gs, nullptr, h->k_max, h->sc, &h->st->M_next, nullptr, 0,
                                                                          h->sp[cur], h->sp[nxt]);
    }
    {
        LaunchTimer lt(h, KC_STEPLOG);
        k_steplog<<<1, 32, 0, h->stream>>>(h->st, h->res, log_dev, h->sc, h->steps, h->dbg_status, nullptr, nullptr, 0);
    }
    CUDA_TRY(cudaGetLastError());
    h->cur = nxt;
    h->ub_live = ub_next;
    h->steps++;
    return PF_OK;
}

// One Chen-Liu observation as six stages; the exchange kernel that closes stage s publishes this shard's scalar and
// (unless the shards of this process run in lock step on one stream) also waits for the peers'; in lock step the wait
// is the first launch of stage s + 1, queued after every shard has published.  Single GPU: stages 0..5 back to back.
template <int D, typename R>
static int32_t chenliu_stage(pf_handle h, int s, const double *y_dev, const double *u_dev /* [2 n_loc] or null */, StepLog *log_dev)
{
    const long long cap = h->cap, N = h->N, n_loc = h->shard.n_loc;
    const Xch *x = h->nranks > 1 ? h->xch : nullptr;
    const unsigned long long seq = (unsigned long long)h->steps + 1;
    const bool split = h->lockstep;
    const int pub = split ? SEL_ST_PUB : (SEL_ST_PUB | SEL_ST_RUN);
    const int g256 = grid_for(n_loc, 256), gtile = grid_for(n_loc, CL_TILE);
    int gstat = h->num_sms * 2;
    if (gstat > 1024) gstat = 1024;
    switch (s) {
    case 0: {
        if (h->hist_cap && h->steps >= h->hist_cap) return set_error(h, PF_ERR_HISTORY, "history capacity exhausted (pf_config.history)");
        {
            LaunchTimer lt(h, KC_PRESTEP);            // (sharded: waits until no peer reads this shard's store any more)
            k_prestep<D><<<1, 32, 0, h->stream>>>(h->st, h->pc, y_dev, h->stc, 1, h->sc, nullptr, x, seq, h->steps);
        }
        LaunchTimer lt(h, KC_CL_PROP);
        k_cl_propagate<D, R><<<g256, 256, 0, h->stream>>>((R *)h->S[0], (R *)h->S[1], h->logw[0], h->logw[1], h->Kc[0], h->Kc[1], cap, h->shard, h->k_max,
                                                         h->tab, h->pc, h->stc, h->st, h->cl, u_dev, h->cfg.seed, (u64)h->steps, h->W,
                                                         h->asg_tmp, h->V, h->sc);
        if (x) { k_cl_x_wmax<<<1, 32, 0, h->stream>>>(h->cl, x, seq, h->steps, pub, h->sc); h->launches++; }
        break;
    }
    case 1: {
        LaunchTimer lt(h, KC_CL_STATS, 2);
        if (x && split) { k_cl_x_wmax<<<1, 32, 0, h->stream>>>(h->cl, x, seq, h->steps, SEL_ST_RUN, h->sc); h->launches++; }
        k_cl_sum<<<gstat, 256, 0, h->stream>>>(h->W, n_loc, h->cl, h->sc);
        k_cl_x_sum<<<1, 32, 0, h->stream>>>(h->cl, x, seq, h->steps, pub, N, h->sc);
        break;
    }
    case 2: {
        LaunchTimer lt(h, KC_CL_STATS, 2);
        if (x && split) { k_cl_x_sum<<<1, 32, 0, h->stream>>>(h->cl, x, seq, h->steps, SEL_ST_RUN, N, h->sc); h->launches++; }
        k_cl_var<<<gstat, 256, 0, h->stream>>>(h->W, n_loc, h->cl, h->sc);
        k_cl_decide<<<1, 32, 0, h->stream>>>(h->cl, x, seq, h->steps, pub, N, h->cfg.rejuv_threshold, h->sc);
        break;
    }
    case 3: {
        LaunchTimer lt(h, KC_CL_STATS, 2);
        if (x && split) { k_cl_decide<<<1, 32, 0, h->stream>>>(h->cl, x, seq, h->steps, SEL_ST_RUN, N, h->cfg.rejuv_threshold, h->sc); h->launches++; }
        k_cl_tile_sums<<<gtile, 256, 0, h->stream>>>(h->W, n_loc, h->cl, h->cl_tiles, h->sc);
        k_cl_tile_prefix<<<1, 1024, 0, h->stream>>>(h->cl_tiles, gtile, h->cl, x, seq, h->steps, pub, h->sc);
        break;
    }
    case 4: {
        LaunchTimer lt(h, KC_CL_STATS, 1);
        if (x && split) { k_cl_tile_prefix<<<1, 1024, 0, h->stream>>>(h->cl_tiles, gtile, h->cl, x, seq, h->steps, SEL_ST_RUN, h->sc); h->launches++; }
        k_cl_tile_apply<<<gtile, 256, 0, h->stream>>>(h->W, n_loc, h->cl, h->cl_tiles, h->cum, h->sc);
        if (x) { k_cl_x_ready<<<1, 32, 0, h->stream>>>(h->cl, x, seq, h->steps, pub, h->sc); h->launches++; }
        break;
    }
    default: {
        unsigned int *ga = nullptr; unsigned char *gs = nullptr;
        if (h->hist_cap) { ga = h->gen_anc + gen_row(h, h->steps); gs = h->gen_asg + gen_row(h, h->steps); }
        {
            LaunchTimer lt(h, KC_CL_FINISH);
            if (x && split) { k_cl_x_ready<<<1, 32, 0, h->stream>>>(h->cl, x, seq, h->steps, SEL_ST_RUN, h->sc); h->launches++; }
            k_cl_finish<D, R><<<g256, 256, 0, h->stream>>>((R *)h->S[0], (R *)h->S[1], h->Kc[0], h->Kc[1], h->logw[0], h->logw[1], cap, h->shard, h->nranks, h->cl,
                                                          h->W, h->clp, u_dev ? u_dev + n_loc : nullptr, h->cfg.seed, (u64)h->steps, h->asg_tmp, ga, gs, h->sc);
        }
        {
            LaunchTimer lt(h, KC_STEPLOG);
            k_cl_steplog<<<1, 32, 0, h->stream>>>(h->st, h->cl, log_dev, N, n_loc, x, seq, h->sc);
        }
        CUDA_TRY(cudaGetLastError());
        h->steps++;
        break;
    }
    }
    return PF_OK;
}

template <int D, typename R>
static int32_t chenliu_step(pf_handle h, const double *y_dev, const double *u_dev /* 2N or null */, StepLog *log_dev)
{
    for (int s = 0; s < 6; s++) {
        int32_t rc = chenliu_stage<D, R>(h, s, y_dev, u_dev, log_dev);
        if (rc != PF_OK) return rc;
    }
    return PF_OK;
}

struct DevBuf {                  // scoped device allocation
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <typename T> T *as() { return static_cast<T *>(p); }
};

// device copy of the generations' offsets for the backtrack kernels (null: dense log)
static int32_t upload_gen_off(pf_handle h, const long long **out)
{
    *out = nullptr;
    if (!h->arena) return PF_OK;
    const long long T = h->steps;
    if (T > h->gen_off_dev_cap) {
        if (h->gen_off_dev) cudaFree(h->gen_off_dev);
        h->gen_off_dev = nullptr; h->gen_off_dev_cap = 0;
        CUDA_TRY(dmalloc(&h->gen_off_dev, (size_t)std::max<long long>(2 * T, 64)));
        h->gen_off_dev_cap = std::max<long long>(2 * T, 64);
    }
    CUDA_TRY(cudaMemcpyAsync(h->gen_off_dev, h->gen_off.data(), sizeof(long long) * T, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));        // (gen_off may be edited by the host right after)
    *out = h->gen_off_dev;
    return PF_OK;
}

// Arena mode: drop the nodes that are no longer an ancestor of the current population (summaries.cuh, "genealogy
// pruning") and pack the generations.  Newest generation first; all lengths the kernels need that changed in this pass
// are read on the device (n_dev), the host learns them once at the end.
static int32_t prune_genealogy(pf_handle h)
{
    const long long G = h->steps;
    if (!h->arena || G < 2) return PF_OK;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    long long maxn = 1;
    for (long long g = 0; g < G; g++) maxn = std::max(maxn, h->gen_n[g]);
    DevBuf dmark, didx, dpart, dtanc, dtasg, dn;
    CUDA_TRY(dmark.alloc(sizeof(unsigned int) * maxn));
    CUDA_TRY(didx.alloc(sizeof(unsigned int) * maxn));
    CUDA_TRY(dpart.alloc(sizeof(unsigned int) * (maxn / RS_SCAN_TILE + 1024)));
    CUDA_TRY(dtanc.alloc(sizeof(unsigned int) * maxn));
    CUDA_TRY(dtasg.alloc((size_t)maxn));
    CUDA_TRY(dn.alloc(sizeof(long long) * G));
    CUDA_TRY(cudaMemcpyAsync(dn.p, h->gen_n.data(), sizeof(long long) * G, cudaMemcpyHostToDevice, h->stream));
    long long *n_dev = dn.as<long long>();
    unsigned int *mark = dmark.as<unsigned int>(), *idx = didx.as<unsigned int>(), *part = dpart.as<unsigned int>();
    const int gmax = h->num_sms * 8;
    DevBuf doff;
    CUDA_TRY(doff.alloc(sizeof(long long) * G));
    CUDA_TRY(cudaMemcpyAsync(doff.p, h->gen_off.data(), sizeof(long long) * G, cudaMemcpyHostToDevice, h->stream));
    for (long long g = G - 1; g >= 1; g--) {
        const long long p = g - 1, np_old = h->gen_n[p], ng_old = h->gen_n[g];
        if (ng_old <= GEN_SMALL) {
            // the stored nodes only get fewer with age: from here on every generation is short -- one launch for the rest
            bool monotone = true;
            for (long long q = 0; q < g && monotone; q++) monotone = h->gen_n[q] <= GEN_SMALL;
            if (monotone) {
                k_gen_prune_small<<<1, 1024, 0, h->stream>>>(h->gen_anc, h->gen_asg, doff.as<long long>(), n_dev, g);
                break;
            }
        }
        if (np_old == 0 || ng_old == 0) continue;
        unsigned int *anc_g = h->gen_anc + h->gen_off[g];
        unsigned int *anc_p = h->gen_anc + h->gen_off[p];
        unsigned char *asg_p = h->gen_asg + h->gen_off[p];
        CUDA_TRY(cudaMemsetAsync(mark, 0, sizeof(unsigned int) * np_old, h->stream));
        k_gen_mark<<<std::min(grid_for(ng_old, 256), gmax), 256, 0, h->stream>>>(anc_g, n_dev + g, mark);
        CUDA_TRY(cudaMemcpyAsync(idx, mark, sizeof(unsigned int) * np_old, cudaMemcpyDeviceToDevice, h->stream));
        const long long nb = (np_old + RS_SCAN_TILE - 1) / RS_SCAN_TILE;
        k_rs_scan_sums<<<(unsigned)nb, 1024, 0, h->stream>>>(idx, np_old, part);
        k_rs_scan_top<<<1, 1024, 0, h->stream>>>(part, nb);
        k_rs_scan_apply<<<(unsigned)nb, 1024, 0, h->stream>>>(idx, np_old, part);
        k_gen_count<<<1, 32, 0, h->stream>>>(mark, idx, np_old, n_dev + p);
        k_gen_compact<<<std::min(grid_for(np_old, 256), gmax), 256, 0, h->stream>>>(anc_p, asg_p, np_old, mark, idx, dtanc.as<unsigned int>(), dtasg.as<unsigned char>());
        k_gen_copy<<<std::min(grid_for(np_old, 256), gmax), 256, 0, h->stream>>>(dtanc.as<unsigned int>(), dtasg.as<unsigned char>(), n_dev + p, anc_p, asg_p);
        k_gen_rewrite<<<std::min(grid_for(ng_old, 256), gmax), 256, 0, h->stream>>>(anc_g, n_dev + g, idx);
    }
    std::vector<long long> n_new(G);
    CUDA_TRY(cudaMemcpyAsync(n_new.data(), dn.p, sizeof(long long) * G, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    // pack: oldest first, every generation moves towards the start of the arena (through the scratch when it overlaps)
    long long run = 0;
    for (long long g = 0; g < G; g++) {
        const long long n = n_new[g], src = h->gen_off[g];
        if (src != run && n > 0) {
            if (run + n <= src) {
                CUDA_TRY(cudaMemcpyAsync(h->gen_anc + run, h->gen_anc + src, sizeof(unsigned int) * n, cudaMemcpyDeviceToDevice, h->stream));
                CUDA_TRY(cudaMemcpyAsync(h->gen_asg + run, h->gen_asg + src, (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
            } else {
                CUDA_TRY(cudaMemcpyAsync(dtanc.p, h->gen_anc + src, sizeof(unsigned int) * n, cudaMemcpyDeviceToDevice, h->stream));
                CUDA_TRY(cudaMemcpyAsync(dtasg.p, h->gen_asg + src, (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
                CUDA_TRY(cudaMemcpyAsync(h->gen_anc + run, dtanc.p, sizeof(unsigned int) * n, cudaMemcpyDeviceToDevice, h->stream));
                CUDA_TRY(cudaMemcpyAsync(h->gen_asg + run, dtasg.p, (size_t)n, cudaMemcpyDeviceToDevice, h->stream));
            }
        }
        h->gen_off[g] = run; h->gen_n[g] = n;
        run += n;
    }
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    h->arena_used = run;
    h->n_prunes++; h->last_prune_step = G;
    return PF_OK;
}

// arena: room for `want` more entries behind arena_used (doubling up to the budget, then exactly what is needed)
static int32_t arena_reserve(pf_handle h, long long want)
{
    if (h->arena_used + want <= h->arena_cap) return PF_OK;
    long long ncap = std::max<long long>(h->arena_used + want, std::min<long long>(2 * h->arena_cap, h->arena_budget));
    ncap = std::max<long long>(ncap, 16 * h->cap);
    unsigned int *na = nullptr; unsigned char *ns = nullptr;
    cudaError_t e = dmalloc(&na, (size_t)ncap);
    if (e == cudaSuccess) e = dmalloc(&ns, (size_t)ncap);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (na) cudaFree(na);
        return set_error(h, PF_ERR_HISTORY, "the genealogy no longer fits in device memory: create the filter with history = 0 (none) or a fixed capacity");
    }
    if (h->arena_used > 0) {
        CUDA_TRY(cudaMemcpyAsync(na, h->gen_anc, sizeof(unsigned int) * (size_t)h->arena_used, cudaMemcpyDeviceToDevice, h->stream));
        CUDA_TRY(cudaMemcpyAsync(ns, h->gen_asg, (size_t)h->arena_used, cudaMemcpyDeviceToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    if (h->gen_anc) cudaFree(h->gen_anc);
    if (h->gen_asg) cudaFree(h->gen_asg);
    h->gen_anc = na; h->gen_asg = ns; h->arena_cap = ncap;
    return PF_OK;
}

// history < 0: make room for the generations [steps, need).  Dense (sharded) log: rows double, the old generations are
// copied over.  Arena: prune the dead lineages when the arena is full, grow it (up to the budget) when that is not enough.
static int32_t ensure_history(pf_handle h, long long need)
{
    if (!h->hist_auto) return PF_OK;
    if (h->arena) {
        const long long rows = need - h->steps;
        if (rows <= 0) return PF_OK;
        const long long want = rows * h->cap;
        // (a pruning pass walks every stored generation: passes are spaced geometrically -- at least steps / 8 observations
        // apart -- so that their total cost stays O(T log T) whatever the budget; in between, the arena grows instead)
        if (h->arena_used + want > h->arena_cap && h->steps - h->last_prune_step >= std::max<long long>(4, h->steps / 8)) {
            int32_t rc = prune_genealogy(h);
            if (rc != PF_OK) return rc;
        }
        int32_t rc2 = arena_reserve(h, want);
        if (rc2 != PF_OK) return rc2;
        h->gen_off.resize(need); h->gen_n.resize(need);
        for (long long t = h->steps; t < need; t++) { h->gen_off[t] = h->arena_used; h->gen_n[t] = h->cap; h->arena_used += h->cap; }
        return PF_OK;
    }
    if (need <= h->hist_cap) return PF_OK;
    long long ncap = std::max<long long>(std::max<long long>(need, 2 * h->hist_cap), 16);
    unsigned int *na = nullptr; unsigned char *ns = nullptr;
    cudaError_t e = dmalloc(&na, (size_t)ncap * h->cap);
    if (e == cudaSuccess) e = dmalloc(&ns, (size_t)ncap * h->cap);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (na) cudaFree(na);
        return set_error(h, PF_ERR_HISTORY, "the genealogy no longer fits in device memory: create the filter with history = 0 (none) or a fixed capacity");
    }
    if (h->steps > 0 && h->gen_anc) {
        CUDA_TRY(cudaMemcpyAsync(na, h->gen_anc, sizeof(unsigned int) * (size_t)h->steps * h->cap, cudaMemcpyDeviceToDevice, h->stream));
        CUDA_TRY(cudaMemcpyAsync(ns, h->gen_asg, (size_t)h->steps * h->cap, cudaMemcpyDeviceToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    if (h->gen_anc) cudaFree(h->gen_anc);
    if (h->gen_asg) cudaFree(h->gen_asg);
    h->gen_anc = na; h->gen_asg = ns; h->hist_cap = ncap;
    return PF_OK;
}

static int32_t run_steps(pf_handle h, const double *ys, long long T, const double *uniforms, long long ups, double *log_evidence,
                         const int32_t *labels = nullptr)
{
    if (!h) return PF_ERR_ARG;
    if (T < 0 || (!ys && T > 0)) return set_error(h, PF_ERR_ARG, "pf_filter: bad ys/T");
    if (T == 0) return PF_OK;
    const bool fh = h->cfg.filter_kind == PF_FEARNHEAD;
    if (labels) {
        if (!fh || h->nranks > 1) return set_error(h, PF_ERR_ARG, "labeled observations run the single-GPU Fearnhead filter");
        for (long long t = 0; t < T; t++)
            if (labels[t] < -1 || labels[t] >= h->k_max) return set_error(h, PF_ERR_ARG, "label out of range (0-based component, or -1 for missing)");
    }
    if (uniforms) {
        long long need = fh ? 1 : 2 * h->N;
        if (ups < need) return set_error(h, PF_ERR_ARG, fh ? "pf_filter: Fearnhead needs 1 uniform per step" : "pf_filter: Chen-Liu needs 2*N uniforms per step");
    }
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    int32_t rc = ensure_table(h, h->steps + T + 2);
    if (rc != PF_OK) return rc;
    rc = ensure_history(h, h->steps + T);
    if (rc != PF_OK) return rc;
    const int d = h->d;
    if (T * d > h->ys_cap) {
        if (h->ys_dev) cudaFree(h->ys_dev);
        h->ys_cap = T * d;
        CUDA_TRY(dmalloc(&h->ys_dev, (size_t)h->ys_cap));
    }
    if (T > h->slog_cap) {
        if (h->slog) cudaFree(h->slog);
        h->slog_cap = T;
        CUDA_TRY(dmalloc(&h->slog, (size_t)T));
    }
    // uniforms: Fearnhead T doubles; Chen-Liu is staged step by step (2N per step)
    long long un = fh ? T : 2 * h->N;
    if (un > h->u_cap) {
        if (h->u_dev) cudaFree(h->u_dev);
        h->u_cap = un;
        CUDA_TRY(dmalloc(&h->u_dev, (size_t)un));
    }
    h->launches = 0;
    CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->ys_dev, ys, sizeof(double) * T * d, cudaMemcpyHostToDevice, h->stream));
    if (fh) {
        std::vector<double> ubuf(T);
        for (long long t = 0; t < T; t++)
            ubuf[t] = uniforms ? uniforms[t * ups] : pf_uniform(h->cfg.seed, (uint64_t)(h->steps + t), 0, 0);
        CUDA_TRY(cudaMemcpyAsync(h->u_dev, ubuf.data(), sizeof(double) * T, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));      // ubuf goes out of scope
        CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    }
    // Every step of the call is queued without a host round trip.  A Fearnhead step whose threshold search
    // needed more rounds than were queued halts the handle on the device (later kernels are no-ops); the
    // host then runs that step again with more rounds queued and resumes behind it.
    const long long steps0 = h->steps;
    const int cur0 = h->cur;
    std::vector<long long> ub_hist(T + 1, h->ub_live);
    long long t_begin = 0;
    for (int attempt = 0; ; attempt++) {
        for (long long t = t_begin; t < T; t++) {
            ub_hist[t] = h->ub_live;
            if (fh) {
                h->cur_label = labels ? labels[t] : -1;
                // the tail of step t may be fused with the expansion of step t + 1 (both ordinary CRP steps of this call,
                // genealogy row available)
                const bool next_ok = h->pipeline && t + 1 < T && !(labels && (labels[t] >= 0 || labels[t + 1] >= 0)) &&
                                     (!h->hist_cap || h->steps + 1 < h->hist_cap);
                DISPATCH_D(d, rc = fearnhead_step<D, R>(h, h->ys_dev + t * d, h->u_dev + t, h->slog + t,
                                                     next_ok ? h->ys_dev + (t + 1) * d : nullptr, next_ok ? h->u_dev + t + 1 : nullptr, ys + t * d));
                h->redo = 0; h->retries = SEL_RETRIES; h->cur_label = -1;
            } else {
                const double *ud = nullptr;
                if (uniforms) {
                    CUDA_TRY(cudaMemcpyAsync(h->u_dev, uniforms + t * ups, sizeof(double) * 2 * h->N, cudaMemcpyHostToDevice, h->stream));
                    ud = h->u_dev;
                }
                DISPATCH_D(d, rc = chenliu_step<D, R>(h, h->ys_dev + t * d, ud, h->slog + t));
            }
            if (rc != PF_OK) {
                // an error while queuing observation t (e.g. the fixed history is exhausted): the observations queued before
                // it have run; bring the host mirrors in line with the device before reporting
                cudaStreamSynchronize(h->stream);
                const long long done = h->steps - steps0;
                if (done > 0) {
                    h->logs.resize(done);
                    if (cudaMemcpy(h->logs.data(), h->slog, sizeof(StepLog) * done, cudaMemcpyDeviceToHost) == cudaSuccess) {
                        h->last = h->logs.back(); h->have_last = true; h->n_host = h->last.n_out;
                        if (h->arena) for (long long k = 0; k < done; k++) h->gen_n[steps0 + k] = h->logs[k].n_out;
                        if (log_evidence) for (long long k = 0; k < done; k++) log_evidence[k] = h->logs[k].log_total;
                    }
                    if (!fh) { ClState cs; if (cudaMemcpy(&cs, h->cl, sizeof cs, cudaMemcpyDeviceToHost) == cudaSuccess) h->cur = cs.cur; }
                }
                return rc;
            }
        }
        long long halt[2] = {0, 0};
        if (fh) CUDA_TRY(cudaMemcpyAsync(halt, &h->sc->halt, sizeof halt, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        if (!halt[0]) break;
        const long long th = halt[0] - 1 - steps0;          // index of the halted step within this call
        const int reason = (int)(halt[1] & 0xffffffffll);
        if (reason == PF_HALT_LABEL) {
            // (the handle stays halted at that observation: like the reference's ArgumentError, the filter is not usable further)
            char buf[160];
            snprintf(buf, sizeof buf, "can't fit the labelled component of observation %lld: it is not a candidate state of every particle", halt[0] - 1);
            return set_error(h, PF_ERR_ARG, buf);           // ArgumentError, src/particle.jl:137-139
        }
        if (th < 0 || th >= T || reason != PF_HALT_SEARCH || attempt >= 3) {
            char buf[160];
            snprintf(buf, sizeof buf, "step %lld could not be completed on the device (reason %d, attempt %d)", halt[0] - 1, reason, attempt);
            return set_error(h, PF_ERR_STATE, buf);
        }
        // rewind the host mirrors to the halted step and run it again with more rounds queued
        h->steps = steps0 + th; h->cur = cur0 ^ (int)(th & 1); h->ub_live = ub_hist[th];
        h->redo = 1; h->retries = 12; h->pre_expanded = false;       // the halted step is expanded again from the store
        CUDA_TRY(cudaMemsetAsync(&h->sc->halt, 0, 2 * sizeof(long long), h->stream));
        t_begin = th;
    }
    h->logs.resize(T);
    CUDA_TRY(cudaMemcpyAsync(h->logs.data(), h->slog, sizeof(StepLog) * T, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    h->last_ms = ms; h->last_launches = h->launches;
    collect_profile(h);
    h->last = h->logs.back(); h->have_last = true;
    h->n_host = h->last.n_out;
    if (h->arena) {
        // the generations of this call were appended at a stride of `cap`; their true lengths are known now, and what the
        // last one left unused goes back to the arena
        for (long long t = 0; t < T; t++) h->gen_n[steps0 + t] = h->logs[t].n_out;
        h->arena_used = h->gen_off[steps0 + T - 1] + h->gen_n[steps0 + T - 1];
    }
    if (!fh) {
        ClState cs;
        CUDA_TRY(cudaMemcpy(&cs, h->cl, sizeof cs, cudaMemcpyDeviceToHost));
        h->cur = cs.cur;
    }
    if (log_evidence) for (long long t = 0; t < T; t++) log_evidence[t] = h->logs[t].log_total;
    return PF_OK;
}

extern "C" int32_t pf_step(pf_handle h, const double *y, const double *uniforms, int64_t n_uniforms)
{
    return run_steps(h, y, 1, uniforms, n_uniforms, nullptr);
}

// Arena mode: a call's generations are appended densely before they can be pruned, so a long stream is filtered in
// chunks whose rows fit in half the genealogy budget (the pipelining of the steps only breaks at the chunk boundaries)
static int32_t run_chunked(pf_handle h, const double *ys, long long T, const double *uniforms, long long ups, double *log_evidence,
                           const int32_t *labels)
{
    if (!h) return PF_ERR_ARG;
    if (!h->arena || T <= 0) return run_steps(h, ys, T, uniforms, ups, log_evidence, labels);
    const long long chunk = std::max<long long>(4, h->arena_budget / (2 * h->cap));
    for (long long t0 = 0; t0 < T; t0 += chunk) {
        const long long n = std::min(chunk, T - t0);
        int32_t rc = run_steps(h, ys + t0 * h->d, n, uniforms ? uniforms + t0 * ups : nullptr, ups, log_evidence ? log_evidence + t0 : nullptr,
                               labels ? labels + t0 : nullptr);
        if (rc != PF_OK) return rc;
    }
    return PF_OK;
}

extern "C" int32_t pf_filter(pf_handle h, const double *ys, int64_t T, const double *uniforms, int64_t uniforms_per_step,
                             double *log_evidence)
{
    return run_chunked(h, ys, T, uniforms, uniforms_per_step, log_evidence, nullptr);
}

extern "C" int32_t pf_filter_labeled(pf_handle h, const double *ys, const int32_t *labels, int64_t T, const double *uniforms,
                                     int64_t uniforms_per_step, double *log_evidence)
{
    return run_chunked(h, ys, T, uniforms, uniforms_per_step, log_evidence, labels);
}

extern "C" int32_t pf_get_stateprior(pf_handle h, int64_t i, double *N, double *Nsame, int32_t *last)
{
    if (!h) return PF_ERR_ARG;
    if (h->sp_kind != PF_SP_STICKY) return set_error(h, PF_ERR_STATE, "pf_get_stateprior: only the StickyCRP keeps counts of its own");
    if (i < 0 || i >= h->n_host) return set_error(h, PF_ERR_ARG, "pf_get_stateprior: index out of range");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    int K = 0, l = 0;
    CUDA_TRY(cudaMemcpyAsync(&K, h->Kc[h->cur] + i, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaMemcpyAsync(&l, h->sp[h->cur].last + i, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (K > 0) {
        if (N) CUDA_TRY(cudaMemcpy2DAsync(N, sizeof(double), h->sp[h->cur].N + i, sizeof(double) * h->cap, sizeof(double), (size_t)K, cudaMemcpyDeviceToHost, h->stream));
        if (Nsame) CUDA_TRY(cudaMemcpy2DAsync(Nsame, sizeof(double), h->sp[h->cur].Nsame + i, sizeof(double) * h->cap, sizeof(double), (size_t)K, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    if (last) *last = l;
    return PF_OK;
}

// ------------------------------------------------------------------ read-out
extern "C" int64_t pf_n_particles(pf_handle h) { return h ? h->n_host : -1; }
extern "C" int64_t pf_n_steps(pf_handle h) { return h ? h->steps : -1; }
extern "C" void *pf_stream(pf_handle h) { return h ? (void *)h->stream : nullptr; }

extern "C" int32_t pf_last_timing(pf_handle h, double *device_ms, int64_t *kernel_launches)
{
    if (!h) return PF_ERR_ARG;
    if (device_ms) *device_ms = h->last_ms;
    if (kernel_launches) *kernel_launches = h->nranks > 1 ? h->launches : h->last_launches;
    return PF_OK;
}

extern "C" int32_t pf_set_profiling(pf_handle h, int32_t on)
{
    if (!h) return PF_ERR_ARG;
    h->profiling = on != 0;
    for (int k = 0; k < PF_KERNEL_CLASSES; k++) { h->cls_ms[k] = 0.0; h->cls_launches[k] = 0; }
    return PF_OK;
}

extern "C" int32_t pf_get_kernel_times(pf_handle h, double *ms, int64_t *launches)
{
    if (!h) return PF_ERR_ARG;
    for (int k = 0; k < PF_KERNEL_CLASSES; k++) { if (ms) ms[k] = h->cls_ms[k]; if (launches) launches[k] = h->cls_launches[k]; }
    return PF_OK;
}

extern "C" const char *pf_kernel_class_name(int32_t k) { return (k >= 0 && k < PF_KERNEL_CLASSES) ? k_class_names[k] : ""; }

extern "C" int32_t pf_get_logweights(pf_handle h, double *out)
{
    if (!h || !out) return PF_ERR_ARG;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaMemcpyAsync(out, h->logw[h->cur], sizeof(double) * h->n_host, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PF_OK;
}

extern "C" int32_t pf_get_ncomponents(pf_handle h, int32_t *out)
{
    if (!h || !out) return PF_ERR_ARG;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaMemcpyAsync(out, h->Kc[h->cur], sizeof(int) * h->n_host, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PF_OK;
}

extern "C" int32_t pf_get_particle(pf_handle h, int64_t i, int32_t *K_out, double *counts, double *means, double *chol, double *logdet)
{
    if (!h) return PF_ERR_ARG;
    if (i < 0 || i >= h->n_host) return set_error(h, PF_ERR_ARG, "pf_get_particle: index out of range");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const int d = h->d, F = h->F;
    int K = 0;
    CUDA_TRY(cudaMemcpyAsync(&K, h->Kc[h->cur] + i, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    std::vector<double> rec((size_t)std::max(K, 1) * F);
    if (K > 0) {
        // strided gather: element (j*F+f) lives at S[(j*F+f)*cap + i]
        std::vector<float> recf(h->f32 ? rec.size() : 0);
        CUDA_TRY(cudaMemcpy2DAsync(h->f32 ? (void *)recf.data() : (void *)rec.data(), h->esz, (const char *)h->S[h->cur] + i * h->esz, h->esz * h->cap, h->esz,
                                   (size_t)K * F, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        if (h->f32) for (size_t k = 0; k < rec.size(); k++) rec[k] = recf[k];
    }
    if (K_out) *K_out = K;
    for (int j = 0; j < K; j++) {
        const double *r = rec.data() + (size_t)j * F;
        if (counts) counts[j] = r[0];
        if (logdet) logdet[j] = r[1];
        if (means) for (int k = 0; k < d; k++) means[(size_t)j * d + k] = r[2 + k];
        if (chol) {
            double *Lj = chol + (size_t)j * d * d;
            std::fill(Lj, Lj + d * d, 0.0);
            for (int k = 0; k < d; k++) Lj[k + k * d] = 1.0 / r[2 + d + k];
            for (int a = 1, o = 0; a < d; a++) for (int b = 0; b < a; b++) Lj[a + b * d] = r[2 + 2 * d + o++];
        }
    }
    return PF_OK;
}


// assignments(ps), src/filters.jl:26-34: the genealogy is walked on the device into a byte-wide time-major block for
// a slice of the particles at a time (k_backtrack8), the slices are widened / transposed on the host
static long long assign_chunk(long long T, long long n)
{
    long long nc = (1ll << 30) / std::max<long long>(T, 1);          // <= 1 GiB of labels per slice
    if (getenv("PF_ASSIGN_CHUNK")) return std::max<long long>(1, std::min<long long>(atoll(getenv("PF_ASSIGN_CHUNK")), n));      // tests: force several slices
    nc = std::max<long long>(nc & ~255ll, 4096);
    return std::min(nc, (n + 15) & ~15ll);
}

extern "C" int32_t pf_get_assignments(pf_handle h, int32_t *out)
{
    if (!h || !out) return PF_ERR_ARG;
    if (!h->hist_cap) return set_error(h, PF_ERR_HISTORY, "pf_get_assignments: history is disabled (pf_config.history = 0)");
    if (h->nranks > 1) return set_error(h, PF_ERR_STATE, "pf_get_assignments: a sharded handle holds global parent indices; use pf_mg_get_assignments");
    const long long T = h->steps, n = h->n_host;
    if (T == 0 || n == 0) return PF_OK;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const long long nc_max = assign_chunk(T, n);
    const long long *goff = nullptr;
    { int32_t rc = upload_gen_off(h, &goff); if (rc != PF_OK) return rc; }
    unsigned char *tmp = nullptr;
    CUDA_TRY(dmalloc(&tmp, (size_t)T * nc_max));
    std::vector<unsigned char> host((size_t)T * nc_max);
    for (long long i0 = 0; i0 < n; i0 += nc_max) {
        const long long nc = std::min(nc_max, n - i0);
        k_backtrack8<<<grid_for(nc, 256), 256, 0, h->stream>>>(h->gen_anc, h->gen_asg, h->cap, T, i0, nc, tmp, goff);
        cudaError_t e = cudaMemcpyAsync(host.data(), tmp, (size_t)T * nc, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { cudaFree(tmp); return set_cuda_error(h, e, "pf_get_assignments", __LINE__); }
        for (long long c = 0; c < nc; c++)                  // -> column-major T x n
            for (long long t = 0; t < T; t++) out[(i0 + c) * T + t] = host[(size_t)t * nc + c];
    }
    cudaFree(tmp);
    return PF_OK;
}

// assignments(p) of ONE particle (src/particle.jl:25-32): its ancestor chain, T labels -- the read-out that stays cheap when
// the full T x n matrix is too large to bring to the host
extern "C" int32_t pf_get_particle_assignments(pf_handle h, int64_t i, int32_t *out)
{
    if (!h || !out) return PF_ERR_ARG;
    if (!h->hist_cap) return set_error(h, PF_ERR_HISTORY, "pf_get_particle_assignments: history is disabled (pf_config.history = 0)");
    if (h->nranks > 1) return set_error(h, PF_ERR_STATE, "pf_get_particle_assignments: use pf_mg_get_assignments on a sharded handle");
    if (i < 0 || i >= h->n_host) return set_error(h, PF_ERR_ARG, "pf_get_particle_assignments: index out of range");
    const long long T = h->steps;
    if (T == 0) return PF_OK;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    DevBuf tmp;
    CUDA_TRY(tmp.alloc((size_t)T));
    const long long *goff = nullptr;
    { int32_t rc = upload_gen_off(h, &goff); if (rc != PF_OK) return rc; }
    k_backtrack8<<<1, 32, 0, h->stream>>>(h->gen_anc, h->gen_asg, h->cap, T, i, 1, tmp.as<unsigned char>(), goff);
    std::vector<unsigned char> host((size_t)T);
    CUDA_TRY(cudaMemcpyAsync(host.data(), tmp.p, (size_t)T, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (long long t = 0; t < T; t++) out[t] = host[t];
    return PF_OK;
}

extern "C" int32_t pf_genealogy_info(pf_handle h, int64_t *out /* [4] */)
{
    if (!h || !out) return PF_ERR_ARG;
    long long nodes = 0;
    if (h->arena) for (long long t = 0; t < h->steps; t++) nodes += h->gen_n[t];
    else if (h->hist_cap) nodes = h->steps * h->cap;
    out[0] = h->hist_cap ? h->steps : 0; out[1] = nodes; out[2] = h->arena ? h->arena_cap : (h->hist_cap ? h->hist_cap * h->cap : 0); out[3] = h->n_prunes;
    return PF_OK;
}

extern "C" int32_t pf_get_last_step(pf_handle h, pf_step_stats *o)
{
    if (!h || !o) return PF_ERR_ARG;
    if (!h->have_last) return set_error(h, PF_ERR_STATE, "pf_get_last_step: no step has run");
    if (h->last_stale) {
        CUDA_TRY(cudaSetDevice(h->cfg.device));
        CUDA_TRY(cudaMemcpyAsync(&h->last, h->slog, sizeof(StepLog), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        h->last_stale = false;
    }
    const StepLog &r = h->last;
    memset(o, 0, sizeof *o);
    o->step = h->steps; o->n_in = r.n_in; o->n_putative = r.M; o->n_kept = r.n_kept; o->n_resamp_req = r.n_req;
    o->n_resamp_got = r.n_got; o->n_out = r.n_out; o->overflow = r.overflow; o->log_total_w = r.log_total;
    o->log_w_resamp = r.lw_resamp; o->threshold = r.thr; o->cv = r.cv; o->resampled = r.resampled;
    o->select_rounds = r.rounds; o->n_candidates = r.n_cand; o->n_plateau = r.n_plateau; o->select_status = r.status;
    for (int k = 0; k < 4; k++) o->select_phase_us[k] = r.sel_us[k];
    return PF_OK;
}

extern "C" int32_t pf_get_putative_weights(pf_handle h, double *out, int64_t cap_out, int64_t *M_out)
{
    if (!h || !M_out) return PF_ERR_ARG;
    if (!h->have_last || h->cfg.filter_kind != PF_FEARNHEAD) return set_error(h, PF_ERR_STATE, "pf_get_putative_weights: needs a Fearnhead step");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const long long n_in = h->last.n_in, slots = h->k_max + 2;
    std::vector<unsigned char> c(n_in);
    CUDA_TRY(cudaMemcpyAsync(c.data(), h->cnt, n_in, cudaMemcpyDeviceToHost, h->stream));
    std::vector<double> v((size_t)slots * n_in);
    CUDA_TRY(cudaMemcpy2DAsync(v.data(), sizeof(double) * n_in, h->V, sizeof(double) * h->cap, sizeof(double) * n_in, slots,
                               cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    long long M = 0;
    for (long long i = 0; i < n_in; i++)
        for (int j = 0; j < c[i]; j++, M++)
            if (out && M < cap_out) out[M] = v[(size_t)j * n_in + i];
    *M_out = M;
    return PF_OK;
}

extern "C" int32_t pf_get_last_ancestors(pf_handle h, int32_t *out)
{
    if (!h || !out) return PF_ERR_ARG;
    if (!h->hist_cap || h->steps == 0) return set_error(h, PF_ERR_HISTORY, "pf_get_last_ancestors: needs history >= 1 and one step");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    static_assert(sizeof(int32_t) == sizeof(unsigned int), "");
    CUDA_TRY(cudaMemcpyAsync(out, h->gen_anc + gen_row(h, h->steps - 1), sizeof(int) * h->n_host, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PF_OK;
}

extern "C" int32_t pf_get_last_assignments(pf_handle h, int32_t *out)
{
    if (!h || !out) return PF_ERR_ARG;
    if (!h->hist_cap || h->steps == 0) return set_error(h, PF_ERR_HISTORY, "pf_get_last_assignments: needs history >= 1 and one step");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    std::vector<unsigned char> a(h->n_host);
    CUDA_TRY(cudaMemcpyAsync(a.data(), h->gen_asg + gen_row(h, h->steps - 1), h->n_host, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    for (long long i = 0; i < h->n_host; i++) out[i] = a[i];
    return PF_OK;
}

// ------------------------------------------------------------------ filter-level summaries on the device
// (src/filters.jl:17-42, 53-80; kernels in summaries.cuh)

static int32_t summary_ready(pf_handle h, const char *who)
{
    if (h->nranks > 1) return set_error(h, PF_ERR_STATE, std::string(who) + ": sharded handles are summarised per shard by pf_mg_* read-outs");
    if (h->n_host < 1) return set_error(h, PF_ERR_STATE, std::string(who) + ": the population is empty");
    return PF_OK;
}

template <int D, typename R>
static void launch_predictive(pf_handle h, int nb, const double *xs_dev, int nq, double *part)
{
    k_predictive<D, R><<<nb, SUM_THREADS, 0, h->stream>>>((const R *)h->S[h->cur], h->logw[h->cur], h->Kc[h->cur], h->cap, h->n_host, h->tab, h->pc, xs_dev, nq, part);
}
template <int D, typename R>
static void launch_pop_summary(pf_handle h, int nb, int slots, double *part)
{
    k_pop_summary<D, R><<<nb, SUM_THREADS, 0, h->stream>>>((const R *)h->S[h->cur], h->logw[h->cur], h->Kc[h->cur], h->cap, h->n_host, slots, h->pc, h->sp[h->cur],
                                                         (double)h->steps, part);
}

// sums[q] = sum over this handle's particles of w_i sum_j (N_ij or alpha) t_ij(x_q), sums[Q] = sum of w_i
static int32_t predictive_sums(pf_handle h, const double *xs, int64_t Q, double *sums)
{
    if (h->sp_kind != PF_SP_CRP)      // components(p) has K + 1 entries, weights(p) one per candidate state: only the CRP lines them up
        return set_error(h, PF_ERR_STATE, "pf_posterior_predictive: defined for the ChineseRestaurantProcess state prior (src/particle.jl:156-172)");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const int d = h->d;
    const long long n = h->n_host;
    const int QB = 64;                                    // query points per pass over the store
    const long long nb = grid_for(n, SUM_THREADS);
    for (int64_t q = 0; q <= Q; q++) sums[q] = 0.0;
    if (n == 0) return PF_OK;
    DevBuf dx, dpart, dout;
    CUDA_TRY(dx.alloc(sizeof(double) * std::max<int64_t>(Q, 1) * d));
    CUDA_TRY(dpart.alloc(sizeof(double) * nb * (QB + 1)));
    CUDA_TRY(dout.alloc(sizeof(double) * (QB + 1)));
    CUDA_TRY(cudaMemcpyAsync(dx.p, xs, sizeof(double) * Q * d, cudaMemcpyHostToDevice, h->stream));
    for (long long q0 = 0; q0 < Q; q0 += QB) {
        const int nq = (int)std::min<long long>(QB, Q - q0);
        DISPATCH_D(d, launch_predictive<D, R>(h, (int)nb, dx.as<double>() + q0 * d, nq, dpart.as<double>()));
        k_sum_partials<<<nq + 1, SUM_THREADS, 0, h->stream>>>(dpart.as<double>(), nb, nq + 1, dout.as<double>());
        double res[QB + 1];
        CUDA_TRY(cudaMemcpyAsync(res, dout.p, sizeof(double) * (nq + 1), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        for (int q = 0; q < nq; q++) sums[q0 + q] = res[q];
        sums[Q] = res[nq];
    }
    return PF_OK;
}

extern "C" int32_t pf_posterior_predictive(pf_handle h, const double *xs, int64_t Q, double *out)
{
    if (!h || (Q > 0 && (!xs || !out)) || Q < 0) return PF_ERR_ARG;
    int32_t rc = summary_ready(h, "pf_posterior_predictive");
    if (rc != PF_OK) return rc;
    if (Q == 0) return PF_OK;
    std::vector<double> sums(Q + 1);
    rc = predictive_sums(h, xs, Q, sums.data());
    if (rc != PF_OK) return rc;
    const double norm = (double)h->steps + h->cfg.alpha;  // sum_j N_j + alpha, the same for every particle under the CRP
    for (int64_t q = 0; q < Q; q++) out[q] = sums[q] / sums[Q] / norm;
    return PF_OK;
}

// [0] sum w, [1] sum w K, [2] sum w H, [3 + k] weight of the particles with K = k
static int32_t pop_summary(pf_handle h, std::vector<double> &rec)
{
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const long long n = h->n_host;
    const int slots = h->k_max + 1, ncol = 3 + slots;
    const long long nb = grid_for(n, SUM_THREADS);
    DevBuf dpart, dout;
    CUDA_TRY(dpart.alloc(sizeof(double) * nb * ncol));
    CUDA_TRY(dout.alloc(sizeof(double) * ncol));
    DISPATCH_D(h->d, launch_pop_summary<D, R>(h, (int)nb, slots, dpart.as<double>()));
    k_sum_partials<<<ncol, SUM_THREADS, 0, h->stream>>>(dpart.as<double>(), nb, ncol, dout.as<double>());
    rec.resize(ncol);
    CUDA_TRY(cudaMemcpyAsync(rec.data(), dout.p, sizeof(double) * ncol, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PF_OK;
}

extern "C" int32_t pf_weighted_mean(pf_handle h, int32_t what, double *out)
{
    if (!h || !out) return PF_ERR_ARG;
    if (what != PF_MEAN_NCOMPONENTS && what != PF_MEAN_STATE_ENTROPY) return set_error(h, PF_ERR_ARG, "pf_weighted_mean: unknown quantity");
    int32_t rc = summary_ready(h, "pf_weighted_mean");
    if (rc != PF_OK) return rc;
    std::vector<double> rec;
    rc = pop_summary(h, rec);
    if (rc != PF_OK) return rc;
    *out = rec[what == PF_MEAN_NCOMPONENTS ? 1 : 2] / rec[0];
    return PF_OK;
}

extern "C" int32_t pf_ncomponents_dist(pf_handle h, double *out)
{
    if (!h || !out) return PF_ERR_ARG;
    int32_t rc = summary_ready(h, "pf_ncomponents_dist");
    if (rc != PF_OK) return rc;
    std::vector<double> rec;
    rc = pop_summary(h, rec);
    if (rc != PF_OK) return rc;
    for (int k = 0; k <= h->k_max; k++) out[k] = rec[3 + k] / rec[0];
    return PF_OK;
}

// number of particles that give observations s and t the same label, for all pairs (counts[T x T])
static int32_t similarity_counts(pf_handle h, std::vector<unsigned long long> &counts)
{
    if (!h->hist_cap) return set_error(h, PF_ERR_HISTORY, "assignment_similarity: history is disabled (pf_config.history = 0)");
    const long long T = h->steps, n = h->n_host;
    counts.assign((size_t)T * T, 0ull);
    if (T == 0) return PF_OK;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const long long nc_max = assign_chunk(T, n);
    const long long *goff = nullptr;
    { int32_t rc = upload_gen_off(h, &goff); if (rc != PF_OK) return rc; }
    DevBuf dA, dC;
    CUDA_TRY(dA.alloc((size_t)T * nc_max));
    CUDA_TRY(dC.alloc(sizeof(unsigned long long) * T * T));
    CUDA_TRY(cudaMemsetAsync(dC.p, 0, sizeof(unsigned long long) * T * T, h->stream));
    const long long ntile = (T + SIM_TILE - 1) / SIM_TILE, npair = ntile * (ntile + 1) / 2;
    if (npair > 0x7fffffffll) return set_error(h, PF_ERR_ARG, "assignment_similarity: too many observations");
    for (long long i0 = 0; i0 < n; i0 += nc_max) {
        const long long nc = std::min(nc_max, n - i0);
        k_backtrack8<<<grid_for(nc, 256), 256, 0, h->stream>>>(h->gen_anc, h->gen_asg, h->cap, T, i0, nc, dA.as<unsigned char>(), goff);
        long long gy = std::max<long long>(1, std::min<long long>((nc / 16 + SUM_THREADS * 4 - 1) / (SUM_THREADS * 4), 4096));
        if (npair * gy < 2 * h->num_sms) gy = std::min<long long>((2 * h->num_sms + npair - 1) / npair, std::max<long long>(1, (nc + SUM_THREADS - 1) / SUM_THREADS));
        k_similarity<<<dim3((unsigned)npair, (unsigned)gy), SUM_THREADS, 0, h->stream>>>(dA.as<unsigned char>(), T, nc, dC.as<unsigned long long>());
    }
    CUDA_TRY(cudaMemcpyAsync(counts.data(), dC.p, sizeof(unsigned long long) * T * T, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return PF_OK;
}

extern "C" int32_t pf_assignment_similarity(pf_handle h, double *out)
{
    if (!h || !out) return PF_ERR_ARG;
    int32_t rc = summary_ready(h, "pf_assignment_similarity");
    if (rc != PF_OK) return rc;
    std::vector<unsigned long long> c;
    rc = similarity_counts(h, c);
    if (rc != PF_OK) return rc;
    const long long T = h->steps;
    const double n = (double)h->n_host;
    for (long long k = 0; k < T * T; k++) out[k] = 1.0 - ((double)(h->n_host - (long long)c[k])) / n;     // 1 - Hamming / n
    return PF_OK;
}

extern "C" int32_t pf_randindex(pf_handle h, const int32_t *truth, int64_t T, int32_t adjusted, double *out)
{
    if (!h || !truth || !out) return PF_ERR_ARG;
    int32_t rc = summary_ready(h, "pf_randindex");
    if (rc != PF_OK) return rc;
    if (T != h->steps) return set_error(h, PF_ERR_ARG, "Ground truth and particles have different number of obs!");   // DimensionMismatch, src/filters.jl:57-58
    std::vector<unsigned long long> c;
    rc = similarity_counts(h, c);
    if (rc != PF_OK) return rc;
    // src/filters.jl:60-79 on ps_sim = 1 - Hamming / n and truth_sim = truth .== truth'
    const double n_part = (double)h->n_host;
    const double Np = (double)T * (double)T;
    double nis = 0.0, njs = 0.0, both_same = 0.0;
    for (long long s = 0; s < T; s++)
        for (long long t = 0; t < T; t++) {
            const double sim = 1.0 - ((double)(h->n_host - (long long)c[s * T + t])) / n_part;
            nis += sim;
            if (truth[s] == truth[t]) { njs += 1.0; both_same += sim; }
        }
    const double both_diff = Np - nis - njs + both_same;
    if (adjusted) {
        const double n = (double)T;
        const double nc = (n * (n * n + 1.0) - (n + 1.0) * nis - (n + 1.0) * njs + 2.0 * (nis * njs) / n) / (2.0 * (n - 1.0));
        *out = (both_same + both_diff - nc) / (Np - nc);
    } else {
        *out = (both_same + both_diff) / Np;
    }
    return PF_OK;
}

// ------------------------------------------------------------------ checkpoint / restore
// The population (live slots only), the control blocks, the genealogy and the host mirrors of a single-GPU handle
// go to one file; pf_load builds a handle that continues bit for bit (same seed and step counter => same device
// uniforms).  Layout: header, pf_config + prior arrays, mirrors, device control blocks, arrays.
static const char CKPT_MAGIC[8] = {'P', 'F', 'B', '2', '0', '0', 'C', '3'};
struct CkptHeader {
    char magic[8];
    uint32_t sz_cfg, sz_st, sz_res, sz_sc, sz_cl, sz_log;
    int32_t d, F, k_max, cur, sp_kind, have_last, max_k, has_hist;
    int64_t N, cap, steps, n_host, ub_live, hist_rows, epoch;
};

static bool dump_dev(FILE *f, const void *dptr, size_t bytes, std::vector<char> &stage)
{
    const char *p = static_cast<const char *>(dptr);
    while (bytes) {
        const size_t c = std::min(bytes, stage.size());
        if (cudaMemcpy(stage.data(), p, c, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
        if (fwrite(stage.data(), 1, c, f) != c) return false;
        p += c; bytes -= c;
    }
    return true;
}
static bool load_dev(FILE *f, void *dptr, size_t bytes, std::vector<char> &stage)
{
    char *p = static_cast<char *>(dptr);
    while (bytes) {
        const size_t c = std::min(bytes, stage.size());
        if (fread(stage.data(), 1, c, f) != c) return false;
        if (cudaMemcpy(p, stage.data(), c, cudaMemcpyHostToDevice) != cudaSuccess) return false;
        p += c; bytes -= c;
    }
    return true;
}

extern "C" int32_t pf_save(pf_handle h, const char *path)
{
    if (!h || !path) return PF_ERR_ARG;
    if (h->nranks > 1) return set_error(h, PF_ERR_STATE, "pf_save: sharded handles are not checkpointed");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    const long long n = h->n_host, cap = h->cap;
    const int cur = h->cur;
    std::vector<int> K(n);
    CUDA_TRY(cudaMemcpy(K.data(), h->Kc[cur], sizeof(int) * n, cudaMemcpyDeviceToHost));
    int max_k = 0;
    for (long long i = 0; i < n; i++) max_k = std::max(max_k, K[i]);
    FILE *f = fopen(path, "wb");
    if (!f) return set_error(h, PF_ERR_ARG, std::string("pf_save: cannot open ") + path);
    CkptHeader hd;
    memset(&hd, 0, sizeof hd);
    memcpy(hd.magic, CKPT_MAGIC, 8);
    hd.sz_cfg = sizeof(pf_config); hd.sz_st = sizeof(StepState); hd.sz_res = sizeof(SelResult); hd.sz_sc = sizeof(SelScratch);
    hd.sz_cl = sizeof(ClState); hd.sz_log = sizeof(StepLog);
    hd.d = h->d; hd.F = h->F; hd.k_max = h->k_max; hd.cur = cur; hd.sp_kind = h->sp_kind; hd.have_last = h->have_last; hd.max_k = max_k;
    hd.has_hist = h->hist_cap > 0; hd.N = h->N; hd.cap = cap; hd.steps = h->steps; hd.n_host = n; hd.ub_live = h->ub_live;
    hd.hist_rows = h->hist_cap > 0 ? h->steps : 0; hd.epoch = h->epoch;
    std::vector<char> stage((size_t)64 << 20);
    bool ok = fwrite(&hd, sizeof hd, 1, f) == 1 && fwrite(&h->cfg, sizeof h->cfg, 1, f) == 1 &&
              fwrite(h->mu0.data(), sizeof(double), h->d, f) == (size_t)h->d &&
              fwrite(h->lam0.data(), sizeof(double), (size_t)h->d * h->d, f) == (size_t)h->d * h->d;
    const int64_t nsp = (int64_t)h->sp_counts.size();
    ok = ok && fwrite(&nsp, sizeof nsp, 1, f) == 1 && (nsp == 0 || fwrite(h->sp_counts.data(), sizeof(double), nsp, f) == (size_t)nsp);
    ok = ok && fwrite(&h->last, sizeof(StepLog), 1, f) == 1;
    ok = ok && dump_dev(f, h->st, sizeof(StepState), stage) && dump_dev(f, h->res, sizeof(SelResult), stage) && dump_dev(f, h->sc, sizeof(SelScratch), stage);
    if (h->cl) ok = ok && dump_dev(f, h->cl, sizeof(ClState), stage);
    ok = ok && dump_dev(f, h->Kc[cur], sizeof(int) * n, stage) && dump_dev(f, h->logw[cur], sizeof(double) * n, stage);
    for (long long r = 0; ok && r < (long long)max_k * h->F; r++) ok = dump_dev(f, (const char *)h->S[cur] + r * cap * h->esz, h->esz * n, stage);
    if (h->sp_kind == PF_SP_STICKY) {
        for (long long r = 0; ok && r < max_k; r++) ok = dump_dev(f, h->sp[cur].N + r * cap, sizeof(double) * n, stage) && dump_dev(f, h->sp[cur].Nsame + r * cap, sizeof(double) * n, stage);
        ok = ok && dump_dev(f, h->sp[cur].last, sizeof(int) * n, stage);
    }
    for (long long t = 0; ok && t < hd.hist_rows; t++) {          // every generation with its own length (pruned arena) or `cap` (dense log)
        const int64_t nt = h->arena ? h->gen_n[t] : cap;
        ok = fwrite(&nt, sizeof nt, 1, f) == 1 && dump_dev(f, h->gen_anc + gen_row(h, t), sizeof(unsigned int) * nt, stage) &&
             dump_dev(f, h->gen_asg + gen_row(h, t), (size_t)nt, stage);
    }
    ok = (fclose(f) == 0) && ok;
    if (!ok) { cudaGetLastError(); return set_error(h, PF_ERR_STATE, std::string("pf_save: write to ") + path + " failed"); }
    return PF_OK;
}

extern "C" int32_t pf_load(const char *path, int32_t device, pf_handle *out)
{
    if (!path || !out) return set_error(nullptr, PF_ERR_ARG, "pf_load: null argument");
    *out = nullptr;
    FILE *f = fopen(path, "rb");
    if (!f) return set_error(nullptr, PF_ERR_ARG, std::string("pf_load: cannot open ") + path);
    CkptHeader hd;
    pf_config cfg;
    bool ok = fread(&hd, sizeof hd, 1, f) == 1 && memcmp(hd.magic, CKPT_MAGIC, 8) == 0 && hd.sz_cfg == sizeof(pf_config) &&
              hd.sz_st == sizeof(StepState) && hd.sz_res == sizeof(SelResult) && hd.sz_sc == sizeof(SelScratch) &&
              hd.sz_cl == sizeof(ClState) && hd.sz_log == sizeof(StepLog) && fread(&cfg, sizeof cfg, 1, f) == 1;
    if (!ok || hd.d < 1 || hd.d > PF_MAX_D) { fclose(f); return set_error(nullptr, PF_ERR_ARG, "pf_load: not a checkpoint of this library version"); }
    std::vector<double> mu0(hd.d), lam0((size_t)hd.d * hd.d), spc;
    int64_t nsp = 0;
    ok = fread(mu0.data(), sizeof(double), hd.d, f) == (size_t)hd.d && fread(lam0.data(), sizeof(double), lam0.size(), f) == lam0.size() &&
         fread(&nsp, sizeof nsp, 1, f) == 1 && nsp >= 0 && nsp <= PF_MAX_SLOTS;
    if (ok && nsp) { spc.resize(nsp); ok = fread(spc.data(), sizeof(double), nsp, f) == (size_t)nsp; }
    StepLog last;
    ok = ok && fread(&last, sizeof last, 1, f) == 1;
    if (!ok) { fclose(f); return set_error(nullptr, PF_ERR_ARG, "pf_load: truncated checkpoint"); }
    cfg.mu0 = mu0.data(); cfg.lambda0 = lam0.data(); cfg.sp_counts = nsp ? spc.data() : nullptr;
    cfg.device = device;
    if (cfg.prior_kind == PF_NIX2) cfg.lambda0 = nullptr;
    if (cfg.history > 0 && cfg.history < hd.steps) cfg.history = hd.steps;
    pf_handle h = nullptr;
    int32_t rc = pf_create(&cfg, &h);
    if (rc != PF_OK) { fclose(f); return rc; }
    auto fail = [&](const char *msg) { fclose(f); cudaGetLastError(); g_create_error = msg; pf_destroy(h); return (int32_t)PF_ERR_STATE; };
    if (h->cap != hd.cap || h->F != hd.F) return fail("pf_load: layout mismatch");
    h->steps = hd.steps; h->n_host = hd.n_host; h->ub_live = hd.ub_live; h->cur = hd.cur; h->epoch = (unsigned int)hd.epoch;
    h->last = last; h->have_last = hd.have_last != 0;
    if (ensure_table(h, h->steps + 4096) != PF_OK) return fail("pf_load: table allocation failed");
    if (hd.has_hist && !h->arena && ensure_history(h, hd.steps) != PF_OK) return fail("pf_load: the genealogy does not fit in device memory");
    std::vector<char> stage((size_t)64 << 20);
    const long long n = hd.n_host, cap = hd.cap;
    const int cur = hd.cur;
    ok = load_dev(f, h->st, sizeof(StepState), stage) && load_dev(f, h->res, sizeof(SelResult), stage) && load_dev(f, h->sc, sizeof(SelScratch), stage);
    if (h->cl) ok = ok && load_dev(f, h->cl, sizeof(ClState), stage);
    ok = ok && load_dev(f, h->Kc[cur], sizeof(int) * n, stage) && load_dev(f, h->logw[cur], sizeof(double) * n, stage);
    for (long long r = 0; ok && r < (long long)hd.max_k * h->F; r++) ok = load_dev(f, (char *)h->S[cur] + r * cap * h->esz, h->esz * n, stage);
    if (h->sp_kind == PF_SP_STICKY) {
        for (long long r = 0; ok && r < hd.max_k; r++) ok = load_dev(f, h->sp[cur].N + r * cap, sizeof(double) * n, stage) && load_dev(f, h->sp[cur].Nsame + r * cap, sizeof(double) * n, stage);
        ok = ok && load_dev(f, h->sp[cur].last, sizeof(int) * n, stage);
    }
    if (hd.has_hist && h->hist_cap < hd.hist_rows) ok = false;
    for (long long t = 0; ok && hd.has_hist && t < hd.hist_rows; t++) {
        int64_t nt = 0;
        ok = fread(&nt, sizeof nt, 1, f) == 1 && nt >= 0 && nt <= cap;
        if (ok && h->arena) {
            ok = arena_reserve(h, nt) == PF_OK;
            if (ok) { h->gen_off.push_back(h->arena_used); h->gen_n.push_back(nt); h->arena_used += nt; }
        } else if (ok && nt != cap) ok = false;
        ok = ok && load_dev(f, h->gen_anc + gen_row(h, t), sizeof(unsigned int) * nt, stage) && load_dev(f, h->gen_asg + gen_row(h, t), (size_t)nt, stage);
    }
    if (!ok) return fail("pf_load: truncated checkpoint or device copy failed");
    fclose(f);
    *out = h;
    return PF_OK;
}

extern "C" int32_t pf_get_config(pf_handle h, pf_config *out)
{
    if (!h || !out) return PF_ERR_ARG;
    *out = h->cfg;                                     // (the array pointers are null: pf_get_prior returns the arrays)
    return PF_OK;
}

extern "C" int32_t pf_get_prior(pf_handle h, double *mu0, double *lambda0, double *sp_counts)
{
    if (!h) return PF_ERR_ARG;
    if (mu0) std::copy(h->mu0.begin(), h->mu0.end(), mu0);
    if (lambda0) std::copy(h->lam0.begin(), h->lam0.end(), lambda0);
    if (sp_counts) std::copy(h->sp_counts.begin(), h->sp_counts.end(), sp_counts);
    return PF_OK;
}

// ------------------------------------------------------------------ stand-alone selection
__global__ void k_vmax(const double *w, long long M, StepState *st)
{
    u64 v = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += (long long)gridDim.x * blockDim.x) {
        u64 b = (u64)__double_as_longlong(w[i]);
        v = b > v ? b : v;
    }
    v = warp_max64(v);
    if ((threadIdx.x & 31) == 0 && v) atomicMax(&st->vmax_bits, v);
}

extern "C" int32_t pf_select(int32_t device, const double *w, int64_t M, int64_t N, double u, int32_t order,
                             int64_t *survivors, uint8_t *resampled, pf_step_stats *stats)
{
    pf_filter_s tmp;             // only for error text / stream bookkeeping
    pf_handle h = &tmp;
    if (!w || M < 1 || N < 1 || !survivors || M > (1ll << 28)) return set_error(nullptr, PF_ERR_ARG, "pf_select: bad argument");
    h->stc = nullptr;
    if (!(u >= 0.0 && u < 1.0)) return set_error(nullptr, PF_ERR_ARG, "pf_select: u must be in [0, 1)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return set_error(nullptr, PF_ERR_CUDA, "pf_select: no CUDA device (no CPU fallback)"); }
    int32_t rc = PF_OK;
    double *dW = nullptr, *dU = nullptr, *cand = nullptr;
    StepState *st = nullptr; SelScratch *sc = nullptr; SelResult *res = nullptr;
    TileSum *tiles = nullptr;
    TileSum *supers = nullptr;            // totals of SUPER_TILES consecutive tile records (level 2 of the survivor scan)
    unsigned int *sp = nullptr; unsigned char *ss = nullptr;
    u64 *keys = nullptr, *keys2 = nullptr; unsigned int *vals = nullptr, *vals2 = nullptr;
    unsigned int *hist = nullptr;
    std::vector<unsigned int> perm_host;
#define ST(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { rc = set_cuda_error(h, e_, #x, __LINE__); g_create_error = h->err; goto done; } } while (0)
    {
        ST(cudaSetDevice(device));
        cudaDeviceProp prop;
        ST(cudaGetDeviceProperties(&prop, device));
        h->num_sms = prop.multiProcessorCount;
        ST(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        {
            int occ = 0;
            ST(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sel_classify, SEL_THREADS, 0));
            h->sel_grid = h->num_sms * (occ > 2 ? 2 : (occ < 1 ? 1 : occ));
            rc = setup_search_kernels(h);
            if (rc != PF_OK) { g_create_error = h->err; goto done; }
            h->sel_margin = getenv("PF_SEL_MARGIN") ? std::max(1, atoi(getenv("PF_SEL_MARGIN"))) : 8;
        }
        h->fuse_stride = std::max<long long>(1, (M + SEL_SAMPLE_TARGET - 1) / SEL_SAMPLE_TARGET);
        long long cand_cap = M + (long long)h->sel_grid * (SEL_THREADS / 32) * SEL_CHUNK * 2 + SEL_CHUNK;
        long long ntiles = (M + SCAN_THREADS - 1) / SCAN_THREADS;
        ST(dmalloc(&dW, (size_t)M)); ST(dmalloc(&dU, 1)); ST(dmalloc(&cand, (size_t)cand_cap));
        ST(dmalloc(&st, 1)); ST(dmalloc(&sc, 1)); ST(dmalloc(&res, 1));
        ST(dmalloc(&tiles, (size_t)ntiles));
        ST(dmalloc(&supers, (size_t)ntiles / SUPER_TILES + 2));
        ST(dmalloc(&sp, (size_t)M)); ST(dmalloc(&ss, (size_t)M));
        ST(cudaMemsetAsync(sc, 0, sizeof(SelScratch), h->stream));
        ST(cudaMemsetAsync(res, 0, sizeof(SelResult), h->stream));
        StepState s0; memset(&s0, 0, sizeof s0);
        s0.n_live = M; s0.M = M;
        ST(cudaMemcpyAsync(st, &s0, sizeof s0, cudaMemcpyHostToDevice, h->stream));
        ST(cudaMemcpyAsync(dW, w, sizeof(double) * M, cudaMemcpyHostToDevice, h->stream));
        ST(cudaMemcpyAsync(dU, &u, sizeof(double), cudaMemcpyHostToDevice, h->stream));
        const double *Vsel = dW;
        if (M <= N) order = PF_ORDER_INDEX;          // exhaustive step: no sort (src/fearnhead.jl:135-138)
        if (order == PF_ORDER_SORTED) {
            // literal reference order: stable ascending sort by weight (src/fearnhead.jl:148)
            ST(dmalloc(&keys, (size_t)M)); ST(dmalloc(&keys2, (size_t)M));
            ST(dmalloc(&vals, (size_t)M)); ST(dmalloc(&vals2, (size_t)M));
            int nblk = grid_for(M, RS_TILE);
            ST(dmalloc(&hist, (size_t)(256 + 1) * nblk + 1024));
            rc = radix_sort_pairs(h->stream, (const u64 *)dW, keys, keys2, vals, vals2, hist, M, &h->launches);
            if (rc != PF_OK) { rc = set_error(nullptr, PF_ERR_CUDA, "pf_select: radix sort failed"); goto done; }
            Vsel = (const double *)keys;           // sorted weights (bit patterns are the doubles)
        }
        k_vmax<<<h->num_sms * 4, 256, 0, h->stream>>>(Vsel, M, st);
        h->retries = 12;                  // stand-alone call: queue enough rounds for any input
        rc = launch_classic_search(h, sel_args(h, Vsel, nullptr, M, &st->n_live, &st->M, &st->vmax_bits, N, dU, cand, cand_cap, sc, res));
        if (rc != PF_OK) { g_create_error = h->err; goto done; }
        k_tile_sums<<<std::min(grid_for(M, SCAN_THREADS), h->num_sms * 8), SCAN_THREADS, 0, h->stream>>>(Vsel, nullptr, M, res, &st->n_live, tiles, 0, sc, nullptr);
        k_tile_scan1<<<grid_for(ntiles, SUPER_TILES), SUPER_TILES, 0, h->stream>>>(tiles, nullptr, nullptr, nullptr, res, nullptr, &st->n_live, nullptr, supers, 0, sc);
        k_tile_prefix<<<1, 1024, 0, h->stream>>>(supers, res, &st->n_live, st, 0, nullptr, 0, sc, 1);
        k_scan_apply<<<grid_for(M, SCAN_THREADS), SCAN_THREADS, 0, h->stream>>>(Vsel, nullptr, M, res, &st->n_live, tiles, supers, sp, ss, 0, nullptr, sc, M);
        ST(cudaGetLastError());
        StepState s1; SelResult r1;
        int phase_end = 0;
        ST(cudaMemcpyAsync(&s1, st, sizeof s1, cudaMemcpyDeviceToHost, h->stream));
        ST(cudaMemcpyAsync(&r1, res, sizeof r1, cudaMemcpyDeviceToHost, h->stream));
        ST(cudaMemcpyAsync(&phase_end, &sc->phase, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        ST(cudaStreamSynchronize(h->stream));
        if (phase_end != 4) {
            SelScratch hs;
            cudaMemcpy(&hs, sc, offsetof(SelScratch, final_vals), cudaMemcpyDeviceToHost);
            char buf[400];
            snprintf(buf, sizeof buf, "pf_select: the threshold search did not converge (phase %d mode %d rounds %d status %d lo %llx hi %llx scale %d "
                     "count_in %lld n_fail %d n_cand %llu A_hi %llu final_n %u chunks %llu)", hs.phase, hs.mode, r1.rounds, r1.status,
                     (unsigned long long)hs.lo, (unsigned long long)hs.hi, hs.scale, hs.count_in, hs.n_fail, hs.n_cand, hs.A_hi, hs.final_n, hs.cand_chunks);
            rc = set_error(nullptr, PF_ERR_STATE, buf); goto done;
        }
        long long n_out = s1.n_next;
        std::vector<unsigned int> p(n_out > 0 ? n_out : 1);
        std::vector<unsigned char> sl(n_out > 0 ? n_out : 1);
        if (n_out > 0) {
            ST(cudaMemcpyAsync(p.data(), sp, sizeof(unsigned int) * n_out, cudaMemcpyDeviceToHost, h->stream));
            ST(cudaMemcpyAsync(sl.data(), ss, n_out, cudaMemcpyDeviceToHost, h->stream));
        }
        if (order == PF_ORDER_SORTED) {
            perm_host.resize(M);
            ST(cudaMemcpyAsync(perm_host.data(), vals, sizeof(unsigned int) * M, cudaMemcpyDeviceToHost, h->stream));
        }
        ST(cudaStreamSynchronize(h->stream));
        if (order == PF_ORDER_SORTED) {
            // output order of the reference: [resampled (ascending) ; kept (ascending)], src/fearnhead.jl:172-179
            long long k = 0;
            for (int pass = 0; pass < 2; pass++)
                for (long long t = 0; t < n_out; t++) {
                    bool picked = sl[t] & 0x80;
                    if (picked == (pass == 0)) { survivors[k] = perm_host[p[t]]; if (resampled) resampled[k] = picked; k++; }
                }
        } else {
            for (long long t = 0; t < n_out; t++) { survivors[t] = p[t]; if (resampled) resampled[t] = (sl[t] & 0x80) ? 1 : 0; }
        }
        if (stats) {
            memset(stats, 0, sizeof *stats);
            stats->n_in = M; stats->n_putative = M; stats->n_kept = s1.n_kept; stats->n_resamp_req = r1.keep_all ? 0 : r1.n;
            stats->n_resamp_got = s1.n_picked; stats->n_out = n_out; stats->log_total_w = r1.log_total;
            stats->log_w_resamp = r1.log_c; stats->threshold = r1.thr_value; stats->resampled = r1.keep_all ? 0 : 1;
            stats->select_rounds = r1.rounds; stats->n_candidates = r1.n_cand_first;
        }
    }
done:
#undef ST
    if (h->stream) { cudaStreamSynchronize(h->stream); }
    {
        void *ptrs[] = {dW, dU, cand, st, sc, res, tiles, supers, sp, ss, keys, keys2, vals, vals2, hist};
        for (void *p : ptrs) if (p) cudaFree(p);
    }
    if (h->stream) cudaStreamDestroy(h->stream);
    h->stream = nullptr;
    return rc;
}

// ------------------------------------------------------------------ sharded run
// Particles block-sharded over ranks (one handle per GPU; one process per GPU, or several handles in
// one process).  Every exchange of a step is done by the kernels themselves through peer memory
// (select.cuh, "peer exchange"); the host only queues launches: no collective library, no host round
// trip inside pf_mg_filter.
#define MG_CHECK(h) do { if (!(h)) return PF_ERR_ARG; if ((h)->nranks < 2) return set_error((h), PF_ERR_STATE, "not a sharded handle (pf_config.nranks < 2)"); } while (0)

extern "C" int32_t pf_set_stream(pf_handle h, void *cuda_stream)
{
    if (!h) return PF_ERR_ARG;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (h->own_stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)cuda_stream; h->own_stream = false;
    return PF_OK;
}

// arrays a peer must be able to write or read: S[0], S[1], Kc[0], Kc[1], logw[0], logw[1], gen_anc, gen_asg, exchange
// region, and (Chen-Liu) the prefix array and the step's assignments
static void *peer_array(pf_handle h, int k)
{
    switch (k) {
    case 0: return h->S[0]; case 1: return h->S[1]; case 2: return h->Kc[0]; case 3: return h->Kc[1];
    case 4: return h->logw[0]; case 5: return h->logw[1]; case 6: return h->gen_anc; case 7: return h->gen_asg;
    case 8: return h->xreg; case 9: return h->cum; case 10: return h->asg_tmp;
    case 11: return h->Vbuf[0]; case 12: return h->Vbuf[1]; case 13: return h->cnt;
    case 14: return h->srec.x; case 15: return h->srec.m;
    default: return h->srec.h;
    }
}

extern "C" int32_t pf_mg_local_arrays(pf_handle h, void **out /* [PF_MG_NARRAYS] */)
{
    MG_CHECK(h);
    for (int k = 0; k < PF_MG_NARRAYS; k++) out[k] = peer_array(h, k);
    return PF_OK;
}

extern "C" int32_t pf_mg_ipc_handles(pf_handle h, void *out /* PF_MG_NARRAYS x 64 bytes */)
{
    MG_CHECK(h);
    static_assert(sizeof(cudaIpcMemHandle_t) == PF_MG_IPC_BYTES, "IPC handle size");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    memset(out, 0, PF_MG_NARRAYS * sizeof(cudaIpcMemHandle_t));
    for (int k = 0; k < PF_MG_NARRAYS; k++)
        if (peer_array(h, k)) CUDA_TRY(cudaIpcGetMemHandle((cudaIpcMemHandle_t *)out + k, peer_array(h, k)));
    return PF_OK;
}

// arrays of rank `peer`: either pointers valid in this process (ptrs != NULL: the peer's handle lives in
// this process) or the IPC handles pf_mg_ipc_handles produced in the peer's process
extern "C" int32_t pf_mg_set_peer(pf_handle h, int32_t peer, void *const *ptrs, const void *ipc_handles)
{
    MG_CHECK(h);
    if (peer < 0 || peer >= h->nranks) return set_error(h, PF_ERR_ARG, "pf_mg_set_peer: bad rank");
    if (peer != h->rank && !ptrs && !ipc_handles) return set_error(h, PF_ERR_ARG, "pf_mg_set_peer: no pointers and no IPC handles");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    void *a[PF_MG_NARRAYS];
    for (int k = 0; k < PF_MG_NARRAYS; k++) {
        a[k] = nullptr;
        if (peer == h->rank) a[k] = peer_array(h, k);
        else if (ptrs) a[k] = ptrs[k];
        else if (peer_array(h, k)) {      // same configuration on every rank: the peer has the array iff we do
            CUDA_TRY(cudaIpcOpenMemHandle(&a[k], ((const cudaIpcMemHandle_t *)ipc_handles)[k], cudaIpcMemLazyEnablePeerAccess));
            h->ipc_opened.push_back(a[k]);
        }
    }
    for (int b = 0; b < 2; b++) {
        h->peers_host[b].S[peer] = a[b]; h->peers_host[b].Kc[peer] = (int *)a[2 + b];
        h->peers_host[b].logw[peer] = (double *)a[4 + b];
        h->peers_host[b].gen_anc[peer] = (unsigned int *)a[6]; h->peers_host[b].gen_asg[peer] = (unsigned char *)a[7];
        h->peers_host[b].V[0][peer] = (double *)a[11]; h->peers_host[b].V[1][peer] = (double *)a[12];
        h->peers_host[b].cnt[peer] = (unsigned char *)a[13];
        h->peers_host[b].recx[peer] = (ulonglong2 *)a[14]; h->peers_host[b].recm[peer] = (unsigned long long *)a[15];
        h->peers_host[b].rech[peer] = (unsigned long long *)a[16];
    }
    h->xch_host.reg[peer] = (XchRegion *)a[8];
    for (int b = 0; b < 2; b++) { h->clp_host.S[b][peer] = a[b]; h->clp_host.Kc[b][peer] = (const int *)a[2 + b]; }
    h->clp_host.cum[peer] = (const u128 *)a[9]; h->clp_host.asg_tmp[peer] = (const unsigned char *)a[10];
    return PF_OK;
}

// All peers are set: upload the peer tables.  hs[0..n_local) are the handles of THIS process; if several of
// them share a device they are driven in lock step on one stream (publish and consume as separate launches:
// a kernel that waits for a peer on the same device could not be overtaken by the kernel it waits for).
extern "C" int32_t pf_mg_connect(pf_handle *hs, int32_t n_local)
{
    if (!hs || n_local < 1) return PF_ERR_ARG;
    bool same_device = false;
    for (int i = 0; i < n_local; i++) {
        MG_CHECK(hs[i]);
        for (int j = 0; j < i; j++) same_device = same_device || hs[i]->cfg.device == hs[j]->cfg.device;
    }
    for (int i = 0; i < n_local; i++) {
        pf_handle h = hs[i];
        CUDA_TRY(cudaSetDevice(h->cfg.device));
        for (int r = 0; r < h->nranks; r++)
            if (!h->xch_host.reg[r]) return set_error(h, PF_ERR_STATE, "pf_mg_connect: a peer has not been set (pf_mg_set_peer)");
        // handles of this process on other devices: plain pointers need peer access
        for (int j = 0; j < n_local; j++)
            if (hs[j]->cfg.device != h->cfg.device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(hs[j]->cfg.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return set_cuda_error(h, e, "cudaDeviceEnablePeerAccess", __LINE__);
                cudaGetLastError();
            }
        for (int b = 0; b < 2; b++) {
            if (!h->peers[b]) CUDA_TRY(dmalloc(&h->peers[b], 1));
            CUDA_TRY(cudaMemcpyAsync(h->peers[b], &h->peers_host[b], sizeof(PeerTable), cudaMemcpyHostToDevice, h->stream));
        }
        if (!h->xch) CUDA_TRY(dmalloc(&h->xch, 1));
        CUDA_TRY(cudaMemcpyAsync(h->xch, &h->xch_host, sizeof(Xch), cudaMemcpyHostToDevice, h->stream));
        if (h->clp) CUDA_TRY(cudaMemcpyAsync(h->clp, &h->clp_host, sizeof(ClPeers), cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        h->lockstep = same_device;
        if (same_device) {                               // one stream for the whole group
            if (i == 0) {
                if (h->own_stream) { h->shared_stream = std::shared_ptr<void>((void *)h->stream, [](void *s) { cudaStreamDestroy((cudaStream_t)s); }); h->own_stream = false; }
            } else {
                if (h->own_stream) cudaStreamDestroy(h->stream);
                h->stream = hs[0]->stream; h->own_stream = false; h->shared_stream = hs[0]->shared_stream;
            }
        }
        h->connected = true;
    }
    return PF_OK;
}

// ---- the stages of one sharded observation (each queues launches on the handle's stream)
// A: roll the control block over (waits for every rank's previous generation), sample expansion, bracket
template <int D, typename R>
static int32_t mg_stage_a(pf_handle h, const double *y_dev, const double *u_dev, int search_stage)
{
    const long long cap = h->cap;
    const int cur = h->cur;
    const unsigned long long seq = (unsigned long long)h->steps + 1;
    if (search_stage & SEL_ST_PUB) {
        {
            LaunchTimer lt(h, KC_PRESTEP);            // (includes the wait for every rank's previous generation)
            k_prestep<D><<<1, 32, 0, h->stream>>>(h->st, h->pc, y_dev, h->stc, h->steps == 0 ? 1 : 0, h->sc, h->res, h->xch, seq, h->steps);
        }
        LaunchTimer lt(h, KC_SELECT);
        const long long ns = 4 * ((cap + 4 * h->fuse_stride - 1) / (4 * h->fuse_stride) + 1);      // groups of 4 particles (+1: the shard's phase)
        k_expand<D, 1, false, R><<<grid_for(ns, EXP_SAMPLE_THREADS), EXP_SAMPLE_THREADS, 0, h->stream>>>((const R *)h->S[cur], h->logw[cur], h->Kc[cur], cap, h->k_max, h->tab, h->pc,
                                                                   h->stc, h->st, h->V, h->cnt, h->sc, h->fuse_stride, h->cand, nullptr, nullptr);
    }
    SelArgs a = sel_args(h, h->V, h->cnt, cap, &h->st->n_live, &h->st->M_pred, &h->st->vmax_sample, h->N, u_dev, h->cand, h->cand_cap, h->sc, h->res);
    a.external_sample = 1; a.x = h->xch; a.stage = search_stage;
    LaunchTimer lt(h, KC_SELECT);
    CUDA_TRY(launch_coop(k_sel_bracket, h->coop_grid, h->stream, a));
    return PF_OK;
}

// B: fused expansion (round 0) / classify pass (retry rounds), then the solve on the gathered candidates
template <int D, typename R>
static int32_t mg_stage_b(pf_handle h, const double *u_dev, int retry, int search_stage)
{
    const long long cap = h->cap;
    const int cur = h->cur;
    SelArgs a = sel_args(h, h->V, h->cnt, cap, &h->st->n_live, &h->st->M, &h->st->vmax_bits, h->N, u_dev, h->cand, h->cand_cap, h->sc, h->res);
    a.external_sample = 1; a.preclassified = 1; a.x = h->xch; a.retry = retry;
    if (search_stage & SEL_ST_PUB) {
        if (!retry && !h->pre_expanded) {
            LaunchTimer lt(h, KC_EXPAND);
            k_expand<D, 2, false, R><<<grid_for(cap, EXP_THREADS), EXP_THREADS, 0, h->stream>>>((const R *)h->S[cur], h->logw[cur], h->Kc[cur], cap, h->k_max, h->tab, h->pc,
                                                                        h->stc, h->st, h->V, h->cnt, h->sc, 1, h->cand, h->tsub, h->tfuse, h->srec);
        }
        LaunchTimer lt(h, KC_SELECT);
        k_sel_classify<<<h->sel_grid, SEL_THREADS, 0, h->stream>>>(a);
    }
    a.stage = search_stage;
    LaunchTimer lt(h, KC_SELECT);
    CUDA_TRY(launch_coop(k_sel_solve, h->coop_grid, h->stream, a));
    return PF_OK;
}

// C: survivor-scan tile records and their prefix; the shard's total goes to every rank
static int32_t mg_stage_c(pf_handle h)
{
    const int g = grid_for(h->cap, SCAN_THREADS);
    const unsigned long long seq = (unsigned long long)h->steps + 1;
    LaunchTimer lt(h, KC_SCAN, 3);
    if (h->pre_expanded) {
        // the particles were expanded by their parents' owners: tile records from the delivered per-particle records
        k_tile_sums<<<std::min(g, h->num_sms * 8), SCAN_THREADS, 0, h->stream>>>(h->V, h->cnt, h->cap, h->res, &h->st->n_live, h->tiles, 0, h->sc, nullptr,
                                                                                  h->srec, h->sc);
        k_tile_scan1<<<grid_for(g, SUPER_TILES), SUPER_TILES, 0, h->stream>>>(h->tiles, nullptr, nullptr, nullptr, h->res, h->sc, &h->st->n_live, h->stc, h->supers, 0, h->sc);
    } else {
        k_tile_sums<<<std::min(g, h->num_sms * 8), SCAN_THREADS, 0, h->stream>>>(h->V, h->cnt, h->cap, h->res, &h->st->n_live, h->tiles, 0, h->sc, h->sc);
        k_tile_scan1<<<grid_for(g, SUPER_TILES), SUPER_TILES, 0, h->stream>>>(h->tiles, h->tsub, h->tfuse, h->cand, h->res, h->sc, &h->st->n_live, h->stc, h->supers, 0, h->sc);
    }
    k_tile_prefix<<<1, 1024, 0, h->stream>>>(h->supers, h->res, &h->st->n_live, h->st, 0, h->xch, seq, h->sc, 1);
    return PF_OK;
}

// D: global offsets from every rank's total, survivor scan; then either the classic tail -- instantiate (children that
// change owner are stored into their new owner's arrays), step record + "generation complete" signal -- or the
// pipelined one (mg_tail): the children are instantiated WHILE they are expanded for the next observation
static int32_t mg_stage_d_scan(pf_handle h)
{
    const unsigned long long seq = (unsigned long long)h->steps + 1;
    LaunchTimer lt(h, KC_SCAN, h->ttooth ? 3 : 2);            // (k_rank_offsets includes the wait for every rank's total)
    k_rank_offsets<<<1, 32, 0, h->stream>>>(h->xch, seq, h->steps, h->res, h->st, h->ro, h->sc, h->surv_cap);
    if (h->ttooth) k_tooth_bases<<<grid_for(grid_for(h->cap, SCAN_THREADS), 256), 256, 0, h->stream>>>(h->tiles, h->supers, h->res, &h->st->n_live, h->ro, h->ttooth, 0, h->sc);
    k_scan_apply<<<grid_for(h->cap, SCAN_THREADS), SCAN_THREADS, 0, h->stream>>>(h->V, h->cnt, h->cap, h->res, &h->st->n_live, h->tiles, h->supers,
                                                                                  h->surv_parent, h->surv_slot, 0, h->ro, h->sc, h->surv_cap, h->srec, h->sc, h->ttooth);
    return PF_OK;
}

template <int D, typename R>
static int32_t mg_stage_d(pf_handle h, StepLog *log_dev)
{
    const int cur = h->cur, nxt = cur ^ 1;
    const unsigned long long seq = (unsigned long long)h->steps + 1;
    unsigned int *ga = nullptr; unsigned char *gs = nullptr;
    if (h->hist_cap) { ga = h->gen_anc + (size_t)h->steps * h->cap; gs = h->gen_asg + (size_t)h->steps * h->cap; }
    {
        LaunchTimer lt(h, KC_INSTANTIATE);
        k_instantiate<D, R><<<grid_for(std::min<long long>(h->surv_cap, h->cap + h->cap / 4), 256), 256, 0, h->stream>>>((const R *)h->S[cur], h->Kc[cur], (R *)h->S[nxt], h->Kc[nxt], h->logw[nxt], h->cap,
                                                                              h->tab, h->pc, h->stc, h->st, h->res, h->V, h->surv_parent,
                                                                              h->surv_slot, ga, gs, h->ro, h->k_max, h->sc, &h->st->M_next, h->peers[nxt],
                                                                              (long long)h->steps * h->cap);
    }
    {
        LaunchTimer lt(h, KC_STEPLOG);
        k_steplog<<<1, 32, 0, h->stream>>>(h->st, h->res, log_dev, h->sc, h->steps, h->dbg_status, h->ro, h->xch, seq);
    }
    CUDA_TRY(cudaGetLastError());
    h->cur = nxt;
    h->steps++;
    h->pre_expanded = false;
    return PF_OK;
}

// Pipelined tail of a sharded step, in three phases (a lock-step group queues each phase for all its shards before
// the next: the bracket's exchange is published in phase 1 and awaited in phase 2):
//   1  step record; roll the control block to the next observation; count its putatives; sample expansion of this
//      rank's sampled children; publish the sample
//   2  bracket on the gathered sample
//   3  instantiate + expansion + classification of every child this rank produced, each delivered with its putatives
//      and scan record to its new owner; "generation delivered"
template <int D, typename R>
static int32_t mg_tail(pf_handle h, int phase, StepLog *log_dev, const double *y_next, const double *u_next, int search_stage,
                       const double *y_host /* this observation and the next */)
{
    const long long cap = h->cap;
    if (phase == 1) {
        {
            LaunchTimer lt(h, KC_STEPLOG);
            k_steplog<<<1, 32, 0, h->stream>>>(h->st, h->res, log_dev, h->sc, h->steps, h->dbg_status, h->ro, nullptr, 0);
        }
        h->steps++;                                   // the host mirrors now describe the next step; h->cur still names the parents' store
        LaunchTimer lt(h, KC_PRESTEP, 2);
        k_prestep<D><<<1, 32, 0, h->stream>>>(h->st, h->pc, y_next, h->stc_alt, 3, h->sc, h->res, nullptr, 0, h->steps, -1);
        k_count_next<<<std::min(grid_for(cap, 1024), h->num_sms * 16), 256, 0, h->stream>>>(h->st, h->Kc[h->cur], h->surv_parent, h->surv_slot, h->k_max,
                                                                                          &h->st->M_pred, h->sc, h->ro);
    }
    const int cur = h->cur, nxt = cur ^ 1;
    InstArgs ia{};
    ia.surv_parent = h->surv_parent; ia.surv_slot = h->surv_slot; ia.V_prev = h->V; ia.scp_prev = h->stc;
    ia.S2 = h->S[nxt]; ia.Kc2 = h->Kc[nxt]; ia.logw2 = h->logw[nxt];
    ia.gen_off = (long long)(h->steps - 1) * cap;
    if (h->hist_cap) { ia.gen_anc = h->gen_anc + ia.gen_off; ia.gen_asg = h->gen_asg + ia.gen_off; }
    ia.ro = h->ro; ia.peers = h->peers[nxt]; ia.vnext = h->vcur ^ 1;
    for (int k = 0; k < h->d; k++) { ia.y_prev[k] = y_host[k]; ia.yc[k] = y_host[h->d + k]; ia.mu0c[k] = h->mu0[k]; }
    if (phase == 1 || phase == 2) {
        LaunchTimer lt(h, KC_SELECT, phase == 1 ? 2 : 1);
        if (phase == 1) {
            const long long ns = 4 * ((h->surv_cap + 4 * h->fuse_stride - 1) / (4 * h->fuse_stride) + 1);      // groups of 4 children (+1: the shard's phase)
            k_expand<D, 1, true, R, true><<<grid_for(ns, EXP_SAMPLE_THREADS), EXP_SAMPLE_THREADS, 0, h->stream>>>((const R *)h->S[cur], nullptr, h->Kc[cur], cap, h->k_max, h->tab, h->pc,
                                                                       h->stc_alt, h->st, h->V_alt, h->cnt, h->sc, h->fuse_stride, h->cand, nullptr,
                                                                       nullptr, ScanRec{nullptr, nullptr, nullptr}, ia);
        }
        if ((phase == 1 && (search_stage & SEL_ST_PUB)) || (phase == 2 && search_stage == SEL_ST_RUN)) {
            std::swap(h->stc, h->stc_alt);            // (sel_args points the search at the next step's plateau value)
            SelArgs a = sel_args(h, h->V_alt, h->cnt, cap, &h->st->n_live, &h->st->M_pred, &h->st->vmax_sample, h->N, u_next, h->cand, h->cand_cap, h->sc, h->res);
            std::swap(h->stc, h->stc_alt);
            a.external_sample = 1; a.x = h->xch; a.stage = search_stage;
            CUDA_TRY(launch_coop(k_sel_bracket, h->coop_grid, h->stream, a));
        }
        return PF_OK;
    }
    {
        LaunchTimer lt(h, KC_INST_EXPAND);
        // main launch: the tiles a balanced share needs (+ 25 %); a small strided launch takes whatever lies beyond
        const int gmain = grid_for(std::min<long long>(h->surv_cap, cap + cap / 4), EXP_THREADS);
        constexpr size_t stage_bytes = instexp_staging_bytes<D, R>();          // asynchronous-copy staging of the slot loop
        if (stage_bytes && !h->instexp_attr) {
            CUDA_TRY(cudaFuncSetAttribute(k_expand<D, 2, true, R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
            CUDA_TRY(cudaFuncSetAttribute(k_expand<D, 2, true, R, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
            h->instexp_attr = true;
        }
        k_expand<D, 2, true, R, true><<<gmain, EXP_THREADS, stage_bytes, h->stream>>>((const R *)h->S[cur], nullptr, h->Kc[cur], cap, h->k_max, h->tab, h->pc,
                                                                   h->stc_alt, h->st, h->V_alt, h->cnt, h->sc, 1, h->cand, nullptr, nullptr, h->srec, ia);
        if ((long long)gmain * EXP_THREADS < h->surv_cap) {
            ia.vb0 = gmain;
            k_expand<D, 2, true, R, true, true><<<h->num_sms, EXP_THREADS, stage_bytes, h->stream>>>((const R *)h->S[cur], nullptr, h->Kc[cur], cap, h->k_max, h->tab, h->pc,
                                                                   h->stc_alt, h->st, h->V_alt, h->cnt, h->sc, 1, h->cand, nullptr, nullptr, h->srec, ia);
            h->launches++;
        }
    }
    {
        LaunchTimer lt(h, KC_STEPLOG);
        k_mg_signal_generation<<<1, 32, 0, h->stream>>>(h->xch, (unsigned long long)h->steps, h->sc);
    }
    CUDA_TRY(cudaGetLastError());
    std::swap(h->V, h->V_alt); h->vcur ^= 1;
    std::swap(h->stc, h->stc_alt);
    h->cur = nxt;
    h->pre_expanded = true;
    return PF_OK;
}

static const char *halt_text(int reason)
{
    switch (reason) {
    case PF_HALT_SEARCH: return "the threshold search needed more rounds than were queued";
    case PF_HALT_TIMEOUT: return "a peer's exchange flag did not arrive in time (PF_MG_TIMEOUT_MS)";
    case PF_HALT_XCH_OVERFLOW: return "a sample / candidate list outgrew its exchange segment";
    case PF_HALT_SURVIVORS: return "a shard produced more survivors than its survivor list holds";
    default: return "unknown";
    }
}

// filter!(ps, ys, false) (src/filters.jl:6-13) on a sharded population: hs[0..n_local) are the shards this
// process drives; every process of the group calls this with the same observations and uniforms.
extern "C" int32_t pf_mg_filter(pf_handle *hs, int32_t n_local, const double *ys, int64_t T, const double *uniforms,
                                int64_t uniforms_per_step, double *log_evidence)
{
    if (!hs || n_local < 1) return PF_ERR_ARG;
    for (int i = 0; i < n_local; i++) {
        MG_CHECK(hs[i]);
        if (!hs[i]->connected) return set_error(hs[i], PF_ERR_STATE, "pf_mg_filter: pf_mg_connect has not been called");
    }
    if (T < 0 || (!ys && T > 0)) return set_error(hs[0], PF_ERR_ARG, "pf_mg_filter: bad ys/T");
    const bool cl = hs[0]->cfg.filter_kind == PF_CHENLIU;
    if (uniforms && uniforms_per_step < (cl ? 2 * hs[0]->N : 1))
        return set_error(hs[0], PF_ERR_ARG, cl ? "pf_mg_filter: Chen-Liu needs 2*N uniforms per step" : "pf_mg_filter: Fearnhead needs 1 uniform per step");
    if (T == 0) return PF_OK;
    const int d = hs[0]->d;
    const bool lockstep = hs[0]->lockstep;
    int32_t rc = PF_OK;
    for (int i = 0; i < n_local; i++) {
        pf_handle h = hs[i];
        CUDA_TRY(cudaSetDevice(h->cfg.device));
        rc = ensure_table(h, h->steps + T + 2);
        if (rc != PF_OK) return rc;
        if (h->hist_cap && h->steps + T > h->hist_cap) return set_error(h, PF_ERR_HISTORY, "history capacity exhausted (pf_config.history)");
        if (T * d > h->ys_cap) { if (h->ys_dev) cudaFree(h->ys_dev); h->ys_cap = T * d; CUDA_TRY(dmalloc(&h->ys_dev, (size_t)h->ys_cap)); }
        if (T > h->slog_cap) { if (h->slog) cudaFree(h->slog); h->slog_cap = T; CUDA_TRY(dmalloc(&h->slog, (size_t)T)); }
        const long long un = cl ? std::max<long long>(T, 2 * h->shard.n_loc) : T;
        if (un > h->u_cap) { if (h->u_dev) cudaFree(h->u_dev); h->u_cap = un; CUDA_TRY(dmalloc(&h->u_dev, (size_t)un)); }
        std::vector<double> ubuf(T);
        for (long long t = 0; t < T; t++)
            ubuf[t] = uniforms ? uniforms[t * uniforms_per_step] : pf_uniform(h->cfg.seed, (uint64_t)(h->steps + t), 0, 0);
        h->launches = 0;
        CUDA_TRY(cudaMemcpyAsync(h->ys_dev, ys, sizeof(double) * T * d, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaMemcpyAsync(h->u_dev, ubuf.data(), sizeof(double) * T, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));      // ubuf goes out of scope
        CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    }
#define FOR_SHARDS(...) for (int i_ = 0; i_ < n_local && rc == PF_OK; i_++) { pf_handle h = hs[i_]; cudaSetDevice(h->cfg.device); DISPATCH_D(d, rc = __VA_ARGS__); }
    const int both = SEL_ST_PUB | SEL_ST_RUN;
    for (long long t = 0; t < T && rc == PF_OK; t++) {
        if (cl) {
            // (uniforms[2N] of the step: [0, N) categorical draws, [N, 2N) rejuvenation; every shard takes its block)
            if (uniforms)
                for (int i = 0; i < n_local; i++) {
                    pf_handle h = hs[i];
                    cudaSetDevice(h->cfg.device);
                    const double *ut = uniforms + t * uniforms_per_step;
                    cudaMemcpyAsync(h->u_dev, ut + h->shard.goff, sizeof(double) * h->shard.n_loc, cudaMemcpyHostToDevice, h->stream);
                    cudaMemcpyAsync(h->u_dev + h->shard.n_loc, ut + h->N + h->shard.goff, sizeof(double) * h->shard.n_loc, cudaMemcpyHostToDevice, h->stream);
                }
            for (int st = 0; st < 6; st++) FOR_SHARDS((chenliu_stage<D, R>(h, st, h->ys_dev + t * d, uniforms ? h->u_dev : nullptr, h->slog + t)));
            continue;
        }
        // A step whose expansion ran in the previous step's pipelined tail starts at the search, once every rank's
        // children (with their putatives and scan records) have been delivered
        const bool pre = hs[0]->pre_expanded;
        if (pre) {
            for (int i = 0; i < n_local && rc == PF_OK; i++) {
                pf_handle h = hs[i];
                cudaSetDevice(h->cfg.device);
                LaunchTimer lt(h, KC_PRESTEP);
                k_mg_await_generation<<<1, 32, 0, h->stream>>>(h->xch, (unsigned long long)h->steps, h->steps, h->sc);
            }
        }
        if (lockstep) {
            if (!pre) {
                FOR_SHARDS((mg_stage_a<D, R>(h, h->ys_dev + t * d, h->u_dev + t, SEL_ST_PUB)));
                FOR_SHARDS((mg_stage_a<D, R>(h, h->ys_dev + t * d, h->u_dev + t, SEL_ST_RUN)));
            }
            for (int r = 0; r <= SEL_RETRIES; r++) {
                FOR_SHARDS((mg_stage_b<D, R>(h, h->u_dev + t, r > 0, SEL_ST_PUB)));
                FOR_SHARDS((mg_stage_b<D, R>(h, h->u_dev + t, r > 0, SEL_ST_RUN)));
            }
        } else {
            if (!pre) FOR_SHARDS((mg_stage_a<D, R>(h, h->ys_dev + t * d, h->u_dev + t, both)));
            for (int r = 0; r <= SEL_RETRIES; r++) FOR_SHARDS((mg_stage_b<D, R>(h, h->u_dev + t, r > 0, both)));
        }
        for (int i = 0; i < n_local && rc == PF_OK; i++) { cudaSetDevice(hs[i]->cfg.device); rc = mg_stage_c(hs[i]); }
        for (int i = 0; i < n_local && rc == PF_OK; i++) { cudaSetDevice(hs[i]->cfg.device); rc = mg_stage_d_scan(hs[i]); }
        const bool next_ok = hs[0]->pipeline_mg && t + 1 < T;
        if (next_ok) {
            if (lockstep) {
                FOR_SHARDS((mg_tail<D, R>(h, 1, h->slog + t, h->ys_dev + (t + 1) * d, h->u_dev + t + 1, SEL_ST_PUB, ys + t * d)));
                FOR_SHARDS((mg_tail<D, R>(h, 2, h->slog + t, h->ys_dev + (t + 1) * d, h->u_dev + t + 1, SEL_ST_RUN, ys + t * d)));
            } else {
                FOR_SHARDS((mg_tail<D, R>(h, 1, h->slog + t, h->ys_dev + (t + 1) * d, h->u_dev + t + 1, both, ys + t * d)));
            }
            FOR_SHARDS((mg_tail<D, R>(h, 3, h->slog + t, h->ys_dev + (t + 1) * d, h->u_dev + t + 1, both, ys + t * d)));
        } else {
            FOR_SHARDS((mg_stage_d<D, R>(h, h->slog + t)));
        }
    }
#undef FOR_SHARDS
    // drain: every shard's records; a halted shard reports why
    int32_t first_err = rc;
    for (int i = 0; i < n_local; i++) {
        pf_handle h = hs[i];
        cudaSetDevice(h->cfg.device);
        h->logs.resize(T);
        long long halt[2] = {0, 0};
        StepState st;
        cudaError_t e = cudaMemcpyAsync(h->logs.data(), h->slog, sizeof(StepLog) * T, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(halt, &h->sc->halt, sizeof halt, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(&st, h->st, sizeof st, cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaEventRecord(h->ev1, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) { if (first_err == PF_OK) first_err = set_cuda_error(h, e, "pf_mg_filter drain", __LINE__); continue; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, h->ev0, h->ev1);
        h->last_ms = ms; h->last_launches = h->launches;
        collect_profile(h);
        if (halt[0]) {
            char buf[200];
            const int reason = (int)(halt[1] & 0xffffffffll);
            snprintf(buf, sizeof buf, "sharded step %lld halted on rank %d: %s", halt[0] - 1, h->rank, halt_text(reason));
            const int32_t code = (reason == PF_HALT_XCH_OVERFLOW || reason == PF_HALT_SURVIVORS) ? PF_ERR_OVERFLOW : PF_ERR_STATE;
            set_error(h, code, buf);
            if (first_err == PF_OK) { first_err = code; if (h != hs[0]) hs[0]->err = buf; }
            continue;
        }
        h->last = h->logs.back(); h->have_last = true; h->last_stale = false;
        h->n_host = st.n_next;
        if (cl) { ClState cs; if (cudaMemcpy(&cs, h->cl, sizeof cs, cudaMemcpyDeviceToHost) == cudaSuccess) h->cur = cs.cur; }
        for (long long t = 0; t < T; t++) h->gen_q.push_back(std::max<long long>(1, (h->logs[t].n_out + h->nranks - 1) / h->nranks));
    }
    if (first_err != PF_OK) return first_err;
    if (log_evidence) for (long long t = 0; t < T; t++) log_evidence[t] = hs[0]->logs[t].log_total;
    return PF_OK;
}

// assignments(ps) (src/filters.jl:26-34, src/particle.jl:25-32) of THIS shard's particles: the ancestor chain of a
// particle runs through the genealogy arrays of whichever rank owned each ancestor (peer memory); generation t's
// parent index is global in the population after step t - 1, whose owner blocks have length q[t - 1]
__global__ void k_mg_wait_generation(const Xch *x, unsigned long long seq, int *ok)
{
    if (threadIdx.x || blockIdx.x) return;
    *ok = xch_wait(x, XCH_D, seq) ? 1 : 0;
}
__global__ void k_backtrack8_mg(const PeerTable *__restrict__ peers, long long cap, long long T, const long long *__restrict__ q,
                                int rank, long long i0, long long nc, unsigned char *__restrict__ out8)
{
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    int r = rank;
    long long k = i0 + c;
    for (long long t = T - 1; t >= 0; t--) {
        out8[t * nc + c] = peers->gen_asg[r][t * cap + k];
        const long long g = peers->gen_anc[r][t * cap + k];
        if (t > 0) { const long long qq = q[t - 1]; r = (int)(g / qq); k = g - (long long)r * qq; }
    }
}

extern "C" int32_t pf_mg_get_assignments(pf_handle h, int32_t *out)
{
    MG_CHECK(h);
    if (!out) return PF_ERR_ARG;
    if (!h->connected) return set_error(h, PF_ERR_STATE, "pf_mg_get_assignments: pf_mg_connect has not been called");
    if (!h->hist_cap) return set_error(h, PF_ERR_HISTORY, "pf_mg_get_assignments: history is disabled (pf_config.history = 0)");
    const long long T = h->steps, n = h->n_host;
    if ((long long)h->gen_q.size() != T) return set_error(h, PF_ERR_STATE, "pf_mg_get_assignments: the last sharded call did not complete");
    if (T == 0 || n == 0) return PF_OK;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const long long nc_max = assign_chunk(T, n);
    DevBuf dq, dtmp, dok;
    CUDA_TRY(dq.alloc(sizeof(long long) * T));
    CUDA_TRY(dtmp.alloc((size_t)T * nc_max));
    CUDA_TRY(dok.alloc(sizeof(int)));
    CUDA_TRY(cudaMemcpyAsync(dq.p, h->gen_q.data(), sizeof(long long) * T, cudaMemcpyHostToDevice, h->stream));
    // every rank has finished its last generation (peers stored children and genealogy entries into this shard, and
    // this shard reads theirs)
    k_mg_wait_generation<<<1, 32, 0, h->stream>>>(h->xch, (unsigned long long)T, dok.as<int>());
    int ok = 0;
    CUDA_TRY(cudaMemcpyAsync(&ok, dok.p, sizeof ok, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (!ok) return set_error(h, PF_ERR_STATE, "pf_mg_get_assignments: a peer has not finished its last generation");
    std::vector<unsigned char> host((size_t)T * nc_max);
    for (long long i0 = 0; i0 < n; i0 += nc_max) {
        const long long nc = std::min(nc_max, n - i0);
        k_backtrack8_mg<<<grid_for(nc, 256), 256, 0, h->stream>>>(h->peers[0], h->cap, T, dq.as<long long>(), h->rank, i0, nc, dtmp.as<unsigned char>());
        CUDA_TRY(cudaMemcpyAsync(host.data(), dtmp.p, (size_t)T * nc, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        for (long long c = 0; c < nc; c++)
            for (long long t = 0; t < T; t++) out[(i0 + c) * T + t] = host[(size_t)t * nc + c];
    }
    return PF_OK;
}

// Filter-level summaries of a SHARDED population (src/filters.jl:17-42, 53-80): every shard returns the unnormalised
// sums over its own particles; added up over the shards (any order: the caller's host-side gather) they give
//   pred[q] / pred[Q] / (steps + alpha)        pdf of posterior_predictive(ps) at x_q          (pred[Q] = sum of weights)
//   pop[1] / pop[0], pop[2] / pop[0]           mean(ncomponents, ps), state_entropy(ps)        (pop[0] = sum of weights)
//   pop[3 + k] / pop[0]                        ncomponents_dist(ps)
//   sim[s * T + t] / n                         assignment_similarity(ps)                        (counts of agreeing particles)
// Any of pred (with xs, Q), pop, sim may be NULL.
extern "C" int32_t pf_mg_summary_partials(pf_handle h, const double *xs, int64_t Q, double *pred /* [Q + 1] */, double *pop /* [3 + k_max + 1] */,
                                          uint64_t *sim /* [T x T] */)
{
    MG_CHECK(h);
    if (!h->connected) return set_error(h, PF_ERR_STATE, "pf_mg_summary_partials: pf_mg_connect has not been called");
    int32_t rc = PF_OK;
    if (pred) {
        if (Q < 0 || (Q > 0 && !xs)) return PF_ERR_ARG;
        rc = predictive_sums(h, xs, Q, pred);
        if (rc != PF_OK) return rc;
    }
    if (pop) {
        std::vector<double> rec(3 + h->k_max + 1, 0.0);
        if (h->n_host > 0) { rc = pop_summary(h, rec); if (rc != PF_OK) return rc; }
        std::copy(rec.begin(), rec.end(), pop);
    }
    if (sim) {
        if (!h->hist_cap) return set_error(h, PF_ERR_HISTORY, "pf_mg_summary_partials: history is disabled (pf_config.history = 0)");
        const long long T = h->steps, n = h->n_host;
        if ((long long)h->gen_q.size() != T) return set_error(h, PF_ERR_STATE, "pf_mg_summary_partials: the last sharded call did not complete");
        for (long long k = 0; k < T * T; k++) sim[k] = 0;
        if (T == 0 || n == 0) return PF_OK;
        CUDA_TRY(cudaSetDevice(h->cfg.device));
        const long long nc_max = assign_chunk(T, n);
        DevBuf dq, dA, dC, dok;
        CUDA_TRY(dq.alloc(sizeof(long long) * T));
        CUDA_TRY(dA.alloc((size_t)T * nc_max));
        CUDA_TRY(dC.alloc(sizeof(unsigned long long) * T * T));
        CUDA_TRY(dok.alloc(sizeof(int)));
        CUDA_TRY(cudaMemcpyAsync(dq.p, h->gen_q.data(), sizeof(long long) * T, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaMemsetAsync(dC.p, 0, sizeof(unsigned long long) * T * T, h->stream));
        k_mg_wait_generation<<<1, 32, 0, h->stream>>>(h->xch, (unsigned long long)T, dok.as<int>());
        int ok = 0;
        CUDA_TRY(cudaMemcpyAsync(&ok, dok.p, sizeof ok, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        if (!ok) return set_error(h, PF_ERR_STATE, "pf_mg_summary_partials: a peer has not finished its last generation");
        const long long ntile = (T + SIM_TILE - 1) / SIM_TILE, npair = ntile * (ntile + 1) / 2;
        for (long long i0 = 0; i0 < n; i0 += nc_max) {
            const long long nc = std::min(nc_max, n - i0);
            k_backtrack8_mg<<<grid_for(nc, 256), 256, 0, h->stream>>>(h->peers[0], h->cap, T, dq.as<long long>(), h->rank, i0, nc, dA.as<unsigned char>());
            long long gy = std::max<long long>(1, std::min<long long>((nc / 16 + SUM_THREADS * 4 - 1) / (SUM_THREADS * 4), 4096));
            k_similarity<<<dim3((unsigned)npair, (unsigned)gy), SUM_THREADS, 0, h->stream>>>(dA.as<unsigned char>(), T, nc, dC.as<unsigned long long>());
        }
        static_assert(sizeof(uint64_t) == sizeof(unsigned long long), "");
        CUDA_TRY(cudaMemcpyAsync(sim, dC.p, sizeof(unsigned long long) * T * T, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    return PF_OK;
}

// population over all shards after the last step
extern "C" int64_t pf_mg_n_global(pf_handle h) { return (h && h->have_last) ? h->last.n_out : (h ? 1 : -1); }

// Convenience for a single process that drives several GPUs (one host thread): creates one shard per entry
// of `devices`, maps them into each other (peer access) and connects them.  cfg->rank / nranks / device are
// ignored.  The same device may be listed several times (lock-step emulation, used by the tests).
extern "C" int32_t pf_mg_create_local(const pf_config *cfg, const int32_t *devices, int32_t n, pf_handle *out)
{
    if (!cfg || !devices || !out || n < 2 || n > PF_MAX_RANKS) return set_error(nullptr, PF_ERR_ARG, "pf_mg_create_local: bad argument");
    for (int i = 0; i < n; i++) out[i] = nullptr;
    int32_t rc = PF_OK;
    for (int i = 0; i < n && rc == PF_OK; i++) {
        pf_config c = *cfg;
        c.rank = i; c.nranks = n; c.device = devices[i];
        rc = pf_create(&c, &out[i]);
    }
    for (int i = 0; i < n && rc == PF_OK; i++)
        for (int j = 0; j < n && rc == PF_OK; j++) {
            void *arr[PF_MG_NARRAYS];
            pf_mg_local_arrays(out[j], arr);
            rc = pf_mg_set_peer(out[i], j, arr, nullptr);
            if (rc != PF_OK) g_create_error = out[i]->err;
        }
    if (rc == PF_OK) { rc = pf_mg_connect(out, n); if (rc != PF_OK) for (int i = 0; i < n; i++) if (!out[i]->err.empty()) g_create_error = out[i]->err; }
    if (rc != PF_OK) { for (int i = 0; i < n; i++) { if (out[i]) pf_destroy(out[i]); out[i] = nullptr; } }
    return rc;
}

extern "C" int32_t pf_debug_dump(pf_handle h, int64_t *out /* [24] */)
{
    if (!h || !out) return PF_ERR_ARG;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    SelResult r; StepState st; RankOff ro; SelScratch *scp = h->sc;
    long long hdr[8];
    memset(&ro, 0, sizeof ro);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaMemcpy(&r, h->res, sizeof r, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(&st, h->st, sizeof st, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(hdr, scp, sizeof hdr, cudaMemcpyDeviceToHost));
    if (h->ro) CUDA_TRY(cudaMemcpy(&ro, h->ro, sizeof ro, cudaMemcpyDeviceToHost));
    int64_t v[24] = {(int64_t)r.thr_bits, r.scale, r.keep_all, (int64_t)r.T.lo, (int64_t)r.T.hi, (int64_t)r.Uoff.lo, (int64_t)r.Uoff.hi,
                     (int64_t)r.Tq.lo, (int64_t)r.Tq.hi, r.Tr, r.A, r.n, r.rounds, st.n_live, st.M, st.n_next,
                     r.status, hdr[2] /* scale | phase */, hdr[5] /* halt */, ro.n_local, ro.g_first, ro.kept_off, (int64_t)ro.S_off.lo, ro.n_global};
    for (int k = 0; k < 24; k++) out[k] = v[k];
    return PF_OK;
}
