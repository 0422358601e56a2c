/*
 * particles_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A literal, single-threaded restatement in plain C of the hot path of
 * kleinschmidt/Particles.jl (the online particle-filter step for Dirichlet-process
 * mixture clustering).  It exists so that the CUDA path can be checked against the
 * reference's algorithm on a machine that has no Julia.  Only tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may
 * load this file's shared object; the product library never links or calls it.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference repository root).  Arithmetic that lives in the reference's un-vendored
 * third-party dependencies is restated from the published algorithm and named here:
 *   - ConjugatePriors.jl 0.3.2 (Manifest.toml:30-34): posterior_canon for
 *     NormalInverseChisq / NormalInverseWishart (Murphy 2007, "Conjugate Bayesian
 *     analysis of the Gaussian distribution", eqs. 141-144 and 250-254).
 *   - Distributions.jl 0.19.2 (Manifest.toml:60-64): NormalStats / MvNormalStats field
 *     meaning (src/component.jl:36-41).
 *   - StatsFuns.jl 0.8.0 (Manifest.toml:181-185): logmvgamma.
 *   - SpecialFunctions.jl 0.7.2 (Manifest.toml:165-169): lgamma (openlibm) -> libm lgamma.
 *   - StatsBase.jl 0.31.0 (Manifest.toml:175-179): sample(::Weights) inverse-CDF draw,
 *     mean_and_std.
 *
 * PINNING STATUS.  The arithmetic (Welford update, marginal likelihoods, CRP prior,
 * weight update, cutoff_ascending) is pinned by the reference's own known-answer and
 * identity tests, ported in tests/test_oracle_*.py (test/component.jl:7-26,
 * test/particle.jl:31-54, test/niw.jl:44-67, test/fearnhead.jl:36-71,
 * test/statepriors.jl:4-30).  WHICH items the samplers pick is pinned by nothing in the
 * reference (test/fearnhead.jl:75-89 only bounds the count; Julia's MersenneTwister
 * stream is not reproducible here): for sampler outputs this oracle is "parity
 * unpinned" -- it is a statement-by-statement restatement with host-supplied uniforms.
 *
 * Random numbers: every rand() site of the reference takes a caller-supplied uniform.
 * Sort: the reference's sort!(alg=QuickSort, by=weight) is unstable
 * (src/fearnhead.jl:148); the declared tie-break here is the putative index (stable).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_DMAX 16
#define ORC_LOGPI 1.1447298858494001741434273513530587116472948129153

enum { ORC_NIX2 = 0, ORC_NIW = 1 };
enum { ORC_FEARNHEAD = 0, ORC_CHENLIU = 1 };
enum { ORC_ORDER_SORTED = 0, ORC_ORDER_INDEX = 1 };
enum { ORC_SP_CRP = 0, ORC_SP_STICKY = 1, ORC_SP_CHANGEPOINT = 2, ORC_SP_NSTATE = 3 };

typedef struct {
    int kind, d;
    double mu0[ORC_DMAX];
    double kappa0, nu0;
    double s20;                       /* NIX2 sigma^2_0 */
    double U0[ORC_DMAX * ORC_DMAX];   /* NIW: upper Cholesky factor of Lambda_0, column-major */
} orc_prior;

/* sufficient statistics of one component, src/component.jl:36-41:
 * layout [tw, s[d], m[d], s2[d*d]] (s2 column-major) */
static int ss_len(int d) { return 1 + 2 * d + d * d; }

/* ------------------------------------------------------------------ linear algebra */

/* upper Cholesky A = U'U of a symmetric matrix given by its UPPER triangle
 * (Julia cholesky(Hermitian(A)) -> LAPACK potrf('U')); column-major d x d. */
static int chol_upper(int d, const double *A, double *U)
{
    memset(U, 0, sizeof(double) * d * d);
    for (int j = 0; j < d; j++) {
        for (int i = 0; i <= j; i++) {
            double acc = A[i + j * d];
            for (int k = 0; k < i; k++) acc -= U[k + i * d] * U[k + j * d];
            if (i == j) {
                if (!(acc > 0.0)) return -1;
                U[i + j * d] = sqrt(acc);
            } else {
                U[i + j * d] = acc / U[i + i * d];
            }
        }
    }
    return 0;
}

/* logdet(::Cholesky) = 2 * sum(log(diag)) (LinearAlgebra) */
static double logdet_chol(int d, const double *U)
{
    double dd = 0.0;
    for (int i = 0; i < d; i++) dd += log(U[i + i * d]);
    return dd + dd;
}

/* StatsFuns.logmvgamma(p, a) = p(p-1)/4 log(pi) + sum_{i=1..p} lgamma(a + (1-i)/2) */
static double logmvgamma(int p, double a)
{
    double res = p * (p - 1) / 4.0 * ORC_LOGPI;
    for (int ii = 1; ii <= p; ii++) res += lgamma(a + (1.0 - ii) / 2.0);
    return res;
}

/* ------------------------------------------------------------------ component.jl */

/* empty_suffstats, src/component.jl:30-34 */
void orc_empty_stats(int d, double *ss) { memset(ss, 0, sizeof(double) * ss_len(d)); }

/* add(ss::NormalStats, x) src/component.jl:44-50 ; add(ss::MvNormalStats, x) :64-70.
 * d-D: s2 = ss.s2 + D*(x.-m)' then Hermitian(s2) -> the UPPER triangle is authoritative;
 * we mirror the upper triangle into the lower one so later reads see a symmetric matrix. */
void orc_stats_add(int d, const double *in, const double *x, double *out)
{
    const double *s = in + 1, *m = in + 1 + d, *s2 = in + 1 + 2 * d;
    double *so = out + 1, *mo = out + 1 + d, *s2o = out + 1 + 2 * d;
    double tw = in[0] + 1.0;
    double delta[ORC_DMAX], mn[ORC_DMAX];
    for (int i = 0; i < d; i++) {
        delta[i] = x[i] - m[i];
        mn[i] = m[i] + delta[i] / tw;
    }
    for (int j = 0; j < d; j++)
        for (int i = 0; i < d; i++)
            s2o[i + j * d] = tw > 1.0 ? s2[i + j * d] + delta[i] * (x[j] - mn[j]) : 0.0;
    for (int j = 0; j < d; j++)      /* Hermitian(): upper triangle wins */
        for (int i = j + 1; i < d; i++) s2o[i + j * d] = s2o[j + i * d];
    for (int i = 0; i < d; i++) { so[i] = s[i] + x[i]; mo[i] = mn[i]; }
    out[0] = tw;
}

/* sub(ss, x) src/component.jl:52-62, :72-83 (Gibbs only; kept for the known-answer tests) */
void orc_stats_sub(int d, const double *in, const double *x, double *out)
{
    double tw = in[0] - 1.0;
    if (tw <= 0.0) { orc_empty_stats(d, out); return; }
    const double *s = in + 1, *m = in + 1 + d, *s2 = in + 1 + 2 * d;
    double *so = out + 1, *mo = out + 1 + d, *s2o = out + 1 + 2 * d;
    double delta[ORC_DMAX], mn[ORC_DMAX];
    for (int i = 0; i < d; i++) {
        delta[i] = x[i] - m[i];
        mn[i] = m[i] - delta[i] / tw;
    }
    for (int j = 0; j < d; j++)
        for (int i = 0; i < d; i++) s2o[i + j * d] = s2[i + j * d] - delta[i] * (x[j] - mn[j]);
    for (int i = 0; i < d; i++) { so[i] = s[i] - x[i]; mo[i] = mn[i]; }
    out[0] = tw;
}

/* posterior_canon(::NormalInverseChisq, ::NormalStats) -- ConjugatePriors 0.3.2.
 * out = (mu_n, sigma2_n, kappa_n, nu_n)  [params order, src/component.jl:123-124] */
void orc_posterior_nix2(const orc_prior *p, const double *ss, double out[4])
{
    double tw = ss[0], s = ss[1], m = ss[2], s2 = ss[3];
    double kn = p->kappa0 + tw, nn = p->nu0 + tw;
    double dm = p->mu0[0] - m;
    out[0] = (p->kappa0 * p->mu0[0] + s) / kn;
    out[1] = (p->nu0 * p->s20 + s2 + tw * p->kappa0 / (tw + p->kappa0) * (dm * dm)) / nn;
    out[2] = kn;
    out[3] = nn;
}

/* posterior_canon(::NormalInverseWishart, ::MvNormalStats) -- ConjugatePriors 0.3.2.
 * Lam_n = Lam_0 + s2 + kappa0*tw/kappa_n * z z',  z = m - mu0 ; returns cholesky(Lam_n).
 * out_mu[d], out_U[d*d] (upper factor), *kn, *nn. */
int orc_posterior_niw(const orc_prior *p, const double *ss, double *out_mu, double *out_U,
                      double *kn_out, double *nn_out)
{
    int d = p->d;
    double tw = ss[0];
    const double *s = ss + 1, *m = ss + 1 + d, *s2 = ss + 1 + 2 * d;
    double kn = p->kappa0 + tw, nn = p->nu0 + tw;
    double Lam[ORC_DMAX * ORC_DMAX], z[ORC_DMAX];
    for (int i = 0; i < d; i++) {
        out_mu[i] = (p->kappa0 * p->mu0[i] + s[i]) / kn;
        z[i] = m[i] - p->mu0[i];
    }
    double c = p->kappa0 * tw / kn;
    for (int j = 0; j < d; j++)
        for (int i = 0; i < d; i++) {
            double l0 = 0.0;                       /* Lam0 = U0' * U0 */
            for (int k = 0; k < d; k++) l0 += p->U0[k + i * d] * p->U0[k + j * d];
            Lam[i + j * d] = l0 + s2[i + j * d] + c * (z[i] * z[j]);
        }
    *kn_out = kn;
    *nn_out = nn;
    return chol_upper(d, Lam, out_U);
}

/* marginal_log_lhood, src/component.jl:122-130 (NIX2) and :132-142 (NIW) */
double orc_mll(const orc_prior *p, const double *ss)
{
    if (p->kind == ORC_NIX2) {
        double post[4];
        orc_posterior_nix2(p, ss, post);
        double s2n = post[1], kn = post[2], nn = post[3];
        double n = nn - p->nu0;
        return lgamma(nn * 0.5) - lgamma(p->nu0 * 0.5) +
               0.5 * (log(p->kappa0) - log(kn)) +
               (0.5 * p->nu0) * (log(p->nu0) + log(p->s20)) - (0.5 * nn) * (log(nn) + log(s2n)) -
               (0.5 * n) * log(M_PI);
    } else {
        int d = p->d;
        double mu[ORC_DMAX], U[ORC_DMAX * ORC_DMAX], kn, nn;
        if (orc_posterior_niw(p, ss, mu, U, &kn, &nn) != 0) return NAN;
        double n = nn - p->nu0;
        return (-0.5 * n * d) * ORC_LOGPI +
               logmvgamma(d, nn * 0.5) - logmvgamma(d, p->nu0 * 0.5) +
               logdet_chol(d, p->U0) * p->nu0 * 0.5 - logdet_chol(d, U) * nn * 0.5 +
               d * 0.5 * (log(p->kappa0) - log(kn));
    }
}

/* log pdf of the posterior predictive, src/component.jl:12-25:
 * NIX2 -> LocationScale(mu_n, sqrt((1+kn) s2n / kn), TDist(nu_n))
 * NIW  -> MvTDist(df = nu - d + 1, mu, Lam (kn+1)/(kn df))
 * Independent second statement of mll(add(c,y)) - mll(c); used by tests only. */
double orc_logpred(const orc_prior *p, const double *ss, const double *y)
{
    if (p->kind == ORC_NIX2) {
        double post[4];
        orc_posterior_nix2(p, ss, post);
        double mu = post[0], s2n = post[1], kn = post[2], nu = post[3];
        double sc = sqrt((1.0 + kn) * s2n / kn);
        double t = (y[0] - mu) / sc;
        return lgamma((nu + 1.0) / 2.0) - lgamma(nu / 2.0) - 0.5 * log(nu * M_PI) - log(sc) -
               (nu + 1.0) / 2.0 * log1p(t * t / nu);
    } else {
        int d = p->d;
        double mu[ORC_DMAX], U[ORC_DMAX * ORC_DMAX], kn, nn;
        if (orc_posterior_niw(p, ss, mu, U, &kn, &nn) != 0) return NAN;
        double df = nn - d + 1.0;
        double f = (kn + 1.0) / (kn * df);           /* Sigma = Lam * f */
        /* z = U'^{-1} (y - mu) ;  quad = |z|^2 / f */
        double z[ORC_DMAX], q = 0.0;
        for (int i = 0; i < d; i++) {
            double acc = y[i] - mu[i];
            for (int k = 0; k < i; k++) acc -= U[k + i * d] * z[k];
            z[i] = acc / U[i + i * d];
            q += z[i] * z[i];
        }
        q /= f;
        double logdetS = logdet_chol(d, U) + d * log(f);
        return lgamma((df + d) / 2.0) - lgamma(df / 2.0) - 0.5 * d * log(df * M_PI) -
               0.5 * logdetS - (df + d) / 2.0 * log1p(q / df);
    }
}

/* ------------------------------------------------------------------ particles */

typedef struct {
    int K;
    double *st;      /* K * ss_len doubles */
    double *crpN;    /* K doubles: ChineseRestaurantProcess.N (src/statepriors.jl:47-50), StickyCRP.N (:95-100),
                        NStatePrior.counts (:203-207) */
    double *Nsame;   /* K doubles: StickyCRP.Nsame (src/statepriors.jl:99) */
    int last;        /* StickyCRP.last (:97), 1-based */
    double total;    /* NStatePrior.total_count (:205) */
    int cp_n;        /* ChangePoint.n (:162-166); ChangePoint.k == K */
    double weight;   /* linear space, src/particle.jl:42 */
    int ancestor;    /* index into the previous generation, -1 for the root */
    int assignment;  /* 1-based component index, 0 for the root (src/particle.jl:62) */
} particle;

typedef struct {
    int parent;      /* index into f->pop */
    int state;       /* 1-based candidate, src/statepriors.jl:54 */
    double weight;   /* exp(logweight), src/particle.jl:114 */
    double *upd;     /* updated component stats, src/particle.jl:111 */
} putative;

typedef struct { int n; int *anc; int *asg; } generation;

typedef struct {
    int filter, N, order;
    orc_prior prior;
    double alpha, rejuv;
    int sp_kind;                 /* ORC_SP_* */
    double sp_rho, sp_logp;      /* StickyCRP.rho ; ChangePoint.logp */
    int sp_n;                    /* NStatePrior.n */
    particle *pop; int n;
    putative *put; int M, put_cap; double *upd_pool; long upd_pool_cap;
    generation *gens; int T, gens_cap;
    int keep_history;
    int forced_label;            /* > 0: the observation being filtered is Labeled(y, label), src/labeled.jl:1-11 */
    /* last-step diagnostics */
    double total_w, w_resamp, cv;
    int n_kept, n_resamp_req, n_resamp_got, kept_i, resampled;
} orc_filter;

static void particle_free(particle *p) { free(p->st); free(p->crpN); free(p->Nsame); }

static void pop_free(particle *pop, int n)
{
    for (int i = 0; i < n; i++) particle_free(&pop[i]);
    free(pop);
}

orc_filter *orc_create(int filter, int N, int kind, int d, const double *mu0, double kappa0,
                       double nu0, double s20, const double *Lam0, double alpha, double rejuv,
                       int order, int keep_history)
{
    if (d < 1 || d > ORC_DMAX) return NULL;
    orc_filter *f = calloc(1, sizeof(*f));
    f->filter = filter; f->N = N; f->order = order; f->alpha = alpha; f->rejuv = rejuv;
    f->keep_history = keep_history;
    f->prior.kind = kind; f->prior.d = d; f->prior.kappa0 = kappa0; f->prior.nu0 = nu0;
    f->prior.s20 = s20;
    for (int i = 0; i < d; i++) f->prior.mu0[i] = mu0[i];
    if (kind == ORC_NIW) {
        if (chol_upper(d, Lam0, f->prior.U0) != 0) { free(f); return NULL; }
    }
    /* FearnheadParticles: ONE empty particle (src/fearnhead.jl:29-30);
     * ChenLiuParticles: N empty particles (src/chenliu.jl:25-26); weight 1.0 (src/particle.jl:62) */
    f->n = filter == ORC_FEARNHEAD ? 1 : N;
    f->pop = calloc(f->n, sizeof(particle));
    for (int i = 0; i < f->n; i++) {
        f->pop[i].weight = 1.0; f->pop[i].ancestor = -1; f->pop[i].assignment = 0;
    }
    return f;
}

void orc_destroy(orc_filter *f)
{
    if (!f) return;
    pop_free(f->pop, f->n);
    free(f->put); free(f->upd_pool);
    if (f->keep_history)
        for (int t = 0; t < f->T; t++) { free(f->gens[t].anc); free(f->gens[t].asg); }
    free(f->gens);
    free(f);
}

/* StatsFuns.log1mexp(x) = x < -log(2) ? log1p(-exp(x)) : log(-expm1(x)) */
static double log1mexp(double x) { return x < -0.6931471805599453 ? log1p(-exp(x)) : log(-expm1(x)); }

/* candidates(stateprior): CRP 1:K+1 (src/statepriors.jl:54); StickyCRP 1:1 if empty else 0:K+1 (:105);
 * ChangePoint max(k,1):k+1 (:177-178); NStatePrior 1:n (:219).  Returns the count, *first = first candidate. */
static int sp_candidates(const orc_filter *f, const particle *p, int *first)
{
    switch (f->sp_kind) {
    case ORC_SP_STICKY: if (p->K == 0) { *first = 1; return 1; } *first = 0; return p->K + 2;
    case ORC_SP_CHANGEPOINT: *first = p->K > 1 ? p->K : 1; return p->K + 1 - *first + 1;
    case ORC_SP_NSTATE: *first = 1; return f->sp_n;
    default: *first = 1; return p->K + 1;
    }
}

/* state_to_index(stateprior, state): identity (src/statepriors.jl:22) except StickyCRP 0 -> last (:103) */
static int sp_state_to_index(const orc_filter *f, const particle *p, int x)
{
    return (f->sp_kind == ORC_SP_STICKY && x == 0) ? p->last : x;
}

/* log_prior(stateprior, x): CRP src/statepriors.jl:65-66; StickyCRP :131-145; ChangePoint :181-183;
 * NStatePrior :220.  NAN = ArgumentError (:141). */
static double sp_log_prior(const orc_filter *f, const particle *p, int x)
{
    switch (f->sp_kind) {
    case ORC_SP_STICKY: {
        if (x == 0) {
            double sumN = 0.0;
            for (int k = 0; k < p->K; k++) sumN += p->crpN[k];
            return log(f->sp_rho / (1.0 - f->sp_rho)) + log(sumN + f->alpha);       /* logit(rho) + log(sum(N) + alpha) */
        }
        if (x >= 1 && x <= p->K) return log(p->crpN[x - 1]);
        if (x == p->K + 1) return log(f->alpha);
        return NAN;
    }
    case ORC_SP_CHANGEPOINT:
        return p->K == 0 ? 0.0 : (x == p->K ? log1mexp(f->sp_logp) : f->sp_logp);
    case ORC_SP_NSTATE:
        return log(p->crpN[x - 1]) - log(p->total);
    default:
        return x == p->K + 1 ? log(f->alpha) : log(p->crpN[x - 1]);
    }
}

/* PutativeParticle(p, obs, state), src/particle.jl:100-116.  upd receives add(comp, obs). */
static double putative_weight(const orc_filter *f, const particle *p, const double *y, int state,
                              double *upd)
{
    int L = ss_len(f->prior.d);
    double logweight = log(p->weight);
    logweight += sp_log_prior(f, p, state);
    int comp_i = sp_state_to_index(f, p, state);
    double empty[1 + 2 * ORC_DMAX + ORC_DMAX * ORC_DMAX];
    const double *comp;
    if (comp_i <= p->K) {
        comp = p->st + (size_t)(comp_i - 1) * L;
        logweight -= orc_mll(&f->prior, comp);
    } else {
        orc_empty_stats(f->prior.d, empty);
        comp = empty;
    }
    orc_stats_add(f->prior.d, comp, y, upd);
    double again[1 + 2 * ORC_DMAX + ORC_DMAX * ORC_DMAX];
    orc_stats_add(f->prior.d, comp, y, again);      /* the reference adds twice, :111-112 */
    logweight += orc_mll(&f->prior, again);
    return exp(logweight);
}

/* instantiate(putative, weight), src/particle.jl:134-149, with add(stateprior, state):
 * CRP src/statepriors.jl:55-63; StickyCRP :106-128; ChangePoint :171-175; NStatePrior :211-217 */
static void instantiate(const orc_filter *f, const putative *q, double weight, particle *out)
{
    const particle *p = &f->pop[q->parent];
    int L = ss_len(f->prior.d);
    int comp_i = sp_state_to_index(f, p, q->state);
    int K = p->K + (comp_i > p->K);
    out->K = K;
    out->st = malloc(sizeof(double) * (size_t)(K > 0 ? K : 1) * L);
    out->crpN = malloc(sizeof(double) * (K > 0 ? K : 1));
    out->Nsame = malloc(sizeof(double) * (K > 0 ? K : 1));
    memcpy(out->st, p->st, sizeof(double) * (size_t)p->K * L);
    memcpy(out->crpN, p->crpN, sizeof(double) * p->K);
    if (p->Nsame) memcpy(out->Nsame, p->Nsame, sizeof(double) * p->K); else memset(out->Nsame, 0, sizeof(double) * p->K);
    memcpy(out->st + (size_t)(comp_i - 1) * L, q->upd, sizeof(double) * L);
    out->last = p->last; out->total = p->total; out->cp_n = p->cp_n;
    switch (f->sp_kind) {
    case ORC_SP_STICKY: {
        const int sticky = q->state == 0;
        if (comp_i > p->K) { out->crpN[comp_i - 1] = 0.0; out->Nsame[comp_i - 1] = 0.0; }
        if (comp_i == p->last) out->Nsame[comp_i - 1] += 1.0;
        if (!sticky) out->crpN[comp_i - 1] += 1.0;
        out->last = comp_i;
        break;
    }
    case ORC_SP_CHANGEPOINT: out->cp_n = p->cp_n + 1; break;
    case ORC_SP_NSTATE: out->crpN[comp_i - 1] += 1.0; out->total = p->total + 1.0; break;
    default:
        if (comp_i > p->K) out->crpN[comp_i - 1] = 1.0; else out->crpN[comp_i - 1] += 1.0;
    }
    out->weight = weight;
    out->ancestor = q->parent;
    out->assignment = comp_i;
}

/* Replace the state prior of a filter that has not seen data yet (StatePrior of the initial particles,
 * src/particle.jl:59-63).  kind = ORC_SP_*: p1 = alpha (CRP, StickyCRP); p2 = rho (StickyCRP, :102) or the
 * change probability (ChangePoint(p), :168-171); counts[n] = NStatePrior(counts) (:209): the particles start
 * with n empty components, as test/labeled.jl:13-18 builds them. */
int orc_set_stateprior(orc_filter *f, int kind, double p1, double p2, const double *counts, int n)
{
    if (f->T != 0) return -1;
    f->sp_kind = kind;
    if (kind == ORC_SP_CRP || kind == ORC_SP_STICKY) f->alpha = p1;
    if (kind == ORC_SP_STICKY) f->sp_rho = p2;
    if (kind == ORC_SP_CHANGEPOINT) { if (!(p2 >= 0.0 && p2 <= 1.0)) return -2; f->sp_logp = log(p2); }
    int L = ss_len(f->prior.d);
    for (int i = 0; i < f->n; i++) {
        particle *p = &f->pop[i];
        particle_free(p);
        p->st = NULL; p->crpN = NULL; p->Nsame = NULL; p->K = 0; p->last = 1; p->total = 0.0; p->cp_n = 0;
        if (kind == ORC_SP_NSTATE) {
            f->sp_n = n;
            p->K = n;
            p->st = calloc((size_t)n * L, sizeof(double));
            p->crpN = malloc(sizeof(double) * n);
            p->Nsame = calloc(n, sizeof(double));
            for (int k = 0; k < n; k++) { p->crpN[k] = counts[k]; p->total += counts[k]; }
        }
    }
    return 0;
}

static void record_generation(orc_filter *f)
{
    if (!f->keep_history) { f->T++; return; }
    if (f->T == f->gens_cap) {
        f->gens_cap = f->gens_cap ? 2 * f->gens_cap : 64;
        f->gens = realloc(f->gens, sizeof(generation) * f->gens_cap);
    }
    generation *g = &f->gens[f->T++];
    g->n = f->n;
    g->anc = malloc(sizeof(int) * f->n);
    g->asg = malloc(sizeof(int) * f->n);
    for (int i = 0; i < f->n; i++) { g->anc[i] = f->pop[i].ancestor; g->asg[i] = f->pop[i].assignment; }
}

/* ------------------------------------------------------------------ fearnhead.jl */

/* cutoff_ascending(ws, N), src/fearnhead.jl:59-77.  Returns the 1-based index of the
 * first kept weight (M+1 when everything is resampled); *w_out = unnormalised weight of
 * the resampled particles.  Returns -1 if ws is not sorted (ArgumentError, :60-61). */
long orc_cutoff_ascending(const double *ws, long M, long N, double *w_out)
{
    for (long i = 1; i < M; i++) if (ws[i] < ws[i - 1]) return -1;
    double tot = 0.0;
    for (long i = 1; i <= M; i++) {
        double w = ws[i - 1];
        if (w != 0.0) {
            long n_geq = M - i + 1;
            if (tot <= (double)(N - n_geq) * w) {
                *w_out = tot / (double)(N - n_geq);
                return i;
            }
            tot += w;
        }
    }
    *w_out = tot / (double)N;
    return M + 1;
}

/* Base.sum over an array: pairwise summation with 1024-element leaves (Julia
 * mapreduce_impl); the leaf is a plain left-to-right loop here (Julia's leaf is @simd). */
static double sum_pairwise(const double *w, long n)
{
    if (n <= 0) return 0.0;
    if (n < 1024) {
        double s = w[0];
        for (long i = 1; i < n; i++) s += w[i];
        return s;
    }
    long h = n >> 1;
    return sum_pairwise(w, h) + sum_pairwise(w + h, n - h);
}
double orc_sum_pairwise(const double *w, long n) { return sum_pairwise(w, n); }

/* sample_stratified!(x, n, w), src/fearnhead.jl:97-114, with rand() := u.  picks[] gets
 * the 0-based positions (into w) of the selected items; returns their count (<= n). */
long orc_sample_stratified(const double *w, long len, long n, double u, long *picks)
{
    double K = sum_pairwise(w, len) / (double)n;
    double U = u * K;
    long i_store = 0;
    for (long i = 0; i < len; i++) {
        U = U - w[i];
        if (U < 0) {
            picks[i_store++] = i;
            U += K;
        }
    }
    return i_store;
}

/* stable merge sort of idx[] by key w[idx] ascending (declared tie-break: index) */
static void msort(int *idx, int *tmp, const double *w, long n)
{
    if (n < 2) return;
    long h = n / 2;
    msort(idx, tmp, w, h);
    msort(idx + h, tmp, w, n - h);
    long a = 0, b = h, k = 0;
    while (a < h && b < n) tmp[k++] = (w[idx[b]] < w[idx[a]]) ? idx[b++] : idx[a++];
    while (a < h) tmp[k++] = idx[a++];
    while (b < n) tmp[k++] = idx[b++];
    memcpy(idx, tmp, sizeof(int) * n);
}

/* putatives for the whole population, particle-major / candidate-minor
 * (src/fearnhead.jl:132, src/particle.jl:68-69, src/statepriors.jl:54) */
int orc_fh_expand(orc_filter *f, const double *y)
{
    int L = ss_len(f->prior.d);
    long M = 0;
    int first;
    for (int i = 0; i < f->n; i++) M += f->forced_label > 0 ? 1 : sp_candidates(f, &f->pop[i], &first);
    if (M > f->put_cap) {
        f->put_cap = (int)M;
        f->put = realloc(f->put, sizeof(putative) * M);
    }
    if (M * L > f->upd_pool_cap) {
        f->upd_pool_cap = M * L;
        free(f->upd_pool);
        f->upd_pool = malloc(sizeof(double) * M * L);
    }
    long k = 0;
    for (int i = 0; i < f->n; i++) {
        /* putatives(p, y), src/particle.jl:68-69; a labeled observation has ONE putative, src/labeled.jl:10-11 */
        int nc = sp_candidates(f, &f->pop[i], &first);
        if (f->forced_label > 0) { first = f->forced_label; nc = 1; }
        for (int j = first; j < first + nc; j++, k++) {
            putative *q = &f->put[k];
            q->parent = i; q->state = j; q->upd = f->upd_pool + k * L;
            q->weight = putative_weight(f, &f->pop[i], y, j, q->upd);
        }
    }
    f->M = (int)M;
    double total = 0.0;                 /* sum(weight(p) for p in putative): left fold, :133 */
    for (long q = 0; q < M; q++) total += f->put[q].weight;
    f->total_w = total;
    return (int)M;
}

/* replace the population by instantiate.(put[idx], w) (used by every selection mode) */
void orc_fh_commit(orc_filter *f, const int *idx, const double *w, int n)
{
    particle *np = calloc(n > 0 ? n : 1, sizeof(particle));
    for (int i = 0; i < n; i++) instantiate(f, &f->put[idx[i]], w[i], &np[i]);
    pop_free(f->pop, f->n);
    f->pop = np; f->n = n;
    record_generation(f);
}

/* fit!(ps::FearnheadParticles, y), src/fearnhead.jl:130-184, rand() := u.
 * order = ORC_ORDER_SORTED is the literal reference (ascending-weight traversal, output
 * [resampled ; kept]); ORC_ORDER_INDEX is the variant the reference's own TODO
 * (src/fearnhead.jl:140-145) proposes: same kept set and same sampler, but the low
 * partition is traversed in putative-index order and survivors stay in index order. */
int orc_fh_step(orc_filter *f, const double *y, double u)
{
    int M = orc_fh_expand(f, y);
    int N = f->N;
    int *idx = malloc(sizeof(int) * (M > N ? M : N));
    double *nw = malloc(sizeof(double) * (M > N ? M : N));
    int n_out;
    f->resampled = 0; f->n_resamp_req = f->n_resamp_got = 0; f->kept_i = 1; f->w_resamp = NAN;
    if (M <= N) {
        for (int k = 0; k < M; k++) { idx[k] = k; nw[k] = f->put[k].weight / f->total_w; }
        n_out = M; f->n_kept = M;
    } else {
        double *w = malloc(sizeof(double) * M);
        double *ws = malloc(sizeof(double) * M);
        int *perm = malloc(sizeof(int) * M), *tmp = malloc(sizeof(int) * M);
        long *picks = malloc(sizeof(long) * M);
        for (int k = 0; k < M; k++) { w[k] = f->put[k].weight; perm[k] = k; }
        msort(perm, tmp, w, M);                              /* sort!, :148 */
        for (int k = 0; k < M; k++) ws[k] = w[perm[k]];      /* :149 */
        double w_resamp;
        long kept_i = orc_cutoff_ascending(ws, M, N, &w_resamp);   /* :151 */
        int n_kept = M - (int)(kept_i - 1);                  /* :154 */
        int n_resamp = N - n_kept;                           /* :157 */
        f->kept_i = (int)kept_i; f->n_kept = n_kept; f->n_resamp_req = n_resamp;
        f->w_resamp = w_resamp; f->resampled = 1;
        if (f->order == ORC_ORDER_SORTED) {
            long got = orc_sample_stratified(ws, kept_i - 1, n_resamp, u, picks);  /* :165-167 */
            n_out = 0;
            for (long k = 0; k < got; k++) { idx[n_out] = perm[picks[k]]; nw[n_out++] = w_resamp / f->total_w; }  /* :172-173 */
            for (long k = kept_i - 1; k < M; k++) { idx[n_out] = perm[k]; nw[n_out++] = ws[k] / f->total_w; }     /* :177-179 */
            f->n_resamp_got = (int)got;
        } else {
            /* kept set = {w >= kappa}; low partition in putative-index order */
            double kappa = kept_i <= M ? ws[kept_i - 1] : INFINITY;
            long nlow = 0;
            for (int k = 0; k < M; k++) if (w[k] < kappa) { ws[nlow] = w[k]; tmp[nlow++] = k; }
            long got = orc_sample_stratified(ws, nlow, n_resamp, u, picks);
            char *take = calloc(M, 1);
            for (long k = 0; k < got; k++) take[tmp[picks[k]]] = 1;
            n_out = 0;
            for (int k = 0; k < M; k++) {
                if (w[k] >= kappa) { idx[n_out] = k; nw[n_out++] = w[k] / f->total_w; }
                else if (take[k]) { idx[n_out] = k; nw[n_out++] = w_resamp / f->total_w; }
            }
            free(take);
            f->n_resamp_got = (int)got;
        }
        free(w); free(ws); free(perm); free(tmp); free(picks);
    }
    orc_fh_commit(f, idx, nw, n_out);
    free(idx); free(nw);
    return n_out;
}

/* fit!(ps, Labeled(y, label)), src/labeled.jl:10-11 through src/fearnhead.jl:130-184: every particle has the single
 * putative fit(p, y, label); label <= 0 = missing (the ordinary step). */
int orc_fh_step_labeled(orc_filter *f, const double *y, double u, int label)
{
    f->forced_label = label > 0 ? label : 0;
    int n = orc_fh_step(f, y, u);
    f->forced_label = 0;
    return n;
}

/* ------------------------------------------------------------------ chenliu.jl */

/* StatsBase.sample(rng, wv::AbstractWeights) (inverse CDF), rand() := u; 0-based result */
static int wsample1(const double *w, int n, double u)
{
    double sum = sum_pairwise(w, n);
    double t = u * sum;
    int i = 1;
    double cw = w[0];
    while (cw < t && i < n) { i++; cw += w[i - 1]; }
    return i - 1;
}

/* coefvar, src/chenliu.jl:49-52 : StatsBase.mean_and_std (two-pass, corrected) */
double orc_coefvar(const double *x, long n)
{
    double mu = sum_pairwise(x, n) / (double)n;
    double *c = malloc(sizeof(double) * n);
    for (long i = 0; i < n; i++) c[i] = (x[i] - mu) * (x[i] - mu);
    double sd = sqrt(sum_pairwise(c, n) / (double)(n - 1));
    free(c);
    return sd / mu * 100.0;
}

/* fit!(ps::ChenLiuParticles, y), src/chenliu.jl:54-69 with propogate_chenliu :28-42.
 * u_prop[N]: one uniform per particle for the categorical draw (:35).
 * u_rejuv[N]: uniforms for the rejuvenation step.  The reference draws N ancestors with
 * replacement through StatsBase's alias sampler on Julia's RNG (:61), which cannot be
 * reproduced outside Julia; following the north-star contract the rejuvenation here is a
 * STRATIFIED inverse-CDF resample: position_i = (i + u_rejuv[i]) / N * sum(w), ancestor =
 * first index whose running weight sum exceeds the position.  Weights reset to 1/N (:61). */
int orc_cl_step(orc_filter *f, const double *y, const double *u_prop, const double *u_rejuv)
{
    int N = f->n;
    int L = ss_len(f->prior.d);
    particle *np = calloc(N, sizeof(particle));
    double *ws = malloc(sizeof(double) * N);
    double upd[(1 + 2 * ORC_DMAX + ORC_DMAX * ORC_DMAX)];
    for (int i = 0; i < N; i++) {
        particle *p = &f->pop[i];
        int first;
        int C = sp_candidates(f, p, &first);
        double *v = malloc(sizeof(double) * C);
        double *u = malloc(sizeof(double) * (size_t)C * L);
        for (int j = 0; j < C; j++) v[j] = putative_weight(f, p, y, first + j, u + (size_t)j * L);
        int pick = wsample1(v, C, u_prop[i]);                      /* :35 */
        putative q = { i, first + pick, v[pick], u + (size_t)pick * L };
        instantiate(f, &q, sum_pairwise(v, C), &np[i]);            /* :35,:41 */
        ws[i] = np[i].weight;
        free(v); free(u);
        (void)upd;
    }
    f->cv = N > 1 ? orc_coefvar(ws, N) : 0.0;                      /* :57-58 */
    f->resampled = 0;
    particle *out = np;
    if (f->cv > f->rejuv) {                                        /* :59 */
        f->resampled = 1;
        double tot = sum_pairwise(ws, N);
        out = calloc(N, sizeof(particle));
        int k = 0; double cw = ws[0];
        for (int i = 0; i < N; i++) {
            double pos = ((double)i + u_rejuv[i]) / (double)N * tot;
            while (cw <= pos && k < N - 1) { k++; cw += ws[k]; }
            particle *src = &np[k];
            out[i].K = src->K;
            out[i].st = malloc(sizeof(double) * (size_t)src->K * L);
            out[i].crpN = malloc(sizeof(double) * (src->K > 0 ? src->K : 1));
            out[i].Nsame = malloc(sizeof(double) * (src->K > 0 ? src->K : 1));
            memcpy(out[i].st, src->st, sizeof(double) * (size_t)src->K * L);
            memcpy(out[i].crpN, src->crpN, sizeof(double) * src->K);
            memcpy(out[i].Nsame, src->Nsame, sizeof(double) * src->K);
            out[i].last = src->last; out[i].total = src->total; out[i].cp_n = src->cp_n;
            out[i].weight = 1.0 / (double)N;
            out[i].ancestor = src->ancestor; out[i].assignment = src->assignment;
        }
        pop_free(np, N);
    } else {
        double tot = sum_pairwise(ws, N);                           /* :65-66 */
        for (int i = 0; i < N; i++) np[i].weight = ws[i] / tot;
    }
    f->total_w = sum_pairwise(ws, N);
    pop_free(f->pop, f->n);
    f->pop = out;
    free(ws);
    record_generation(f);
    return N;
}

/* ------------------------------------------------------------------ read-out */

int orc_n_particles(const orc_filter *f) { return f->n; }
int orc_n_steps(const orc_filter *f) { return f->T; }
int orc_n_putatives(const orc_filter *f) { return f->M; }
double orc_total_w(const orc_filter *f) { return f->total_w; }
double orc_w_resamp(const orc_filter *f) { return f->w_resamp; }
double orc_cv(const orc_filter *f) { return f->cv; }
int orc_resampled(const orc_filter *f) { return f->resampled; }
int orc_n_kept(const orc_filter *f) { return f->n_kept; }
int orc_n_resamp_req(const orc_filter *f) { return f->n_resamp_req; }
int orc_n_resamp_got(const orc_filter *f) { return f->n_resamp_got; }

void orc_get_weights(const orc_filter *f, double *out) { for (int i = 0; i < f->n; i++) out[i] = f->pop[i].weight; }
void orc_get_ncomponents(const orc_filter *f, int *out) { for (int i = 0; i < f->n; i++) out[i] = f->pop[i].K; }
void orc_get_last_ancestors(const orc_filter *f, int *out) { for (int i = 0; i < f->n; i++) out[i] = f->pop[i].ancestor; }
void orc_get_last_assignments(const orc_filter *f, int *out) { for (int i = 0; i < f->n; i++) out[i] = f->pop[i].assignment; }
void orc_get_putative_weights(const orc_filter *f, double *out) { for (int k = 0; k < f->M; k++) out[k] = f->put[k].weight; }
void orc_get_putative_parents(const orc_filter *f, int *parent, int *state)
{
    for (int k = 0; k < f->M; k++) { parent[k] = f->put[k].parent; state[k] = f->put[k].state; }
}

/* one particle's sufficient statistics: out has K * ss_len(d) doubles, crpN K doubles */
int orc_get_particle(const orc_filter *f, int i, double *st, double *crpN)
{
    const particle *p = &f->pop[i];
    memcpy(st, p->st, sizeof(double) * (size_t)p->K * ss_len(f->prior.d));
    memcpy(crpN, p->crpN, sizeof(double) * p->K);
    return p->K;
}

/* StickyCRP book-keeping of one particle (src/statepriors.jl:95-100): Nsame[K]; returns `last` (1-based) */
int orc_get_sticky(const orc_filter *f, int i, double *Nsame)
{
    const particle *p = &f->pop[i];
    for (int k = 0; k < p->K; k++) Nsame[k] = p->Nsame ? p->Nsame[k] : 0.0;
    return p->last;
}
/* log_prior(p.stateprior, x) of particle i (NAN = ArgumentError) and the candidate range */
double orc_sp_log_prior(const orc_filter *f, int i, int x) { return sp_log_prior(f, &f->pop[i], x); }
int orc_sp_candidates(const orc_filter *f, int i, int *first) { return sp_candidates(f, &f->pop[i], first); }

/* assignments(ps), src/filters.jl:26-34 via assignments(p), src/particle.jl:25-32:
 * out is T x n column-major (column i = particle i), 1-based component labels. */
int orc_get_assignments(const orc_filter *f, int64_t *out)
{
    if (!f->keep_history) return -1;
    int T = f->T;
    for (int i = 0; i < f->n; i++) {
        int k = i;
        for (int t = T - 1; t >= 0; t--) {
            out[(size_t)i * T + t] = f->gens[t].asg[k];
            k = f->gens[t].anc[k];
        }
    }
    return 0;
}

/* marginal_log_posterior(p), src/particle.jl:182-183 with marginal_log_prior(crp),
 * src/statepriors.jl:68-69 (the prior component contributes mll(empty)). */
double orc_marginal_log_posterior(const orc_filter *f, int i)
{
    const particle *p = &f->pop[i];
    int L = ss_len(f->prior.d);
    double empty[1 + 2 * ORC_DMAX + ORC_DMAX * ORC_DMAX];
    orc_empty_stats(f->prior.d, empty);
    double s = 0.0;
    for (int k = 0; k < p->K; k++) s += orc_mll(&f->prior, p->st + (size_t)k * L);
    s += orc_mll(&f->prior, empty);
    double lp = 0.0;
    for (int k = 0; k < p->K; k++) lp += lgamma(p->crpN[k]);
    lp += p->K * log(f->alpha);
    return s + lp;
}

/* stand-alone helpers so tests can drive mll / logpred / posterior without a filter */
void orc_make_prior(orc_prior *p, int kind, int d, const double *mu0, double kappa0, double nu0,
                    double s20, const double *Lam0)
{
    memset(p, 0, sizeof(*p));
    p->kind = kind; p->d = d; p->kappa0 = kappa0; p->nu0 = nu0; p->s20 = s20;
    for (int i = 0; i < d; i++) p->mu0[i] = mu0[i];
    if (kind == ORC_NIW) chol_upper(d, Lam0, p->U0);
}
int orc_prior_sizeof(void) { return (int)sizeof(orc_prior); }
const orc_prior *orc_filter_prior(const orc_filter *f) { return &f->prior; }
