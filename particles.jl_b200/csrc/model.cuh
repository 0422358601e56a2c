// model.cuh -- particle store layout and the collapsed posterior-predictive arithmetic.
//
// Store (structure of arrays, FP64): for particle i, cluster slot j, field f
//     S[(j * F + f) * cap + i]
// so that a warp working on 32 consecutive particles reads 256 contiguous bytes per
// (slot, field).  Fields of one slot (F = 2 + d + d(d+1)/2, the same byte count b(d) as
// SURVEY 8(d)):
//     0            n        observation count = CRP N_j (src/statepriors.jl:49) = NormalStats.tw
//     1            logdet   log det Lambda_n
//     2 .. 2+d-1   m        running sample mean (Welford m, src/component.jl:44-50, 64-70)
//     2+d ..       1/L_ii   reciprocal diagonal of L = chol(Lambda_n) (lower)
//     2+2d ..      L_ik     strict lower triangle, row-major: (1,0),(2,0),(2,1),...
// Lambda_n = Lambda_0 + s2 + kappa0 n/kappa_n (m-mu0)(m-mu0)' is the NIW posterior scale of
// ConjugatePriors' posterior_canon; NIX2(mu, s2, kappa, nu) is stored as the d = 1 case with
// Lambda_0 = nu0 * sigma2_0 (the identity test/niw.jl:44-67 pins).  The factor is carried
// forward by a rank-1 update Lambda_{n+1} = Lambda_n + kappa_n/(kappa_n+1) (y-mu_n)(y-mu_n)'
// instead of being refactorised from s2 as the reference does twice per putative
// (src/particle.jl:107,112 -> src/component.jl:134).
#pragma once
#include "common.cuh"

// Per-count constants (depend only on the integer n): one 32-byte row per n.
//   C = log N_j-or-alpha + lgamma((nu_n+1)/2) - lgamma((nu_n+1-d)/2) - d/2 log(pi) + d/2 log(r)
//   r = kappa_n / (kappa_n + 1),  h = (nu_n + 1) / 2,  g = n / kappa_n
struct __align__(32) NRow {
    double C, r, h, g;
};

#include <math.h>
// one table row, computed on the host in long double (the terms that depend on the integer count only)
// with_lp = false leaves the state prior's term out of C (state priors other than the CRP add theirs per candidate)
static inline NRow make_nrow(double kappa0, double nu0, double alpha, int dim, long long n, bool with_lp = true)
{
    const long double LOGPI_L = 1.1447298858494001741434273513530587116473L;
    long double k0 = kappa0, d = dim;
    long double kn = k0 + n, nn = (long double)nu0 + n;
    long double h = (nn + 1.0L) / 2.0L;
    long double r = kn / (kn + 1.0L);
    long double lp = !with_lp ? 0.0L : (n == 0 ? logl((long double)alpha) : logl((long double)n));
    long double C = lp + lgammal(h) - lgammal(h - d / 2.0L) - d / 2.0L * LOGPI_L + d / 2.0L * logl(r);
    NRow row;
    row.C = (double)C; row.r = (double)r; row.h = (double)h; row.g = (double)((long double)n / kn);
    return row;
}

template <int D>
struct Layout {
    static constexpr int F = 2 + D + D * (D + 1) / 2;
    static constexpr int F_N = 0, F_LOGDET = 1, F_MEAN = 2, F_DINV = 2 + D, F_OFF = 2 + 2 * D;
    static constexpr int NOFF = D * (D - 1) / 2;
};

// Per-step constants produced by k_prestep from the observation and the prior.
struct StepConst {
    double y[PF_MAX_D];
    double lw_new;                         // log alpha + log predictive of y under the prior
    unsigned long long plateau_bits;       // bit pattern of the new-cluster putative of every RESAMPLED particle: they all
                                           // carry the same weight c/total (src/fearnhead.jl:173), so these putatives are
                                           // exactly tied -- the one mass tie of the algorithm, counted instead of listed
    double newslot[2 + PF_MAX_D + PF_MAX_D * (PF_MAX_D + 1) / 2];   // fields of prior (+) y
    int forced;                            // >= 0: Labeled(y, label) -- every particle has the single putative "slot forced"
                                           // (src/labeled.jl:10-11); -1: label missing
    int pad;
    double prev_lw_resamp, prev_log_total; // the previous step's log(c / total) and log(total) (src/fearnhead.jl:172-179): the
                                           // pipelined step builds a child's log-weight while it expands it for this observation
};

struct PriorConst {
    double mu0[PF_MAX_D];
    double kappa0, nu0, alpha;
    double L0_dinv[PF_MAX_D];
    double L0_off[PF_MAX_D * (PF_MAX_D - 1) / 2];
    double logdet0;
    NRow row0;                             // n = 0 row (C uses log alpha under the CRP)
    // state prior (src/statepriors.jl): PF_SP_CRP N_j-or-alpha is folded into the table rows; the others add their
    // log-prior per candidate
    int sp_kind, sp_n;                     // PF_SP_* ; NStatePrior: number of states
    double sp_log_alpha;                   // StickyCRP: log alpha
    double sp_logit_rho;                   // StickyCRP: logit(rho), src/statepriors.jl:135
    double sp_lp_stay, sp_lp_change;       // ChangePoint: log1mexp(logp), logp, src/statepriors.jl:181-183
    double sp_counts0[PF_MAX_SLOTS];       // NStatePrior: initial counts (src/statepriors.jl:203-209)
    double sp_total0;
};

enum { SPK_CRP = 0, SPK_STICKY = 1, SPK_CHANGEPOINT = 2, SPK_NSTATE = 3 };      // = PF_SP_* of the C header

// StickyCRP book-keeping of a population (src/statepriors.jl:95-100), slot-major like the store
struct SpArrays {
    double *N;                             // [slot * cap + i]: non-sticky occupations
    double *Nsame;                         // [slot * cap + i]: transitions to the same state
    int *last;                             // [i]: component of the previous observation (0-based)
};

// z = L^{-1} dx (forward substitution), returns |z|^2
template <int D>
__host__ __device__ __forceinline__ double quad_form(const double *dx, const double *dinv, const double *off, double *z)
{
    double q = 0.0;
    int o = 0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        double acc = dx[i];
#pragma unroll
        for (int k = 0; k < i; k++) acc -= off[o + k] * z[k];
        o += i;
        z[i] = acc * dinv[i];
        q += z[i] * z[i];
    }
    return q;
}

// Rank-1 update of the lower Cholesky factor: L L' + x x'  (x is overwritten).
// Works on (1/L_ii, strict lower) storage; returns nothing, logdet handled by the caller.
template <int D>
__host__ __device__ __forceinline__ void chol_rank1(double *dinv, double *off, double *x)
{
#pragma unroll
    for (int k = 0; k < D; k++) {
        double Lkk = 1.0 / dinv[k];
        double r = sqrt(Lkk * Lkk + x[k] * x[k]);
        double c = r * dinv[k];            // r / Lkk
        double s = x[k] * dinv[k];         // x_k / Lkk
        dinv[k] = 1.0 / r;
#pragma unroll
        for (int i = k + 1; i < D; i++) {
            int idx = i * (i - 1) / 2 + k;
            double Lik = (off[idx] + s * x[i]) / c;
            x[i] = c * x[i] - s * Lik;
            off[idx] = Lik;
        }
    }
}

// add(comp, y) (src/component.jl:44-50, 64-70, 85) on one slot's fields in registers: Welford mean update and rank-1
// update of the carried Cholesky factor of Lambda_n
template <int D>
__host__ __device__ __forceinline__ void slot_add(double *fl, const NRow &row, const double *y, const double *mu0)
{
    using L = Layout<D>;
    double dx[D], z[D];
    double *m = fl + L::F_MEAN, *dinv = fl + L::F_DINV, *off = fl + L::F_OFF;
#pragma unroll
    for (int k = 0; k < D; k++) dx[k] = y[k] - (mu0[k] + row.g * (m[k] - mu0[k]));
    const double q = quad_form<D>(dx, dinv, off, z);
    const double sr = sqrt(row.r);
#pragma unroll
    for (int k = 0; k < D; k++) dx[k] *= sr;
    chol_rank1<D>(dinv, off, dx);
    const double tw = fl[L::F_N] + 1.0;
    fl[L::F_N] = tw;
    fl[L::F_LOGDET] = fl[L::F_LOGDET] + log1p(row.r * q);
#pragma unroll
    for (int k = 0; k < D; k++) {
        const double delta = y[k] - m[k];                          // Welford, src/component.jl:46-47
        m[k] = m[k] + delta / tw;
    }
}

// log-weight increment of assigning y to an existing cluster (without the parent's log w):
//   log N_j + log t_{nu_n-d+1}(y | mu_n, Lambda_n (kappa_n+1)/(kappa_n (nu_n-d+1)))
// = C(n) - logdet/2 - h(n) log1p(r(n) |L^{-1}(y-mu_n)|^2),   mu_n = mu0 + g(n) (m - mu0)
// (src/component.jl:18-23 in the form of SURVEY 8(a4); equals mll(add(c,y)) - mll(c) of
// src/particle.jl:104-112 plus log_prior of src/statepriors.jl:65-66).
template <int D>
__host__ __device__ __forceinline__ double slot_logw(const NRow &row, double logdet, const double *m, const double *dinv,
                                            const double *off, const double *y, const double *mu0, double *z_out,
                                            double *q_out)
{
    double dx[D];
#pragma unroll
    for (int i = 0; i < D; i++) dx[i] = y[i] - (mu0[i] + row.g * (m[i] - mu0[i]));
    double q = quad_form<D>(dx, dinv, off, z_out);
    *q_out = q;
    return row.C - 0.5 * logdet - row.h * log1p(row.r * q);
}
