/* abi_smoke.c -- drives libparticles_b200.so exactly as the Julia `ccall`s of
 * particles.jl_b200/julia/particles_b200.jl do (plain C, no Python, no torch):
 *   pf_create -> pf_filter -> pf_n_particles / pf_get_logweights / pf_get_ncomponents ->
 *   pf_get_assignments -> pf_get_particle -> pf_get_last_step -> pf_destroy
 * and checks the invariants the reference's own test asserts on the result
 * (test/clustering.jl:9-13: every column of the assignment matrix starts at label 1, its maximum is
 * the particle's component count) plus weight normalisation.  Compiled with gcc and run by
 * tests/test_gpu_abi_c.py on the GPU box.  Exit code 0 = ok. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "particles_b200.h"

#define CHECK(call)                                                                       \
    do {                                                                                  \
        int32_t rc_ = (call);                                                             \
        if (rc_ != PF_OK) {                                                               \
            fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, pf_last_error(h));        \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

static double lcg(unsigned long long *s)
{
    *s = *s * 6364136223846793005ull + 1442695040888963407ull;
    return (double)(*s >> 11) * (1.0 / 9007199254740992.0);
}

int main(int argc, char **argv)
{
    const int kind = argc > 1 ? atoi(argv[1]) : PF_FEARNHEAD;
    const long long N = argc > 2 ? atoll(argv[2]) : 2000;
    const int T = 60, d = 2;
    pf_handle h = NULL;
    double mu0[2] = {0.0, 0.0}, lam0[4] = {1.0, 0.0, 0.0, 1.0};
    pf_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.filter_kind = kind; cfg.prior_kind = PF_NIW; cfg.d = d; cfg.precision = PF_F64;
    cfg.n_particles = N; cfg.k_max = 24; cfg.order = PF_ORDER_INDEX; cfg.history = -1; cfg.device = 0;
    cfg.alpha = 1.0; cfg.rejuv_threshold = 50.0; cfg.kappa0 = 0.1; cfg.nu0 = 3.0; cfg.mu0 = mu0; cfg.lambda0 = lam0;
    cfg.seed = 7;
    CHECK(pf_create(&cfg, &h));

    unsigned long long s = 12345;
    double *ys = malloc(sizeof(double) * d * T), *le = malloc(sizeof(double) * T);
    for (int t = 0; t < T; t++) {                      /* 4 clusters on a circle of radius 6, Box-Muller noise */
        int k = (int)(lcg(&s) * 4);
        double r = sqrt(-2.0 * log(1.0 - lcg(&s))), a = 6.283185307179586 * lcg(&s);
        ys[d * t] = 6.0 * cos(1.5707963267948966 * k) + r * cos(a);
        ys[d * t + 1] = 6.0 * sin(1.5707963267948966 * k) + r * sin(a);
    }
    CHECK(pf_filter(h, ys, T, NULL, 0, le));          /* uniforms == NULL: the library's documented device stream */

    const long long n = pf_n_particles(h);
    if (n < 1 || n > N || pf_n_steps(h) != T) { fprintf(stderr, "bad sizes: n=%lld steps=%lld\n", n, (long long)pf_n_steps(h)); return 1; }
    double *lw = malloc(sizeof(double) * n);
    int32_t *K = malloc(sizeof(int32_t) * n), *A = malloc(sizeof(int32_t) * (size_t)n * T);
    CHECK(pf_get_logweights(h, lw));
    CHECK(pf_get_ncomponents(h, K));
    CHECK(pf_get_assignments(h, A));
    double sw = 0.0;
    for (long long i = 0; i < n; i++) sw += exp(lw[i]);
    if (fabs(sw - 1.0) > 1e-9) { fprintf(stderr, "weights sum to %.17g\n", sw); return 1; }
    for (long long i = 0; i < n; i++) {
        int mx = 0;
        for (int t = 0; t < T; t++) { int a = A[i * T + t]; if (a < 0) { fprintf(stderr, "negative label\n"); return 1; } if (a > mx) mx = a; }
        if (A[i * T] != 0 || mx + 1 != K[i]) { fprintf(stderr, "assignment invariant broken at particle %lld: first=%d max=%d K=%d\n", i, A[i * T], mx, K[i]); return 1; }
    }
    int32_t Kp = 0;
    double counts[24], means[48], chol[96], logdet[24];
    CHECK(pf_get_particle(h, 0, &Kp, counts, means, chol, logdet));
    double nobs = 0.0;
    for (int j = 0; j < Kp; j++) nobs += counts[j];
    if (Kp != K[0] || nobs != (double)T) { fprintf(stderr, "particle 0: K=%d nobs=%g\n", Kp, nobs); return 1; }
    pf_step_stats st;
    CHECK(pf_get_last_step(h, &st));
    if (st.step != T || st.n_out != n || st.overflow != 0) { fprintf(stderr, "bad step stats\n"); return 1; }
    double ev = 0.0;
    for (int t = 0; t < T; t++) ev += le[t];
    printf("abi_smoke ok: kind=%d n=%lld K[0]=%d log-evidence=%.6f %s\n", kind, n, Kp, ev, pf_version());
    CHECK(pf_destroy(h));
    free(ys); free(le); free(lw); free(K); free(A);
    return 0;
}
