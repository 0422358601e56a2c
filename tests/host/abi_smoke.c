/* abi_smoke.c -- drives libparticles_b200.so exactly as the Julia `ccall`s of
 * particles.jl_b200/julia/device.jl do (plain C, no Python, no torch):
 *   pf_create -> pf_filter -> pf_n_particles / pf_get_logweights / pf_get_ncomponents ->
 *   pf_get_assignments -> pf_get_particle -> pf_get_last_step -> the device summaries (pf_weighted_mean,
 *   pf_ncomponents_dist, pf_posterior_predictive, pf_assignment_similarity, pf_randindex, pf_get_particle_assignments,
 *   pf_genealogy_info) -> pf_save / pf_load (the loaded handle continues identically) -> pf_destroy
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
    /* filter-level summaries on the device */
    double meanK = 0.0, H = 0.0, dist[25], dens[3], xs[6] = {6.0, 0.0, 0.0, 6.0, 40.0, 40.0}, ri = 0.0;
    CHECK(pf_weighted_mean(h, PF_MEAN_NCOMPONENTS, &meanK));
    CHECK(pf_weighted_mean(h, PF_MEAN_STATE_ENTROPY, &H));
    CHECK(pf_ncomponents_dist(h, dist));
    double sd = 0.0, mk = 0.0;
    for (int k = 0; k <= 24; k++) { sd += dist[k]; mk += k * dist[k]; }
    if (fabs(sd - 1.0) > 1e-9 || fabs(mk - meanK) > 1e-9 || !(H > 0.0)) { fprintf(stderr, "summaries: sum %.17g mean %.17g / %.17g H %.17g\n", sd, mk, meanK, H); return 1; }
    CHECK(pf_posterior_predictive(h, xs, 3, dens));
    if (!(dens[0] > dens[2] && dens[1] > dens[2] && dens[2] > 0.0)) { fprintf(stderr, "predictive density: %g %g %g\n", dens[0], dens[1], dens[2]); return 1; }
    double *sim = malloc(sizeof(double) * T * T);
    CHECK(pf_assignment_similarity(h, sim));
    for (int t = 0; t < T; t++) if (sim[t * T + t] != 1.0 || sim[t * T] != sim[t]) { fprintf(stderr, "similarity matrix\n"); return 1; }
    int32_t truth[60], chain[60];
    CHECK(pf_get_particle_assignments(h, 0, chain));
    for (int t = 0; t < T; t++) { truth[t] = chain[t]; if (chain[t] != A[t]) { fprintf(stderr, "chain of particle 0\n"); return 1; } }
    CHECK(pf_randindex(h, truth, T, 1, &ri));
    if (!(ri > 0.0 && ri <= 1.0 + 1e-12)) { fprintf(stderr, "rand index %g\n", ri); return 1; }
    int64_t ginfo[4];
    CHECK(pf_genealogy_info(h, ginfo));
    if (ginfo[0] != T || ginfo[1] < T) { fprintf(stderr, "genealogy info\n"); return 1; }
    /* checkpoint: a loaded handle filters the next observations exactly like the original */
    const char *path = argc > 3 ? argv[3] : "/tmp/abi_smoke.ckpt";
    CHECK(pf_save(h, path));
    pf_handle h2 = NULL;
    if (pf_load(path, 0, &h2) != PF_OK) { fprintf(stderr, "pf_load: %s\n", pf_last_error(NULL)); return 1; }
    double le1[5], le2[5];
    CHECK(pf_filter(h, ys, 5, NULL, 0, le1));
    if (pf_filter(h2, ys, 5, NULL, 0, le2) != PF_OK) { fprintf(stderr, "pf_filter on the loaded handle: %s\n", pf_last_error(h2)); return 1; }
    for (int t = 0; t < 5; t++) if (le1[t] != le2[t]) { fprintf(stderr, "the loaded handle diverges at observation %d\n", t); return 1; }
    pf_destroy(h2);
    remove(path);
    free(sim);
    double ev = 0.0;
    for (int t = 0; t < T; t++) ev += le[t];
    printf("abi_smoke ok: kind=%d n=%lld K[0]=%d log-evidence=%.6f %s\n", kind, n, Kp, ev, pf_version());
    CHECK(pf_destroy(h));
    free(ys); free(le); free(lw); free(K); free(A);
    return 0;
}
