// Host-side harness for the __host__ __device__ model arithmetic of csrc/model.cuh (the same source
// the kernels compile): table rows, forward substitution, rank-1 Cholesky update, the collapsed
// predictive log-weight, and the slot update exactly as k_instantiate / k_cl_propagate sequence it.
// Test infrastructure only (tests/test_host_model.py).
#include "../../particles.jl_b200/csrc/model.cuh"

template <int D>
static double slot_logw_t(const double *row4, double logdet, const double *m, const double *dinv, const double *off, const double *y,
                          const double *mu0, double *q_out)
{
    NRow row; row.C = row4[0]; row.r = row4[1]; row.h = row4[2]; row.g = row4[3];
    double z[D];
    return slot_logw<D>(row, logdet, m, dinv, off, y, mu0, z, q_out);
}

// fields: n, logdet, m[D], 1/L_ii [D], strict lower triangle row-major (the store's slot record)
template <int D>
static void slot_add_t(double *f, const double *y, const double *mu0, const double *row4)
{
    using L = Layout<D>;
    NRow row; row.C = row4[0]; row.r = row4[1]; row.h = row4[2]; row.g = row4[3];
    double n = f[L::F_N], logdet = f[L::F_LOGDET];
    double m[D], dinv[D], off[L::NOFF > 0 ? L::NOFF : 1], dx[D], z[D];
    for (int k = 0; k < D; k++) { m[k] = f[L::F_MEAN + k]; dinv[k] = f[L::F_DINV + k]; }
    for (int k = 0; k < L::NOFF; k++) off[k] = f[L::F_OFF + k];
    for (int k = 0; k < D; k++) dx[k] = y[k] - (mu0[k] + row.g * (m[k] - mu0[k]));
    double q = quad_form<D>(dx, dinv, off, z);
    double sr = sqrt(row.r);
    for (int k = 0; k < D; k++) dx[k] *= sr;
    chol_rank1<D>(dinv, off, dx);
    double tw = n + 1.0;
    f[L::F_N] = tw;
    f[L::F_LOGDET] = logdet + log1p(row.r * q);
    for (int k = 0; k < D; k++) {
        double delta = y[k] - m[k];
        f[L::F_MEAN + k] = m[k] + delta / tw;
        f[L::F_DINV + k] = dinv[k];
    }
    for (int k = 0; k < L::NOFF; k++) f[L::F_OFF + k] = off[k];
}

template <int D>
static void rank1_t(double *dinv, double *off, double *x) { chol_rank1<D>(dinv, off, x); }

#define DISPATCH(dd, CALL)                                   \
    switch (dd) {                                            \
    case 1: { constexpr int D = 1; CALL; } break;            \
    case 2: { constexpr int D = 2; CALL; } break;            \
    case 3: { constexpr int D = 3; CALL; } break;            \
    case 4: { constexpr int D = 4; CALL; } break;            \
    case 8: { constexpr int D = 8; CALL; } break;            \
    default: break;                                          \
    }

extern "C" {
void h_make_nrow(double kappa0, double nu0, double alpha, int d, long long n, double *out4)
{
    NRow r = make_nrow(kappa0, nu0, alpha, d, n);
    out4[0] = r.C; out4[1] = r.r; out4[2] = r.h; out4[3] = r.g;
}
double h_slot_logw(int d, const double *row4, double logdet, const double *m, const double *dinv, const double *off, const double *y,
                   const double *mu0, double *q_out)
{
    double r = 0.0;
    DISPATCH(d, r = slot_logw_t<D>(row4, logdet, m, dinv, off, y, mu0, q_out));
    return r;
}
void h_slot_add(int d, double *fields, const double *y, const double *mu0, const double *row4) { DISPATCH(d, slot_add_t<D>(fields, y, mu0, row4)); }
void h_chol_rank1(int d, double *dinv, double *off, double *x) { DISPATCH(d, rank1_t<D>(dinv, off, x)); }
int h_fields(int d) { int r = 0; DISPATCH(d, r = Layout<D>::F); return r; }
}
