"""The model arithmetic of csrc/model.cuh compiled for the host (tests/host/host_model.cu) against
independent implementations: table rows vs scipy gammaln, the rank-1 Cholesky update vs a fresh
factorisation, the predictive log-weight vs scipy's multivariate t (src/component.jl:18-23), and
the slot update carried through hundreds of observations vs the closed-form NIW posterior
(ConjugatePriors posterior_canon).  No GPU needed: the functions are __host__ __device__."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest
from scipy import stats

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "host_model.cu")
OUT = os.path.join(HERE, "host", "_build", "libhost_model.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
pytestmark = pytest.mark.skipif(not (os.path.exists(NVCC) or shutil.which("nvcc")), reason="nvcc not available")
DP = C.POINTER(C.c_double)


def _p(a):
    return a.ctypes.data_as(DP)


@pytest.fixture(scope="module")
def H():
    deps = [SRC] + [os.path.join(HERE, "..", "particles.jl_b200", "csrc", f) for f in ("model.cuh", "common.cuh")]
    if not os.path.exists(OUT) or os.path.getmtime(OUT) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call([NVCC if os.path.exists(NVCC) else "nvcc", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                               "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, SRC])
    L = C.CDLL(OUT)
    L.h_slot_logw.restype = C.c_double
    L.h_make_nrow.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, C.c_longlong, DP]
    L.h_slot_logw.argtypes = [C.c_int, DP, C.c_double, DP, DP, DP, DP, DP, DP]
    L.h_slot_add.argtypes = [C.c_int, DP, DP, DP, DP]
    L.h_chol_rank1.argtypes = [C.c_int, DP, DP, DP]
    return L


def _row(H, kappa0, nu0, alpha, d, n):
    out = np.zeros(4)
    H.h_make_nrow(kappa0, nu0, alpha, d, n, _p(out))
    return out


def _pack(Lo):
    """lower Cholesky factor -> (1/diag, strict lower triangle row-major), the store's encoding"""
    d = Lo.shape[0]
    off = np.array([Lo[i, k] for i in range(d) for k in range(i)] or [0.0])
    return 1.0 / np.diag(Lo).copy(), off


def _unpack(dinv, off, d):
    Lo = np.diag(1.0 / dinv)
    o = 0
    for i in range(d):
        for k in range(i):
            Lo[i, k] = off[o]; o += 1
    return Lo


@pytest.mark.parametrize("d", [1, 2, 3, 4, 8])
def test_table_row(H, d):
    kappa0, nu0, alpha = 0.1, d + 1.0, 1.5
    for n in (0, 1, 2, 17, 1000, 123456):
        C_, r, h, g = _row(H, kappa0, nu0, alpha, d, n)
        kn, nn = kappa0 + n, nu0 + n
        mp = pytest.importorskip("mpmath")
        mp.mp.dps = 40                                  # (the difference of two large lgammas cancels in double)
        kq, nq = mp.mpf(kappa0) + n, mp.mpf(nu0) + n
        want = (mp.log(alpha if n == 0 else n) + mp.loggamma((nq + 1) / 2) - mp.loggamma((nq + 1 - d) / 2)
                - mp.mpf(d) / 2 * mp.log(mp.pi) + mp.mpf(d) / 2 * mp.log(kq / (kq + 1)))
        assert C_ == pytest.approx(float(want), rel=1e-14, abs=2e-13)      # long double: ~1e-19 of two lgammas of ~1e6
        assert r == pytest.approx(kn / (kn + 1), rel=1e-15) and h == pytest.approx((nn + 1) / 2, rel=1e-15)
        assert g == pytest.approx(n / kn, rel=1e-15)


@pytest.mark.parametrize("d", [1, 2, 3, 4, 8])
def test_rank1_update_equals_fresh_cholesky(H, d):
    rng = np.random.default_rng(d)
    for _ in range(50):
        A = rng.standard_normal((d, d))
        Lam = A @ A.T + np.eye(d) * rng.uniform(0.01, 2.0)
        x = rng.standard_normal(d) * rng.uniform(0.01, 30.0)
        dinv, off = _pack(np.linalg.cholesky(Lam))
        xx = x.copy()
        H.h_chol_rank1(d, _p(dinv), _p(off), _p(xx))
        want = np.linalg.cholesky(Lam + np.outer(x, x))
        np.testing.assert_allclose(_unpack(dinv, off, d), want, rtol=1e-11, atol=1e-12)


def _posterior(ys, mu0, kappa0, nu0, Lam0):
    n = len(ys)
    xbar = ys.mean(axis=0)
    S = (ys - xbar).T @ (ys - xbar)
    kn, nn = kappa0 + n, nu0 + n
    Lamn = Lam0 + S + kappa0 * n / kn * np.outer(xbar - mu0, xbar - mu0)
    return kn, nn, (kappa0 * mu0 + n * xbar) / kn, Lamn, xbar


@pytest.mark.parametrize("d", [1, 2, 3, 4, 8])
def test_predictive_logweight_is_multivariate_t(H, d):
    rng = np.random.default_rng(10 + d)
    mu0, kappa0, nu0, alpha = rng.standard_normal(d) * 0.3, 0.1, d + 1.0, 1.0
    A = rng.standard_normal((d, d)) * 0.3
    Lam0 = np.eye(d) + A @ A.T
    for n in (1, 2, 5, 40, 700):
        ys = rng.standard_normal((n, d)) * 1.3 + 2.0
        kn, nn, mun, Lamn, xbar = _posterior(ys, mu0, kappa0, nu0, Lam0)
        dinv, off = _pack(np.linalg.cholesky(Lamn))
        logdet = np.linalg.slogdet(Lamn)[1]
        y = rng.standard_normal(d) * 2.0 + 1.0
        q = np.zeros(1)
        got = H.h_slot_logw(d, _p(_row(H, kappa0, nu0, alpha, d, n)), logdet, _p(xbar.copy()), _p(dinv), _p(off), _p(y), _p(mu0.copy()), _p(q))
        df = nn - d + 1
        want = np.log(n) + stats.multivariate_t(loc=mun, shape=Lamn * (kn + 1) / (kn * df), df=df).logpdf(y)
        assert got == pytest.approx(want, rel=1e-11, abs=1e-11)


@pytest.mark.parametrize("d", [1, 2, 3, 4, 8])
def test_slot_update_tracks_the_closed_form_posterior(H, d):
    """n -> n+1 by Welford + rank-1 update, 600 times, against posterior_canon from scratch."""
    rng = np.random.default_rng(20 + d)
    mu0, kappa0, nu0, alpha = np.zeros(d), 0.1, d + 1.0, 1.0
    Lam0 = np.eye(d)
    F = H.h_fields(d)
    ys = rng.standard_normal((601, d)) * 1.5 + rng.standard_normal(d) * 4
    # a new cluster holding ys[0] (what k_prestep prepares): n = 1, mean = y, Lambda_1 = Lambda_0 + k0/(k0+1) (y-mu0)(y-mu0)'
    _, _, _, Lam1, _ = _posterior(ys[:1], mu0, kappa0, nu0, Lam0)
    dinv, off = _pack(np.linalg.cholesky(Lam1))
    f = np.concatenate([[1.0, np.linalg.slogdet(Lam1)[1]], ys[0], dinv, off[:d * (d - 1) // 2]])
    assert len(f) == F
    for t in range(1, 601):
        H.h_slot_add(d, _p(f), _p(ys[t].copy()), _p(mu0), _p(_row(H, kappa0, nu0, alpha, d, t)))
        if t in (1, 2, 10, 100, 600):
            kn, nn, mun, Lamn, xbar = _posterior(ys[:t + 1], mu0, kappa0, nu0, Lam0)
            assert f[0] == t + 1
            np.testing.assert_allclose(f[2:2 + d], xbar, rtol=1e-12, atol=1e-12)
            Lo = _unpack(f[2 + d:2 + 2 * d], np.concatenate([f[2 + 2 * d:], [0.0]]), d)
            np.testing.assert_allclose(Lo @ Lo.T, Lamn, rtol=1e-10, atol=1e-10)
            assert f[1] == pytest.approx(np.linalg.slogdet(Lamn)[1], rel=1e-11, abs=1e-11)
