"""Pins the CPU oracle's arithmetic against the reference's own known-answer and identity
tests (ports of test/component.jl, test/niw.jl, test/nix2.jl, test/particle.jl,
test/statepriors.jl) and against independent implementations (scipy / mpmath)."""
import math

import numpy as np
import pytest
from scipy import stats as sps

from oracle import oracle as orc


def test_welford_nix2_known_answer():
    # test/component.jl:7-11 : reduce(add, 1:10) == suffstats(Normal, 1:10) exactly
    p = orc.Prior.nix2(0.0, 1.0, 1.0, 1.0)
    ss = p.empty()
    for x in range(1, 11):
        ss = p.add(ss, float(x))
    tw, s, m, s2 = ss
    assert (tw, s, m, s2) == (10.0, 55.0, 5.5, 82.5)
    # sub inverse: reduce(sub, 1:5, init=suffstats(1:10)) == suffstats(6:10)
    for x in range(1, 6):
        ss = p.sub(ss, float(x))
    assert tuple(ss) == (5.0, 40.0, 8.0, 10.0)
    one = p.add(p.empty(), 1.0)
    assert np.all(p.sub(one, 1.0) == 0.0)


def test_welford_mv_known_answer():
    # test/component.jl:13-26
    p = orc.Prior.niw([0.0, 0.0], 1.0, np.eye(2), 3.0)
    vecs = [np.array([i, i + 1.0]) for i in range(1, 11)]
    ss = p.empty()
    for v in vecs:
        ss = p.add(ss, v)
    X = np.stack(vecs, axis=1)
    m = X.mean(axis=1)
    Z = X - m[:, None]
    assert ss[0] == 10.0
    assert np.all(ss[1:3] == X.sum(axis=1))
    assert np.all(ss[3:5] == m)
    assert np.all(ss[5:9].reshape(2, 2) == Z @ Z.T)
    for v in vecs[:5]:
        ss = p.sub(ss, v)
    X2 = np.stack(vecs[5:], axis=1)
    Z2 = X2 - X2.mean(axis=1)[:, None]
    assert ss[0] == 5.0
    assert np.all(ss[3:5] == X2.mean(axis=1))
    assert np.all(ss[5:9].reshape(2, 2) == Z2 @ Z2.T)


def test_empty_suffstats_shapes():
    # test/component.jl:28-36
    assert np.all(orc.Prior.nix2(0, 1, 1, 1).empty() == np.zeros(4))
    assert orc.Prior.niw(np.zeros(3), 1.0, np.eye(3), 1.0).empty().shape == (1 + 3 + 3 + 9,)


def _suff(p, xs):
    ss = p.empty()
    for x in xs:
        ss = p.add(ss, x)
    return ss


def test_niw_d1_equals_nix2():
    # test/niw.jl:44-67 : NIW([mu], kappa, nu*[s2], nu) == NIX2(mu, s2, kappa, nu)
    rng = np.random.default_rng(7)
    nix2 = orc.Prior.nix2(1.0, 2.0, 3.0, 4.0)
    niw = orc.Prior.niw([1.0], 3.0, [[4.0 * 2.0]], 4.0)
    x = 3 + 2 * rng.standard_normal(100)
    ss1 = _suff(nix2, x)
    ssm = _suff(niw, [[v] for v in x])
    mu, s2n, kn, nn = nix2.posterior(ss1)
    mu_v, U, kn_v, nn_v = niw.posterior(ssm)
    assert mu == pytest.approx(mu_v[0], rel=1e-14)
    assert (kn, nn) == (kn_v, nn_v)
    assert s2n == pytest.approx(U[0, 0] ** 2 / nn_v, rel=1e-13)
    assert nix2.mll(ss1) == pytest.approx(niw.mll(ssm), rel=1e-12)


def _mll_nix2_mp(prior, xs):
    """50-digit evaluation of src/component.jl:122-130 from raw data."""
    import mpmath as mp
    mp.mp.dps = 50
    xs = [mp.mpf(float(x)) for x in xs]
    n = len(xs)
    mu0, s20, k0, nu0 = [mp.mpf(v) for v in (prior.mu0[0], prior.s20, prior.kappa0, prior.nu0)]
    xbar = sum(xs) / n
    s2 = sum((x - xbar) ** 2 for x in xs)
    kn, nn = k0 + n, nu0 + n
    s2n = (nu0 * s20 + s2 + n * k0 / (n + k0) * (mu0 - xbar) ** 2) / nn
    return (mp.loggamma(nn / 2) - mp.loggamma(nu0 / 2) + (mp.log(k0) - mp.log(kn)) / 2
            + nu0 / 2 * mp.log(nu0 * s20) - nn / 2 * mp.log(nn * s2n) - mp.mpf(n) / 2 * mp.log(mp.pi))


def test_mll_nix2_vs_mpmath():
    rng = np.random.default_rng(3)
    p = orc.Prior.nix2(0.0, 1.0, 0.1, 1.0)
    for n in (1, 2, 7, 50, 400):
        x = rng.standard_normal(n) * 1.7 + 0.4
        got = p.mll(_suff(p, x))
        want = float(_mll_nix2_mp(p, x))
        assert got == pytest.approx(want, rel=1e-12, abs=1e-12)


def test_logpred_is_mll_difference_and_matches_scipy_nix2():
    # the identity the expansion relies on: mll(c+y) - mll(c) == logpdf(posterior_predictive(c), y)
    rng = np.random.default_rng(11)
    p = orc.Prior.nix2(0.3, 1.5, 0.1, 1.0)
    for n in (0, 1, 5, 60):
        ss = _suff(p, rng.standard_normal(n) * 2 + 1)
        for y in (-3.0, 0.2, 4.0):
            lp = p.logpred(ss, y)
            diff = p.mll(p.add(ss, y)) - p.mll(ss)
            assert lp == pytest.approx(diff, rel=1e-10, abs=1e-11)
            mu, s2n, kn, nn = p.posterior(ss)
            want = sps.t.logpdf(y, df=nn, loc=mu, scale=math.sqrt((1 + kn) * s2n / kn))
            assert lp == pytest.approx(want, rel=1e-12, abs=1e-12)


@pytest.mark.parametrize("d", [1, 2, 3, 8])
def test_logpred_is_mll_difference_and_matches_scipy_niw(d):
    rng = np.random.default_rng(100 + d)
    A = rng.standard_normal((d, d))
    Lam0 = A @ A.T + d * np.eye(d)
    p = orc.Prior.niw(rng.standard_normal(d), 0.1, Lam0, d + 1.0)
    for n in (0, 1, 4, 40):
        ss = _suff(p, rng.standard_normal((n, d)) * 1.3 + 0.5)
        for _ in range(3):
            y = rng.standard_normal(d) * 3
            lp = p.logpred(ss, y)
            diff = p.mll(p.add(ss, y)) - p.mll(ss)
            assert lp == pytest.approx(diff, rel=1e-9, abs=1e-10)
            mu, U, kn, nn = p.posterior(ss)
            df = nn - d + 1
            Sigma = (U.T @ U) * (kn + 1) / (kn * df)
            want = sps.multivariate_t(loc=mu, shape=Sigma, df=df).logpdf(y)
            assert lp == pytest.approx(want, rel=1e-11, abs=1e-11)


def test_niw_posterior_matches_direct_formula():
    # Murphy (2007) eq. 250-254 from raw data, independent of the Welford path
    rng = np.random.default_rng(5)
    d = 3
    Lam0 = np.array([[4.0, 2.0, 0.5], [2.0, 3.0, 0.1], [0.5, 0.1, 2.0]])
    mu0 = np.array([0.0, 1.0, -1.0])
    p = orc.Prior.niw(mu0, 10.0, Lam0, 11.0)
    X = rng.standard_normal((25, d)) + 2
    mu, U, kn, nn = p.posterior(_suff(p, X))
    xbar = X.mean(0)
    S = (X - xbar).T @ (X - xbar)
    Lam = Lam0 + S + 10.0 * 25 / 35.0 * np.outer(xbar - mu0, xbar - mu0)
    np.testing.assert_allclose(U.T @ U, Lam, rtol=1e-13)
    np.testing.assert_allclose(mu, (10 * mu0 + X.sum(0)) / 35, rtol=1e-14)
    assert (kn, nn) == (35.0, 36.0)


def test_weights_equal_marginal_posterior_over_all_partitions():
    # test/particle.jl:46-54 : exhaustive expansion over x = 1.:5. ; 52 = Bell(5) particles,
    # weight.(ps) ~ marginal_posterior.(ps) = exp(sum mll + sum lgamma(n) + K log alpha)
    prior = orc.Prior.nix2(0.0, 1.0, 0.1, 0.1)
    f = orc.OracleFilter(orc.FEARNHEAD, 10 ** 6, prior, alpha=2.0)
    bell = [1, 2, 5, 15, 52]
    for t, x in enumerate([1.0, 2.0, 3.0, 4.0, 5.0]):
        M = f.fh_expand(x)
        assert M == bell[t]
        f.fh_commit(np.arange(M), f.putative_weights())      # instantiate every putative
        w = f.weights()
        mp = np.array([math.exp(f.marginal_log_posterior(i)) for i in range(f.n)])
        np.testing.assert_allclose(w, mp, rtol=1e-11)
    # every particle is a distinct partition and the labels are in order of appearance
    A = f.assignments()
    assert A.shape == (5, 52)
    assert len({tuple(c) for c in A.T}) == 52
    assert np.all(A[0] == 1)
    assert np.all(A.max(axis=0) == f.ncomponents())


def test_putatives_equal_old_fit():
    # test/particle.jl:4-44 : second, independent statement of the one-step update
    rng = np.random.default_rng(0)
    prior = orc.Prior.nix2(0.0, 1.0, 0.1, 0.1)
    alpha = 2.0
    f = orc.OracleFilter(orc.FEARNHEAD, 10 ** 6, prior, alpha=alpha)
    # python-side mirror population: list of (components[list of ss], counts, weight)
    pop = [([], [], 1.0)]
    for y in rng.random(6):
        M = f.fh_expand(y)
        wput = f.putative_weights()
        new = []
        for comps, counts, w in pop:
            for x in range(1, len(comps) + 2):
                dlw = math.log(alpha) if x == len(comps) + 1 else math.log(counts[x - 1])
                c2 = [c.copy() for c in comps]
                n2 = list(counts)
                if x <= len(comps):
                    dlw -= prior.mll(c2[x - 1])
                    n2[x - 1] += 1.0
                else:
                    c2.append(prior.empty())
                    n2.append(1.0)
                c2[x - 1] = prior.add(c2[x - 1], y)
                dlw += prior.mll(c2[x - 1])
                new.append((c2, n2, math.exp(math.log(w) + dlw)))
        assert len(new) == M
        np.testing.assert_allclose(wput, [p[2] for p in new], rtol=1e-13)
        f.fh_commit(np.arange(M), wput)
        pop = new
        for i in (0, f.n // 2, f.n - 1):
            st, crp = f.particle(i)
            assert np.all(st == np.stack(pop[i][0]))
            assert np.all(crp == np.array(pop[i][1]))


def test_crp_counts():
    # test/statepriors.jl:4-30 through the filter: counts follow assignments
    prior = orc.Prior.nix2(0.0, 1.0, 1.0, 1.0)
    f = orc.OracleFilter(orc.FEARNHEAD, 100, prior, alpha=0.5)
    f.fh_expand(0.1)
    f.fh_commit([0], [1.0])            # -> N = [1.]
    assert f.particle(0)[1].tolist() == [1.0]
    assert f.fh_expand(0.2) == 2       # candidates 1:2
    f.fh_commit([0], [1.0])            # -> N = [2.]
    assert f.particle(0)[1].tolist() == [2.0]
    f.fh_expand(0.3)
    f.fh_commit([1], [1.0])            # new table -> N = [2., 1.]
    assert f.particle(0)[1].tolist() == [2.0, 1.0]
    assert f.fh_expand(0.4) == 3       # candidates 1:3


# ------------------------------------------------------------------ other state priors (src/statepriors.jl:93-221)
def _nix2():
    return orc.Prior.nix2(0.0, 1.0, 2.0, 3.0)


def test_sticky_crp_with_rho_0_is_the_crp():
    """test/statepriors.jl:36-51: with rho = 0 the sticky CRP's non-sticky candidates carry the CRP's log-priors
    (the sticky candidate has log-prior -inf) and the filters agree."""
    rng = np.random.default_rng(3)
    ys = np.where(rng.random(40) < 0.5, -2.0, 2.0) + rng.standard_normal(40)
    us = rng.random(40)
    a = orc.OracleFilter(orc.FEARNHEAD, 30, _nix2(), 0.5, order=orc.ORDER_INDEX)
    b = orc.OracleFilter(orc.FEARNHEAD, 30, _nix2(), 0.5, order=orc.ORDER_INDEX, stateprior=("sticky", 0.5, 0.0))
    for t, (y, u) in enumerate(zip(ys, us)):
        if t:
            i0 = int(np.argmax(b.weights() > 0))          # first live particle of the sticky filter = particle 0 of the CRP one
            assert b.sp_candidates(i0) == [0] + a.sp_candidates(0)
            assert b.sp_log_prior(i0, 0) == -np.inf
            assert [b.sp_log_prior(i0, x) for x in a.sp_candidates(0)] == [a.sp_log_prior(0, x) for x in a.sp_candidates(0)]
        else:
            assert b.sp_candidates(0) == [1] == a.sp_candidates(0)            # isempty: 1:1, src/statepriors.jl:105
        a.fh_step(y, u)
        b.fh_step(y, u)
        live = b.weights() > 0                # (the sticky putatives have weight exp(-inf) = 0; they are kept while M <= N)
        assert np.array_equal(a.ncomponents(), b.ncomponents()[live])
        np.testing.assert_allclose(a.weights(), b.weights()[live], rtol=1e-12)
        if t > 8:
            break


def test_sticky_crp_bookkeeping():
    """test/statepriors.jl:53-62 and the rules of src/statepriors.jl:106-128: a sticky visit leaves N alone, any
    repeat of the previous state counts in Nsame, `last` follows the assignment; log-prior of the sticky candidate
    is logit(rho) + log(sum(N) + alpha) (:133-135); an invalid state is an ArgumentError (:141)."""
    f = orc.OracleFilter(orc.FEARNHEAD, 1000, _nix2(), 0.5, order=orc.ORDER_INDEX, stateprior=("sticky", 0.5, 0.9))
    f.fh_step(0.3, 0.5)                      # first observation: one candidate, a new component, not sticky
    st, N = f.particle(0)
    assert N.tolist() == [1.0] and f.sticky(0)[1] == 1 and f.sticky(0)[0].tolist() == [1.0]      # x == last (initially 1): Nsame counts it, :119-121
    f.fh_step(0.1, 0.5)                      # candidates 0 (stick to 1), 1, 2
    assert f.n == 3
    (_, N0), (_, N1), (_, N2) = f.particle(0), f.particle(1), f.particle(2)
    assert N0.tolist() == [1.0] and N1.tolist() == [2.0] and N2.tolist() == [1.0, 1.0]      # sticky: N unchanged
    assert f.sticky(0)[0].tolist() == [2.0] and f.sticky(1)[0].tolist() == [2.0] and f.sticky(2)[0].tolist() == [1.0, 0.0]
    assert [f.sticky(i)[1] for i in range(3)] == [1, 1, 2]
    assert [int(a) for a in f.last_assignments()] == [1, 1, 2]
    assert f.sp_log_prior(1, 0) == pytest.approx(math.log(0.9 / 0.1) + math.log(2.0 + 0.5))
    assert f.sp_log_prior(1, 1) == pytest.approx(math.log(2.0)) and f.sp_log_prior(1, 2) == pytest.approx(math.log(0.5))
    assert math.isnan(f.sp_log_prior(1, 3))


def test_changepoint_prior():
    """src/statepriors.jl:162-190: candidates max(k,1):k+1, log-prior 0 for the first point, then log(1-p) to stay
    and log(p) to change; p outside [0, 1] is an ArgumentError (:169)."""
    f = orc.OracleFilter(orc.FEARNHEAD, 100, _nix2(), 1.0, order=orc.ORDER_INDEX, stateprior=("changepoint", 0.2))
    assert f.sp_candidates(0) == [1] and f.sp_log_prior(0, 1) == 0.0
    f.fh_step(0.0, 0.5)
    assert f.sp_candidates(0) == [1, 2]
    assert f.sp_log_prior(0, 1) == pytest.approx(math.log(0.8)) and f.sp_log_prior(0, 2) == pytest.approx(math.log(0.2))
    for y in (0.1, 5.0, 5.2):
        f.fh_step(y, 0.5)
    assert f.n == 8                           # 2^(T-1) segmentations
    A = f.assignments()
    assert np.all(np.diff(A, axis=0) >= 0) and np.all(np.diff(A, axis=0) <= 1) and np.all(A[0] == 1)
    # the change between the 2nd and 3rd observation is the most probable segmentation
    assert A[:, int(np.argmax(f.weights()))].tolist() == [1, 1, 2, 2]
    with pytest.raises(ValueError):
        orc.OracleFilter(orc.FEARNHEAD, 10, _nix2(), 1.0, stateprior=("changepoint", 1.5))


def test_nstate_prior_and_labeled_observations():
    """test/labeled.jl:21-64: NStatePrior([1, 1]) particles with two empty components; a labeled observation has the
    single putative fit(p, y, label) (src/labeled.jl:10-11); missing labels filter like plain observations."""
    rng = np.random.default_rng(11)
    ys = np.where(rng.random(20) < 0.5, -2.0, 2.0) + rng.standard_normal(20)
    labels = np.where(ys > 0, 1, 2)
    us = rng.random(20)
    mk = lambda: orc.OracleFilter(orc.FEARNHEAD, 10, _nix2(), 1.0, order=orc.ORDER_SORTED, stateprior=("nstate", [1.0, 1.0]))
    f1, f2, f3 = mk(), mk(), mk()
    assert f1.sp_candidates(0) == [1, 2]
    assert f1.sp_log_prior(0, 1) == pytest.approx(math.log(0.5))
    for y, u, lab in zip(ys, us, labels):
        f1.fh_step(y, u)
        f2.fh_step_labeled(y, u, None)        # "filter equivalence with missing"
        f3.fh_step_labeled(y, u, int(lab))    # "filtering with labels"
    np.testing.assert_array_equal(f1.weights(), f2.weights())
    assert np.array_equal(f1.assignments(), f2.assignments())
    assert f3.n == 1                          # only one particle
    st, counts = f3.particle(0)
    assert [int(r[0]) for r in st] == [int((labels == 1).sum()), int((labels == 2).sum())]     # nobs.(pp.components)
    assert counts.tolist() == [1.0 + (labels == 1).sum(), 1.0 + (labels == 2).sum()]
    assert f3.assignments()[:, 0].tolist() == labels.tolist()
