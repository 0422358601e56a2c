"""The oracle's filter-level summaries (src/filters.jl:17-42, 53-80) pinned on the CPU: the reference's own
invariants for the clustering metrics (test/clustering.jl:6-20) and an independent scipy restatement of the
posterior predictive mixture (src/particle.jl:156-172, src/component.jl:12-23)."""
import numpy as np
import pytest
from scipy import stats

from oracle import oracle as orc


def _run(kind, n=100, T=50, seed=3):
    rng = np.random.default_rng(seed)
    x = rng.random(T)                                        # x = rand(50), test/clustering.jl:2
    f = orc.OracleFilter(kind, n, orc.Prior.nix2(0.0, 1.0, 2.0, 2.0), 1.0)
    for t in range(T):
        if kind == orc.FEARNHEAD:
            f.fh_step(x[t], rng.random())
        else:
            f.cl_step(x[t], rng.random(n), rng.random(n))
    return f, x


@pytest.mark.parametrize("kind", [orc.FEARNHEAD, orc.CHENLIU])
def test_clustering_metrics_invariants(kind):
    """test/clustering.jl:9-20."""
    f, _ = _run(kind)
    a = f.assignments()
    assert a.shape == (50, 100) and a.dtype == np.int64
    assert np.array_equal(a.max(axis=0), f.ncomponents())
    assert (a.min(axis=0) == 1).all()
    sim = f.assignment_similarity()
    assert sim.shape == (50, 50)
    assert np.array_equal(sim, sim.T)
    assert ((0 <= sim) & (sim <= 1)).all()
    assert (np.diag(sim) == 1).all()
    # Rand index of a clustering with itself is 1; against a constant labelling the plain index is mean(sim)
    assert f.randindex(np.zeros(50, dtype=int), "plain") == pytest.approx(sim.mean(), rel=1e-14)
    p = f.ncomponents_dist()
    assert p.sum() == pytest.approx(1.0, rel=1e-12) and len(p) == f.ncomponents().max()


def test_randindex_identical_clusterings():
    """All particles carrying the SAME assignment vector as the truth give (adjusted) Rand index 1."""
    f, x = _run(orc.FEARNHEAD, n=1, T=12)
    truth = f.assignments()[:, 0]
    assert f.randindex(truth, "plain") == pytest.approx(1.0, abs=1e-15)
    assert f.randindex(truth, "adjusted") == pytest.approx(1.0, abs=1e-12)
    with pytest.raises(ValueError):
        f.randindex(truth[:-1])


def test_posterior_predictive_against_scipy():
    """Mixture over particles and components of Student-t densities, restated with scipy from the particle's
    sufficient statistics: NIX2 posterior (mu_n, sigma2_n, kappa_n, nu_n) -> t_{nu_n}(mu_n, sigma2_n (1 + 1/kappa_n))."""
    f, x = _run(orc.FEARNHEAD, n=30, T=15, seed=5)
    xs = np.linspace(-2.0, 3.0, 11)
    got = f.posterior_predictive(xs)
    w = f.weights()
    w = w / w.sum()
    alpha, T = 1.0, 15
    want = np.zeros_like(xs)
    pr = f.prior
    for i in range(f.n):
        st, crpN = f.particle(i)
        comps = [pr.posterior(s) for s in st] + [(pr.mu0[0], pr.s20, pr.kappa0, pr.nu0)]
        pis = np.concatenate([crpN, [alpha]]) / (T + alpha)
        for (mu, s2, k, nu), pi in zip(comps, pis):
            want += w[i] * pi * stats.t.pdf(xs, df=nu, loc=mu, scale=np.sqrt(s2 * (1 + k) / k))
    np.testing.assert_allclose(got, want, rtol=1e-11)
    # a density: integrates to one
    grid = np.linspace(-40, 40, 3201)
    assert np.trapezoid(f.posterior_predictive(grid), grid) == pytest.approx(1.0, abs=5e-3)


def test_state_entropy_crp_closed_form():
    f, _ = _run(orc.FEARNHEAD, n=40, T=20, seed=7)
    w = f.weights()
    H = []
    for i in range(f.n):
        _, crpN = f.particle(i)
        p = np.concatenate([crpN, [1.0]]) / (20 + 1.0)
        H.append(-(p * np.log(p)).sum())
    assert f.state_entropy() == pytest.approx(np.dot(H, w) / w.sum(), rel=1e-13)
