"""Filter-level summaries evaluated on the device (pf_posterior_predictive, pf_weighted_mean, pf_ncomponents_dist,
pf_assignment_similarity, pf_randindex; src/filters.jl:17-42, 53-80) against the CPU oracle's restatement, and
checkpoint / restore (pf_save, pf_load): a loaded filter continues bit for bit."""
import os

import numpy as np
import pytest

import particles_b200 as pb
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _model(d):
    if d == 1:
        return pb.NormalInverseChisq(0.0, 1.0, 0.1, 1.0), orc.Prior.nix2(0.0, 1.0, 0.1, 1.0)
    return pb.NormalInverseWishart(np.zeros(d), 0.1, np.eye(d), d + 1.0), orc.Prior.niw(np.zeros(d), 0.1, np.eye(d), d + 1.0)


def _stream(rng, T, d):
    if d == 1:
        return (np.array([-4.0, 0.0, 4.0])[rng.integers(0, 3, T)] + rng.standard_normal(T)).reshape(-1, 1)
    ang = 2 * np.pi * rng.integers(0, 4, T) / 4
    pts = 5 * np.stack([np.cos(ang), np.sin(ang)], axis=1)
    if d > 2:
        pts = np.concatenate([pts, np.zeros((T, d - 2))], axis=1)
    return pts + rng.standard_normal((T, d))


def _pair(kind, d, n, T, seed, sp=None, order=pb.PF_ORDER_INDEX):
    rng = np.random.default_rng(seed)
    ys = _stream(rng, T, d)
    pprior, oprior = _model(d)
    if kind == "fearnhead":
        us = rng.random(T)
        psp = {None: pb.ChineseRestaurantProcess(1.0), "sticky": pb.StickyCRP(1.0, 0.4), "changepoint": pb.ChangePoint(0.2),
               "nstate": pb.NStatePrior([1.0, 2.0, 1.0])}[sp]
        osp = {None: None, "sticky": ("sticky", 1.0, 0.4), "changepoint": ("changepoint", 0.2), "nstate": ("nstate", [1.0, 2.0, 1.0])}[sp]
        dev = pb.FearnheadParticles(n, pprior, psp, k_max=40, order=order)
        ref = orc.OracleFilter(orc.FEARNHEAD, n, oprior, 1.0, order=orc.ORDER_INDEX if order == pb.PF_ORDER_INDEX else orc.ORDER_SORTED,
                               stateprior=osp)
        dev.filter_(ys, uniforms=us)
        for y, u in zip(ys, us):
            ref.fh_step(y, u)
    else:
        U = rng.random((T, 2 * n))
        dev = pb.ChenLiuParticles(n, pprior, pb.ChineseRestaurantProcess(1.0), k_max=40)
        ref = orc.OracleFilter(orc.CHENLIU, n, oprior, 1.0)
        dev.filter_(ys, uniforms=U)
        for t in range(T):
            ref.cl_step(ys[t], U[t, :n], U[t, n:])
    assert np.array_equal(dev.assignments(), ref.assignments())
    return dev, ref, ys


@pytest.mark.parametrize("kind,d,n,T", [("fearnhead", 1, 150, 30), ("fearnhead", 2, 200, 36), ("fearnhead", 4, 64, 20),
                                        ("chenliu", 1, 128, 30), ("chenliu", 2, 96, 24)])
def test_summaries_match_oracle(kind, d, n, T):
    dev, ref, ys = _pair(kind, d, n, T, seed=11 + d)
    rng = np.random.default_rng(1)
    xs = np.concatenate([ys[:7] + 0.3, 8.0 * rng.standard_normal((70, d))])          # > one 64-point pass
    got, want = pb.posterior_predictive(dev, xs), ref.posterior_predictive(xs)
    assert (got > 0).all()
    np.testing.assert_allclose(np.log(got), np.log(want), rtol=RTOL, atol=RTOL)
    assert pb.state_entropy(dev) == pytest.approx(ref.state_entropy(), rel=RTOL)
    assert pb.mean(pb.ncomponents, dev) == pytest.approx(ref.mean_ncomponents(), rel=RTOL)
    np.testing.assert_allclose(pb.ncomponents_dist(dev), ref.ncomponents_dist(), rtol=RTOL, atol=1e-300)
    A = dev.assignments()
    for i in (0, len(dev) // 3, len(dev) - 1):
        assert np.array_equal(dev.particle_assignments(i), A[:, i])
    # the similarity counts are integers: the matrix is exact
    assert np.array_equal(pb.assignment_similarity(dev), ref.assignment_similarity())
    truth = np.random.default_rng(2).integers(0, 4, T)
    for k in ("adjusted", "plain"):
        assert pb.randindex(dev, truth, k) == pytest.approx(ref.randindex(truth, k), rel=1e-12)
    with pytest.raises(ValueError):                      # DimensionMismatch, src/filters.jl:57-58
        pb.randindex(dev, truth[:-1])
    # mean(f, ps) with a host function of the particle views agrees with the device reduction
    assert pb.mean(lambda p: float(p.K), dev) == pytest.approx(pb.mean(pb.ncomponents, dev), rel=1e-12)


@pytest.mark.parametrize("sp", ["sticky", "changepoint", "nstate"])
def test_state_entropy_other_priors(sp):
    dev, ref, _ = _pair("fearnhead", 1, 120, 25, seed=21, sp=sp)
    assert pb.state_entropy(dev) == pytest.approx(ref.state_entropy(), rel=RTOL, abs=1e-14)
    assert pb.mean(pb.ncomponents, dev) == pytest.approx(ref.mean_ncomponents(), rel=RTOL)
    # the host view of one particle's state prior gives the same entropy as the oracle's
    ps = dev.particles()
    w = dev.weights()
    host = np.dot([pb.state_entropy(p) for p in ps], w) / w.sum()
    assert host == pytest.approx(ref.state_entropy(), rel=1e-9, abs=1e-14)
    with pytest.raises(pb.PFError):                      # components(p) and weights(p) only line up under the CRP
        pb.posterior_predictive(dev, np.zeros((1, 1)))


def test_similarity_large_population_and_slices(monkeypatch):
    """2^16 particles: the 16-bytes-per-load path, and the same matrix when the particles are processed in several
    slices (PF_ASSIGN_CHUNK) and when the slice length is not a multiple of 16."""
    rng = np.random.default_rng(5)
    T, n = 44, 1 << 16
    ys, us = _stream(rng, T, 2), rng.random(T)
    pprior, _ = _model(2)
    dev = pb.FearnheadParticles(n, pprior, pb.ChineseRestaurantProcess(1.0), k_max=24)
    dev.filter_(ys, uniforms=us)
    a = dev.assignments()
    want = np.empty((T, T))
    for t in range(T):
        want[t] = 1.0 - (a != a[t]).sum(axis=1) / n
    assert np.array_equal(dev.assignment_similarity(), want)
    for chunk in ("8192", "10000"):
        monkeypatch.setenv("PF_ASSIGN_CHUNK", chunk)
        assert np.array_equal(dev.assignment_similarity(), want)
        assert np.array_equal(dev.assignments(), a)
    monkeypatch.delenv("PF_ASSIGN_CHUNK")
    # predictive: a density on a grid (integrates to one), positive everywhere
    g = np.linspace(-14, 14, 57)
    X = np.stack(np.meshgrid(g, g), axis=-1).reshape(-1, 2)
    dens = pb.posterior_predictive(dev, X)
    assert (dens > 0).all() and dens.sum() * (g[1] - g[0]) ** 2 == pytest.approx(1.0, abs=0.02)
    assert pb.ncomponents_dist(dev).sum() == pytest.approx(1.0, rel=1e-12)


@pytest.mark.parametrize("case", ["fh_index", "fh_sorted", "fh_sticky", "fh_device_uniforms", "chenliu", "fh_fixed_history"])
def test_checkpoint_continues_bit_for_bit(case, tmp_path):
    rng = np.random.default_rng(9)
    d, T1, T2 = 2, 22, 18
    ys = _stream(rng, T1 + T2, d)
    pprior, _ = _model(d)

    def make():
        if case == "chenliu":
            return pb.ChenLiuParticles(500, pprior, pb.ChineseRestaurantProcess(1.0), k_max=32, seed=4)
        sp = pb.StickyCRP(1.0, 0.3) if case == "fh_sticky" else pb.ChineseRestaurantProcess(1.0)
        return pb.FearnheadParticles(3000, pprior, sp, k_max=32, seed=4,
                                     order=pb.PF_ORDER_SORTED if case == "fh_sorted" else pb.PF_ORDER_INDEX,
                                     history=T1 + T2 if case == "fh_fixed_history" else -1)
    n_u = 2 * 500 if case == "chenliu" else 1
    us = None if case == "fh_device_uniforms" else rng.random((T1 + T2, n_u))
    whole = make()
    whole.filter_(ys, uniforms=us)
    first = make()
    first.filter_(ys[:T1], uniforms=None if us is None else us[:T1])
    path = os.path.join(tmp_path, "ckpt.bin")
    pb.save(first, path)
    first.close()
    second = pb.load(path)
    assert type(second) is type(whole) and second.n_steps == T1 and len(second) > 0
    assert second.prior.kappa == pprior.kappa and np.array_equal(second.prior.Lam, pprior.Lam)
    second.filter_(ys[T1:], uniforms=None if us is None else us[T1:])
    assert len(second) == len(whole)
    assert np.array_equal(second.logweights(), whole.logweights())
    assert np.array_equal(second.ncomponents(), whole.ncomponents())
    assert np.array_equal(second.assignments(), whole.assignments())
    assert np.array_equal(second.log_evidence, whole.log_evidence[T1:])
    p0, p1 = second.particles()[0], whole.particles()[0]
    for c0, c1 in zip(p0.components, p1.components):
        assert c0.n == c1.n and np.array_equal(c0.mean, c1.mean) and np.array_equal(c0.chol, c1.chol)
    assert second.last_step()["overflow"] == whole.last_step()["overflow"]


def test_load_rejects_garbage(tmp_path):
    path = os.path.join(tmp_path, "junk.bin")
    with open(path, "wb") as f:
        f.write(b"not a checkpoint" * 100)
    with pytest.raises(ValueError):
        pb.load(path)
