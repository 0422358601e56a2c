"""State priors other than the CRP (src/statepriors.jl:93-221) and Labeled observations (src/labeled.jl) on the
device against the CPU oracle: putative weights (1e-10 relative in log space), ancestors, assignments and the
priors' own counts, observation by observation.  Fixtures ported from test/statepriors.jl:34-62 and
test/labeled.jl:21-64."""
import math

import numpy as np
import pytest

import particles_b200 as pb
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _logclose(dev, ref, what):
    dev, ref = np.asarray(dev), np.asarray(ref)
    assert dev.shape == ref.shape, what
    pos = ref > 0
    assert np.array_equal(dev > 0, pos), what
    err = np.abs(np.log(dev[pos]) - np.log(ref[pos])) / np.maximum(np.abs(np.log(ref[pos])), 1.0)
    assert err.size == 0 or err.max() <= RTOL, f"{what}: max rel err {err.max():.3e}"


def _priors(d):
    if d == 1:
        return pb.NormalInverseChisq(0.0, 1.0, 2.0, 3.0), orc.Prior.nix2(0.0, 1.0, 2.0, 3.0)
    return pb.NormalInverseWishart(np.zeros(d), 0.1, np.eye(d), d + 1.0), orc.Prior.niw(np.zeros(d), 0.1, np.eye(d), d + 1.0)


def _stream(rng, T, d):
    if d == 1:
        return (np.array([-4.0, 0.0, 4.0])[rng.integers(0, 3, T)] + rng.standard_normal(T)).reshape(-1, 1)
    ang = 2 * np.pi * rng.integers(0, 4, T) / 4
    return 5 * np.stack([np.cos(ang), np.sin(ang)], axis=1) + rng.standard_normal((T, 2))


def _compare_stepwise(dev, ref, ys, us, labels=None):
    for t, (y, u) in enumerate(zip(ys, us)):
        lab = None if labels is None else labels[t]
        dev.fit_(pb.Labeled(y, lab) if labels is not None else y, [u])
        ref.fh_step_labeled(y, u, lab)
        assert len(dev) == ref.n, f"population differs at step {t}"
        _logclose(dev.putative_weights(), ref.putative_weights(), f"putatives step {t}")
        assert np.array_equal(dev.last_ancestors(), ref.last_ancestors()), f"ancestors differ at step {t}"
        assert np.array_equal(dev.last_assignments() + 1, ref.last_assignments()), f"assignments differ at step {t}"
        assert np.array_equal(dev.ncomponents(), ref.ncomponents())
        _logclose(dev.weights(), ref.weights(), f"weights step {t}")
        s = dev.last_step()
        assert s["log_total_w"] == pytest.approx(math.log(ref.stat("total_w")), rel=RTOL, abs=1e-10)
    assert np.array_equal(dev.assignments(), ref.assignments())


@pytest.mark.parametrize("d,order", [(1, orc.ORDER_INDEX), (1, orc.ORDER_SORTED), (2, orc.ORDER_INDEX)])
@pytest.mark.parametrize("sp", ["sticky", "sticky0", "changepoint", "nstate"])
def test_stateprior_filter_matches_oracle(sp, d, order):
    if sp == "nstate" and order == orc.ORDER_SORTED:
        # states that start out exchangeable give particles whose weights agree to the last bits: the ascending-weight
        # traversal then depends on roundings the device (t-form predictive) and the oracle (mll differences) do not share
        pytest.skip("near-tied weights: the sorted traversal is not comparable bit for bit")
    rng = np.random.default_rng(40 + d)
    T, N = 45, 300
    ys, us = _stream(rng, T, d), rng.random(T)
    pprior, oprior = _priors(d)
    dsp, osp = {"sticky": (pb.StickyCRP(0.7, 0.4), ("sticky", 0.7, 0.4)),
                "sticky0": (pb.StickyCRP(0.5, 0.0), ("sticky", 0.5, 0.0)),
                "changepoint": (pb.ChangePoint(0.15), ("changepoint", 0.15)),
                "nstate": (pb.NStatePrior([1.0, 2.0, 0.5]), ("nstate", [1.0, 2.0, 0.5]))}[sp]
    dorder = pb.PF_ORDER_SORTED if order == orc.ORDER_SORTED else pb.PF_ORDER_INDEX
    dev = pb.FearnheadParticles(N, pprior, dsp, history=T, order=dorder, k_max=30)
    ref = orc.OracleFilter(orc.FEARNHEAD, N, oprior, 1.0, order=order, stateprior=osp)
    _compare_stepwise(dev, ref, ys, us)
    if sp.startswith("sticky"):
        for i in (0, len(dev) // 2, len(dev) - 1):
            p = dev.particles()[i]
            st, N_o = ref.particle(i)
            Nsame_o, last_o = ref.sticky(i)
            q = p.stateprior
            assert q.N.tolist() == N_o.tolist() and q.Nsame.tolist() == Nsame_o.tolist() and q.last == last_o
            assert [c.n for c in p.components] == [int(r[0]) for r in st]
    if sp == "changepoint":
        A = dev.assignments()
        assert np.all(np.diff(A, axis=0) >= 0) and np.all(np.diff(A, axis=0) <= 1)       # states only ever step up by one


def test_labeled_observations_fixture():
    """test/labeled.jl:21-64 on the device."""
    rng = np.random.default_rng(7)
    ys = np.where(rng.random(20) < 0.5, -2.0, 2.0) + rng.standard_normal(20)
    labels = np.where(ys > 0, 1, 2)
    us = rng.random(20)
    prior = pb.NormalInverseChisq(0.0, 1.0, 2.0, 3.0)
    mk = lambda: pb.FearnheadParticles(10, prior, pb.NStatePrior([1.0, 1.0]), order=pb.PF_ORDER_SORTED, k_max=8)
    # "filter equivalence with missing"
    f1, f2 = mk(), mk()
    f1.filter_(ys, uniforms=us)
    f2.filter_([pb.Labeled(y) for y in ys], uniforms=us)
    assert np.array_equal(f1.logweights(), f2.logweights()) and np.array_equal(f1.assignments(), f2.assignments())
    # "filtering with labels": only one particle, its components hold the labelled observations
    f3 = mk()
    f3.filter_([pb.Labeled(y, int(l)) for y, l in zip(ys, labels)], uniforms=us)
    assert len(f3) == 1
    pp = f3.particles()[0]
    assert pb.nobs(pp) == len(ys)
    assert [c.n for c in pp.components] == [int((labels == 1).sum()), int((labels == 2).sum())]
    assert f3.assignments()[:, 0].tolist() == labels.tolist()
    assert pp.stateprior.counts.tolist() == [1.0 + (labels == 1).sum(), 1.0 + (labels == 2).sum()]
    # same component parameters whether the unlabelled observations come before or after the labelled ones
    extra = [pb.Labeled(v) for v in (-1.0, 0.0, 1.0)]
    lab = [pb.Labeled(y, int(l)) for y, l in zip(ys, labels)]
    fb, fa = mk(), mk()
    ub = rng.random(23)
    fb.filter_(extra + lab, uniforms=ub)
    fa.filter_(lab + extra, uniforms=ub)
    def params(ps):
        return sorted([tuple(np.round([c.params()[0], c.params()[1], c.params()[2], c.params()[3]], 9)) for p in ps.particles() for c in p.components])
    assert params(fb) == params(fa)


@pytest.mark.parametrize("sp", ["crp", "sticky", "nstate"])
def test_labeled_stream_matches_oracle(sp):
    """A half-labelled stream through the device and the oracle, step by step."""
    rng = np.random.default_rng(99)
    T, N = 40, 200
    ys = _stream(rng, T, 1)[:, 0]
    us = rng.random(T)
    truth = np.digitize(ys, [-2.0, 2.0]) + 1                 # 1, 2, 3
    labels = [int(truth[t]) if (t % 2 == 0 and t > 6) else None for t in range(T)]
    if sp != "nstate":
        # under a CRP a label can only name an existing component or the next new one: use labels the population can take
        labels = [1 if (t % 3 == 0 and t > 0) else None for t in range(T)]
    pprior, oprior = _priors(1)
    dsp, osp = {"crp": (pb.ChineseRestaurantProcess(1.0), None), "sticky": (pb.StickyCRP(1.0, 0.3), ("sticky", 1.0, 0.3)),
                "nstate": (pb.NStatePrior([1.0, 1.0, 1.0]), ("nstate", [1.0, 1.0, 1.0]))}[sp]
    dev = pb.FearnheadParticles(N, pprior, dsp, history=T, k_max=30)
    ref = orc.OracleFilter(orc.FEARNHEAD, N, oprior, 1.0, order=orc.ORDER_INDEX, stateprior=osp)
    _compare_stepwise(dev, ref, ys, us, labels)


def test_invalid_label_is_an_argument_error():
    prior = pb.NormalInverseChisq(0.0, 1.0, 2.0, 3.0)
    ps = pb.FearnheadParticles(50, prior, pb.ChineseRestaurantProcess(1.0), k_max=8)
    ps.fit_(0.3, [0.5])
    with pytest.raises(ValueError):
        ps.fit_(pb.Labeled(0.1, 3), [0.5])                   # one component so far: valid labels are 1 and 2
    with pytest.raises(ValueError):
        pb.ChangePoint(1.5)
