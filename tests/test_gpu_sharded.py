"""Sharded run == unsharded run, bit for bit, with all ranks emulated inside one
process on one GPU (LocalGroup, lock step): the same kernels and the same peer-memory exchanges
(sample / candidates / survivor totals / children that change owner) as the one-process-per-GPU run,
with "publish" and "consume" queued as separate launches on one stream."""
import numpy as np
import pytest

import particles_b200 as pb
from particles_b200.distributed import LocalGroup, ShardedFearnheadParticles, equal_blocks

pytestmark = pytest.mark.gpu


def _mix2d(rng, T, k=8, radius=6.0):
    ang = 2 * np.pi * rng.integers(0, k, T) / k
    return radius * np.stack([np.cos(ang), np.sin(ang)], axis=1) + rng.standard_normal((T, 2))


def _same_population(sh, ref, what):
    assert len(sh) == len(ref), f"population differs {what}"
    assert np.array_equal(sh.ncomponents(), ref.ncomponents()), f"K differs {what}"
    assert np.array_equal(sh.last_ancestors(), ref.last_ancestors()), f"ancestors differ {what}"
    assert np.array_equal(sh.last_assignments(), ref.last_assignments()), f"assignments differ {what}"
    assert np.array_equal(sh.logweights(), ref.logweights()), f"log-weights differ {what}"
    a, b = sh.last_step(), ref.last_step()
    for key in ("n_out", "log_total_w", "log_w_resamp", "threshold", "resampled", "n_resamp_req"):
        assert a[key] == b[key] or (np.isnan(a[key]) and np.isnan(b[key])), (key, what)


@pytest.mark.parametrize("G,N,T", [(2, 1000, 40), (4, 1000, 40), (2, 1 << 16, 30), (8, 3000, 25), (3, 5000, 30), (16, 700, 20)])
def test_sharded_equals_unsharded_stepwise(G, N, T):
    rng = np.random.default_rng(5)
    ys = _mix2d(rng, T)
    us = rng.random(T)
    prior = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0)
    crp = pb.ChineseRestaurantProcess(1.0)
    ref = pb.FearnheadParticles(N, prior, crp, history=T, k_max=24)
    sh = ShardedFearnheadParticles(N, prior, crp, LocalGroup(G), history=T, k_max=24)
    for t in range(T):
        ref.fit_(ys[t], [us[t]])
        sh.fit_(ys[t], us[t])
        _same_population(sh, ref, f"at step {t}")
        # equal blocks after every step
        assert sh.local_sizes() == np.diff(equal_blocks(len(sh), G)).tolist()
    assert np.array_equal(sh.assignments(), ref.assignments())


@pytest.mark.parametrize("G,N,T,alpha", [(4, 1 << 17, 40, 1.0), (8, 1 << 18, 30, 3.0), (2, 1 << 18, 30, 0.3)])
def test_sharded_equals_unsharded_one_call(G, N, T, alpha):
    """The whole stream in ONE pf_mg_filter call (no host round trip between observations), at sizes where the
    sample stride is > 1 and the candidate lists hold 10^4..10^5 values per rank."""
    rng = np.random.default_rng(50 + G)
    ys = _mix2d(rng, T)
    us = rng.random(T)
    prior = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0)
    crp = pb.ChineseRestaurantProcess(alpha)
    ref = pb.FearnheadParticles(N, prior, crp, history=T, k_max=24)
    ref.filter_(ys, uniforms=us)
    sh = ShardedFearnheadParticles(N, prior, crp, LocalGroup(G), history=T, k_max=24)
    sh.filter_(ys, us)
    _same_population(sh, ref, "after the stream")
    assert np.array_equal(sh.log_evidence, ref.log_evidence)
    # the ancestor chains cross the shards: every shard walks them through the peers' genealogy arrays
    assert np.array_equal(sh.assignments(), ref.assignments())
    # filter-level summaries from per-shard sums: integer similarity counts are exact, the floating-point sums differ from
    # the unsharded reduction only in the order of their block partials
    assert np.array_equal(sh.assignment_similarity(), ref.assignment_similarity())
    xs = np.concatenate([ys[:5], 9.0 * rng.standard_normal((6, 2))])
    np.testing.assert_allclose(sh.posterior_predictive(xs), ref.posterior_predictive(xs), rtol=1e-12)
    assert sh.state_entropy() == pytest.approx(pb.state_entropy(ref), rel=1e-12)
    assert sh.mean_ncomponents() == pytest.approx(pb.mean(pb.ncomponents, ref), rel=1e-12)
    np.testing.assert_allclose(sh.ncomponents_dist(), ref.ncomponents_dist(), rtol=1e-12)


def test_sharded_1d_nix2():
    rng = np.random.default_rng(9)
    T, N, G = 60, 20000, 4
    ys = np.array([-6.0, -2.0, 2.0, 6.0])[rng.integers(0, 4, T)] + rng.standard_normal(T)
    us = rng.random(T)
    prior = pb.NormalInverseChisq(0.0, 1.0, 0.1, 1.0)
    crp = pb.ChineseRestaurantProcess(1.0)
    ref = pb.FearnheadParticles(N, prior, crp, history=T, k_max=24)
    ref.filter_(ys, uniforms=us)
    sh = ShardedFearnheadParticles(N, prior, crp, LocalGroup(G), history=T, k_max=24)
    sh.filter_(ys.reshape(-1, 1), us)
    _same_population(sh, ref, "after the stream")


def test_failed_bracket_retries_on_the_device(monkeypatch):
    """A sample bracket that misses the threshold (forced by shrinking its margin from 8 sigma to 1) is moved /
    widened and the round is repeated by the queued retry launches, including the second exchange of the
    candidates; decisions equal the default run, the total weight agrees to rounding (a different bracket splits
    it differently)."""
    rng = np.random.default_rng(77)
    G, N, T = 4, 1 << 17, 30
    ys = _mix2d(rng, T)
    us = rng.random(T)
    prior = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0)
    crp = pb.ChineseRestaurantProcess(1.0)
    ref = pb.FearnheadParticles(N, prior, crp, history=T, k_max=16)
    ref.filter_(ys, uniforms=us)
    monkeypatch.setenv("PF_SEL_MARGIN", "1")
    sh = ShardedFearnheadParticles(N, prior, crp, LocalGroup(G), history=T, k_max=16)
    rounds = []
    for t in range(T):
        sh.fit_(ys[t], us[t])
        rounds.append(sh.last_step()["select_rounds"])
    assert max(rounds) >= 2
    assert np.array_equal(sh.ncomponents(), ref.ncomponents())
    assert np.array_equal(sh.last_ancestors(), ref.last_ancestors())
    assert np.array_equal(sh.last_assignments(), ref.last_assignments())
    np.testing.assert_allclose(sh.logweights(), ref.logweights(), rtol=0, atol=1e-12)


def test_exchange_overflow_is_reported_not_silent(monkeypatch):
    """A candidate list that outgrows its exchange segment halts the step on the device and comes back as
    PF_ERR_OVERFLOW (never a truncated gather)."""
    monkeypatch.setenv("PF_MG_CAPB", "64")
    rng = np.random.default_rng(3)
    G, N, T = 2, 1 << 15, 24
    ys = _mix2d(rng, T)
    prior = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0)
    sh = ShardedFearnheadParticles(N, prior, pb.ChineseRestaurantProcess(1.0), LocalGroup(G), k_max=16)
    with pytest.raises(pb.PFError) as e:
        sh.filter_(ys, rng.random(T))
    assert e.value.code == 5 and "exchange segment" in str(e.value)


# ------------------------------------------------------------------ Chen-Liu (src/chenliu.jl:54-69), sharded
def _same_chenliu(sh, ref, what):
    assert np.array_equal(sh.ncomponents(), ref.ncomponents()), f"K differs {what}"
    assert np.array_equal(sh.last_ancestors(), ref.last_ancestors()), f"ancestors differ {what}"
    assert np.array_equal(sh.last_assignments(), ref.last_assignments()), f"assignments differ {what}"
    assert np.array_equal(sh.logweights(), ref.logweights()), f"log-weights differ {what}"
    a, b = sh.last_step(), ref.last_step()
    for key in ("n_out", "log_total_w", "cv", "resampled"):
        assert a[key] == b[key], (key, what)


@pytest.mark.parametrize("G,N,T,d", [(2, 1000, 60, 1), (4, 1003, 60, 2), (8, 4096, 40, 2), (3, 600, 30, 8), (16, 5000, 30, 1)])
def test_sharded_chenliu_equals_unsharded_stepwise(G, N, T, d):
    """All decisions (categorical draws, CV trigger, stratified rejuvenation across shards) and the weights are those
    of the one-GPU run, bit for bit; N not divisible by G leaves the last shard shorter."""
    from particles_b200.distributed import ShardedChenLiuParticles
    rng = np.random.default_rng(40 + G)
    if d == 1:
        ys = (np.array([-6.0, -2.0, 2.0, 6.0])[rng.integers(0, 4, T)] + rng.standard_normal(T)).reshape(-1, 1)
        prior = pb.NormalInverseChisq(0.0, 1.0, 0.1, 1.0)
    else:
        ys = _mix2d(rng, T) if d == 2 else 6.0 * np.sign(rng.standard_normal((8, d)))[rng.integers(0, 8, T)] + rng.standard_normal((T, d))
        prior = pb.NormalInverseWishart(np.zeros(d), 0.1, np.eye(d), d + 1.0)
    U = rng.random((T, 2 * N))
    crp = pb.ChineseRestaurantProcess(1.0)
    ref = pb.ChenLiuParticles(N, prior, crp, history=T, k_max=40)
    sh = ShardedChenLiuParticles(N, prior, crp, LocalGroup(G), history=T, k_max=40)
    n_rejuv = 0
    for t in range(T):
        ref.fit_(ys[t], U[t])
        sh.fit_(ys[t], U[t])
        _same_chenliu(sh, ref, f"at step {t}")
        n_rejuv += ref.last_step()["resampled"]
    assert n_rejuv >= (2 if d < 8 else 0)
    assert np.array_equal(sh.assignments(), ref.assignments())


@pytest.mark.parametrize("G,N,T", [(4, 1 << 17, 40), (8, 1 << 18, 30)])
def test_sharded_chenliu_one_call_device_uniforms(G, N, T):
    """The whole stream in one pf_mg_filter call, uniforms drawn on the device from the global particle index."""
    from particles_b200.distributed import ShardedChenLiuParticles
    rng = np.random.default_rng(60 + G)
    ys = _mix2d(rng, T)
    prior = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0)
    crp = pb.ChineseRestaurantProcess(1.0)
    ref = pb.ChenLiuParticles(N, prior, crp, history=T, k_max=32, seed=11)
    ref.filter_(ys)
    sh = ShardedChenLiuParticles(N, prior, crp, LocalGroup(G), history=T, k_max=32, seed=11)
    sh.filter_(ys)
    _same_chenliu(sh, ref, "after the stream")
    assert np.array_equal(sh.log_evidence, ref.log_evidence)
    assert np.array_equal(sh.assignments(), ref.assignments())
    assert np.array_equal(sh.assignment_similarity(), ref.assignment_similarity())
    np.testing.assert_allclose(sh.posterior_predictive(ys[:4]), ref.posterior_predictive(ys[:4]), rtol=1e-12)
