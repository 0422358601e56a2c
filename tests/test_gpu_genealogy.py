"""history < 0 on a single GPU: the genealogy is an arena pruned of dead lineages when it fills up (the reference's
ancestor chain is pruned by the garbage collector, src/particle.jl:25-32).  Pruning renumbers and compacts the stored
generations; every read-out that walks the ancestor chains must be unchanged by it."""
import os

import numpy as np
import pytest

import particles_b200 as pb

pytestmark = pytest.mark.gpu


def _stream(rng, T, d):
    if d == 1:
        return (np.array([-6.0, -2.0, 2.0, 6.0])[rng.integers(0, 4, T)] + rng.standard_normal(T)).reshape(-1, 1)
    ang = 2 * np.pi * rng.integers(0, 8, T) / 8
    return 6 * np.stack([np.cos(ang), np.sin(ang)], axis=1) + rng.standard_normal((T, 2))


def _prior(d):
    return pb.NormalInverseChisq(0.0, 1.0, 0.1, 1.0) if d == 1 else pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0)


@pytest.mark.parametrize("kind,d,N,T", [("fearnhead", 2, 3000, 150), ("fearnhead", 1, 500, 400), ("fearnhead_sorted", 1, 400, 120),
                                        ("chenliu", 1, 1024, 200), ("sticky", 1, 600, 120)])
def test_pruned_genealogy_equals_dense(kind, d, N, T, monkeypatch, tmp_path):
    rng = np.random.default_rng(17)
    ys = _stream(rng, T, d)
    us = rng.random((T, 2 * N if kind == "chenliu" else 1))

    def make(history):
        if kind == "chenliu":
            return pb.ChenLiuParticles(N, _prior(d), pb.ChineseRestaurantProcess(1.0), k_max=40, history=history)
        sp = pb.StickyCRP(1.0, 0.3) if kind == "sticky" else pb.ChineseRestaurantProcess(1.0)
        return pb.FearnheadParticles(N, _prior(d), sp, k_max=40, history=history,
                                     order=pb.PF_ORDER_SORTED if kind == "fearnhead_sorted" else pb.PF_ORDER_INDEX)
    dense = make(T)
    dense.filter_(ys, uniforms=us)
    A = dense.assignments()
    # a budget of a few generations: the stream is filtered in chunks of 4 observations and the arena pruned every ~12
    monkeypatch.setenv("PF_GENEALOGY_BYTES", str(5 * 20 * (N + 64)))
    pr = make(-1)
    cuts = [0, T // 3, T // 3 + 1, T - 7, T]
    for a, b in zip(cuts[:-1], cuts[1:]):
        pr.filter_(ys[a:b], uniforms=us[a:b])
        if b == T // 3 + 1:
            path = str(tmp_path / "pruned.ckpt")                # a checkpoint of the pruned arena continues identically
            pr.save(path)
            pr.close()
            pr = pb.load(path)
    info = pr.genealogy_info()
    assert info["generations"] == T and info["prunes"] >= 3, info
    if kind != "chenliu":
        assert info["nodes"] < 0.6 * T * len(pr), info          # dead lineages are gone (Fearnhead resamples every step)
    assert np.array_equal(pr.logweights(), dense.logweights())
    assert np.array_equal(pr.assignments(), A)
    assert np.array_equal(pr.last_ancestors(), dense.last_ancestors())
    assert np.array_equal(pr.last_assignments(), dense.last_assignments())
    for i in (0, len(pr) // 2, len(pr) - 1):
        assert np.array_equal(pr.particle_assignments(i), A[:, i])
    assert np.array_equal(pr.assignment_similarity(), dense.assignment_similarity())
    truth = rng.integers(0, 4, T)
    assert pr.randindex(truth) == dense.randindex(truth)


def test_long_stream_keeps_its_history_in_a_small_arena(monkeypatch):
    """2^16 particles x 1 500 observations would be 490 MB dense; pruned, the arena (budget 60 MB, exceeded only by what
    the live lineages need) stays a fraction of that."""
    monkeypatch.setenv("PF_GENEALOGY_BYTES", str(60 << 20))
    rng = np.random.default_rng(3)
    N, T = 1 << 16, 1500
    ys, us = _stream(rng, T, 2), rng.random(T)
    ps = pb.FearnheadParticles(N, _prior(2), pb.ChineseRestaurantProcess(1.0), k_max=32, history=-1, on_overflow="raise")
    ps.filter_(ys, uniforms=us)
    info = ps.genealogy_info()
    assert info["generations"] == T and info["prunes"] >= 5
    print("arena capacity", info["capacity"] * 5 / 1e6, "MB; live nodes", info["nodes"] * 5 / 1e6, "MB; dense", 5 * T * N / 1e6, "MB; prunes", info["prunes"])
    assert info["capacity"] * 5 <= 0.35 * 5 * T * N and info["nodes"] <= 0.3 * T * N      # (up to T / 8 unpruned generations at the end)
    a = ps.particle_assignments(int(np.argmax(ps.logweights())))           # the MAP particle's clustering of the stream
    assert a.shape == (T,) and a.min() == 1 and a.max() == ps.ncomponents()[int(np.argmax(ps.logweights()))]
    sim = ps.assignment_similarity()
    assert sim.shape == (T, T) and np.array_equal(sim, sim.T) and (np.diag(sim) == 1).all()
    print("live genealogy nodes per generation x particle:", info["nodes"] / (T * N))
