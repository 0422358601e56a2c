"""Sharded run == unsharded run, bit for bit, with all ranks emulated inside one
process on one GPU (LocalComm): same kernels, same stage boundaries, the collectives replaced
by local concatenation."""
import numpy as np
import pytest

import particles_b200 as pb
from particles_b200.distributed import LocalComm, ShardedFearnheadParticles, partition_plan

pytestmark = pytest.mark.gpu


def _mix2d(rng, T, k=8, radius=6.0):
    ang = 2 * np.pi * rng.integers(0, k, T) / k
    return radius * np.stack([np.cos(ang), np.sin(ang)], axis=1) + rng.standard_normal((T, 2))


@pytest.mark.parametrize("G,N,T,direct", [(2, 1000, 40, True), (4, 1000, 40, True), (2, 1 << 16, 30, True), (8, 3000, 25, True),
                                          (4, 1000, 40, False), (2, 1 << 16, 30, False)])
def test_sharded_equals_unsharded(G, N, T, direct):
    rng = np.random.default_rng(5)
    ys = _mix2d(rng, T)
    us = rng.random(T)
    prior = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0)
    crp = pb.ChineseRestaurantProcess(1.0)
    ref = pb.FearnheadParticles(N, prior, crp, history=T)
    sh = ShardedFearnheadParticles(N, prior, crp, LocalComm(G), history=T, direct=direct)
    for t in range(T):
        ref.fit_(ys[t], [us[t]])
        sh.fit_(ys[t], us[t])
        assert len(sh) == len(ref), f"population differs at step {t}"
        assert np.array_equal(sh.ncomponents(), ref.ncomponents()), f"K differs at step {t}"
        assert np.array_equal(sh.last_ancestors(), ref.last_ancestors()), f"ancestors differ at step {t}"
        assert np.array_equal(sh.last_assignments(), ref.last_assignments()), f"assignments differ at step {t}"
        assert np.array_equal(sh.logweights(), ref.logweights()), f"log-weights differ at step {t}"
        a, b = sh.last_step(), ref.last_step()
        for key in ("n_out", "log_total_w", "log_w_resamp", "threshold", "resampled"):
            assert a[key] == b[key] or (np.isnan(a[key]) and np.isnan(b[key])), (key, t)
    # shards stay balanced (children stay with their parents only while shares are within q/8 + 64)
    sizes = [int(sh.L.pf_n_particles(s.h)) for s in sh.shards]
    q = -(-len(sh) // G)
    assert sum(sizes) == len(sh) and max(sizes) <= q + q // 2 + 64 + G
    if direct:       # equal blocks after every step, nothing through the send / recv buffers
        assert sizes == [min(max(len(sh) - r * q, 0), q) for r in range(G)] and sh.exchanges == 0
    assert sh.exchanges < T


def test_device_plan_matches_host_plan():
    rng = np.random.default_rng(6)
    G, N, T = 4, 5000, 20
    ys = _mix2d(rng, T)
    prior = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0)
    sh = ShardedFearnheadParticles(N, prior, pb.ChineseRestaurantProcess(1.0), LocalComm(G), history=T, direct=False)
    import ctypes as C
    for t in range(T):
        sh.fit_(ys[t], rng.random())
    # recompute the last plan on the host from the device-reported local counts
    from particles_b200._lib import pf_mg_ptrs
    info = pf_mg_ptrs()
    sh.L.pf_mg_info(sh.shards[0].h, C.byref(info))
    res = [s.plan(G) for s in sh.shards]
    n_local = [r[2] for r in res]
    for r in range(G):
        send, recv, q, mine, mode = partition_plan(n_local, r, stay_cap=info.local_capacity, window=info.window)
        assert res[r][0].tolist() == send.tolist() and res[r][1].tolist() == recv.tolist() and mode == res[r][5]
        assert mine == int(sh.L.pf_n_particles(sh.shards[r].h))


@pytest.mark.parametrize("kw", [{"cand_entries": 64}, {"sample_entries": 64}, {"sample_entries": 64, "cand_entries": 64}])
def test_slow_path_truncated_gather_and_rebracketing(kw):
    """A candidate list that outgrows its fixed gather size (phase 7) and a sample too small to
    bracket the threshold (verification failure, phase 3) both fall back to exact-size gathers
    and still reproduce the unsharded run."""
    rng = np.random.default_rng(11)
    G, N, T = 4, 20000, 30
    ys = _mix2d(rng, T)
    us = rng.random(T)
    prior = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0)
    crp = pb.ChineseRestaurantProcess(1.0)
    ref = pb.FearnheadParticles(N, prior, crp, history=T)
    ref.filter_(ys, uniforms=us)
    sh = ShardedFearnheadParticles(N, prior, crp, LocalComm(G), history=T, **kw)
    sh.filter_(ys, us)
    assert sh.slow_steps > 0
    # every decision is identical; the total weight is (exact mass above the bracket) + (fixed-point
    # mass below it), so with a deliberately different bracket the log-weights agree to rounding
    assert np.array_equal(sh.ncomponents(), ref.ncomponents())
    assert np.array_equal(sh.last_ancestors(), ref.last_ancestors())
    assert np.array_equal(sh.last_assignments(), ref.last_assignments())
    np.testing.assert_allclose(sh.logweights(), ref.logweights(), rtol=0, atol=1e-12)


@pytest.mark.parametrize("lookahead,window", [(1, 0), (4, 0), (16, 0), (4, 50), (8, 400)])
def test_lookahead_batches_equal_stepwise(lookahead, window, monkeypatch):
    """Queuing several observations ahead of the host (device-side halt mark for the steps that
    need it) gives the same population as stepping one at a time."""
    monkeypatch.setenv("PF_MG_WINDOW", str(window))          # > 0: fixed-capacity neighbour exchange every step
    rng = np.random.default_rng(21)
    G, N, T = 4, 8000, 48
    ys = _mix2d(rng, T)
    us = rng.random(T)
    prior = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0)
    crp = pb.ChineseRestaurantProcess(1.0)
    ref = pb.FearnheadParticles(N, prior, crp, history=T)
    ref.filter_(ys, uniforms=us)
    sh = ShardedFearnheadParticles(N, prior, crp, LocalComm(G), history=T)
    sh.filter_(ys, us, lookahead=lookahead)
    assert len(sh) == len(ref)
    assert np.array_equal(sh.logweights(), ref.logweights())
    assert np.array_equal(sh.ncomponents(), ref.ncomponents())
    assert np.array_equal(sh.last_ancestors(), ref.last_ancestors())
    assert sh.host_syncs < 2 * T if lookahead > 1 else True
