"""PF_F32: the particle store holds floats (half the store traffic); the predictive arithmetic, the putative weights
and the integer selection are those of the FP64 filter.  BASELINE north_star: log-weights and predictive log-densities
within 1e-5 relative of the reference in FP32; given identical weights and uniforms the survivors are exact.

Protocol: per observation the FP64 CPU oracle expands ITS population, the FP32 device's putative log-weights must agree
to 1e-5 relative one to one; the device's survivors must be, bit for bit, the exact selection specification evaluated
on the DEVICE's weights; the oracle then commits the device's survivors so that both populations stay aligned."""
import math

import numpy as np
import pytest

import particles_b200 as pb
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
RTOL32 = 1e-5


def _logclose(dev, ref, what, rtol=RTOL32):
    dev, ref = np.asarray(dev), np.asarray(ref)
    assert dev.shape == ref.shape, what
    pos = ref > 0
    assert np.array_equal(dev > 0, pos), what
    err = np.abs(np.log(dev[pos]) - np.log(ref[pos])) / np.maximum(np.abs(np.log(ref[pos])), 1.0)
    assert err.size == 0 or err.max() <= rtol, f"{what}: max rel err {err.max():.3e}"
    return float(err.max()) if err.size else 0.0


def _model(d):
    if d == 1:
        return pb.NormalInverseChisq(0.0, 1.0, 0.1, 1.0), orc.Prior.nix2(0.0, 1.0, 0.1, 1.0)
    return pb.NormalInverseWishart(np.zeros(d), 0.1, np.eye(d), d + 1.0), orc.Prior.niw(np.zeros(d), 0.1, np.eye(d), d + 1.0)


def _stream(rng, T, d):
    if d == 1:
        return (np.array([-6.0, -2.0, 2.0, 6.0])[rng.integers(0, 4, T)] + rng.standard_normal(T)).reshape(-1, 1)
    c = 6.0 * np.sign(rng.standard_normal((8, d)))
    return c[rng.integers(0, 8, T)] + rng.standard_normal((T, d))


@pytest.mark.parametrize("d,N,T", [(1, 400, 120), (2, 3000, 80), (4, 1000, 60), (8, 400, 40), (2, 1 << 16, 40)])
def test_f32_fearnhead_against_fp64_oracle(d, N, T):
    rng = np.random.default_rng(500 + d)
    ys, us = _stream(rng, T, d), rng.random(T)
    pprior, oprior = _model(d)
    dev = pb.FearnheadParticles(N, pprior, pb.ChineseRestaurantProcess(1.0), history=T, k_max=40, on_overflow="raise", precision="f32")
    ref = orc.OracleFilter(orc.FEARNHEAD, N, oprior, 1.0, order=orc.ORDER_INDEX)
    worst, checked = 0.0, 0
    for t in range(T):
        K_before = dev.ncomponents() if t else np.zeros(1, dtype=np.int32)
        dev.fit_(ys[t], [us[t]])
        s = dev.last_step()
        M = ref.fh_expand(ys[t])
        wo, wd = ref.putative_weights(), dev.putative_weights()
        assert len(wd) == M == s["n_putative"]
        worst = max(worst, _logclose(wd, wo, f"putative weights, step {t}"))
        assert s["log_total_w"] == pytest.approx(math.log(ref.stat("total_w")), rel=RTOL32, abs=1e-5)
        off = np.concatenate([[0], np.cumsum(K_before.astype(np.int64) + 1)])
        anc, asg = dev.last_ancestors().astype(np.int64), dev.last_assignments().astype(np.int64)
        idx = off[anc] + asg
        lw_dev = dev.logweights()
        if M > N and (t % 3 == 0 or M < 50000):
            # identical weights and uniforms => identical survivors: the exact specification on the DEVICE's weights
            s_ex, _f, _Tn, kept_ex = orc.select_exact_fast(wd, N, us[t], orc.ORDER_INDEX)
            assert s["n_kept"] == kept_ex
            assert np.array_equal(np.asarray(s_ex), idx), f"survivors differ from the exact specification at step {t}"
            checked += 1
        _logclose(np.exp(lw_dev), np.exp(lw_dev), "self")
        ref.fh_commit(idx, np.exp(lw_dev))
        assert np.array_equal(dev.ncomponents(), ref.ncomponents())
    assert checked >= 3
    print(f"d={d} N={N}: worst relative log-weight error {worst:.2e} over {T} observations")
    assert np.array_equal(dev.assignments(), ref.assignments())
    # the store really is FP32: cluster means are floats
    m = dev.particles()[0].components[0].mean
    assert np.array_equal(m, m.astype(np.float32).astype(np.float64))


@pytest.mark.parametrize("d,N,T", [(1, 512, 150), (2, 256, 80), (8, 128, 40)])
def test_f32_chenliu_against_fp64_oracle(d, N, T):
    """Chen-Liu in PF_F32: the per-particle weights track the FP64 oracle to 1e-5 while both make the same draws (checked
    up to the first draw that the weight differences flip, which the test requires to be rare)."""
    rng = np.random.default_rng(600 + d)
    ys = _stream(rng, T, d)
    pprior, oprior = _model(d)
    dev = pb.ChenLiuParticles(N, pprior, pb.ChineseRestaurantProcess(1.0), history=T, k_max=40, precision="f32")
    ref = orc.OracleFilter(orc.CHENLIU, N, oprior, 1.0)
    agree = 0
    for t in range(T):
        u = rng.random(2 * N)
        dev.fit_(ys[t], u)
        ref.cl_step(ys[t], u[:N], u[N:])
        if not (np.array_equal(dev.last_ancestors(), ref.last_ancestors()) and np.array_equal(dev.last_assignments() + 1, ref.last_assignments())):
            break
        _logclose(dev.weights(), ref.weights(), f"weights step {t}")
        assert dev.last_step()["cv"] == pytest.approx(ref.stat("cv"), rel=1e-4, abs=1e-6)
        agree += 1
    assert agree >= min(T, 30), f"FP32 and FP64 draws diverged after {agree} observations"


def test_f32_pipelined_sharded_and_checkpoint_agree(tmp_path):
    """Within PF_F32 every path computes the same filter bit for bit: one pf_filter call (pipelined steps), observation
    by observation, sharded over 4 ranks, and across a checkpoint."""
    from particles_b200.distributed import LocalGroup, ShardedFearnheadParticles
    rng = np.random.default_rng(77)
    N, T = 1 << 15, 40
    ys, us = _stream(rng, T, 2), rng.random(T)
    pprior, _ = _model(2)
    crp = pb.ChineseRestaurantProcess(1.0)
    a = pb.FearnheadParticles(N, pprior, crp, history=T, k_max=32, precision="f32")
    a.filter_(ys, uniforms=us)
    b = pb.FearnheadParticles(N, pprior, crp, history=T, k_max=32, precision="f32")
    for t in range(T):
        b.fit_(ys[t], [us[t]])
    sh = ShardedFearnheadParticles(N, pprior, crp, LocalGroup(4), history=T, k_max=32, precision="f32")
    sh.filter_(ys, us)
    c = pb.FearnheadParticles(N, pprior, crp, history=T, k_max=32, precision="f32")
    c.filter_(ys[:17], uniforms=us[:17])
    path = str(tmp_path / "f32.ckpt")
    c.save(path)
    c2 = pb.load(path)
    assert c2.precision == "f32"
    c2.filter_(ys[17:], uniforms=us[17:])
    for other, name in ((b, "stepwise"), (sh, "sharded"), (c2, "checkpoint")):
        assert np.array_equal(a.logweights(), other.logweights()), name
        assert np.array_equal(a.ncomponents(), other.ncomponents()), name
        assert np.array_equal(a.assignments(), other.assignments()), name
    # and it is a different (coarser) filter than FP64, by less than the tolerance
    f64 = pb.FearnheadParticles(N, pprior, crp, history=T, k_max=32)
    f64.filter_(ys, uniforms=us)
    assert np.abs(a.log_evidence - f64.log_evidence).max() < 1e-4
    assert not np.array_equal(a.log_evidence, f64.log_evidence)
