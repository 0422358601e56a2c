"""Parity of the CUDA path (through the C ABI) against the CPU oracle.

Tolerances (BASELINE.json north_star): log-weights / predictive log-densities within 1e-10
relative in FP64; ancestors, assignments and selected indices bit-exact."""
import math

import numpy as np
import pytest

import particles_b200 as pb
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _mixture_1d(rng, T, centers=(-6.0, -2.0, 2.0, 6.0)):
    z = rng.integers(0, len(centers), T)
    return np.asarray(centers)[z] + rng.standard_normal(T)


def _mixture_nd(rng, T, d, k, radius=6.0):
    if d == 2:
        ang = 2 * np.pi * np.arange(k) / k
        centers = radius * np.stack([np.cos(ang), np.sin(ang)], axis=1)
    else:
        centers = radius * np.sign(rng.standard_normal((k, d)))
    z = rng.integers(0, k, T)
    return centers[z] + rng.standard_normal((T, d))


def _priors(kind, d):
    if kind == "nix2":
        return pb.NormalInverseChisq(0.0, 1.0, 0.1, 1.0), orc.Prior.nix2(0.0, 1.0, 0.1, 1.0)
    rng = np.random.default_rng(d)
    A = rng.standard_normal((d, d)) * 0.3
    Lam = np.eye(d) + A @ A.T
    mu = np.zeros(d)
    return pb.NormalInverseWishart(mu, 0.1, Lam, d + 1.0), orc.Prior.niw(mu, 0.1, Lam, d + 1.0)


def _assert_logclose(dev, ref, what):
    dev, ref = np.asarray(dev), np.asarray(ref)
    assert dev.shape == ref.shape, what
    pos = ref > 0
    assert np.all((dev > 0) == pos), what
    ld, lr = np.log(dev[pos]), np.log(ref[pos])
    err = np.abs(ld - lr) / np.maximum(np.abs(lr), 1.0)
    assert err.max() <= RTOL, f"{what}: max rel err {err.max():.3e}"


# ------------------------------------------------------------------ expansion
@pytest.mark.parametrize("kind,d", [("nix2", 1), ("niw", 1), ("niw", 2), ("niw", 3), ("niw", 4), ("niw", 8)])
def test_expansion_matches_oracle_through_enumeration(kind, d):
    """Every putative weight exp(logweight) (src/particle.jl:100-116) of the exhaustive phase
    (M <= N, src/fearnhead.jl:135-138): populations stay aligned, so weights compare 1:1."""
    rng = np.random.default_rng(10 + d)
    pprior, oprior = _priors(kind, d)
    N = 6000
    ys = _mixture_1d(rng, 8).reshape(-1, 1) if d == 1 else _mixture_nd(rng, 8, d, 4)
    dev = pb.FearnheadParticles(N, pprior, pb.ChineseRestaurantProcess(1.0), history=8)
    ref = orc.OracleFilter(orc.FEARNHEAD, N, oprior, 1.0, order=orc.ORDER_INDEX)
    bell = [1, 2, 5, 15, 52, 203, 877, 4140]
    for t, y in enumerate(ys):
        dev.fit_(y, [0.5])
        ref.fh_step(y, 0.5)
        _assert_logclose(dev.putative_weights(), ref.putative_weights(), f"putatives step {t}")
        assert len(dev) == ref.n == bell[t]
        _assert_logclose(dev.weights(), ref.weights(), f"weights step {t}")
        assert np.array_equal(dev.ncomponents(), ref.ncomponents())
        assert np.array_equal(dev.last_ancestors(), ref.last_ancestors())
        assert np.array_equal(dev.last_assignments() + 1, ref.last_assignments())
        s = dev.last_step()
        assert s["log_total_w"] == pytest.approx(math.log(ref.stat("total_w")), rel=RTOL, abs=1e-10)
    assert np.array_equal(dev.assignments(), ref.assignments())
    # component statistics of a few particles against the oracle's sufficient statistics
    for i in (0, 17, len(dev) - 1):
        comps = dev.particles()[i].components
        st, crp = ref.particle(i)
        assert [c.n for c in comps] == [int(x) for x in crp]
        for c, row in zip(comps, st):
            np.testing.assert_array_equal(c.mean, row[1 + d:1 + 2 * d])        # Welford mean is bit-exact
            mu, U, kn, nn = (oprior.posterior(row) if oprior.kind == orc.NIW else (None, None, None, None))
            if U is not None:
                np.testing.assert_allclose(c.chol @ c.chol.T, U.T @ U, rtol=1e-11, atol=1e-12)
            else:
                post = oprior.posterior(row)
                assert c.chol[0, 0] ** 2 == pytest.approx(post[1] * post[3], rel=1e-11)


# ------------------------------------------------------------------ selection on injected weights
def _check_select(w, N, u, order):
    s_dev, f_dev, st = pb.select(w, N, u, order)
    o = orc.ORDER_INDEX if order == pb.PF_ORDER_INDEX else orc.ORDER_SORTED
    if len(w) <= N:
        assert np.array_equal(s_dev, np.arange(len(w))) and not f_dev.any()
        return st
    s_ex, f_ex, Tn, n_kept = orc.select_exact(w, N, u, o)
    assert st["n_kept"] == n_kept
    assert np.array_equal(s_dev, s_ex), f"survivors differ (M={len(w)}, N={N})"
    assert np.array_equal(f_dev, f_ex)
    if Tn[1] > 0 and Tn[0] > 0:
        assert st["log_w_resamp"] == pytest.approx(math.log(orc.exact_c_value(Tn)), rel=1e-12, abs=1e-12)
    assert st["log_total_w"] == pytest.approx(math.log(math.fsum(w)), rel=1e-12, abs=1e-12)
    return st


@pytest.mark.parametrize("order", [pb.PF_ORDER_INDEX, pb.PF_ORDER_SORTED])
def test_select_reference_edge_cases(order):
    # test/fearnhead.jl:52-71 and the worked examples of SURVEY 8(a11/a12)
    _check_select(np.ones(20), 10, 0.3, order)                                   # resample all, c = 2
    st = _check_select(np.array([1.0] * 10 + [0.0] * 10), 10, 0.3, order)        # keep the 10 ones
    assert st["n_kept"] == 10 and st["n_resamp_got"] == 0
    w = np.array([.05, .05, .1, .1, .2, .5])
    for N in (2, 3, 4, 5):
        for u in (0.0, 0.25, 0.5, 0.999):
            _check_select(w, N, u, order)
    s, f, st = pb.select(w, 4, 0.5, pb.PF_ORDER_SORTED)
    assert s.tolist() == [1, 3, 4, 5] and f.tolist() == [True, True, False, False]
    _check_select(np.array([3.0, 1.0, 2.0]), 5, 0.1, order)                      # M <= N: keep all
    _check_select(np.array([1e-300, 1e-310, 5e-324, 1.0, 2.0, 3.0]), 4, 0.7, order)   # subnormals


@pytest.mark.parametrize("order", [pb.PF_ORDER_INDEX, pb.PF_ORDER_SORTED])
def test_select_matches_exact_oracle_random(order):
    rng = np.random.default_rng(2024)
    for trial in range(40):
        M = int(rng.integers(20, 3000))
        N = int(rng.integers(2, M))
        w = np.exp(rng.standard_normal(M) * rng.choice([0.5, 3.0, 12.0]))
        if trial % 4 == 0:
            w[rng.integers(0, M, M // 3)] = 0.0
        if trial % 5 == 0:
            w = np.round(w, 1)                                                   # many exact ties
        if trial % 7 == 0:
            w[rng.integers(0, M)] = 1e12                                         # one dominant item
        _check_select(w, N, rng.random(), order)


@pytest.mark.parametrize("order", [pb.PF_ORDER_INDEX, pb.PF_ORDER_SORTED])
def test_select_large_goes_through_bracket_and_descent(order):
    # M above the direct-solve size: sample -> bracket -> classify -> radix descent -> solve
    rng = np.random.default_rng(5)
    for M, N, spread in ((60_000, 7_000, 4.0), (150_000, 16_384, 9.0), (100_000, 99_000, 2.0), (50_000, 3, 5.0)):
        w = np.exp(rng.standard_normal(M) * spread)
        st = _check_select(w, N, rng.random(), order)
        assert st["select_rounds"] >= 1
    # wide dynamic range with a dust tail and heavy ties around the threshold
    w = np.concatenate([np.exp(rng.standard_normal(40_000) * 20 - 60), np.full(30_000, 0.125), rng.random(30_000)])
    rng.shuffle(w)
    _check_select(w, 45_000, rng.random(), order)
    _check_select(np.full(70_000, 3.0), 1000, rng.random(), order)              # everything tied


def test_select_agrees_with_literal_reference_order():
    # the literal floating-point restatement picks the same items at these sizes
    rng = np.random.default_rng(77)
    for _ in range(10):
        M = int(rng.integers(200, 5000))
        N = int(rng.integers(10, M // 2))
        w = np.exp(rng.standard_normal(M) * 3)
        u = rng.random()
        s_dev, f_dev, st = pb.select(w, N, u, pb.PF_ORDER_SORTED)
        s_lit, f_lit, c_lit, n_kept = orc.select_literal(w, N, u, orc.ORDER_SORTED)
        assert np.array_equal(s_dev, s_lit) and np.array_equal(f_dev, f_lit)
        assert st["log_w_resamp"] == pytest.approx(math.log(c_lit), rel=1e-12)


# ------------------------------------------------------------------ end-to-end Fearnhead
def _run_fearnhead_pair(pprior, oprior, ys, N, order, k_max=16):
    dorder = pb.PF_ORDER_SORTED if order == orc.ORDER_SORTED else pb.PF_ORDER_INDEX
    dev = pb.FearnheadParticles(N, pprior, pb.ChineseRestaurantProcess(1.0), history=len(ys), order=dorder, k_max=k_max)
    ref = orc.OracleFilter(orc.FEARNHEAD, N, oprior, 1.0, order=order)
    rng = np.random.default_rng(99)
    for t, y in enumerate(ys):
        u = rng.random()
        dev.fit_(y, [u])
        ref.fh_step(y, u)
        assert len(dev) == ref.n, f"population size differs at step {t}"
        assert np.array_equal(dev.last_ancestors(), ref.last_ancestors()), f"ancestors differ at step {t}"
        assert np.array_equal(dev.last_assignments() + 1, ref.last_assignments()), f"assignments differ at step {t}"
        s = dev.last_step()
        assert s["n_kept"] == ref.stat("n_kept") and s["n_resamp_got"] == ref.stat("n_resamp_got")
    _assert_logclose(dev.weights(), ref.weights(), "final weights")
    assert np.array_equal(dev.ncomponents(), ref.ncomponents())
    A = dev.assignments()
    assert np.array_equal(A, ref.assignments())
    assert np.all(A.max(axis=0) == dev.ncomponents()) and np.all(A.min(axis=0) == 1)   # test/clustering.jl:9-13
    return dev, ref


@pytest.mark.parametrize("order", [orc.ORDER_SORTED, orc.ORDER_INDEX])
def test_fearnhead_1d_nix2_config1(order):
    """BASELINE config 1 (bench/1d.jl:20-24): N = 100, 4-cluster mixture; every step's
    ancestors and assignments are identical to the literal restatement."""
    rng = np.random.default_rng(100)
    ys = _mixture_1d(rng, 400)
    pprior, oprior = _priors("nix2", 1)
    _run_fearnhead_pair(pprior, oprior, ys, 100, order)


@pytest.mark.parametrize("order", [orc.ORDER_SORTED, orc.ORDER_INDEX])
def test_fearnhead_2d_niw(order):
    rng = np.random.default_rng(100)
    ys = _mixture_nd(rng, 150, 2, 8)
    pprior = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0)          # bench/2d.jl:18
    oprior = orc.Prior.niw(np.zeros(2), 0.1, np.eye(2), 3.0)
    _run_fearnhead_pair(pprior, oprior, ys, 1000, order)


def test_fearnhead_batch_equals_stepwise_and_kmax_overflow():
    rng = np.random.default_rng(3)
    ys = _mixture_1d(rng, 60)
    us = rng.random(60)
    pprior, _ = _priors("nix2", 1)
    a = pb.FearnheadParticles(500, pprior, pb.ChineseRestaurantProcess(1.0), history=60)
    a.filter_(ys, uniforms=us)
    b = pb.FearnheadParticles(500, pprior, pb.ChineseRestaurantProcess(1.0), history=60)
    for y, u in zip(ys, us):
        b.fit_(y, [u])
    assert np.array_equal(a.logweights(), b.logweights())
    assert np.array_equal(a.assignments(), b.assignments())
    # k_max = 3: the new-cluster candidate is dropped at K = 3 and counted
    c = pb.FearnheadParticles(500, pprior, pb.ChineseRestaurantProcess(5.0), history=60, k_max=3)
    c.filter_(ys, uniforms=us)
    assert c.ncomponents().max() <= 3
    assert c.last_step()["overflow"] > 0
    assert np.exp(c.logweights()).sum() == pytest.approx(1.0, rel=1e-9)


# ------------------------------------------------------------------ end-to-end Chen-Liu
@pytest.mark.parametrize("kind,d,N,T", [("nix2", 1, 300, 250), ("niw", 2, 256, 120), ("niw", 8, 128, 60)])
def test_chenliu_matches_oracle(kind, d, N, T):
    rng = np.random.default_rng(7 + d)
    ys = _mixture_1d(rng, T).reshape(-1, 1) if d == 1 else _mixture_nd(rng, T, d, 4)
    pprior, oprior = _priors(kind, d)
    dev = pb.ChenLiuParticles(N, pprior, pb.ChineseRestaurantProcess(1.0), history=T, rejuv=50.0, k_max=24)
    ref = orc.OracleFilter(orc.CHENLIU, N, oprior, 1.0, rejuv=50.0)
    n_rejuv = 0
    for t, y in enumerate(ys):
        u = rng.random(2 * N)
        dev.fit_(y, u)
        ref.cl_step(y, u[:N], u[N:])
        s = dev.last_step()
        assert s["resampled"] == ref.stat("resampled"), f"rejuvenation decision differs at step {t}"
        assert s["cv"] == pytest.approx(ref.stat("cv"), rel=1e-8, abs=1e-10)
        n_rejuv += s["resampled"]
        assert np.array_equal(dev.last_ancestors(), ref.last_ancestors()), f"ancestors differ at step {t}"
        assert np.array_equal(dev.last_assignments() + 1, ref.last_assignments()), f"assignments differ at step {t}"
        _assert_logclose(dev.weights(), ref.weights(), f"weights step {t}")
    if d <= 2:
        assert n_rejuv > 0          # well-separated 8-D clusters keep all particles identical
    assert np.array_equal(dev.ncomponents(), ref.ncomponents())
    assert np.array_equal(dev.assignments(), ref.assignments())


def test_long_stream_rank1_drift_stays_inside_tolerance():
    """2 000 observations through one Chen-Liu population: the rank-1-updated Cholesky
    factors (n up to ~1000 per cluster) still reproduce the oracle's refactorised weights."""
    rng = np.random.default_rng(21)
    T, N = 2000, 32
    ys = _mixture_nd(rng, T, 2, 3, radius=5.0)
    pprior = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0)
    oprior = orc.Prior.niw(np.zeros(2), 0.1, np.eye(2), 3.0)
    dev = pb.ChenLiuParticles(N, pprior, pb.ChineseRestaurantProcess(1.0), history=T, k_max=32)
    ref = orc.OracleFilter(orc.CHENLIU, N, oprior, 1.0)
    U = rng.random((T, 2 * N))
    dev.filter_(ys, uniforms=U)
    for t in range(T):
        ref.cl_step(ys[t], U[t, :N], U[t, N:])
    assert np.array_equal(dev.assignments(), ref.assignments())
    _assert_logclose(dev.weights(), ref.weights(), "weights after 2000 obs")


# ------------------------------------------------------------------ full-size properties
def test_properties_at_2_pow_18():
    """Size-independent properties at a size the oracle cannot reach in seconds."""
    rng = np.random.default_rng(100)
    N, T = 1 << 18, 40
    ys = _mixture_nd(rng, T, 2, 8)
    prior = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0)
    ps = pb.FearnheadParticles(N, prior, pb.ChineseRestaurantProcess(1.0), history=T)
    ps.filter_(ys)
    n = len(ps)
    assert 0 < n <= N
    s = ps.last_step()
    assert s["resampled"] == 1 and s["n_kept"] + s["n_resamp_got"] == n and s["n_resamp_got"] <= s["n_resamp_req"]
    w = ps.weights()
    assert w.sum() == pytest.approx(1.0, rel=1e-6)
    lw = ps.logweights()
    # resampled survivors all carry log(c/total); kept ones are at least the threshold
    assert np.isclose(lw, s["log_w_resamp"], rtol=0, atol=1e-12).sum() == s["n_resamp_got"]
    A = ps.assignments()
    assert A.shape == (T, n)
    assert np.all(A.max(axis=0) == ps.ncomponents()) and np.all(A.min(axis=0) == 1)
    assert np.all(A[0] == 1)
    # labels appear in order: the first occurrence of label k+1 comes after label k
    first = np.array([[np.argmax(A[:, i] == k) if (A[:, i] == k).any() else T for k in range(1, 6)] for i in range(0, n, n // 50)])
    assert np.all(np.diff(first, axis=1) >= 0)
    # determinism: same inputs, same uniforms (device stream) -> identical population
    ps2 = pb.FearnheadParticles(N, prior, pb.ChineseRestaurantProcess(1.0), history=T)
    ps2.filter_(ys)
    assert np.array_equal(ps2.logweights(), lw)
    assert np.array_equal(ps2.assignments(), A)


@pytest.mark.parametrize("margin", [1])
def test_failed_bracket_is_widened_not_opened(margin, monkeypatch):
    """A sample bracket that misses the threshold (forced here by shrinking its margin from 8 sigma
    to 1-2) is moved to the failing side and widened geometrically; decisions are those of the
    default run, the total weight agrees to rounding (a different bracket splits it differently)."""
    rng = np.random.default_rng(77)
    N, T = 1 << 17, 30
    ys = _mixture_nd(rng, T, 2, 8)
    us = rng.random(T)
    prior = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0)
    ref = pb.FearnheadParticles(N, prior, pb.ChineseRestaurantProcess(1.0), history=T, k_max=16)
    ref.filter_(ys, uniforms=us)
    monkeypatch.setenv("PF_SEL_MARGIN", str(margin))
    ps = pb.FearnheadParticles(N, prior, pb.ChineseRestaurantProcess(1.0), history=T, k_max=16)
    rounds = []
    for t in range(T):
        ps.filter_(ys[t:t + 1], uniforms=us[t:t + 1])
        rounds.append(ps.last_step()["select_rounds"])
    assert max(rounds) >= 2                      # the verification did fail somewhere
    assert np.array_equal(ps.ncomponents(), ref.ncomponents())
    assert np.array_equal(ps.assignments(), ref.assignments())
    np.testing.assert_allclose(ps.logweights(), ref.logweights(), rtol=0, atol=1e-12)
    # and the stand-alone selection against the exact oracle
    w = np.exp(rng.standard_normal(300_000) * 3.0)
    for order in (pb.PF_ORDER_INDEX, pb.PF_ORDER_SORTED):
        _check_select(w, 40_000, rng.random(), order)


# ------------------------------------------------------------------ pipelined steps of pf_filter
@pytest.mark.parametrize("d,N,T", [(1, 300, 80), (2, 5000, 60), (2, 1 << 17, 40), (4, 2000, 40), (8, 600, 30)])
def test_pipelined_filter_equals_stepwise(d, N, T, monkeypatch):
    """Inside one pf_filter call the tail of every observation (instantiate) is fused with the expansion of the next
    one; observation-by-observation calls run the separate kernels.  Same arithmetic, same decisions: weights,
    genealogy, evidence and store are bitwise equal -- also with the pipelining switched off, and when the stream is
    cut into several calls."""
    rng = np.random.default_rng(300 + d)
    ys = _mixture_1d(rng, T).reshape(-1, 1) if d == 1 else _mixture_nd(rng, T, d, 4)
    us = rng.random(T)
    pprior, _ = _priors("nix2" if d == 1 else "niw", d)
    crp = pb.ChineseRestaurantProcess(1.0)

    def run(mode):
        ps = pb.FearnheadParticles(N, pprior, crp, history=T, k_max=40, on_overflow="raise")
        if mode == "one_call":
            ps.filter_(ys, uniforms=us)
            le = ps.log_evidence
        elif mode == "three_calls":
            cuts, le = [0, T // 3, T // 3 + 1, T], []
            for a, b in zip(cuts[:-1], cuts[1:]):
                ps.filter_(ys[a:b], uniforms=us[a:b])
                le.append(ps.log_evidence)
            le = np.concatenate(le)
        else:
            le = []
            for t in range(T):
                ps.fit_(ys[t], [us[t]])
                le.append(ps.last_step()["log_total_w"])
            le = np.array(le)
        return ps, le

    ref, le_ref = run("stepwise")
    for mode in ("one_call", "three_calls"):
        ps, le = run(mode)
        assert len(ps) == len(ref), mode
        assert np.array_equal(le, le_ref), mode
        assert np.array_equal(ps.logweights(), ref.logweights()), mode
        assert np.array_equal(ps.ncomponents(), ref.ncomponents()), mode
        assert np.array_equal(ps.assignments(), ref.assignments()), mode
        assert np.array_equal(ps.putative_weights(), ref.putative_weights()), mode
        for i in (0, len(ps) // 2, len(ps) - 1):
            for c0, c1 in zip(ps.particles()[i].components, ref.particles()[i].components):
                assert c0.n == c1.n and np.array_equal(c0.mean, c1.mean) and np.array_equal(c0.chol, c1.chol) and c0.logdet == c1.logdet
        assert ps.last_step()["n_putative"] == ref.last_step()["n_putative"]
    monkeypatch.setenv("PF_NO_PIPELINE", "1")
    ps, le = run("one_call")
    assert np.array_equal(le, le_ref) and np.array_equal(ps.assignments(), ref.assignments())


def test_history_exhausted_mid_call_leaves_a_consistent_filter():
    """A fixed genealogy capacity that runs out in the middle of a pf_filter call: the error is PF_ERR_HISTORY, the
    observations before it have been filtered, and every read-out describes that state (stale host mirrors)."""
    rng = np.random.default_rng(8)
    T, cap_hist = 30, 17
    ys, us = _mixture_nd(rng, T, 2, 4), rng.random(T)
    pprior, _ = _priors("niw", 2)
    crp = pb.ChineseRestaurantProcess(1.0)
    ps = pb.FearnheadParticles(2000, pprior, crp, history=cap_hist, k_max=40)
    with pytest.raises(pb.PFError) as e:
        ps.filter_(ys, uniforms=us)
    assert e.value.code == 4
    ref = pb.FearnheadParticles(2000, pprior, crp, history=cap_hist, k_max=40)
    ref.filter_(ys[:cap_hist], uniforms=us[:cap_hist])
    assert ps.n_steps == cap_hist and len(ps) == len(ref)
    assert np.array_equal(ps.logweights(), ref.logweights())
    assert np.array_equal(ps.assignments(), ref.assignments())
    assert ps.last_step()["n_out"] == ref.last_step()["n_out"]
