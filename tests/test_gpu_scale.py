"""Oracle parity AT SCALE: the code paths that only exist for large populations (strided-sample
bracket with stride > 1, the 8-sigma margin, plateau counting with many exact ties, candidate
staging, tile-record reuse) against the CPU oracle at N = 2^17 and the exact selection
specification at M > 4 * 2^20.

Protocol of the Fearnhead comparison (per observation):
  1. the oracle expands ITS population (orc_fh_expand: src/particle.jl:100-116, two fresh
     marginal_log_lhood per putative) -> putative weights; the device's putative weights must agree
     to 1e-10 relative in log space (BASELINE north_star), one to one;
  2. the device's survivors (ancestor, assignment) are mapped to putative indices and must equal,
     bit for bit, (a) the literal FP64 restatement of src/fearnhead.jl:146-179 on the oracle's
     weights and, wherever (a) differs or on every 4th step, (b) the exact-arithmetic specification
     (oracle.select_exact_fast) evaluated on the DEVICE's putative weights -- "identical weights and
     uniforms => identical survivors".  Literal-vs-exact disagreements ("flips": the literal FP64
     recurrences round, the exact evaluation does not) are counted and printed;
  3. the oracle commits the device's survivors (orc_fh_commit = instantiate.(...)), so both
     populations stay aligned for the next observation; the committed weights are the oracle's own
     (w / total, c / total), which the device's new log-weights must match to 1e-10.
"""
import math

import numpy as np
import pytest

import particles_b200 as pb
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-10        # FP64 tolerance of BASELINE.json's north_star (relative, on log-weights)


def _mix2d(rng, T, k=8, radius=6.0):
    ang = 2 * np.pi * rng.integers(0, k, T) / k
    return radius * np.stack([np.cos(ang), np.sin(ang)], axis=1) + rng.standard_normal((T, 2))


def _logclose(dev, ref, what, rtol=RTOL):
    dev, ref = np.asarray(dev), np.asarray(ref)
    assert dev.shape == ref.shape, what
    pos = ref > 0
    assert np.array_equal(dev > 0, pos), what
    ld, lr = np.log(dev[pos]), np.log(ref[pos])
    err = np.abs(ld - lr) / np.maximum(np.abs(lr), 1.0)
    assert err.max() <= rtol, f"{what}: max rel err {err.max():.3e}"


def _fearnhead_vs_oracle(N, T, alpha, seed, k_max=40, every=4):
    rng = np.random.default_rng(seed)
    ys = _mix2d(rng, T)
    us = rng.random(T)
    dev = pb.FearnheadParticles(N, pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0),
                                pb.ChineseRestaurantProcess(alpha), history=T, k_max=k_max, on_overflow="raise")
    ref = orc.OracleFilter(orc.FEARNHEAD, N, orc.Prior.niw(np.zeros(2), 0.1, np.eye(2), 3.0), alpha, order=orc.ORDER_INDEX)
    flips_total, resampling_steps, exact_checks, plateau_steps, stride_gt1 = 0, 0, 0, 0, 0
    for t in range(T):
        K_before = dev.ncomponents() if t else np.zeros(1, dtype=np.int32)
        dev.fit_(ys[t], [us[t]])
        s = dev.last_step()
        assert s["overflow"] == 0
        M = ref.fh_expand(ys[t])
        wo = ref.putative_weights()
        wd = dev.putative_weights()
        assert len(wd) == M == s["n_putative"]
        _logclose(wd, wo, f"putative weights, step {t}")
        total_o = ref.stat("total_w")
        assert s["log_total_w"] == pytest.approx(math.log(total_o), rel=RTOL, abs=1e-10)
        off = np.concatenate([[0], np.cumsum(K_before.astype(np.int64) + 1)])
        anc, asg = dev.last_ancestors().astype(np.int64), dev.last_assignments().astype(np.int64)
        idx = off[anc] + asg
        assert np.all(asg <= K_before[anc]) and np.all(np.diff(idx) > 0)
        lw_dev = dev.logweights()
        if M <= N:
            assert np.array_equal(idx, np.arange(M))
            w_new = wo / total_o
        else:
            resampling_steps += 1
            plateau_steps += s["n_plateau"] > 0
            s_lit, f_lit, c_lit, kept_lit = orc.select_literal(wo, N, us[t], orc.ORDER_INDEX)
            same = len(s_lit) == len(idx) and np.array_equal(s_lit, idx)
            if not same or resampling_steps % every == 0 or t == T - 1:
                s_ex, f_ex, Tn, kept_ex = orc.select_exact_fast(wd, N, us[t], orc.ORDER_INDEX)
                exact_checks += 1
                assert s["n_kept"] == kept_ex, f"kept count differs from the exact specification at step {t}"
                assert np.array_equal(s_ex, idx), f"survivors differ from the exact specification at step {t}"
                c = orc.exact_c_value(Tn)
                assert s["log_w_resamp"] == pytest.approx(math.log(c) - math.log(total_o), rel=RTOL, abs=1e-10)
                flag = f_ex
            else:
                c, flag = c_lit, f_lit
                assert s["n_kept"] == kept_lit
            if not same:
                nf = len(set(s_lit.tolist()) ^ set(idx.tolist()))
                flips_total += nf
                print(f"step {t}: literal FP64 restatement and exact specification differ in {nf} survivors (M = {M})")
            assert s["n_kept"] + s["n_resamp_got"] == len(idx)
            w_new = np.where(flag, c / total_o, wo[idx] / total_o)
        _logclose(np.exp(lw_dev), w_new, f"new weights, step {t}")
        ref.fh_commit(idx, w_new)
        assert np.array_equal(dev.ncomponents(), ref.ncomponents())
    print(f"N={N}: {resampling_steps} resampling steps, {exact_checks} exact-spec checks, {flips_total} literal-vs-exact flips, "
          f"{plateau_steps} steps with the plateau inside the bracket")
    A = dev.assignments()
    assert np.array_equal(A, ref.assignments())
    return resampling_steps, flips_total


@pytest.mark.parametrize("log2n,T,alpha", [(17, 36, 1.0), (17, 30, 4.0)])
def test_fearnhead_2d_niw_at_2_pow_17_against_oracle(log2n, T, alpha):
    """BASELINE configs 3/4 model at 2^17 particles (fused path: sample expansion with stride > 1 because
    N * (k_max + 1) > 2^21, bracket, fused classification, plateau counting, tile-record reuse)."""
    n_res, flips = _fearnhead_vs_oracle(1 << log2n, T, alpha, seed=100 + log2n)
    assert n_res >= 20
    assert flips <= 10 * n_res


def test_fearnhead_2d_niw_at_2_pow_18_against_oracle():
    n_res, flips = _fearnhead_vs_oracle(1 << 18, 26, 1.0, seed=218, every=5)
    assert n_res >= 12


def _check_select_exact(w, N, u):
    s_dev, f_dev, st = pb.select(w, N, u, pb.PF_ORDER_INDEX)
    s_ex, f_ex, Tn, n_kept = orc.select_exact_fast(w, N, u, orc.ORDER_INDEX)
    assert st["n_kept"] == n_kept
    assert np.array_equal(s_dev, s_ex) and np.array_equal(f_dev, f_ex)
    if Tn[1] > 0 and Tn[0] > 0:
        assert st["log_w_resamp"] == pytest.approx(math.log(orc.exact_c_value(Tn)), rel=1e-12, abs=1e-12)
    return st


def test_select_stride_gt_1_and_mass_ties_against_exact_spec():
    """pf_select with M >= 4 * 2^20 (sample stride 5) and >= 10^5 exactly tied weights at the threshold."""
    rng = np.random.default_rng(31)
    M = 4 * (1 << 20) + 12345
    w = np.exp(rng.standard_normal(M) * 3.0)
    st = _check_select_exact(w, 1 << 19, rng.random())
    assert st["select_rounds"] >= 1
    # 150 000 copies of one value sitting where the threshold falls (the plateau of the filter is this case)
    w2 = np.exp(rng.standard_normal(M) * 2.0)
    thr = np.sort(w2)[-(1 << 19)]
    w2[rng.choice(M, 150_000, replace=False)] = thr
    _check_select_exact(w2, 1 << 19, rng.random())
    # ... and just below it
    w2[w2 == thr] = np.nextafter(thr, 0.0)
    _check_select_exact(w2, 1 << 19, rng.random())


def test_chenliu_1d_nix2_at_2_pow_17_against_oracle():
    """BASELINE config 2 model (1-D NIX2 Chen-Liu, rejuv = 50) at 2^17 particles, rejuvenation steps included."""
    rng = np.random.default_rng(172)
    N, T = 1 << 17, 40
    centers = np.array([-6.0, -2.0, 2.0, 6.0])
    ys = centers[rng.integers(0, 4, T)] + rng.standard_normal(T)
    dev = pb.ChenLiuParticles(N, pb.NormalInverseChisq(0.0, 1.0, 0.1, 1.0), pb.ChineseRestaurantProcess(1.0), history=T,
                              rejuv=50.0, k_max=40, on_overflow="raise")
    ref = orc.OracleFilter(orc.CHENLIU, N, orc.Prior.nix2(0.0, 1.0, 0.1, 1.0), 1.0, rejuv=50.0)
    n_rejuv = 0
    for t in range(T):
        u = rng.random(2 * N)
        dev.fit_(ys[t], u)
        ref.cl_step(ys[t], u[:N], u[N:])
        s = dev.last_step()
        assert s["resampled"] == ref.stat("resampled"), f"rejuvenation decision differs at step {t}"
        assert s["cv"] == pytest.approx(ref.stat("cv"), rel=1e-8, abs=1e-10)
        n_rejuv += s["resampled"]
        assert np.array_equal(dev.last_ancestors(), ref.last_ancestors()), f"ancestors differ at step {t}"
        assert np.array_equal(dev.last_assignments() + 1, ref.last_assignments()), f"assignments differ at step {t}"
        _logclose(dev.weights(), ref.weights(), f"weights step {t}")
    assert n_rejuv >= 3
    assert np.array_equal(dev.ncomponents(), ref.ncomponents())
    assert np.array_equal(dev.assignments(), ref.assignments())


def test_chenliu_8d_niw_at_2_pow_14_against_oracle():
    """BASELINE config 5 model (8-D NIW(0, 0.1, I, 9), 32 clusters at 6 * sign patterns) at 2^14 particles:
    the Cholesky-update-heavy path, with rejuvenation steps."""
    rng = np.random.default_rng(85)
    N, T, d = 1 << 14, 48, 8
    centers = 6.0 * np.sign(rng.standard_normal((32, d)))
    ys = centers[rng.integers(0, 32, T)] + rng.standard_normal((T, d))
    dev = pb.ChenLiuParticles(N, pb.NormalInverseWishart(np.zeros(d), 0.1, np.eye(d), d + 1.0), pb.ChineseRestaurantProcess(1.0),
                              history=T, rejuv=50.0, k_max=63, on_overflow="raise")
    ref = orc.OracleFilter(orc.CHENLIU, N, orc.Prior.niw(np.zeros(d), 0.1, np.eye(d), d + 1.0), 1.0, rejuv=50.0)
    n_rejuv = 0
    for t in range(T):
        u = rng.random(2 * N)
        dev.fit_(ys[t], u)
        ref.cl_step(ys[t], u[:N], u[N:])
        s = dev.last_step()
        assert s["resampled"] == ref.stat("resampled"), f"rejuvenation decision differs at step {t}"
        n_rejuv += s["resampled"]
        assert np.array_equal(dev.last_ancestors(), ref.last_ancestors()), f"ancestors differ at step {t}"
        assert np.array_equal(dev.last_assignments() + 1, ref.last_assignments()), f"assignments differ at step {t}"
        _logclose(dev.weights(), ref.weights(), f"weights step {t}")
    assert n_rejuv >= 1
    assert np.array_equal(dev.assignments(), ref.assignments())
