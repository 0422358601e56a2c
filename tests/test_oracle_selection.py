"""Ports of test/fearnhead.jl (cutoff + stratified sampler) and test/clustering.jl for the
CPU oracle, plus literal-vs-exact agreement of the selection step."""
import numpy as np
import pytest

from oracle import oracle as orc


def cutoff(ws, N):
    # test/fearnhead.jl:8-12
    ws = np.sort(ws)
    i_keep, w = orc.cutoff_ascending(ws, N)
    return len(ws) - i_keep + 2, w, ws.sum()


def cutoff_normalized(ws, N):
    # test/fearnhead.jl:14-23 (expects descending input)
    ws = ws / ws.sum()
    total = ws.sum()
    for i in range(1, len(ws) + 1):
        if total / ws[i - 1] + (i - 1) >= N:
            return i, total / (N - i + 1)
        total -= ws[i - 1]
    return None


def test_cutoff_property_and_alternatives():
    # test/fearnhead.jl:36-50
    rng = np.random.default_rng(1)
    for trial in range(20):
        ws = np.sort(np.concatenate([[1e10] if trial % 2 == 0 else [], np.exp(rng.standard_normal(19 + trial % 2))]))[::-1]
        keep, cut, tot = cutoff(ws, 10)
        assert np.minimum(ws / cut, 1).sum() == pytest.approx(10, abs=1e-5)
        assert cutoff_normalized(ws, 10)[1] == pytest.approx(cut / tot, abs=1e-10)


def test_cutoff_edge_cases():
    # test/fearnhead.jl:52-71 and the worked examples of SURVEY 8(a11)
    keep, cut, tot = cutoff(np.ones(20), 10)
    assert keep == 1 and cut / tot == pytest.approx(1 / 10)
    assert orc.cutoff_ascending(np.ones(20), 10) == (21, 2.0)
    ws = np.array([1.0] * 10 + [0.0] * 10)
    assert cutoff(ws, 10)[0] == 11
    i, w = orc.cutoff_ascending(np.sort(ws), 10)
    assert i == 11 and np.isnan(w)
    rng = np.random.default_rng(2)
    ws = np.concatenate([rng.random(10), np.zeros(10)])
    assert cutoff(ws, 10)[0] == 11
    ws = np.array([.05, .05, .1, .1, .2, .5])
    assert orc.cutoff_ascending(ws, 2) == (6, 0.5)
    assert orc.cutoff_ascending(ws, 3) == (6, 0.25)
    assert orc.cutoff_ascending(ws, 4) == (5, 0.15000000000000002)
    assert orc.cutoff_ascending(ws, 5)[0] == 3
    with pytest.raises(ValueError):
        orc.cutoff_ascending(ws[::-1], 3)


def test_sample_stratified_count_bound_and_example():
    # test/fearnhead.jl:75-89
    rng = np.random.default_rng(1999)
    w = np.exp(2 * rng.random(100))
    w /= w.sum()
    for n in (1, 8, 64, 90, 99):
        for _ in range(100):
            assert len(orc.sample_stratified(w, n, rng.random())) <= n
    # SURVEY 8(a12) worked example: low set of the N=4 case, n=2, u=0.5 -> items 2 and 4
    picks = orc.sample_stratified([.05, .05, .1, .1], 2, 0.5)
    assert picks.tolist() == [1, 3]


@pytest.mark.parametrize("order", [orc.ORDER_SORTED, orc.ORDER_INDEX])
def test_exact_selection_agrees_with_literal(order):
    rng = np.random.default_rng(42)
    ndiff = 0
    for trial in range(60):
        M = int(rng.integers(30, 400))
        N = int(rng.integers(5, M))
        w = np.exp(3 * rng.standard_normal(M))
        if trial % 5 == 0:
            w[rng.integers(0, M, M // 4)] = 0.0
        if trial % 7 == 0:
            w = np.round(w, 1)              # heavy ties
        u = rng.random()
        s_lit, f_lit, c_lit, k_lit = orc.select_literal(w, N, u, order)
        s_ex, f_ex, Tn, k_ex = orc.select_exact(w, N, u, order)
        assert k_lit == k_ex
        if not (np.array_equal(s_lit, s_ex) and np.array_equal(f_lit, f_ex)):
            ndiff += 1
        if Tn[1] > 0:
            assert orc.exact_c_value(Tn) == pytest.approx(c_lit, rel=1e-13)
        assert len(s_ex) <= N
    assert ndiff == 0


def test_exact_selection_inclusion_probabilities():
    # each low item is selected with probability w_i / c ; kept items always
    w = np.array([.05, .05, .1, .1, .2, .5])
    hits = np.zeros(6)
    us = (np.arange(2000) + 0.5) / 2000
    for u in us:
        s, f, Tn, nk = orc.select_exact(w, 4, u, orc.ORDER_INDEX)
        hits[s] += 1
    c = 0.15
    np.testing.assert_allclose(hits / len(us), np.minimum(w / c, 1), atol=2e-3)


@pytest.mark.parametrize("kind", [orc.FEARNHEAD, orc.CHENLIU])
def test_clustering_invariants(kind):
    # test/clustering.jl:1-22 : 100 particles x 50 obs
    rng = np.random.default_rng(9)
    x = rng.random(50)
    prior = orc.Prior.nix2(0.0, 1.0, 2.0, 2.0)
    f = orc.OracleFilter(kind, 100, prior, alpha=1.0)
    for y in x:
        if kind == orc.FEARNHEAD:
            f.fh_step(y, rng.random())
        else:
            f.cl_step(y, rng.random(100), rng.random(100))
    A = f.assignments()
    assert A.shape == (50, f.n) and A.dtype == np.int64
    if kind == orc.CHENLIU:
        assert f.n == 100
    assert np.all(A.max(axis=0) == f.ncomponents())
    assert np.all(A.min(axis=0) == 1)
    assert f.weights().sum() == pytest.approx(1.0, rel=1e-9)


def test_fearnhead_population_growth_and_orders():
    rng = np.random.default_rng(4)
    prior = orc.Prior.nix2(0.0, 1.0, 0.1, 1.0)
    fs = orc.OracleFilter(orc.FEARNHEAD, 100, prior, 1.0, order=orc.ORDER_SORTED)
    fi = orc.OracleFilter(orc.FEARNHEAD, 100, prior, 1.0, order=orc.ORDER_INDEX)
    sizes = []
    for t in range(30):
        y = rng.standard_normal() + (4 if t % 2 else -4)
        u = rng.random()
        fs.fh_step(y, u)
        fi.fh_step(y, u)
        sizes.append(fs.n)
        if fs.stat("resampled"):
            w = fs.weights()
            assert np.all(np.diff(w[fs.stat("n_resamp_got"):]) >= 0)      # kept part ascending
        assert fi.n <= 100 and fs.n <= 100
    assert sizes[:5] == [1, 2, 5, 15, 52]
    # the first resampling step sees identical inputs in both orders: same kept set size
    assert max(sizes) <= 100


def test_select_exact_fast_equals_select_exact():
    """The large-M evaluation of the exact specification (numpy sort + C-speed integer prefix sums +
    binary search over distinct values) is the same function as the straightforward one."""
    rng = np.random.default_rng(4242)
    cases = []
    for trial in range(60):
        M = int(rng.integers(5, 1500))
        N = int(rng.integers(1, M))
        w = np.exp(rng.standard_normal(M) * rng.choice([0.5, 3.0, 12.0]))
        if trial % 4 == 0:
            w[rng.integers(0, M, M // 3)] = 0.0
        if trial % 5 == 0:
            w = np.round(w, 1)
        if trial % 7 == 0:
            w[rng.integers(0, M)] = 1e12
        cases.append((w, N, rng.random()))
    cases.append((np.ones(20), 10, 0.3))
    cases.append((np.array([1.0] * 10 + [0.0] * 10), 10, 0.3))
    cases.append((np.array([.05, .05, .1, .1, .2, .5]), 4, 0.5))
    cases.append((np.array([1e-300, 1e-310, 5e-324, 1.0, 2.0, 3.0]), 4, 0.7))
    cases.append((np.full(300, 3.0), 7, 0.9))
    for w, N, u in cases:
        for order in (orc.ORDER_INDEX, orc.ORDER_SORTED):
            a = orc.select_exact(w, N, u, order)
            b = orc.select_exact_fast(w, N, u, order)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[3] == b[3]
            if a[2][1] > 0 and a[2][0] > 0:
                assert orc.exact_c_value(a[2]) == orc.exact_c_value(b[2])
