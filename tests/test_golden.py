"""Golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py from the CPU
oracle): the oracle must keep reproducing them (CPU, a regression pin on the checker), and the
CUDA path must reproduce them through the public API (GPU) -- ancestors, assignments and
selected indices bit-exact, weights within 1e-10 relative in log space."""
import glob
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FH = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "fearnhead_*.npz")))
CL = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "chenliu_*.npz")))
RTOL = 1e-10


def _load(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


def _oracle_prior(g):
    d = int(g["d"])
    if str(g["prior_kind"]) == "nix2":
        return orc.Prior.nix2(0.0, 1.0, float(g["kappa0"]), float(g["nu0"]))
    return orc.Prior.niw(np.zeros(d), float(g["kappa0"]), np.eye(d), float(g["nu0"]))


def _device_prior(pb, g):
    d = int(g["d"])
    if str(g["prior_kind"]) == "nix2":
        return pb.NormalInverseChisq(0.0, 1.0, float(g["kappa0"]), float(g["nu0"]))
    return pb.NormalInverseWishart(np.zeros(d), float(g["kappa0"]), np.eye(d), float(g["nu0"]))


def _logclose(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and np.all((a > 0) == (b > 0))
    pos = b > 0
    la, lb = np.log(a[pos]), np.log(b[pos])
    assert (np.abs(la - lb) / np.maximum(np.abs(lb), 1.0)).max() <= RTOL


def test_fixtures_exist():
    assert len(FH) >= 5 and len(CL) >= 2 and os.path.exists(os.path.join(GOLD, "selection_cases.npz"))


# ------------------------------------------------------------------ CPU: the oracle against its fixtures
@pytest.mark.parametrize("name", FH)
def test_oracle_reproduces_fearnhead_fixture(name):
    g = _load(name)
    f = orc.OracleFilter(orc.FEARNHEAD, int(g["N"]), _oracle_prior(g), float(g["alpha"]), order=int(g["order"]))
    anc, asg, counts = [], [], []
    for y, u in zip(g["ys"], g["uniforms"]):
        f.fh_step(y, u)
        anc.append(f.last_ancestors()); asg.append(f.last_assignments())
        counts.append([f.n, f.stat("n_kept"), f.stat("n_resamp_got")])
    assert np.array_equal(np.concatenate(anc), g["ancestors"])
    assert np.array_equal(np.concatenate(asg), g["assignments_step"])
    assert np.array_equal(np.array(counts), g["counts"])
    assert np.array_equal(f.assignments(), g["assignments"])
    np.testing.assert_allclose(f.weights(), g["weights"], rtol=1e-12, atol=0)


@pytest.mark.parametrize("name", CL)
def test_oracle_reproduces_chenliu_fixture(name):
    g = _load(name)
    N = int(g["N"])
    f = orc.OracleFilter(orc.CHENLIU, N, _oracle_prior(g), float(g["alpha"]), rejuv=float(g["rejuv"]))
    for t, (y, u) in enumerate(zip(g["ys"], g["uniforms"])):
        f.cl_step(y, u[:N], u[N:])
        assert f.stat("resampled") == g["resampled"][t]
        assert np.array_equal(f.last_ancestors(), g["ancestors"][t])
    assert np.array_equal(f.assignments(), g["assignments"])
    np.testing.assert_allclose(f.weights(), g["weights"], rtol=1e-12, atol=0)


def test_oracle_reproduces_selection_fixture():
    g = _load("selection_cases")
    for c in range(int(g["n_cases"])):
        w, N, u = g[f"c{c}_w"], int(g[f"c{c}_N"]), float(g[f"c{c}_u"])
        for oname, order in (("sorted", orc.ORDER_SORTED), ("index", orc.ORDER_INDEX)):
            surv, flag, Tn, n_kept = orc.select_exact(w, N, u, order)
            assert np.array_equal(surv, g[f"c{c}_{oname}_survivors"]) and n_kept == int(g[f"c{c}_n_kept"])
            assert np.array_equal(flag.astype(np.int8), g[f"c{c}_{oname}_resampled"])
            lsurv, _, lc, _ = orc.select_literal(w, N, u, order)
            assert np.array_equal(lsurv, g[f"c{c}_{oname}_literal_survivors"])


# ------------------------------------------------------------------ GPU: the CUDA path against the fixtures
@pytest.mark.gpu
@pytest.mark.parametrize("name", FH)
def test_device_reproduces_fearnhead_fixture(name):
    import particles_b200 as pb
    g = _load(name)
    T = len(g["ys"])
    order = pb.PF_ORDER_SORTED if int(g["order"]) == orc.ORDER_SORTED else pb.PF_ORDER_INDEX
    ps = pb.FearnheadParticles(int(g["N"]), _device_prior(pb, g), pb.ChineseRestaurantProcess(float(g["alpha"])),
                               history=T, order=order, k_max=40)      # (the reference's component vector is unbounded)
    anc, asg, counts = [], [], []
    for y, u in zip(g["ys"], g["uniforms"]):
        ps.fit_(y, [u])
        s = ps.last_step()
        anc.append(ps.last_ancestors()); asg.append(ps.last_assignments() + 1)
        counts.append([len(ps), s["n_kept"], s["n_resamp_got"]])
    assert np.array_equal(np.concatenate(anc), g["ancestors"])
    assert np.array_equal(np.concatenate(asg), g["assignments_step"])
    assert np.array_equal(np.array(counts), g["counts"])
    assert np.array_equal(ps.assignments(), g["assignments"])
    assert np.array_equal(ps.ncomponents(), g["ncomponents"])
    _logclose(ps.weights(), g["weights"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CL)
def test_device_reproduces_chenliu_fixture(name):
    import particles_b200 as pb
    g = _load(name)
    N, T = int(g["N"]), len(g["ys"])
    ps = pb.ChenLiuParticles(N, _device_prior(pb, g), pb.ChineseRestaurantProcess(float(g["alpha"])), history=T,
                             rejuv=float(g["rejuv"]), k_max=24)
    for t, (y, u) in enumerate(zip(g["ys"], g["uniforms"])):
        ps.fit_(y, u)
        assert ps.last_step()["resampled"] == g["resampled"][t]
        assert np.array_equal(ps.last_ancestors(), g["ancestors"][t])
    assert np.array_equal(ps.assignments(), g["assignments"])
    assert np.array_equal(ps.ncomponents(), g["ncomponents"])
    _logclose(ps.weights(), g["weights"])


@pytest.mark.gpu
def test_device_reproduces_selection_fixture():
    import particles_b200 as pb
    g = _load("selection_cases")
    for c in range(int(g["n_cases"])):
        w, N, u = g[f"c{c}_w"], int(g[f"c{c}_N"]), float(g[f"c{c}_u"])
        for oname, order in (("sorted", pb.PF_ORDER_SORTED), ("index", pb.PF_ORDER_INDEX)):
            surv, flag, st = pb.select(w, N, u, order)
            assert np.array_equal(surv, g[f"c{c}_{oname}_survivors"])
            assert np.array_equal(flag.astype(np.int8), g[f"c{c}_{oname}_resampled"])
            assert st["n_kept"] == int(g[f"c{c}_n_kept"])
            if np.isfinite(g[f"c{c}_log_c"]):
                assert st["log_w_resamp"] == pytest.approx(float(g[f"c{c}_log_c"]), rel=1e-12, abs=1e-12)
