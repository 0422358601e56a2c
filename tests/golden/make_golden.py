"""Generates the golden fixtures of tests/golden/ from the CPU oracle.

The reference is pure Julia and Julia is not installed here, so these are NOT outputs of the
reference itself: they are outputs of oracle/particles_oracle.c, whose arithmetic is pinned
against the reference's own known-answer / identity tests (tests/test_oracle_*.py).  The
fixtures freeze the oracle (a regression pin) and let the GPU parity tests run against
committed vectors.  Every input is stored in the fixture, so no RNG has to be reproduced.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as orc  # noqa: E402


def mixture(rng, T, d, k, radius=6.0):
    if d == 1:
        return (rng.integers(0, k, T) * 4.0 - 2.0 * (k - 1) + rng.standard_normal(T)).reshape(T, 1)
    c = rng.standard_normal((k, d)) * radius
    return c[rng.integers(0, k, T)] + rng.standard_normal((T, d))


def prior_of(kind, d):
    if kind == "nix2":
        return orc.Prior.nix2(0.0, 1.0, 0.1, 1.0), dict(kind="nix2", mu=0.0, s2=1.0, kappa=0.1, nu=1.0)
    return orc.Prior.niw(np.zeros(d), 0.1, np.eye(d), d + 1.0), dict(kind="niw", kappa=0.1, nu=d + 1.0)


def fearnhead(name, kind, d, N, T, k, order, seed):
    rng = np.random.default_rng(seed)
    ys = mixture(rng, T, d, k)
    us = rng.random(T)
    prior, meta = prior_of(kind, d)
    f = orc.OracleFilter(orc.FEARNHEAD, N, prior, 1.0, order=order)
    anc, asg, nk = [], [], []
    for y, u in zip(ys, us):
        f.fh_step(y, u)
        anc.append(f.last_ancestors().astype(np.int32))
        asg.append(f.last_assignments().astype(np.int16))
        nk.append([f.n, f.stat("n_kept"), f.stat("n_resamp_got")])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), filter="fearnhead", order=order, d=d, N=N, alpha=1.0,
                        prior_kind=meta["kind"], kappa0=meta["kappa"], nu0=meta["nu"], ys=ys, uniforms=us,
                        ancestors=np.concatenate(anc), assignments_step=np.concatenate(asg),
                        counts=np.array(nk, dtype=np.int64), weights=f.weights(), ncomponents=f.ncomponents(),
                        assignments=f.assignments().astype(np.int16))


def chenliu(name, kind, d, N, T, k, seed):
    rng = np.random.default_rng(seed)
    ys = mixture(rng, T, d, k)
    us = rng.random((T, 2 * N))
    prior, meta = prior_of(kind, d)
    f = orc.OracleFilter(orc.CHENLIU, N, prior, 1.0, rejuv=50.0)
    res, cv, anc = [], [], []
    for y, u in zip(ys, us):
        f.cl_step(y, u[:N], u[N:])
        res.append(f.stat("resampled")); cv.append(f.stat("cv")); anc.append(f.last_ancestors().astype(np.int32))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), filter="chenliu", d=d, N=N, alpha=1.0, rejuv=50.0,
                        prior_kind=meta["kind"], kappa0=meta["kappa"], nu0=meta["nu"], ys=ys, uniforms=us,
                        resampled=np.array(res, dtype=np.int8), cv=np.array(cv), ancestors=np.stack(anc),
                        weights=f.weights(), ncomponents=f.ncomponents(), assignments=f.assignments().astype(np.int16))


def selection(name, seed):
    """cutoff_ascending + sample_stratified! (src/fearnhead.jl:59-114) on fixed weight vectors, in
    exact rational arithmetic (oracle.select_exact) and by the literal FP64 restatement."""
    rng = np.random.default_rng(seed)
    out = {}
    cases = [(np.exp(rng.standard_normal(4000) * 3.0), 700), (rng.random(1000), 999), (np.exp(rng.standard_normal(3000) * 12.0), 50),
             (np.concatenate([np.full(500, 0.25), rng.random(1500)]), 600)]
    for c, (w, N) in enumerate(cases):
        u = float(rng.random())
        for oname, order in (("sorted", orc.ORDER_SORTED), ("index", orc.ORDER_INDEX)):
            surv, flag, Tn, n_kept = orc.select_exact(w, N, u, order)
            out[f"c{c}_{oname}_survivors"] = np.asarray(surv, dtype=np.int64)
            out[f"c{c}_{oname}_resampled"] = np.asarray(flag, dtype=np.int8)
            out[f"c{c}_n_kept"] = n_kept
            out[f"c{c}_log_c"] = np.log(orc.exact_c_value(Tn)) if Tn[0] > 0 and Tn[1] > 0 else np.nan
            lsurv, lflag, lc, lkept = orc.select_literal(w, N, u, order)
            out[f"c{c}_{oname}_literal_survivors"] = lsurv
            out[f"c{c}_literal_c"] = lc
        out[f"c{c}_w"] = w; out[f"c{c}_N"] = N; out[f"c{c}_u"] = u
    out["n_cases"] = len(cases)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)


if __name__ == "__main__":
    fearnhead("fearnhead_nix2_1d_n100", "nix2", 1, 100, 300, 4, orc.ORDER_SORTED, 1)          # BASELINE configs[0] (bench/1d.jl)
    fearnhead("fearnhead_nix2_1d_n100_index", "nix2", 1, 100, 300, 4, orc.ORDER_INDEX, 1)
    fearnhead("fearnhead_niw_2d_n500", "niw", 2, 500, 120, 8, orc.ORDER_SORTED, 2)            # bench/2d.jl prior
    fearnhead("fearnhead_niw_2d_n500_index", "niw", 2, 500, 120, 8, orc.ORDER_INDEX, 2)
    fearnhead("fearnhead_niw_4d_n200_index", "niw", 4, 200, 80, 5, orc.ORDER_INDEX, 3)
    chenliu("chenliu_nix2_1d_n200", "nix2", 1, 200, 150, 4, 4)
    chenliu("chenliu_niw_2d_n128", "niw", 2, 128, 100, 4, 5)
    selection("selection_cases", 6)
    print("written:", sorted(p for p in os.listdir(HERE) if p.endswith(".npz")))
