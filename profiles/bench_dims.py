#!/usr/bin/env python
"""Fearnhead filter at other record sizes (d = 3, 4, 8 NIW; the headline is d = 2): ms per observation of the pipelined
pf_filter path on one GPU, per kernel class.  Usage: python profiles/bench_dims.py [d ...] [--log2n 20] [--steps 20]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "particles.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import particles_b200 as pb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("dims", nargs="*", type=int, default=[3, 4, 8])
ap.add_argument("--log2n", type=int, default=20)
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--burnin", type=int, default=60)
ap.add_argument("--precision", default="f64")
a = ap.parse_args()
for d in a.dims:
    rng = np.random.default_rng(7 + d)
    T = a.burnin + 2 * a.steps
    centers = 6.0 * np.sign(rng.standard_normal((8, d)))
    ys = centers[rng.integers(0, 8, T)] + rng.standard_normal((T, d))
    us = rng.random(T)
    ps = pb.FearnheadParticles(1 << a.log2n, pb.NormalInverseWishart(np.zeros(d), 0.1, np.eye(d), d + 1.0),
                               pb.ChineseRestaurantProcess(1.0), k_max=24, history=0, precision=a.precision, on_overflow="ignore")
    ps.filter_(ys[:a.burnin], uniforms=us[:a.burnin])
    ps.filter_(ys[a.burnin:a.burnin + a.steps], uniforms=us[a.burnin:a.burnin + a.steps])
    ms, _ = ps.timing()
    ps.set_profiling(True)
    ps.filter_(ys[a.burnin + a.steps:], uniforms=us[a.burnin + a.steps:])
    kt = ps.kernel_times()
    print(json.dumps({"d": d, "n_particles": 1 << a.log2n, "precision": a.precision, "ms_per_obs": round(ms / a.steps, 4),
                      "mean_clusters": round(float(ps.ncomponents().mean()), 2), "log_evidence_sum": float(np.sum(ps.log_evidence)),
                      "class_ms_per_obs": {k: round(v[0] / a.steps, 4) for k, v in kt.items() if v[1]}}), flush=True)
