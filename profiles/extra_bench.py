"""Secondary configurations of BASELINE.json (configs[1], [2], [4]-like) on one B200: device
time per observation from the library's CUDA events.  Prints one JSON line per configuration."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "particles.jl_b200"))
import particles_b200 as pb  # noqa: E402


def mix1d(rng, T):
    return np.asarray([-6.0, -2.0, 2.0, 6.0])[rng.integers(0, 4, T)] + rng.standard_normal(T)


def mixnd(rng, T, d, k, radius=6.0):
    if d == 2:
        ang = 2 * np.pi * np.arange(k) / k
        c = radius * np.stack([np.cos(ang), np.sin(ang)], axis=1)
    else:
        c = radius * np.sign(np.random.default_rng(1).standard_normal((k, d)))
    return c[rng.integers(0, k, T)] + rng.standard_normal((T, d))


def run(name, ps, ys, burn, K, per_step_uniforms=None):
    ps.filter_(ys[:burn])
    ps.filter_(ys[burn:burn + 3])
    kb0 = float(ps.ncomponents().mean())
    t0 = time.perf_counter()
    ps.filter_(ys[burn + 3:burn + 3 + K])
    wall = time.perf_counter() - t0
    ms, launches = ps.timing()
    n = len(ps)
    kb = 0.5 * (kb0 + float(ps.ncomponents().mean()))
    print(json.dumps({"config": name, "particles": n, "steps": K, "ms_per_step": ms / K, "updates_per_s": n * K / (ms / 1e3),
                      "e2e_updates_per_s": n * K / wall, "mean_K": kb, "launches_per_step": launches / K,
                      "last_step": {k: v for k, v in ps.last_step().items() if k in ("resampled", "cv", "select_rounds")}}), flush=True)


def main():
    rng = np.random.default_rng(100)
    crp = pb.ChineseRestaurantProcess(1.0)
    # configs[1]: 1-D NIX2 Chen-Liu, 2^20 particles
    ys = mix1d(rng, 200)
    run("1-D NIX2 Chen-Liu 2^20 (configs[1])", pb.ChenLiuParticles(1 << 20, pb.NormalInverseChisq(0., 1., 0.1, 1.0), crp, k_max=16), ys, 100, 50)
    # configs[2]: 2-D NIW Fearnhead, 2^20 particles, index and reference-literal (sorted) order
    ys = mixnd(rng, 200, 2, 8)
    niw2 = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0)
    run("2-D NIW Fearnhead 2^20 index order (configs[2])", pb.FearnheadParticles(1 << 20, niw2, crp, k_max=16), ys, 100, 50)
    run("2-D NIW Fearnhead 2^20 sorted (reference-literal) order", pb.FearnheadParticles(1 << 20, niw2, crp, k_max=16, order=pb.PF_ORDER_SORTED), ys, 100, 20)
    # configs[4]-like on one GPU: 8-D NIW Chen-Liu, 2^19 particles (1/8 of 2^22), 32-cluster mixture, k_max 48
    ys = mixnd(rng, 260, 8, 32)
    niw8 = pb.NormalInverseWishart(np.zeros(8), 0.1, np.eye(8), 9.0)
    run("8-D NIW Chen-Liu 2^19 (one GPU's share of configs[4])", pb.ChenLiuParticles(1 << 19, niw8, crp, k_max=48), ys, 200, 30)


if __name__ == "__main__":
    main()
