"""A/B consistency of the step's code paths at a size the oracle cannot reach: the same stream through
(a) the default fused path, (b) the classic path (PF_NO_FUSE=1), (c) a narrow sample bracket that forces
retry rounds (PF_SEL_MARGIN=2), (d) 4 shards in lock step.  All decisions are exact integer arithmetic, so
every arm must give identical populations; log-evidence may differ in the last bits only when the bracket
differs.  Usage: python profiles/ab_check.py [log2n] [T]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "particles.jl_b200")):
    sys.path.insert(0, p)
import particles_b200 as pb  # noqa: E402
from particles_b200.distributed import LocalGroup, ShardedFearnheadParticles  # noqa: E402

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
N = 1 << log2n
rng = np.random.default_rng(100)
ang = 2 * np.pi * rng.integers(0, 8, T) / 8
ys = 6 * np.stack([np.cos(ang), np.sin(ang)], axis=1) + rng.standard_normal((T, 2))
us = rng.random(T)
prior, crp = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0), pb.ChineseRestaurantProcess(1.0)


def run(env, sharded=0):
    for k in ("PF_NO_FUSE", "PF_SEL_MARGIN"):
        os.environ.pop(k, None)
    os.environ.update(env)
    os.environ["PF_DEBUG_STATUS"] = "1"
    if sharded:
        ps = ShardedFearnheadParticles(N, prior, crp, LocalGroup(sharded), k_max=24, history=0)
        ps.filter_(ys, us)
        rounds = None
    else:
        ps = pb.FearnheadParticles(N, prior, crp, k_max=24, history=0, on_overflow="ignore")
        rounds = []
        for t in range(T):
            ps.filter_(ys[t:t + 1], uniforms=us[t:t + 1])
            s = ps.last_step()
            rounds.append((s["select_rounds"] % 1000, s["select_status"], s["n_candidates"]))
        ps.log_evidence = None
    out = (ps.logweights(), ps.ncomponents(), rounds)
    ps.close()
    return out


base = run({})
print("default: rounds/status/cands of the last 5 steps", base[2][-5:], "max rounds", max(r[0] for r in base[2]))
for name, env, sh in (("no_fuse", {"PF_NO_FUSE": "1"}, 0), ("margin2", {"PF_SEL_MARGIN": "2"}, 0), ("sharded4", {}, 4)):
    o = run(env, sh)
    same_k = np.array_equal(o[1], base[1])
    dlw = float(np.max(np.abs(o[0] - base[0]))) if len(o[0]) == len(base[0]) else float("nan")
    print(f"{name}: n {len(o[0])} vs {len(base[0])}, K equal {same_k}, max |dlogw| {dlw:.3e}",
          ("max rounds %d" % max(r[0] for r in o[2])) if o[2] else "")
