"""Per-observation log-evidence and mean K of the bench workload, dumped for an A/B of two library builds
(PARTICLES_B200_LIB selects the build).  Usage: python profiles/dump_evidence.py out.npz [log2n] [T]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "particles.jl_b200")):
    sys.path.insert(0, p)
import bench  # noqa: E402
import particles_b200 as pb  # noqa: E402

out = sys.argv[1]
log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
T = int(sys.argv[3]) if len(sys.argv) > 3 else 161
ys, us = bench.stream()
ps = pb.FearnheadParticles(1 << log2n, pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0), pb.ChineseRestaurantProcess(1.0),
                           k_max=24, history=0, on_overflow="ignore")
le, kb, st = [], [], []
for t in range(T):
    ps.filter_(ys[t:t + 1], uniforms=us[t:t + 1])
    s = ps.last_step()
    le.append(ps.log_evidence[0]); st.append((s["select_rounds"], s.get("select_status", -1), s["n_candidates"], s["n_kept"], s["n_resamp_got"]))
    if t % 20 == 0 or t == T - 1:
        kb.append((t, float(ps.ncomponents().mean())))
np.savez(out, le=np.array(le), kb=np.array(kb), st=np.array(st, dtype=np.int64))
print(out, "sum le", float(np.sum(le)), "kbar", kb[-1])
