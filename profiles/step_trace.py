"""Per-observation trace of the single-GPU Fearnhead step at the bench workload: wall time of
each pf_step (synchronised), select rounds / candidates / phase timings and the per-class kernel
times.  Diagnostic only (synchronises every step)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "particles.jl_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import particles_b200 as pb
from bench import mixture
import torch

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
T = int(sys.argv[2]) if len(sys.argv) > 2 else 159
rng = np.random.default_rng(100)
ys = mixture(rng, T)
us = rng.random(T)          # the same stream bench.py draws for T = burnin + warmup + 3 K
ps = pb.FearnheadParticles(1 << log2n, pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0), pb.ChineseRestaurantProcess(1.0),
                           k_max=16, history=T)
ps.set_profiling(True)
prev = {k: v[0] for k, v in ps.kernel_times().items()}
for t in range(T):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ps.filter_(ys[t:t + 1], uniforms=us[t:t + 1])
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    st = ps.last_step()
    kt = {k: v[0] for k, v in ps.kernel_times().items()}
    d = {k: round(kt[k] - prev[k], 3) for k in kt if kt[k] - prev[k] > 0}
    prev = kt
    if t >= 90:
        print(json.dumps({"t": t, "ms": round(dt, 3), "rounds": st["select_rounds"], "cand": st["n_candidates"], "n_put": st["n_putative"],
                          "kept": st["n_kept"], "phase_us": [round(x, 1) for x in st["select_phase_us"]], **d}))
