"""Timeline inside the cluster search kernels (globaltimer stamps, PF_DEBUG_STAMPS=1) at the bench state."""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "particles.jl_b200")):
    sys.path.insert(0, p)
os.environ["PF_DEBUG_STAMPS"] = "1"
import bench, particles_b200 as pb
ys, us = bench.stream()
ps = pb.FearnheadParticles(1 << 24, pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0), pb.ChineseRestaurantProcess(1.0), k_max=24, history=0, on_overflow="ignore")
ps.filter_(ys[:100], uniforms=us[:100])
out = (C.c_int64 * 24)()
for t in range(100, 104):
    ps.filter_(ys[t:t + 1], uniforms=us[t:t + 1])
    ps._L.pf_debug_dump(ps._h, out)
