"""Two shards of a 2^23-particle Fearnhead filter emulated on ONE GPU (LocalGroup, lock step): the sharded kernels with
every "remote" store landing in local memory -- separates the cost of the sharded kernel variants from the cost of NVLink."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "particles.jl_b200")]
import bench
import particles_b200 as pb
from particles_b200.distributed import LocalGroup, ShardedFearnheadParticles
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 23)
ys, us = bench.stream()
ps = ShardedFearnheadParticles(N, pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0), pb.ChineseRestaurantProcess(1.0), LocalGroup(G), k_max=24, on_overflow="ignore")
ps.filter_(ys[:100], us[:100])
ps.set_profiling(True)
ps.filter_(ys[100:120], us[100:120])
ms, n = ps.timing()
print("ms/step", ms / 20, {k: round(v[0] / 20, 4) for k, v in ps.kernel_times().items() if v[1]})
