#!/usr/bin/env python
"""BASELINE configs[2] at FULL length: 2-D NIW Fearnhead filter, 2^20 particles, 10^4 observations, with the whole
genealogy kept (history < 0: pruned arena) -- then the chain read-outs.  PF_GENEALOGY_BYTES bounds the arena so that
pruning really happens (dense, this history is 52 GB)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "particles.jl_b200")]
import particles_b200 as pb

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
os.environ.setdefault("PF_GENEALOGY_BYTES", str(8 << 30))
N = 1 << log2n
rng = np.random.default_rng(100)
ang = 2 * np.pi * rng.integers(0, 8, T) / 8
truth = (ang / (2 * np.pi / 8)).round().astype(int)
ys = 6 * np.stack([np.cos(ang), np.sin(ang)], axis=1) + rng.standard_normal((T, 2))
ps = pb.FearnheadParticles(N, pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0), pb.ChineseRestaurantProcess(1.0), k_max=32,
                           history=-1, on_overflow="raise")
t0 = time.perf_counter()
ps.filter_(ys)                                   # device uniforms
t_filter = time.perf_counter() - t0
info = ps.genealogy_info()
t0 = time.perf_counter()
imap = int(np.argmax(ps.logweights()))
a_map = ps.particle_assignments(imap)
t_chain = time.perf_counter() - t0
t0 = time.perf_counter()
ri = ps.randindex(truth)
t_ri = time.perf_counter() - t0
print(json.dumps({"workload": f"2-D NIW Fearnhead, 2^{log2n} particles, {T} observations, genealogy kept", "filter_s": t_filter,
                  "updates_per_s_incl_pruning": N * T / t_filter, "genealogy": info, "dense_nodes": N * T,
                  "live_fraction": info["nodes"] / (N * T), "arena_GB": info["capacity"] * 5 / 1e9,
                  "map_particle_clusters": int(a_map.max()), "map_chain_readout_s": t_chain,
                  "adjusted_rand_index_vs_truth": ri, "randindex_s (T x T co-clustering over all particles on the device)": t_ri,
                  "mean_ncomponents": ps.mean(pb.ncomponents), "kmax_overflow": ps.last_step()["overflow"]}))
