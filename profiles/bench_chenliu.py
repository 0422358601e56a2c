#!/usr/bin/env python
"""Chen-Liu lines for BASELINE configs[1] (1-D NIX2, 2^20 particles, 1 GPU) and configs[4] (8-D NIW, 2^22 particles,
32 clusters, sharded over the GPUs of the node).  Same JSON shape as bench.py.

    python profiles/bench_chenliu.py --config 2 [--log2n 20] [--steps 200]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        profiles/bench_chenliu.py --config 5 --log2n 22 --steps 40

A step is one observation through fit!(::ChenLiuParticles, y) (src/chenliu.jl:54-69): propagation of every particle with
one categorical draw, the coefficient of variation, and -- when it exceeds 50 % -- the stratified rejuvenation.  Uniforms
come from the device stream (pf_uniform).  Algorithmic bytes per update: B_CL = (Kbar + 1) b + 37, plus
2 Kbar b on a rejuvenation step; Kbar is measured."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "particles.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2, choices=[2, 5])
    ap.add_argument("--log2n", type=int, default=None)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--burnin", type=int, default=None)
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    import particles_b200 as pb
    rng = np.random.default_rng(100)
    if a.config == 2:
        d, log2n, K, burn, k_max = 1, a.log2n or 20, a.steps or 200, a.burnin or 200, 32
        ys = (np.array([-6.0, -2.0, 2.0, 6.0])[rng.integers(0, 4, burn + 2 * K)] + rng.standard_normal(burn + 2 * K)).reshape(-1, 1)
        prior, name = pb.NormalInverseChisq(0.0, 1.0, 0.1, 1.0), "1-D NIX2 Chen-Liu, 4-cluster mixture (BASELINE configs[1])"
    else:
        d, log2n, K, burn, k_max = 8, a.log2n or (22 if world > 1 else 19), a.steps or 40, a.burnin or 120, 63
        centers = 6.0 * np.sign(rng.standard_normal((32, d)))
        ys = centers[rng.integers(0, 32, burn + 2 * K)] + rng.standard_normal((burn + 2 * K, d))
        prior, name = pb.NormalInverseWishart(np.zeros(d), 0.1, np.eye(d), d + 1.0), "8-D NIW Chen-Liu, 32-cluster mixture (BASELINE configs[4])"
    N = 1 << log2n
    crp = pb.ChineseRestaurantProcess(1.0)
    if world > 1:
        import torch
        import torch.distributed as dist
        from particles_b200.distributed import ShardedChenLiuParticles, TorchGroup
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ps = ShardedChenLiuParticles(N, prior, crp, TorchGroup(), k_max=k_max, history=0, device=local, seed=5, precision=a.precision,
                                     on_overflow="ignore")
        barrier = dist.barrier
    else:
        ps = pb.ChenLiuParticles(N, prior, crp, k_max=k_max, history=0, seed=5, precision=a.precision, on_overflow="ignore")
        barrier = lambda: None
    ps.filter_(ys[:burn])
    barrier()
    ps.filter_(ys[burn:burn + K])
    dev_ms, launches = ps.timing()
    le = ps.log_evidence.copy()
    barrier()
    t0 = time.perf_counter()
    ps.filter_(ys[burn + K:burn + 2 * K])
    e2e_s = time.perf_counter() - t0
    if world > 1:
        import torch
        t = torch.tensor([dev_ms, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s = float(t[0]), float(t[1])
        Kc = ps._local(ps.L.pf_get_ncomponents, np.int32, __import__("ctypes").c_int32)[0]
        s = torch.tensor([float(Kc.sum()), float(len(Kc))], dtype=torch.float64, device="cuda")
        dist.all_reduce(s)
        kbar = float(s[0] / s[1])
    else:
        kbar = float(ps.ncomponents().mean())
    st = ps.last_step()
    if rank == 0:
        sb = 4 if a.precision == "f32" else 8
        b = sb * (2 + d + d * (d + 1) // 2)
        bcl = (kbar + 1) * b + 37
        value = N * K / (dev_ms / 1e3)
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            peak = 6650.0
        print(json.dumps({"metric": "particle_obs_updates_per_sec", "value": value, "unit": "updates/s", "n_gpus": world, "steps": K,
                          "warmup": burn, "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                          "dtype": "f64", "data": "synthetic",
                          "config": {"workload": f"{name}, 2^{log2n} particles", "n_particles": N, "d": d, "k_max": k_max, "rejuv": 50.0,
                                     "store_precision": a.precision, "mean_clusters_per_particle": kbar, "cv_last": st["cv"],
                                     "kmax_overflow": int(st["overflow"]), "log_evidence_sum": float(le.sum()),
                                     "uniforms": "device stream (pf_uniform)"},
                          "e2e": {"value": N * K / e2e_s, "unit": "updates/s", "h2d_bytes_per_step": 8 * d, "d2h_bytes_per_step": 136},
                          "gpu_launches": int(launches) * world,
                          "roofline": {"bound": "hbm", "kernel": "whole step (per GPU)", "achieved": bcl * value / 1e9 / world, "peak": peak, "unit": "GB/s",
                                       "frac": bcl * value / 1e9 / world / peak, "traffic": None,
                                       "algorithmic_bytes_per_update": bcl, "note": "B_CL = (Kbar + 1) b + 37 (no rejuvenation term)"}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
