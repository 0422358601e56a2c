#!/usr/bin/env python
"""bench.py -- particle-observation updates/s of the Fearnhead 2-D NIW filter (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W]             # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...      # the CPU restatement arm

A "step" is one observation filtered through the whole particle population
(fit!(ps, y), src/fearnhead.jl:130-184), i.e. N particle-observation updates.  The
workload is BASELINE.json configs[3] (the configuration the metric is quoted on): 2-D
NormalInverseWishart(0, 0.1, I, 3) prior, CRP(1), synthetic 8-cluster Gaussian mixture on a
circle of radius 6, 2^24 particles in total (strong scaling over the GPUs).  The population
is first brought to steady state (burn-in observations, untimed: the exhaustive phase and
the growth of the cluster count), then W warm-up steps, then EXACTLY K timed steps.

value  : device time of the K steps (CUDA events on the library's stream, max over ranks),
         state resident in HBM.
e2e    : the same K steps through the public host API (FearnheadParticles.filter_) with
         host (numpy) observation / uniform buffers: H2D copies of the inputs and the D2H
         read of the per-step results (log-evidence, counts) are inside the timed region.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "particles.jl_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "particle_obs_updates_per_sec"
UNIT = "updates/s"
D = 2
K_MAX = 24
T_MAX = 1024         # ONE synthetic stream for every arm (N = 1, N > 1, CPU): observation t is the same everywhere
SEED = 100           # echoes bench/2d.jl:12


def stream():
    rng = np.random.default_rng(SEED)
    return mixture(rng, T_MAX), rng.random(T_MAX)


def mixture(rng, T, k=8, radius=6.0):
    ang = 2 * np.pi * rng.integers(0, k, T) / k
    return radius * np.stack([np.cos(ang), np.sin(ang)], axis=1) + rng.standard_normal((T, 2))


STORE_BYTES = 8      # bytes per cluster field in the particle store (4 under --precision f32)


def rec_bytes(d):
    return STORE_BYTES * (2 + d + d * (d + 1) // 2)


def bytes_per_update(kbar, d=D):
    """Algorithmic bytes of one Fearnhead particle-observation update:
    B_F = (3 Kbar + 1) b + 16 Kbar + 53."""
    b = rec_bytes(d)
    return (3 * kbar + 1) * b + 16 * kbar + 53


def kernel_bytes(kbar, d=D):
    """Algorithmic bytes per particle of each kernel of the step."""
    b = rec_bytes(d)
    return {
        "expand": kbar * b + 16 + 8 * (kbar + 1),          # read live slots + header, write putative weights
        "select": 8 * (kbar + 1),                          # read putative weights once
        "scan": 8 * (kbar + 1) + 5,                        # read putative weights, write survivor codes
        "instantiate": kbar * b + 8 + 5 + (kbar + 1) * b + 16 + 5,   # gather parent, write child + header + genealogy
        # pipelined steps: instantiate fused with the next observation's expansion (the children are expanded from
        # registers: the expansion's read of the new population disappears) -- instantiate + the putative weights written
        "inst_expand": kbar * b + 8 + 5 + (kbar + 1) * b + 16 + 5 + 8 * (kbar + 1),
    }


class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.samples, self.stop_flag = device, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm, reasons, mx, pw = [], set(), None, []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = float(s[1])
                pw.append(float(s[2]))
                for nm, v in zip(names, s[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "sm_min_mhz": float(np.min(sm)) if sm else None, "power_w": float(np.median(pw)) if pw else None,
                "samples": len(sm)}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed
    `ncu --set full` capture of this workload (profiles/ncu_traffic.json; bench.py cannot run ncu)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)[kernel]
        return float(t["dram_bytes_per_launch"]), t["source"]
    except Exception:
        return None, "no committed ncu capture for this kernel"


def cpu_leg(n_cpu, burn, timed, seed=100):
    """The reference's CPU path (the literal C restatement in oracle/, 1 thread because
    the reference is single-threaded) on a bounded sample of the same workload."""
    from oracle import oracle as orc
    ys, us = stream()
    assert burn + timed <= T_MAX
    f = orc.OracleFilter(orc.FEARNHEAD, n_cpu, orc.Prior.niw(np.zeros(2), 0.1, np.eye(2), 3.0), 1.0,
                         order=orc.ORDER_SORTED, keep_history=False)
    for t in range(burn):
        f.fh_step(ys[t], us[t])
    t0 = time.perf_counter()
    for t in range(burn, burn + timed):
        f.fh_step(ys[t], us[t])
    dt = time.perf_counter() - t0
    kbar = float(f.ncomponents().mean())
    return {"value": n_cpu * timed / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{n_cpu} particles x {timed} observations after {burn} burn-in (same model and data stream), "
                      f"literal C restatement of src/fearnhead.jl (sort + 2 mll per putative), mean K {kbar:.2f}",
            "seconds": dt}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def evidence_check(le, first_obs):
    """Fingerprint of the log-evidence increments of the device-timed steps: identical on every arm
    (N = 1 and N > 1 filter the same stream and must agree bit for bit)."""
    import zlib
    le = np.ascontiguousarray(le, dtype=np.float64)
    return {"first_obs": int(first_obs), "n": int(len(le)), "sum": float(le.sum()), "crc32": int(zlib.crc32(le.tobytes()))}


def sharded_arm(a, dist, rank, world, local, N, K, W, config):
    """N > 1: particles block-sharded over the ranks, one process per GPU.  Every exchange of a step is done by
    the library's kernels through peer memory over NVLink (CUDA IPC); torch.distributed (NCCL) only carries the
    IPC handles at start-up and the barriers / max-over-ranks of the timing."""
    import torch
    import particles_b200 as pb
    from particles_b200.distributed import ShardedFearnheadParticles, TorchGroup
    dev = torch.device("cuda", local)
    ys, us = stream()
    T_all = a.burnin + W + 3 * K
    assert T_all <= T_MAX
    prior, crp = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0), pb.ChineseRestaurantProcess(1.0)
    ps = ShardedFearnheadParticles(N, prior, crp, TorchGroup(), k_max=a.k_max, history=0, device=local, on_overflow="ignore",
                                   precision=a.precision)

    def mean_k():
        K_loc = ps._local(ps.L.pf_get_ncomponents, np.int32, __import__("ctypes").c_int32)[0]
        t = torch.tensor([float(K_loc.sum()), float(len(K_loc)), float(K_loc.max(initial=0))], dtype=torch.float64, device=dev)
        mx = t[2:3].clone()
        dist.all_reduce(t[:2]); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        return float(t[0] / t[1]), int(mx.item())

    les = []
    o = 0
    ps.filter_(ys[o:o + a.burnin + W], us[o:o + a.burnin + W]); o += a.burnin + W
    les.append(ps.log_evidence.copy())
    kbar0, _ = mean_k()
    n0 = len(ps)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # (1) device-timed K steps: CUDA events on the library's stream around ONE pf_mg_filter call, max over ranks
    dist.barrier(); torch.cuda.synchronize(dev)
    ps.filter_(ys[o:o + K], us[o:o + K]); o += K
    torch.cuda.synchronize(dev); dist.barrier()
    les.append(ps.log_evidence.copy())
    dev_ms_local, launches = ps.timing()
    ms = torch.tensor([dev_ms_local], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dev_ms = float(ms.item())
    # (2) end to end: wall clock around the same public call with host observation buffers
    dist.barrier(); torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    ps.filter_(ys[o:o + K], us[o:o + K]); o += K
    le_e2e = ps.log_evidence.copy()
    torch.cuda.synchronize(dev); dist.barrier()
    wall = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(wall, op=dist.ReduceOp.MAX)
    e2e_s = float(wall.item())
    les.append(le_e2e)
    sampler.stop_flag = True
    kbar1, kmax_seen = mean_k()
    kbar = 0.5 * (kbar0 + kbar1)
    # (3) per-kernel-class device times over K more steps (the bracket / solve / offset kernels include the wait for
    # the peers' data: "select" and "scan" are where the replicated search and the exchanges show up)
    ps.set_profiling(True)
    ps.filter_(ys[o:o + K], us[o:o + K]); o += K
    les.append(ps.log_evidence.copy())
    prof_ms, _ = ps.timing()
    kt = ps.kernel_times()
    ps.set_profiling(False)
    stats = ps.last_step()
    ov = torch.tensor([float(stats["overflow"])], dtype=torch.float64, device=dev)
    dist.all_reduce(ov)
    overflow = int(ov.item())
    by_rank = [None] * world
    try:
        dist.all_gather_object(by_rank, {"rank": rank, "dev_ms_timed": round(dev_ms_local, 4), "n_local": ps.local_sizes()[0],
                                         "class_ms_per_step": {k: round(v[0] / K, 4) for k, v in kt.items() if v[1]}})
    except Exception as e:                                   # diagnostics must never cost the line
        by_rank = [{"error": str(e)[:200]}]
    # (4) correctness of the timed run: rank 0 filters the same observations unsharded; the log-evidence
    # increments of EVERY observation must be bit-identical
    verify = None
    if not a.no_verify:
        if rank == 0:
            ref = pb.FearnheadParticles(N, prior, crp, k_max=a.k_max, history=0, device=local, on_overflow="ignore", precision=a.precision)
            ref.filter_(ys[:o], uniforms=us[:o])
            le_all = np.concatenate(les)
            same = bool(np.array_equal(le_all, ref.log_evidence))
            verify = {"unsharded_run_on_rank0": True, "observations": int(o), "log_evidence_bitwise_equal": same,
                      "max_abs_diff": float(np.max(np.abs(le_all - ref.log_evidence)))}
            ref.close()
        dist.barrier()
    if rank != 0:
        return 0
    sampler.join(timeout=2)
    value = n0 * K / (dev_ms / 1e3)
    peak, peak_src = peaks()
    bpu = bytes_per_update(kbar)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(config, mean_clusters_per_particle=kbar, population=n0, kmax_overflow=overflow,
                                                max_clusters_in_a_particle=kmax_seen, host_syncs_per_step=0,
                                                exchange="peer stores + sequence flags inside the kernels (CUDA IPC over NVLink); "
                                                         "no collective call on the data path",
                                                select_rounds_last=stats["select_rounds"], by_rank=by_rank, verify=verify,
                                                evidence_check=evidence_check(les[1], a.burnin + W)),
            "e2e": {"value": n0 * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": (8 * D + 8) * world, "d2h_bytes_per_step": 136 * world,
                    "ms_per_step": 1e3 * e2e_s / K, "log_evidence_last": float(le_e2e[-1])},
            "gpu_launches": int(launches) * world,
            "roofline": {"bound": "hbm", "kernel": "whole step (per GPU)", "achieved": bpu * value / 1e9 / world, "peak": peak,
                         "unit": "GB/s", "frac": bpu * value / 1e9 / world / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_update": bpu,
                         "class_ms_per_step": {k: round(v[0] / K, 4) for k, v in kt.items() if v[1]}},
            "clocks": sampler.summary()}
    print(json.dumps(line))
    if overflow and not a.allow_overflow:
        print(f"k_max = {a.k_max} dropped {overflow} new-cluster candidates; raise --k-max", file=sys.stderr)
        return 1
    if verify is not None and not verify["log_evidence_bitwise_equal"]:
        print("the sharded run does not reproduce the unsharded run", file=sys.stderr)
        return 1
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--log2n", type=int, default=24, help="log2 of the TOTAL particle budget")
    ap.add_argument("--burnin", type=int, default=96)
    ap.add_argument("--cpu-n", type=int, default=16384)
    ap.add_argument("--cpu-steps", type=int, default=300)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-n2", type=int, default=1 << 20, help="reference arm: population of the second (out-of-cache) CPU sample, 0 = none")
    ap.add_argument("--cpu-burn2", type=int, default=20)
    ap.add_argument("--cpu-steps2", type=int, default=3)
    ap.add_argument("--order", default="index", choices=["index", "sorted"])
    ap.add_argument("--k-max", type=int, default=K_MAX, help="cluster slots per particle (the line fails if a candidate was dropped)")
    ap.add_argument("--allow-overflow", action="store_true")
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"],
                    help="storage type of the particle store (f32: PF_F32, floats in the store, FP64 arithmetic and selection)")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the unsharded re-run on rank 0 (outside the timed region)")
    a = ap.parse_args()
    global STORE_BYTES
    STORE_BYTES = 4 if a.precision == "f32" else 8
    K, W = a.steps, max(a.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    N = 1 << a.log2n
    config = {"workload": f"2-D NIW Fearnhead filter, 2^{a.log2n} particles, 8-cluster synthetic mixture "
                          f"(BASELINE configs[3]); inputs >> L2 (particle store {N * 8 * rec_bytes(D) / 1e9:.1f} GB at K=8)",
              "n_particles": N, "d": D, "k_max": a.k_max, "prior": "NIW(0, 0.1, I, 3)", "alpha": 1.0,
              "burnin_obs": a.burnin, "order": a.order, "store_precision": a.precision, "sharding": f"particles block-sharded over {a.gpus} GPU(s)",
              "l2_policy": "inputs larger than L2"}

    # ---------------------------------------------------------------- reference arm
    if a.impl == "reference":
        if rank != 0:
            return 0
        cb = cpu_leg(a.cpu_n, 48, max(K, 1) * 15)
        # the CPU arm runs a bounded sample of the workload: say which population it really filtered, and time a second,
        # out-of-cache sample (2^20 particles: ~0.5 GB of AoS state, a few observations) beside the in-cache one
        config = dict(config, n_particles_run=a.cpu_n, same_population_as_gpu_arm=False)
        big = None
        if a.cpu_n2 > 0:
            cb2 = cpu_leg(a.cpu_n2, a.cpu_burn2, a.cpu_steps2)
            big = {k: cb2[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": K,
                "warmup": W, "ms_per_step": 1e3 * cb["seconds"] / (max(K, 1) * 15), "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "cpu_baseline_out_of_cache": big,
                "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "Julia is not installed in this image: the reference arm is the literal C restatement "
                        "(oracle/particles_oracle.c), single-threaded like the reference"}
        print(json.dumps(line))
        return 0

    # ---------------------------------------------------------------- this repo's arm
    import particles_b200 as pb
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    if world > 1:
        return sharded_arm(a, dist, rank, world, local, N, K, W, config)

    ys, us = stream()
    T_all = a.burnin + W + 3 * K
    assert T_all <= T_MAX
    order = pb.PF_ORDER_INDEX if a.order == "index" else pb.PF_ORDER_SORTED
    ps = pb.FearnheadParticles(N, pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0),
                               pb.ChineseRestaurantProcess(1.0), k_max=a.k_max, history=T_all, device=local, order=order,
                               on_overflow="ignore", precision=a.precision)
    o = 0
    ps.filter_(ys[o:o + a.burnin], uniforms=us[o:o + a.burnin]); o += a.burnin
    if W:
        ps.filter_(ys[o:o + W], uniforms=us[o:o + W]); o += W
    kbar0 = float(ps.ncomponents().mean())
    n_live0 = len(ps)

    sampler = ClockSampler(local)
    sampler.start()
    # (1) device-timed K steps
    ps.filter_(ys[o:o + K], uniforms=us[o:o + K]); o += K
    le_timed = ps.log_evidence.copy()
    dev_ms, launches = ps.timing()
    # (2) end to end through the public API with host buffers
    t0 = time.perf_counter()
    ps.filter_(ys[o:o + K], uniforms=us[o:o + K])
    le = ps.log_evidence.copy()
    e2e_s = time.perf_counter() - t0
    o += K
    # (3) per-kernel CUDA-event times over K more steps
    ps.set_profiling(True)
    ps.filter_(ys[o:o + K], uniforms=us[o:o + K]); o += K
    prof_ms, _ = ps.timing()
    kt = ps.kernel_times()
    ps.set_profiling(False)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    kbar1 = float(ps.ncomponents().mean())
    kbar = 0.5 * (kbar0 + kbar1)
    stats = ps.last_step()

    value = n_live0 * K / (dev_ms / 1e3)
    e2e = n_live0 * K / e2e_s
    peak, peak_src = peaks()
    kb = kernel_bytes(kbar)
    per_kernel = {}
    for name in ("expand", "select", "scan", "instantiate", "inst_expand"):
        ms, n = kt.get(name, (0.0, 0))
        if n:
            avg = ms / n
            # a class may have several launches per step: its algorithmic bytes are per STEP IN WHICH IT RAN, so the
            # rate divides those bytes by the class's time in those steps (the separate expand / instantiate kernels
            # run only in the first / last step of a pipelined call); avg_ms stays the mean duration of one launch
            steps_active = min(n, K)
            per_kernel[name] = {"avg_ms": avg, "launches_per_step": n / K, "ms_per_step": ms / K, "share": ms / prof_ms,
                                "steps_active": steps_active, "alg_bytes_per_active_step": kb[name] * n_live0,
                                "achieved_gbs": kb[name] * n_live0 * steps_active / (ms * 1e-3) / 1e9}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_per_step"])      # (a single-launch class on this path)
    traffic, traffic_src = ncu_traffic("k_" + dom)
    kname = {"inst_expand": "k_expand<2,2,INST> (instantiate fused with the next observation's expansion)"}.get(dom, "k_" + dom)
    roof = {"bound": "hbm", "kernel": kname, "achieved": per_kernel[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
            "frac": per_kernel[dom]["achieved_gbs"] / peak, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": peak_src,
            "avg_launch_ms": per_kernel[dom]["avg_ms"], "share_of_step": per_kernel[dom]["share"],
            "algorithmic_bytes_per_update": kb[dom], "alg_bytes_per_launch": kb[dom] * n_live0, "per_kernel": per_kernel,
            "whole_step": {"algorithmic_bytes_per_update": bytes_per_update(kbar),
                           "achieved": bytes_per_update(kbar) * value / 1e9, "frac": bytes_per_update(kbar) * value / 1e9 / peak,
                           "frac_of_nominal_8tbs": bytes_per_update(kbar) * value / 1e9 / 8000.0}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": K, "warmup": W,
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict(config, mean_clusters_per_particle=kbar, population=n_live0,
                                                select_rounds_last=stats["select_rounds"], candidates_last=stats["n_candidates"],
                                                select_phase_us_last=stats["select_phase_us"], kmax_overflow=int(stats["overflow"]),
                                                evidence_check=evidence_check(le_timed, a.burnin + W),
                                                max_clusters_in_a_particle=int(ps.ncomponents().max())),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 8 * D + 8, "d2h_bytes_per_step": 136,
                    "ms_per_step": 1e3 * e2e_s / K, "log_evidence_last": float(le[-1])},
            "gpu_launches": int(launches), "roofline": roof, "clocks": sampler.summary()}
    if not a.no_cpu:
        cb = cpu_leg(a.cpu_n, 48, a.cpu_steps)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line))
    if stats["overflow"] and not a.allow_overflow:
        print(f"k_max = {a.k_max} dropped {stats['overflow']} new-cluster candidates: the run no longer computes the reference's "
              "filter; raise --k-max", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
