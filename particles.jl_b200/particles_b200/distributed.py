"""Sharded Fearnhead and Chen-Liu filters: particles block-sharded over the GPUs of one node.

All device work, including every exchange between the shards, happens inside
libparticles_b200.so: the kernels write what other ranks need straight into the peers' device
memory over NVLink and synchronise through sequence flags (include/particles_b200.h,
"multi-GPU").  This module only does the one-time plumbing a host program owes the library:

  * one process per GPU (torchrun / mpirun ...): all-gather the 1088-byte CUDA-IPC blobs of
    `pf_mg_ipc_handles` once (any side channel; `TorchGroup` uses torch.distributed, which is
    also what the tests use over gloo for the host-side logic), `pf_mg_set_peer`, `pf_mg_connect`;
  * one process driving several GPUs (`LocalGroup`): `pf_mg_create_local` does all of it.

`ShardedFearnheadParticles.filter_` is then ONE library call per rank (`pf_mg_filter`)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import PF_CHENLIU, PF_FEARNHEAD, PF_MG_IPC_BYTES, PF_MG_NARRAYS, PF_NIW, PF_NIX2, pf_config, pf_step_stats

_dp = C.POINTER(C.c_double)


def equal_blocks(n_global, G):
    """Ownership after a step: rank r owns global particles [r q, (r+1) q) with q = ceil(n / G)
    (mirrors k_rank_offsets).  Returns the G + 1 boundaries."""
    q = max((int(n_global) + G - 1) // G, 1)
    return [min(r * q, int(n_global)) for r in range(G + 1)]


def owner_of(g, n_global, G):
    """Rank that owns global particle g after a step."""
    q = max((int(n_global) + G - 1) // G, 1)
    return np.asarray(g) // q


class TorchGroup:
    """One shard per process; torch.distributed (NCCL or gloo) carries the one-time IPC blobs and
    the host-side gathers of the read-outs.  The data path never touches it."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.size = dist.get_world_size()
        self.rank = dist.get_rank()
        self.local_ranks = [self.rank]

    def all_gather_blobs(self, blob):
        out = [None] * self.size
        self.dist.all_gather_object(out, blob)
        return out

    def gather_host(self, arrays):
        out = [None] * self.size
        self.dist.all_gather_object(out, arrays[0])
        return out

    def barrier(self):
        self.dist.barrier()


class LocalGroup:
    """All `size` shards live in this process: on `devices` (one per shard; peer access), or all on one
    device (lock-step emulation: the single-GPU tests of "sharded run == unsharded run")."""

    def __init__(self, size, devices=None):
        self.size = int(size)
        self.rank = 0
        self.local_ranks = list(range(self.size))
        self.devices = list(devices) if devices is not None else None

    def gather_host(self, arrays):
        return list(arrays)

    def barrier(self):
        pass


LocalComm = LocalGroup          # (name used by earlier tests)


def _config(n, prior, stateprior, k_max, history, seed, keep, kind=PF_FEARNHEAD, rejuv=50.0, precision="f64"):
    from . import NormalInverseChisq, NormalInverseWishart
    cfg = pf_config()
    cfg.filter_kind, cfg.precision, cfg.n_particles, cfg.k_max = kind, (1 if precision == "f32" else 0), int(n), int(k_max)
    cfg.order, cfg.history = 0, int(history)
    cfg.alpha, cfg.rejuv_threshold, cfg.seed = stateprior.alpha, float(rejuv), seed
    if isinstance(prior, NormalInverseChisq):
        cfg.prior_kind, cfg.d = PF_NIX2, 1
        mu0, lam0 = np.array([prior.mu]), np.zeros((1, 1))
        cfg.kappa0, cfg.nu0, cfg.sigma2_0 = prior.kappa, prior.nu, prior.sigma2
    elif isinstance(prior, NormalInverseWishart):
        cfg.prior_kind, cfg.d = PF_NIW, prior.dim
        mu0, lam0 = prior.mu.copy(), np.asfortranarray(prior.Lam)
        cfg.kappa0, cfg.nu0, cfg.sigma2_0 = prior.kappa, prior.nu, 0.0
    else:
        raise ValueError("prior must be NormalInverseChisq or NormalInverseWishart")
    keep.extend([mu0, lam0])
    cfg.mu0 = mu0.ctypes.data_as(_dp)
    cfg.lambda0 = lam0.ctypes.data_as(_dp)
    return cfg


class ShardedFearnheadParticles:
    """FearnheadParticles (src/fearnhead.jl:13-30) with the population sharded over group.size GPUs.
    Same results as the unsharded filter, particle for particle (global order = rank-major)."""

    _kind = PF_FEARNHEAD

    def __init__(self, n, prior, stateprior, group, *, k_max=32, history=0, device=None, seed=0, on_overflow="warn", rejuv=50.0,
                 precision="f64"):
        self.group, self.N, self.k_max = group, int(n), int(k_max)
        self.on_overflow, self._overflow_seen = on_overflow, 0
        L = _lib.load()
        self.L = L
        self._keep = []
        cfg = _config(n, prior, stateprior, k_max, history, seed, self._keep, self._kind, rejuv, precision)
        self.d = int(cfg.d)
        self._alpha = float(stateprior.alpha)
        G = group.size
        if isinstance(group, LocalGroup):
            devs = group.devices if group.devices is not None else [0 if device is None else int(device)] * G
            hs = (C.c_void_p * G)()
            _lib.check(L.pf_mg_create_local(C.byref(cfg), (C.c_int32 * G)(*devs), G, hs))
            self.handles = [C.c_void_p(hs[i]) for i in range(G)]
        else:
            cfg.rank, cfg.nranks = group.rank, G
            cfg.device = 0 if device is None else int(device)
            h = C.c_void_p()
            _lib.check(L.pf_create(C.byref(cfg), C.byref(h)))
            self.handles = [h]
            blob = C.create_string_buffer(PF_MG_NARRAYS * PF_MG_IPC_BYTES)
            _lib.check(L.pf_mg_ipc_handles(h, blob), h)
            blobs = group.all_gather_blobs(blob.raw)
            for r, b in enumerate(blobs):
                if r != group.rank:
                    _lib.check(L.pf_mg_set_peer(h, r, None, b), h)
            _lib.check(L.pf_mg_set_peer(h, group.rank, None, None), h)
            arr = (C.c_void_p * 1)(h)
            _lib.check(L.pf_mg_connect(arr, 1), h)
            group.barrier()              # every rank has opened every peer's memory before anyone stores into it
        self._harr = (C.c_void_p * len(self.handles))(*[h.value for h in self.handles])
        self.log_evidence = np.zeros(0)

    def close(self):
        for h in getattr(self, "handles", []):
            if h:
                self.L.pf_destroy(h)
        self.handles = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- stepping
    def fit_(self, y, u=None):
        """fit!(ps, y) (src/fearnhead.jl:130-184); u = the rand() of src/fearnhead.jl:99."""
        return self.filter_(np.atleast_2d(np.asarray(y, dtype=np.float64).reshape(1, -1)), None if u is None else [u])

    def filter_(self, ys, uniforms=None):
        """filter!(ps, ys, false) (src/filters.jl:6-13): one library call, every exchange on the device."""
        ys = np.ascontiguousarray(np.asarray(ys, dtype=np.float64).reshape(-1, self.d))
        T = ys.shape[0]
        le = np.empty(T)
        if uniforms is None:
            pu, ups = None, 0
        else:
            us = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(T, -1)
            pu, ups = us.ctypes.data_as(_dp), us.shape[1]
        rc = self.L.pf_mg_filter(self._harr, len(self.handles), ys.ctypes.data_as(_dp), T, pu, ups, le.ctypes.data_as(_dp))
        _lib.check(rc, self.handles[0])
        self.log_evidence = le
        self._check_overflow()
        return self

    def _check_overflow(self):
        if self.on_overflow == "ignore":
            return
        ov = self.last_step()["overflow"]          # this shard's cumulative count
        if ov > self._overflow_seen:
            import warnings
            msg = (f"{ov - self._overflow_seen} new-cluster candidates were dropped because a particle already had k_max = "
                   f"{self.k_max} components; construct the filter with a larger k_max")
            self._overflow_seen = ov
            if self.on_overflow == "raise":
                raise _lib.PFError(_lib.PF_ERR_OVERFLOW, msg)
            warnings.warn(msg, RuntimeWarning, stacklevel=3)

    # ---- read-out (concatenated over ranks in global order)
    def _local(self, fn, dtype, ctype):
        outs = []
        for h in self.handles:
            n = int(self.L.pf_n_particles(h))
            out = np.empty(n, dtype=dtype)
            _lib.check(fn(h, out.ctypes.data_as(C.POINTER(ctype))), h)
            outs.append(out)
        return outs

    def __len__(self):
        return int(self.L.pf_mg_n_global(self.handles[0]))

    @property
    def n_steps(self):
        return int(self.L.pf_n_steps(self.handles[0]))

    def local_sizes(self):
        return [int(self.L.pf_n_particles(h)) for h in self.handles]

    def logweights(self):
        return np.concatenate(self.group.gather_host(self._local(self.L.pf_get_logweights, np.float64, C.c_double)))

    def ncomponents(self):
        return np.concatenate(self.group.gather_host(self._local(self.L.pf_get_ncomponents, np.int32, C.c_int32)))

    def last_ancestors(self):
        return np.concatenate(self.group.gather_host(self._local(self.L.pf_get_last_ancestors, np.int32, C.c_int32)))

    def last_assignments(self):
        return np.concatenate(self.group.gather_host(self._local(self.L.pf_get_last_assignments, np.int32, C.c_int32)))

    def assignments(self):
        """assignments(ps) (src/filters.jl:26-34): T x n matrix of 1-based labels, columns in global particle order.
        Each shard follows its particles' ancestor chains through the peers' genealogy arrays on the device."""
        T = self.n_steps
        outs = []
        for h in self.handles:
            out = np.empty((int(self.L.pf_n_particles(h)), T), dtype=np.int32)
            _lib.check(self.L.pf_mg_get_assignments(h, out.ctypes.data_as(C.POINTER(C.c_int32))), h)
            outs.append(out)
        if len(outs) > 1:
            outs = [np.concatenate(outs, axis=0)]
        return np.concatenate(self.group.gather_host(outs), axis=0).T.astype(np.int64) + 1

    # ---- filter-level summaries (src/filters.jl:17-42, 53-80): per-shard sums on the device, added up on the host
    def _partials(self, xs=None, pop=False, sim=False):
        Q = 0 if xs is None else len(xs)
        T = self.n_steps
        pred = np.zeros(Q + 1)
        popv = np.zeros(3 + self.k_max + 1)
        simv = np.zeros((T, T), dtype=np.uint64)
        for h in self.handles:
            p1, p2, p3 = np.zeros(Q + 1), np.zeros(3 + self.k_max + 1), np.zeros((T, T), dtype=np.uint64)
            _lib.check(self.L.pf_mg_summary_partials(h, None if xs is None else xs.ctypes.data_as(_dp), Q,
                                                     p1.ctypes.data_as(_dp) if xs is not None else None,
                                                     p2.ctypes.data_as(_dp) if pop else None,
                                                     p3.ctypes.data_as(C.POINTER(C.c_uint64)) if sim else None), h)
            pred += p1; popv += p2; simv += p3
        parts = self.group.gather_host([(pred, popv, simv)])
        return (sum(p[0] for p in parts), sum(p[1] for p in parts), sum(p[2] for p in parts))

    def posterior_predictive(self, xs):
        xs = np.ascontiguousarray(np.asarray(xs, dtype=np.float64).reshape(-1, self.d))
        pred, _, _ = self._partials(xs=xs)
        return pred[:-1] / pred[-1] / (self.n_steps + self._alpha)

    def state_entropy(self):
        _, pop, _ = self._partials(pop=True)
        return float(pop[2] / pop[0])

    def mean_ncomponents(self):
        _, pop, _ = self._partials(pop=True)
        return float(pop[1] / pop[0])

    def ncomponents_dist(self):
        _, pop, _ = self._partials(pop=True)
        p = pop[3:] / pop[0]
        kmax = int(np.nonzero(p)[0].max()) if p.any() else 0
        return p[1:kmax + 1].copy()

    def assignment_similarity(self):
        _, _, sim = self._partials(sim=True)
        n = len(self)
        return 1.0 - (n - sim.astype(np.float64)) / n

    def last_step(self):
        st = pf_step_stats()
        _lib.check(self.L.pf_get_last_step(self.handles[0], C.byref(st)), self.handles[0])
        return st.asdict()

    def set_profiling(self, on=True):
        for h in self.handles:
            _lib.check(self.L.pf_set_profiling(h, int(on)), h)

    def kernel_times(self):
        """{class name: (device ms, launches)} of this process's first shard since set_profiling."""
        ms = np.zeros(_lib.PF_KERNEL_CLASSES)
        n = np.zeros(_lib.PF_KERNEL_CLASSES, dtype=np.int64)
        _lib.check(self.L.pf_get_kernel_times(self.handles[0], ms.ctypes.data_as(_dp), n.ctypes.data_as(C.POINTER(C.c_int64))),
                   self.handles[0])
        return {self.L.pf_kernel_class_name(k).decode(): (float(ms[k]), int(n[k])) for k in range(_lib.PF_KERNEL_CLASSES)}

    def timing(self):
        """(device ms of the last filter_ call on this process's first shard, its kernel launches)."""
        ms, n = C.c_double(), C.c_int64()
        _lib.check(self.L.pf_last_timing(self.handles[0], C.byref(ms), C.byref(n)), self.handles[0])
        return ms.value, n.value


class ShardedChenLiuParticles(ShardedFearnheadParticles):
    """ChenLiuParticles (src/chenliu.jl:7-26) with the N particles block-sharded over group.size GPUs (rank r owns the
    global particles [r q, (r + 1) q), q = ceil(N / G)).  The coefficient of variation is formed from exact integer
    sums exchanged through peer memory, so every rank takes the 1-GPU run's rejuvenation decisions; a rejuvenating
    child pulls its ancestor's record from whichever rank owns it.  uniforms: T x 2N (the categorical draws, then the
    stratified rejuvenation uniforms), or None for the device stream."""
    _kind = PF_CHENLIU

    def fit_(self, y, u=None):
        return self.filter_(np.atleast_2d(np.asarray(y, dtype=np.float64).reshape(1, -1)), None if u is None else np.atleast_2d(u))

    def __len__(self):
        return self.N
