"""Sharded Fearnhead filter: particles block-sharded over GPUs, one process per GPU.

The library runs the same kernels on every shard and exposes the stage boundaries of one
observation (include/particles_b200.h, "multi-GPU"); this module supplies the collectives in
between through a small communicator interface:

  TorchComm   torch.distributed (NCCL over NVLink on the GPU box; gloo in the CPU tests of
              the host-side logic), one local rank per process;
  LocalComm   all ranks inside one process on one device, executed in lock step -- used by
              the single-GPU tests to check "sharded run == unsharded run" bit for bit.

What crosses the ranks per observation: two scalars (all-reduce), the strided sample and the
bracket's candidates (all-gather, a few MB), one 32-byte record per rank (all-gather) and the
records of the particles that change owner in the re-partitioning (all-to-all-v)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import PF_FEARNHEAD, PF_MG_BRACKET, PF_MG_CLASSIFY, PF_MG_SAMPLE, PF_MG_SOLVE, PF_NIW, PF_NIX2, pf_config, pf_mg_ptrs, \
    pf_step_stats

_dp = C.POINTER(C.c_double)


# ------------------------------------------------------------------ host-side plan (pure)
def ownership(n_local_all, stay_cap=0, window=0):
    """New ownership boundaries of a rank-major population with n_local_all[r] children born on
    rank r.  The boundary between ranks r-1 and r moves toward the balanced position r q by at
    most `window` particles and never past a neighbour's range (only adjacent ranks exchange,
    fixed-capacity messages: mode 2); if a share would still exceed `stay_cap`, equal blocks of
    q = ceil(n / G) (mode 1, sizes known to the host only).  Mirrors k_rank_offsets."""
    n_local_all = [int(x) for x in n_local_all]
    G = len(n_local_all)
    first = [0]
    for x in n_local_all:
        first.append(first[-1] + x)
    n = first[-1]
    q = max((n + G - 1) // G, 1)
    bnd = [0] * (G + 1)
    bnd[G] = n
    for r in range(1, G):
        cur, tgt = first[r], min(r * q, n)
        b = min(max(tgt, cur - window), cur + window)
        b = min(max(b, first[r - 1]), first[r] + n_local_all[r])
        bnd[r] = max(b, bnd[r - 1])
    fits = all(bnd[r + 1] - bnd[r] <= stay_cap for r in range(G))
    if not fits:
        bnd = [min(r * q, n) for r in range(G + 1)]
    return bnd, first, q, (2 if fits else 1)


def partition_plan(n_local_all, rank, stay_cap=0, window=0):
    """(send_counts by destination, recv_counts by source, q, n_mine, mode) for `rank`."""
    bnd, first, q, mode = ownership(n_local_all, stay_cap, window)
    G = len(n_local_all)
    send = np.zeros(G, dtype=np.int64)
    recv = np.zeros(G, dtype=np.int64)
    for d in range(G):
        lo, hi = max(first[rank], bnd[d]), min(first[rank + 1], bnd[d + 1])
        send[d] = 0 if d == rank else max(hi - lo, 0)
        lo, hi = max(first[d], bnd[rank]), min(first[d + 1], bnd[rank + 1])
        recv[d] = 0 if d == rank else max(hi - lo, 0)
    return send, recv, q, bnd[rank + 1] - bnd[rank], mode


# ------------------------------------------------------------------ communicators
class TorchComm:
    """torch.distributed default process group; one shard per process."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.size = dist.get_world_size()
        self.local_ranks = [dist.get_rank()]

    def all_reduce(self, tensors, op):
        import torch.distributed as dist
        dist.all_reduce(tensors[0], op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX)

    def all_gather(self, tensors):
        import torch
        t = tensors[0].contiguous().reshape(-1)
        out = torch.empty(self.size * t.numel(), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, t)
        return [out.reshape((self.size,) + tuple(tensors[0].shape))]

    def max_int(self, values, device):
        import torch
        t = torch.tensor([int(values[0])], dtype=torch.int64, device=device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return int(t.item())

    def all_to_all_v(self, sends, send_counts, recvs, recv_counts, rows):
        ins = [int(c) * rows for c in send_counts[0]]
        outs = [int(c) * rows for c in recv_counts[0]]
        self.dist.all_to_all_single(recvs[0][:sum(outs)], sends[0][:sum(ins)], outs, ins)

    def device_barrier(self, device):
        """Stream-ordered barrier: a 1-element all-reduce.  Its completion on this rank implies every
        rank has finished the kernels queued before it (their stores into peer memory included)."""
        import torch
        if getattr(self, "_bar", None) is None:
            self._bar = torch.zeros(1, dtype=torch.int32, device=device)
        self.dist.all_reduce(self._bar)

    def exchange_handles(self, blob):
        out = [None] * self.size
        self.dist.all_gather_object(out, blob)
        return out

    def neighbour_exchange(self, sends, recvs, slot):
        """Fixed-size exchange with the two adjacent ranks: send[0:slot] -> left, send[slot:2 slot]
        -> right; recv[0:slot] <- left, recv[slot:2 slot] <- right."""
        dist, r, ops = self.dist, self.local_ranks[0], []
        if r > 0:
            ops += [dist.P2POp(dist.isend, sends[0][:slot], r - 1), dist.P2POp(dist.irecv, recvs[0][:slot], r - 1)]
        if r < self.size - 1:
            ops += [dist.P2POp(dist.isend, sends[0][slot:2 * slot], r + 1), dist.P2POp(dist.irecv, recvs[0][slot:2 * slot], r + 1)]
        for req in dist.batch_isend_irecv(ops) if ops else []:
            req.wait()

    def gather_host(self, arrays):
        out = [None] * self.size
        self.dist.all_gather_object(out, arrays[0])
        return out


class LocalComm:
    """All `size` ranks live in this process (on one device) and advance in lock step."""

    def __init__(self, size):
        self.size = size
        self.local_ranks = list(range(size))

    def all_reduce(self, tensors, op):
        import torch
        st = torch.stack([t.clone() for t in tensors])
        r = st.sum(0) if op == "sum" else st.max(0).values
        for t in tensors:
            t.copy_(r)

    def all_gather(self, tensors):
        import torch
        g = torch.stack([t.contiguous() for t in tensors])
        return [g] * self.size

    def max_int(self, values, device):
        return max(int(v) for v in values)

    def all_to_all_v(self, sends, send_counts, recvs, recv_counts, rows):
        G = self.size
        soff = [np.concatenate([[0], np.cumsum(send_counts[s])]) * rows for s in range(G)]
        for d in range(G):
            o = 0
            for s in range(G):
                assert int(send_counts[s][d]) == int(recv_counts[d][s]), "inconsistent exchange plan"
                n = int(send_counts[s][d]) * rows
                if n:
                    recvs[d][o:o + n].copy_(sends[s][int(soff[s][d]):int(soff[s][d]) + n])
                o += n

    def device_barrier(self, device):
        pass                                  # one process, one stream: already ordered

    def neighbour_exchange(self, sends, recvs, slot):
        for r in range(self.size):
            if r > 0:
                recvs[r][:slot].copy_(sends[r - 1][slot:2 * slot])
            if r < self.size - 1:
                recvs[r][slot:2 * slot].copy_(sends[r + 1][:slot])

    def gather_host(self, arrays):
        return list(arrays)


# ------------------------------------------------------------------ device views
class _DevArray:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _view(ptr, n, typestr, device):
    import torch
    return torch.as_tensor(_DevArray(ptr, n, typestr), device=device)


class _Shard:
    def __init__(self, L, cfg, device):
        import torch
        self.L = L
        self.rank = int(cfg.rank)
        h = C.c_void_p()
        _lib.check(L.pf_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.device = torch.device("cuda", device)
        _lib.check(L.pf_set_stream(h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), h)
        p = pf_mg_ptrs()
        _lib.check(L.pf_mg_info(h, C.byref(p)), h)
        self.rows = int(p.rows_per_particle)
        self.scalars = _view(p.scalars, 2, "<i8", self.device)
        self.acc = _view(p.acc, p.acc_bytes // 8, "<i8", self.device)
        self.list = _view(p.list, p.list_capacity, "<i8", self.device)          # bit patterns of the doubles
        self.list_capacity = int(p.list_capacity)
        self.total = _view(p.local_total, p.total_bytes // 8, "<i8", self.device)
        self.send = _view(p.send_buffer, p.buffer_capacity, "<f8", self.device)
        self.recv = _view(p.recv_buffer, p.buffer_capacity, "<f8", self.device)
        self.sample_entries, self.cand_entries = int(p.sample_entries), int(p.cand_entries)
        self.window, self.capacity = int(p.window), int(p.local_capacity)
        self.cand_entries0 = self.cand_entries

    def state(self):
        st = (C.c_int64 * 4)()
        _lib.check(self.L.pf_mg_state(self.h, st), self.h)
        return list(st)

    def select(self, stage, rounds, gacc, glist, each, pad_to=-1):
        _lib.check(self.L.pf_mg_select(self.h, stage, rounds,
                                       C.c_void_p(gacc.data_ptr()) if gacc is not None else None,
                                       C.c_void_p(glist.data_ptr()) if glist is not None and glist.numel() else None,
                                       int(each), int(pad_to)), self.h)

    def local_arrays(self):
        out = (C.c_void_p * 8)()
        _lib.check(self.L.pf_mg_local_arrays(self.h, out), self.h)
        return out

    def ipc_handles(self):
        buf = C.create_string_buffer(8 * 64)
        _lib.check(self.L.pf_mg_ipc_handles(self.h, buf), self.h)
        return buf.raw

    def plan(self, G):
        sc, rc = (C.c_int64 * G)(), (C.c_int64 * G)()
        nl, ng, ph, ex = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        _lib.check(self.L.pf_mg_plan(self.h, sc, rc, C.byref(nl), C.byref(ng), C.byref(ph), C.byref(ex)), self.h)
        return np.array(sc[:], dtype=np.int64), np.array(rc[:], dtype=np.int64), nl.value, ng.value, ph.value, ex.value


class ShardedFearnheadParticles:
    """FearnheadParticles (src/fearnhead.jl:13-30) with the population sharded over
    comm.size ranks.  Same results as the unsharded filter, particle for particle."""

    def __init__(self, n, prior, stateprior, comm, *, k_max=16, history=0, devices=None, seed=0, sample_entries=None,
                 cand_entries=None, direct=True):
        from . import NormalInverseChisq, NormalInverseWishart
        import torch
        self.comm, self.N, self.k_max = comm, int(n), int(k_max)
        L = _lib.load()
        self.L = L
        self.shards = []
        devices = devices if devices is not None else [torch.cuda.current_device()] * len(comm.local_ranks)
        for r, dev in zip(comm.local_ranks, devices):
            cfg = pf_config()
            cfg.filter_kind, cfg.precision, cfg.n_particles, cfg.k_max = PF_FEARNHEAD, 0, self.N, self.k_max
            cfg.order, cfg.history, cfg.device, cfg.rank, cfg.nranks = 0, int(history), dev, r, comm.size
            cfg.alpha, cfg.rejuv_threshold, cfg.seed = stateprior.alpha, 50.0, seed
            if isinstance(prior, NormalInverseChisq):
                cfg.prior_kind, cfg.d = PF_NIX2, 1
                self._mu0, self._lam0 = np.array([prior.mu]), np.zeros((1, 1))
                cfg.kappa0, cfg.nu0, cfg.sigma2_0 = prior.kappa, prior.nu, prior.sigma2
            elif isinstance(prior, NormalInverseWishart):
                cfg.prior_kind, cfg.d = PF_NIW, prior.dim
                self._mu0, self._lam0 = prior.mu.copy(), np.asfortranarray(prior.Lam)
                cfg.kappa0, cfg.nu0, cfg.sigma2_0 = prior.kappa, prior.nu, 0.0
            else:
                raise ValueError("prior must be NormalInverseChisq or NormalInverseWishart")
            cfg.mu0 = self._mu0.ctypes.data_as(_dp)
            cfg.lambda0 = self._lam0.ctypes.data_as(_dp)
            self.d = cfg.d
            with torch.cuda.device(dev):
                self.shards.append(_Shard(L, cfg, dev))
        for s in self.shards:             # test hooks: force the fixed gather sizes (exercises the slow path)
            if sample_entries is not None:
                s.sample_entries = int(sample_entries)
            if cand_entries is not None:
                s.cand_entries = s.cand_entries0 = int(cand_entries)
        # direct exchange: map every rank's arrays into every other rank (same process: plain pointers;
        # one process per GPU: CUDA IPC), so that children that change owner are stored remotely
        self.direct = bool(direct) and comm.size > 1
        if self.direct:
            if len(self.shards) == comm.size:           # all ranks in this process
                for s in self.shards:
                    for p in self.shards:
                        if p is not s:
                            _lib.check(L.pf_mg_set_peer(s.h, p.rank, p.local_arrays(), None), s.h)
            else:
                s = self.shards[0]
                blobs = comm.exchange_handles(s.ipc_handles())
                for r, blob in enumerate(blobs):
                    if r != s.rank:
                        _lib.check(L.pf_mg_set_peer(s.h, r, None, blob), s.h)
            for s in self.shards:
                _lib.check(L.pf_mg_set_peer(s.h, s.rank, None, None), s.h)
                _lib.check(L.pf_mg_enable_direct(s.h), s.h)
        self.n_global = 1
        self.steps = 0
        self.host_syncs = 0
        self.bytes_moved = 0
        self.slow_steps = 0
        self.exchanges = 0
        self.halts = []                  # (step, search phase) of every step that needed the host
        self.profile = None              # set to [] to collect (label, cuda event) marks per step

    def close(self):
        for s in self.shards:
            if s.h:
                self.L.pf_destroy(s.h)
                s.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- one observation
    def _gather_exact(self):
        """Slow path: gather the lists at their exact (host-read) length."""
        comm, sh = self.comm, self.shards
        lens = [s.state()[1] for s in sh]
        self.host_syncs += 1
        mx = comm.max_int(lens, sh[0].device)
        for s, ln in zip(sh, lens):
            if ln < mx:
                s.list[ln:mx].fill_(-1)                       # all-ones pattern: outside every bracket
        gl = comm.all_gather([s.list[:mx] for s in sh]) if mx else [None] * len(sh)
        ga = comm.all_gather([s.acc for s in sh])
        self.bytes_moved += 8 * mx * comm.size + 8 * sh[0].acc.numel() * comm.size
        return gl, ga, mx

    def _tail(self):
        """Survivor scan + instantiate, launched speculatively behind the threshold search;
        the one host synchronisation of the step reads the search's outcome and the plan."""
        comm, sh, L = self.comm, self.shards, self.L
        for s in sh:
            _lib.check(L.pf_mg_tiles(s.h), s.h)
        gt = comm.all_gather([s.total for s in sh])
        for s, g in zip(sh, gt):
            _lib.check(L.pf_mg_finish(s.h, C.c_void_p(g.data_ptr())), s.h)
        plans = [s.plan(comm.size) for s in sh]
        self.host_syncs += 1
        return plans

    def _mark(self, label):
        if self.profile is not None:
            import time
            import torch
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.profile.append((label, ev, time.perf_counter()))

    def profile_summary(self):
        """Mean device and host milliseconds between consecutive marks, by label."""
        import torch
        torch.cuda.synchronize()
        dev, host, cnt = {}, {}, {}
        for (l0, e0, t0), (l1, e1, t1) in zip(self.profile[:-1], self.profile[1:]):
            dev[l1] = dev.get(l1, 0.0) + e0.elapsed_time(e1)
            host[l1] = host.get(l1, 0.0) + 1e3 * (t1 - t0)
            cnt[l1] = cnt.get(l1, 0) + 1
        return {k: (dev[k] / cnt[k], host[k] / cnt[k]) for k in dev}

    def _enqueue(self, y, u):
        """Queue every kernel and collective of one observation; no host synchronisation.  If
        the step turns out to need the host (search unfinished / particles change owner), the
        device marks it (k_close) and everything queued behind it is a no-op."""
        comm, sh, L = self.comm, self.shards, self.L
        self._mark("begin")
        y = np.ascontiguousarray(np.atleast_1d(y), dtype=np.float64)
        py = y.ctypes.data_as(_dp)
        Bs, Bc = sh[0].sample_entries, sh[0].cand_entries
        for s in sh:
            _lib.check(L.pf_mg_begin(s.h, py, float(u)), s.h)
            self._mark("expand")
            s.select(PF_MG_SAMPLE, 0, None, None, 0, Bs)
        self._mark("sample")
        ga = comm.all_gather([s.acc for s in sh])
        gl = comm.all_gather([s.list[:Bs] for s in sh])
        self._mark("gather_sample")
        for s, g, a in zip(sh, gl, ga):
            s.select(PF_MG_BRACKET | PF_MG_CLASSIFY, 0, a, g, Bs, Bc)
        self._mark("bracket_classify")
        ga = comm.all_gather([s.acc for s in sh])
        gl = comm.all_gather([s.list[:Bc] for s in sh])
        self._mark("gather_cand")
        for s, g, a in zip(sh, gl, ga):
            s.select(PF_MG_SOLVE, 0, a, g, Bc)
        self._mark("solve")
        self.bytes_moved += 8 * (Bs + Bc + 2 * sh[0].acc.numel()) * comm.size
        for s in sh:
            _lib.check(L.pf_mg_tiles(s.h), s.h)
        gt = comm.all_gather([s.total for s in sh])
        for s, g in zip(sh, gt):
            _lib.check(L.pf_mg_finish(s.h, C.c_void_p(g.data_ptr())), s.h)
            _lib.check(L.pf_mg_close(s.h), s.h)
        self._mark("tail")
        slot = sh[0].window * sh[0].rows
        if self.direct:
            comm.device_barrier(sh[0].device)        # remote stores of the re-partitioned children are complete
        elif slot:                                                                       # boundaries shift by <= window
            comm.neighbour_exchange([s.send for s in sh], [s.recv for s in sh], slot)
            self.bytes_moved += 2 * slot * 8
        for s in sh:
            _lib.check(L.pf_mg_unpack(s.h, C.c_void_p(s.recv.data_ptr())), s.h)     # no-op if the step halted
        self._mark("exchange+unpack")

    def _drain(self):
        """The host synchronisation: (0, ...) if every queued step completed, else (1 + index of
        the step that needs the host, its search phase)."""
        outs = []
        for s in self.shards:
            o = (C.c_int64 * 4)()
            _lib.check(self.L.pf_mg_drain(s.h, o), s.h)
            outs.append(list(o))
        self.host_syncs += 1
        self._mark("drain")
        if not outs[0][0]:
            self.n_global = outs[0][3]
        return outs[0][0], outs[0][1]

    def _complete(self, y, u, step, phase):
        """Finish a halted step through the synchronous stages."""
        comm, sh, L = self.comm, self.shards, self.L
        self.halts.append((int(step), int(phase)))
        if phase == 7:
            # a candidate list outgrew its fixed gather size (typically a plateau of exactly tied weights:
            # the new-cluster putatives of all resampled particles): gather more from now on
            for s in sh:
                s.cand_entries = min(2 * s.cand_entries, s.list_capacity - 1024)
            self._clean_batches = 0
        y = np.ascontiguousarray(np.atleast_1d(y), dtype=np.float64)
        for s in sh:
            _lib.check(L.pf_mg_resume(s.h, int(step)), s.h)
        if phase != 4:
            # rare: bracket verification failed, a finer scale is needed, or a list outgrew its
            # fixed gather size -- finish the search with exact-size gathers, then redo the tail
            self.slow_steps += 1
            for s in sh:
                _lib.check(L.pf_mg_set_obs(s.h, y.ctypes.data_as(_dp), float(u)), s.h)
            rounds = sh[0].state()[3]
            while phase != 4:
                if phase in (3, 7):      # 3: new bracket / scale; 7: a fixed-size gather truncated the list
                    for s in sh:
                        s.select(PF_MG_CLASSIFY, rounds, None, None, 0, -1)
                gl, ga, mx = self._gather_exact()
                for s, g, a in zip(sh, gl, ga):
                    s.select(PF_MG_SOLVE, rounds, a, g, mx)
                st = sh[0].state()
                self.host_syncs += 1
                phase, rounds = st[0], st[3]
                if rounds > 40:
                    raise RuntimeError("threshold search did not converge")
            plans = self._tail()
        else:
            plans = [s.plan(comm.size) for s in sh]
            self.host_syncs += 1
        self.n_global = plans[0][3]
        moving = int(sum(int(p[0].sum()) for p in plans))
        if plans[0][5] == 1:               # some share outgrew its shard: equal re-partition, sizes from the plan
            comm.all_to_all_v([s.send for s in sh], [p[0] for p in plans], [s.recv for s in sh], [p[1] for p in plans], sh[0].rows)
            self.exchanges += 1
            self.bytes_moved += moving * sh[0].rows * 8
        elif self.direct:
            comm.device_barrier(sh[0].device)
        elif sh[0].window:
            slot = sh[0].window * sh[0].rows
            comm.neighbour_exchange([s.send for s in sh], [s.recv for s in sh], slot)
            self.bytes_moved += 2 * slot * 8
        for s in sh:
            _lib.check(L.pf_mg_unpack(s.h, C.c_void_p(s.recv.data_ptr())), s.h)
        self._mark("complete")

    def fit_(self, y, u):
        """fit!(ps, y) (src/fearnhead.jl:130-184) on the sharded population; u = the rand()
        of src/fearnhead.jl:99."""
        return self.filter_([y], [u])

    def filter_(self, ys, uniforms, lookahead=8):
        """filter!(ps, ys, false) (src/filters.jl:6-13): observations are queued `lookahead` at a
        time and the host synchronises once per batch; a step that needs the host (re-bracketing,
        re-partitioning) is completed synchronously and the rest of its batch is queued again."""
        ys = np.asarray(ys, dtype=np.float64).reshape(len(uniforms), -1)
        T, t = len(uniforms), 0
        base = self.steps
        while t < T:
            tend = min(T, t + max(int(lookahead), 1))
            for k in range(t, tend):
                self._enqueue(ys[k], uniforms[k])
            halt, phase = self._drain()
            if not halt:
                # the gather size grown by an overflow decays back once the lists are small again
                self._clean_batches = getattr(self, "_clean_batches", 0) + 1
                if self._clean_batches >= 2:
                    for s in self.shards:
                        s.cand_entries = max(s.cand_entries0, s.cand_entries // 2)
                    self._clean_batches = 0
            if halt:
                th = halt - 1 - base                     # index of the halted observation
                self._complete(ys[th], uniforms[th], halt - 1, phase)
                t = th + 1
            else:
                t = tend
            self.steps = base + t
        return self

    # ---- read-out (concatenated over ranks in global order)
    def _local(self, fn, dtype, ctype):
        outs = []
        for s in self.shards:
            n = int(self.L.pf_n_particles(s.h))
            out = np.empty(n, dtype=dtype)
            _lib.check(fn(s.h, out.ctypes.data_as(C.POINTER(ctype))), s.h)
            outs.append(out)
        return outs

    def __len__(self):
        return int(self.n_global)

    def logweights(self):
        return np.concatenate(self.comm.gather_host(self._local(self.L.pf_get_logweights, np.float64, C.c_double)))

    def ncomponents(self):
        return np.concatenate(self.comm.gather_host(self._local(self.L.pf_get_ncomponents, np.int32, C.c_int32)))

    def last_ancestors(self):
        return np.concatenate(self.comm.gather_host(self._local(self.L.pf_get_last_ancestors, np.int32, C.c_int32)))

    def last_assignments(self):
        return np.concatenate(self.comm.gather_host(self._local(self.L.pf_get_last_assignments, np.int32, C.c_int32)))

    def last_step(self):
        s = self.shards[0]
        st = pf_step_stats()
        import torch
        torch.cuda.synchronize(s.device)
        _lib.check(self.L.pf_get_last_step(s.h, C.byref(st)), s.h)
        return st.asdict()

    def synchronize(self):
        import torch
        for s in self.shards:
            torch.cuda.synchronize(s.device)
