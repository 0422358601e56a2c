"""Host-side mirror of the Particles.jl API for the accelerated path.

Julia is not available in this image, so the reference-facing host code is this Python
module; the Julia shim with the same surface is particles.jl_b200/julia/device.jl.
Names and argument meaning follow the reference (src/Particles.jl:23-56); Julia's `fit!` /
`filter!` are spelled `fit_` / `filter_` here.  All arithmetic happens in
libparticles_b200.so on the GPU; nothing in this module computes the filter on the CPU.

    ps = FearnheadParticles(100, NormalInverseChisq(0., 1., 0.1, 1.0), ChineseRestaurantProcess(1.))
    filter_(ps, ys)
    assignments(ps)          # T x n matrix of 1-based component labels

Two things the reference does not have, both stated in include/particles_b200.h:
`k_max` (cluster slots per particle; the reference's component vector is unbounded,
src/particle.jl:141-146) -- a step that had to drop a new-cluster candidate because K == k_max is
reported (`on_overflow="warn"` by default, "raise" to make it an error); and `history`
(generations of genealogy kept for `assignments`; the default -1 grows on demand like the
reference's ancestor chain, 0 disables it, n > 0 is a fixed capacity).
"""
from __future__ import annotations

import ctypes as C
import math
import warnings

import numpy as np

from . import _lib
from ._lib import (PF_CHENLIU, PF_FEARNHEAD, PF_NIW, PF_NIX2, PF_ORDER_INDEX, PF_ORDER_SORTED, PFError, pf_config,
                   pf_step_stats)

__all__ = ["NormalInverseChisq", "NormalInverseWishart", "ChineseRestaurantProcess", "StickyCRP", "ChangePoint", "NStatePrior",
           "Labeled", "Component",
           "InfiniteParticle", "ParticleFilter", "FearnheadParticles", "ChenLiuParticles", "fit_", "filter_",
           "particles", "weight", "nobs", "ncomponents", "components", "assignments", "ncomponents_dist",
           "assignment_similarity", "randindex", "state_entropy", "mean", "posterior_predictive", "save", "load", "select", "PF_ORDER_INDEX", "PF_ORDER_SORTED", "PFError"]

_dp = C.POINTER(C.c_double)


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


# ------------------------------------------------------------------ priors (ConjugatePriors.jl)
class NormalInverseChisq:
    """NormalInverseChisq(mu, sigma2, kappa, nu) (re-exported by the reference, src/Particles.jl:29)."""

    def __init__(self, mu=0.0, sigma2=1.0, kappa=1.0, nu=1.0):
        self.mu, self.sigma2, self.kappa, self.nu = float(mu), float(sigma2), float(kappa), float(nu)
        self.dim = 1

    def params(self):
        return self.mu, self.sigma2, self.kappa, self.nu


class NormalInverseWishart:
    """NormalInverseWishart(mu, kappa, Lambda, nu) (ConjugatePriors; used by bench/2d.jl:18)."""

    def __init__(self, mu, kappa, Lam, nu):
        self.mu = np.atleast_1d(np.asarray(mu, dtype=np.float64)).copy()
        self.dim = len(self.mu)
        self.kappa, self.nu = float(kappa), float(nu)
        self.Lam = np.asarray(Lam, dtype=np.float64).reshape(self.dim, self.dim).copy()


class ChineseRestaurantProcess:
    """ChineseRestaurantProcess(alpha) (src/statepriors.jl:47-52)."""

    def __init__(self, alpha, N=None):
        self.alpha = float(alpha)
        self.N = np.zeros(0) if N is None else np.asarray(N, dtype=np.float64)


class StickyCRP:
    """StickyCRP(alpha, rho) (src/statepriors.jl:93-145): with probability rho the state repeats the previous one,
    otherwise it is drawn from a CRP.  N counts the non-sticky occupations, Nsame the transitions to the same
    state, `last` is the previous component (1-based here, as in the reference)."""

    def __init__(self, alpha, rho, last=1, N=None, Nsame=None):
        self.alpha, self.rho, self.last = float(alpha), float(rho), int(last)
        self.N = np.zeros(0) if N is None else np.asarray(N, dtype=np.float64)
        self.Nsame = np.zeros(0) if Nsame is None else np.asarray(Nsame, dtype=np.float64)


class ChangePoint:
    """ChangePoint(p) (src/statepriors.jl:162-190): the state either stays (log(1-p)) or moves to a new one (log p)."""

    def __init__(self, p, k=0, n=0):
        if not 0.0 <= float(p) <= 1.0:
            raise ValueError(f"ChangePoint probability p must be between 0 and 1 (got p={p})")        # ArgumentError, :169
        self.p, self.logp, self.k, self.n = float(p), (math.log(p) if p > 0 else -math.inf), int(k), int(n)


class NStatePrior:
    """NStatePrior(counts) (src/statepriors.jl:203-221): a fixed number of states; the particles start with one empty
    component per state (test/labeled.jl:13-18)."""

    def __init__(self, counts):
        self.counts = np.asarray(counts, dtype=np.float64).copy()
        self.total_count, self.n = float(self.counts.sum()), len(self.counts)


class Labeled:
    """Labeled(obs, label) (src/labeled.jl:1-8): label = 1-based component, or None for missing."""

    def __init__(self, obs, label=None):
        self.obs, self.label = obs, (None if label is None else int(label))


class Component:
    """A cluster: prior + posterior statistics read back from the device store
    (src/component.jl:1-4).  `n` = nobs, `mean` = sample mean, `chol` = lower Cholesky factor
    of the posterior scale Lambda_n, from which the reference's suffstats follow."""

    def __init__(self, prior, n=0, mean=None, chol=None, logdet=None):
        self.prior, self.n = prior, int(n)
        self.mean, self.chol, self.logdet = mean, chol, logdet

    def nobs(self):
        return self.n

    def isempty(self):
        return self.n == 0

    def params(self):
        """params(posterior_canon(prior, suffstats)) (src/component.jl:28)."""
        p = self.prior
        if self.n == 0:
            return p.params() if isinstance(p, NormalInverseChisq) else (p.mu, p.Lam, p.kappa, p.nu)
        kn, nn = p.kappa + self.n, p.nu + self.n
        if isinstance(p, NormalInverseChisq):
            mun = (p.kappa * p.mu + self.n * self.mean[0]) / kn
            return mun, self.chol[0, 0] ** 2 / nn, kn, nn
        mun = (p.kappa * p.mu + self.n * self.mean) / kn
        return mun, self.chol @ self.chol.T, kn, nn


class InfiniteParticle:
    """Lazy view of one particle of a device-resident population (src/particle.jl:38-45)."""

    def __init__(self, ps, i, logweight, K):
        self._ps, self._i = ps, i
        self.logweight = float(logweight)
        self.K = int(K)
        self._comps = None

    @property
    def weight(self):
        return math.exp(self.logweight)

    @property
    def components(self):
        if self._comps is None:
            self._comps = self._ps._read_particle(self._i)
        return self._comps

    @property
    def stateprior(self):
        sp = self._ps.stateprior
        if isinstance(sp, StickyCRP):
            N, Nsame, last = self._ps._read_stateprior(self._i, self.K)
            return StickyCRP(sp.alpha, sp.rho, last + 1, N, Nsame)
        if isinstance(sp, ChangePoint):
            return ChangePoint(sp.p, self.K, self._ps.n_steps)
        if isinstance(sp, NStatePrior):
            return NStatePrior(sp.counts + np.array([c.n for c in self.components], dtype=np.float64))
        return ChineseRestaurantProcess(sp.alpha, [c.n for c in self.components])


# ------------------------------------------------------------------ filters
class ParticleFilter:
    """abstract type ParticleFilter (src/filters.jl:4) over a device-resident population."""

    _kind = None

    def __init__(self, n, prior, stateprior, *, k_max=32, order=PF_ORDER_INDEX, history=-1, device=0, seed=0,
                 rejuv=50.0, on_overflow="warn", precision="f64"):
        if isinstance(prior, tuple):
            prior = NormalInverseChisq(*prior)                      # Component(params::NTuple), src/component.jl:8
        if not isinstance(stateprior, (ChineseRestaurantProcess, StickyCRP, ChangePoint, NStatePrior)):
            raise ValueError("stateprior must be a ChineseRestaurantProcess, StickyCRP, ChangePoint or NStatePrior")
        self.N, self.prior, self.stateprior = int(n), prior, stateprior
        self.k_max, self.order, self.history = int(k_max), order, int(history)
        self.rejuvination_threshold = float(rejuv)
        if on_overflow not in ("warn", "raise", "ignore"):
            raise ValueError("on_overflow must be 'warn', 'raise' or 'ignore'")
        self.on_overflow, self._overflow_seen = on_overflow, 0
        L = _lib.load()
        cfg = pf_config()
        cfg.filter_kind = self._kind
        if precision not in ("f64", "f32"):
            raise ValueError("precision must be 'f64' or 'f32'")
        self.precision = precision
        cfg.precision = _lib.PF_F32 if precision == "f32" else _lib.PF_F64      # storage type of the particle store
        cfg.n_particles = self.N
        cfg.k_max = self.k_max
        cfg.order = order
        cfg.history = self.history
        cfg.device = device
        cfg.alpha = getattr(stateprior, "alpha", 1.0)
        if isinstance(stateprior, StickyCRP):
            cfg.stateprior_kind, cfg.sp_param = _lib.PF_SP_STICKY, stateprior.rho
        elif isinstance(stateprior, ChangePoint):
            cfg.stateprior_kind, cfg.sp_param = _lib.PF_SP_CHANGEPOINT, stateprior.p
        elif isinstance(stateprior, NStatePrior):
            self._spc = stateprior.counts.copy()
            cfg.stateprior_kind, cfg.sp_n, cfg.sp_counts = _lib.PF_SP_NSTATE, stateprior.n, self._spc.ctypes.data_as(_dp)
        cfg.rejuv_threshold = self.rejuvination_threshold
        cfg.seed = seed
        if isinstance(prior, NormalInverseChisq):
            cfg.prior_kind, cfg.d = PF_NIX2, 1
            self._mu0 = np.array([prior.mu])
            self._lam0 = np.zeros((1, 1))
            cfg.kappa0, cfg.nu0, cfg.sigma2_0 = prior.kappa, prior.nu, prior.sigma2
        elif isinstance(prior, NormalInverseWishart):
            cfg.prior_kind, cfg.d = PF_NIW, prior.dim
            self._mu0 = prior.mu.copy()
            self._lam0 = np.asfortranarray(prior.Lam)
            cfg.kappa0, cfg.nu0, cfg.sigma2_0 = prior.kappa, prior.nu, 0.0
        else:
            raise ValueError("prior must be NormalInverseChisq or NormalInverseWishart")
        self.d = cfg.d
        cfg.mu0 = self._mu0.ctypes.data_as(_dp)
        cfg.lambda0 = self._lam0.ctypes.data_as(_dp)
        h = C.c_void_p()
        _lib.check(L.pf_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self._L = L

    def close(self):
        if getattr(self, "_h", None):
            self._L.pf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- stepping
    def fit_(self, y, uniforms=None):
        """fit!(ps, y) (src/fearnhead.jl:130, src/chenliu.jl:54); y may be Labeled(obs, label) (src/labeled.jl)."""
        if isinstance(y, Labeled):
            return self.filter_([y], uniforms=None if uniforms is None else np.atleast_2d(uniforms))
        y, py = _d(np.atleast_1d(y))
        if len(y) != self.d:
            raise ValueError(f"observation has {len(y)} entries, the prior has dimension {self.d}")
        if uniforms is None:
            _lib.check(self._L.pf_step(self._h, py, None, 0), self._h)
        else:
            u, pu = _d(np.atleast_1d(uniforms))
            _lib.check(self._L.pf_step(self._h, py, pu, len(u)), self._h)
        self._check_overflow()
        return self

    def _check_overflow(self):
        """The reference always proposes a new component (src/particle.jl:141-146); the store has k_max
        slots, and a particle at K == k_max loses that candidate.  Say so when it happens."""
        if self.on_overflow == "ignore":
            return
        s = pf_step_stats()
        if self._L.pf_get_last_step(self._h, C.byref(s)) != 0:
            return
        if s.overflow > self._overflow_seen:
            msg = (f"{s.overflow - self._overflow_seen} new-cluster candidates were dropped because a particle already "
                   f"had k_max = {self.k_max} components (the reference's component vector is unbounded): "
                   "results differ from the reference from this step on; construct the filter with a larger k_max")
            self._overflow_seen = s.overflow
            if self.on_overflow == "raise":
                raise PFError(_lib.PF_ERR_OVERFLOW, msg)
            warnings.warn(msg, RuntimeWarning, stacklevel=3)

    def filter_(self, ys, progress=False, cb=None, uniforms=None):
        """filter!(ps, ys, progress; cb) (src/filters.jl:6-13).  Without a callback the whole
        stream is one library call; with one, the callback sees the filter after every y."""
        labels = None
        if len(ys) and isinstance(ys[0], Labeled):
            # Labeled.(ys, labels): 1-based labels, None = missing  ->  0-based / -1 across the ABI
            labels = np.array([-1 if y.label is None else y.label - 1 for y in ys], dtype=np.int32)
            ys = [y.obs for y in ys]
        ys = np.asarray(ys, dtype=np.float64)
        if ys.ndim == 1 and self.d == 1:
            ys = ys.reshape(-1, 1)
        if ys.ndim != 2 or ys.shape[1] != self.d:
            raise ValueError(f"ys must be T x {self.d}")
        T = ys.shape[0]
        if uniforms is not None:
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(T, -1)
        if cb is not None:
            for t in range(T):
                yt = ys[t] if labels is None else Labeled(ys[t], None if labels[t] < 0 else int(labels[t]) + 1)
                self.fit_(yt, None if uniforms is None else uniforms[t])
                cb(self, ys[t])
            return self
        ysc, pys = _d(ys)                    # row t = observation t == column-major d x T
        le = np.empty(T)
        if labels is not None:
            pu, ups = (None, 0) if uniforms is None else (uniforms.ctypes.data_as(_dp), uniforms.shape[1])
            _lib.check(self._L.pf_filter_labeled(self._h, pys, labels.ctypes.data_as(C.POINTER(C.c_int32)), T, pu, ups,
                                                 le.ctypes.data_as(_dp)), self._h)
        elif uniforms is None:
            _lib.check(self._L.pf_filter(self._h, pys, T, None, 0, le.ctypes.data_as(_dp)), self._h)
        else:
            _lib.check(self._L.pf_filter(self._h, pys, T, uniforms.ctypes.data_as(_dp), uniforms.shape[1],
                                         le.ctypes.data_as(_dp)), self._h)
        self.log_evidence = le
        self._check_overflow()
        return self

    # --- read-out
    def __len__(self):
        return int(self._L.pf_n_particles(self._h))

    @property
    def n_steps(self):
        return int(self._L.pf_n_steps(self._h))

    def logweights(self):
        out = np.empty(len(self))
        _lib.check(self._L.pf_get_logweights(self._h, out.ctypes.data_as(_dp)), self._h)
        return out

    def weights(self):
        return np.exp(self.logweights())

    def ncomponents(self):
        out = np.empty(len(self), dtype=np.int32)
        _lib.check(self._L.pf_get_ncomponents(self._h, out.ctypes.data_as(C.POINTER(C.c_int32))), self._h)
        return out

    def _read_particle(self, i):
        d, km = self.d, self.k_max
        K = C.c_int32()
        counts, means = np.empty(km), np.empty((km, d))
        chol, logdet = np.empty((km, d, d)), np.empty(km)
        _lib.check(self._L.pf_get_particle(self._h, i, C.byref(K), counts.ctypes.data_as(_dp),
                                           means.ctypes.data_as(_dp), chol.ctypes.data_as(_dp),
                                           logdet.ctypes.data_as(_dp)), self._h)
        return [Component(self.prior, counts[j], means[j].copy(), chol[j].T.copy(), logdet[j])
                for j in range(K.value)]

    def _read_stateprior(self, i, K):
        N, Nsame = np.zeros(max(K, 1)), np.zeros(max(K, 1))
        last = C.c_int32()
        _lib.check(self._L.pf_get_stateprior(self._h, i, N.ctypes.data_as(_dp), Nsame.ctypes.data_as(_dp), C.byref(last)), self._h)
        return N[:K].copy(), Nsame[:K].copy(), last.value

    def particles(self):
        """particles(ps) (src/filters.jl:15)."""
        lw, K = self.logweights(), self.ncomponents()
        return [InfiniteParticle(self, i, lw[i], K[i]) for i in range(len(self))]

    def assignments(self):
        """assignments(ps) (src/filters.jl:26-34): T x n matrix, 1-based labels."""
        T, n = self.n_steps, len(self)
        out = np.empty((n, T), dtype=np.int32)
        _lib.check(self._L.pf_get_assignments(self._h, out.ctypes.data_as(C.POINTER(C.c_int32))), self._h)
        return out.T.astype(np.int64) + 1

    # --- filter-level summaries, reduced on the device (src/filters.jl:17-42, 53-80)
    def mean(self, f):
        """mean(f, ps) (src/filters.jl:17-18).  ncomponents and state_entropy are reduced on the device; any other f
        is a host function of InfiniteParticle views, as in the reference."""
        what = {ncomponents: _lib.PF_MEAN_NCOMPONENTS, state_entropy: _lib.PF_MEAN_STATE_ENTROPY}.get(f)
        if what is None:
            ps, w = self.particles(), self.weights()
            return float(np.dot([f(p) for p in ps], w) / w.sum())
        out = C.c_double()
        _lib.check(self._L.pf_weighted_mean(self._h, what, C.byref(out)), self._h)
        return out.value

    def ncomponents_dist(self):
        out = np.zeros(self.k_max + 1)
        _lib.check(self._L.pf_ncomponents_dist(self._h, out.ctypes.data_as(_dp)), self._h)
        kmax = int(np.nonzero(out)[0].max()) if out.any() else 0
        return out[1:kmax + 1].copy()

    def posterior_predictive(self, xs):
        xs = np.asarray(xs, dtype=np.float64)
        xs = xs.reshape(-1, 1) if (xs.ndim == 1 and self.d == 1) else np.atleast_2d(xs)
        if xs.shape[1] != self.d:
            raise ValueError(f"query points must be Q x {self.d}")
        xs, px = _d(xs)
        out = np.empty(len(xs))
        _lib.check(self._L.pf_posterior_predictive(self._h, px, len(xs), out.ctypes.data_as(_dp)), self._h)
        return out

    def assignment_similarity(self):
        T = self.n_steps
        out = np.empty((T, T))
        _lib.check(self._L.pf_assignment_similarity(self._h, out.ctypes.data_as(_dp)), self._h)
        return out

    def randindex(self, truth, kind="adjusted"):
        t = np.ascontiguousarray(truth, dtype=np.int32)
        out = C.c_double()
        rc = self._L.pf_randindex(self._h, t.ctypes.data_as(C.POINTER(C.c_int32)), len(t), int(kind == "adjusted"), C.byref(out))
        _lib.check(rc, self._h)                           # a length mismatch is the reference's DimensionMismatch
        return out.value

    # --- checkpoint / restore
    def save(self, path):
        _lib.check(self._L.pf_save(self._h, str(path).encode()), self._h)

    @classmethod
    def load(cls, path, device=0):
        """Rebuild a filter from pf_save's file; it continues bit for bit where the saved one stopped."""
        L = _lib.load()
        h = C.c_void_p()
        _lib.check(L.pf_load(str(path).encode(), device, C.byref(h)))
        cfg = pf_config()
        _lib.check(L.pf_get_config(h, C.byref(cfg)), h)
        self = object.__new__(FearnheadParticles if cfg.filter_kind == PF_FEARNHEAD else ChenLiuParticles)
        self._h, self._L = h, L
        self.N, self.k_max, self.order, self.history, self.d = cfg.n_particles, cfg.k_max, cfg.order, cfg.history, cfg.d if cfg.prior_kind == PF_NIW else 1
        self.rejuvination_threshold = cfg.rejuv_threshold
        self.precision = "f32" if cfg.precision == _lib.PF_F32 else "f64"
        self.on_overflow, self._overflow_seen = "warn", 0
        mu0, lam0, spc = np.empty(self.d), np.empty((self.d, self.d)), np.zeros(64)
        _lib.check(L.pf_get_prior(h, mu0.ctypes.data_as(_dp), lam0.ctypes.data_as(_dp), spc.ctypes.data_as(_dp)), h)
        self.prior = (NormalInverseChisq(mu0[0], cfg.sigma2_0, cfg.kappa0, cfg.nu0) if cfg.prior_kind == PF_NIX2
                      else NormalInverseWishart(mu0, cfg.kappa0, lam0.T, cfg.nu0))
        k = cfg.stateprior_kind
        self.stateprior = (StickyCRP(cfg.alpha, cfg.sp_param) if k == _lib.PF_SP_STICKY else
                           ChangePoint(cfg.sp_param) if k == _lib.PF_SP_CHANGEPOINT else
                           NStatePrior(spc[:cfg.sp_n]) if k == _lib.PF_SP_NSTATE else ChineseRestaurantProcess(cfg.alpha))
        return self

    def particle_assignments(self, i):
        """assignments(particles(ps)[i]) (src/particle.jl:25-32): the 1-based labels along particle i's ancestor chain."""
        out = np.empty(self.n_steps, dtype=np.int32)
        _lib.check(self._L.pf_get_particle_assignments(self._h, int(i), out.ctypes.data_as(C.POINTER(C.c_int32))), self._h)
        return out.astype(np.int64) + 1

    def genealogy_info(self):
        """{generations, nodes, capacity, prunes} of the genealogy log (pf_genealogy_info)."""
        out = (C.c_int64 * 4)()
        _lib.check(self._L.pf_genealogy_info(self._h, out), self._h)
        return dict(zip(("generations", "nodes", "capacity", "prunes"), (int(x) for x in out)))

    def last_step(self):
        s = pf_step_stats()
        _lib.check(self._L.pf_get_last_step(self._h, C.byref(s)), self._h)
        return s.asdict()

    def putative_weights(self):
        M = C.c_int64()
        cap = len(self) * (self.k_max + 2) + self.N * (self.k_max + 2)
        out = np.empty(cap)
        _lib.check(self._L.pf_get_putative_weights(self._h, out.ctypes.data_as(_dp), cap, C.byref(M)), self._h)
        return out[:M.value].copy()

    def last_ancestors(self):
        out = np.empty(len(self), dtype=np.int32)
        _lib.check(self._L.pf_get_last_ancestors(self._h, out.ctypes.data_as(C.POINTER(C.c_int32))), self._h)
        return out

    def last_assignments(self):
        out = np.empty(len(self), dtype=np.int32)
        _lib.check(self._L.pf_get_last_assignments(self._h, out.ctypes.data_as(C.POINTER(C.c_int32))), self._h)
        return out

    def set_profiling(self, on=True):
        _lib.check(self._L.pf_set_profiling(self._h, int(on)), self._h)

    def kernel_times(self):
        """{class name: (device ms, launches)} accumulated since set_profiling."""
        ms = np.zeros(_lib.PF_KERNEL_CLASSES)
        n = np.zeros(_lib.PF_KERNEL_CLASSES, dtype=np.int64)
        _lib.check(self._L.pf_get_kernel_times(self._h, ms.ctypes.data_as(_dp), n.ctypes.data_as(C.POINTER(C.c_int64))), self._h)
        return {self._L.pf_kernel_class_name(k).decode(): (float(ms[k]), int(n[k])) for k in range(_lib.PF_KERNEL_CLASSES)}

    def timing(self):
        ms, n = C.c_double(), C.c_int64()
        _lib.check(self._L.pf_last_timing(self._h, C.byref(ms), C.byref(n)), self._h)
        return ms.value, n.value


class FearnheadParticles(ParticleFilter):
    """FearnheadParticles(n, prior, stateprior) (src/fearnhead.jl:13-30)."""
    _kind = PF_FEARNHEAD


class ChenLiuParticles(ParticleFilter):
    """ChenLiuParticles(n, prior, stateprior; rejuv=50.) (src/chenliu.jl:7-26)."""
    _kind = PF_CHENLIU


# ------------------------------------------------------------------ generic functions
def fit_(ps, y, uniforms=None):
    return ps.fit_(y, uniforms)


def filter_(ps, ys, progress=False, cb=None, uniforms=None):
    return ps.filter_(ys, progress, cb, uniforms)


def particles(ps):
    return ps.particles()


def weight(p):
    return p.weight


def components(p):
    return p.components


def ncomponents(p):
    return p.K


def nobs(x):
    return x.nobs() if isinstance(x, Component) else sum(c.n for c in x.components)


def assignments(ps):
    return ps.assignments()


def ncomponents_dist(ps):
    """ncomponents_dist(ps) (src/filters.jl:36-37): weighted distribution of K; entry k-1 = P(K = k), k = 1..max K
    (fit(Categorical, ...) has max(K) categories).  Reduced on the device."""
    return ps.ncomponents_dist()


def state_entropy(x):
    """state_entropy(ps) = mean(state_entropy, ps) (src/filters.jl:20) on the device; state_entropy(p) for one particle
    (src/particle.jl:162) from its state prior."""
    if isinstance(x, ParticleFilter):
        return x.mean(state_entropy)
    sp = x.stateprior
    if isinstance(sp, ChangePoint):
        lp = np.zeros(1) if sp.k == 0 else np.array([math.log1p(-sp.p) if sp.p < 1 else -math.inf, sp.logp])
    elif isinstance(sp, NStatePrior):
        lp = np.log(sp.counts) - math.log(sp.total_count)
    elif isinstance(sp, StickyCRP):
        lp = (np.zeros(1) if len(sp.N) == 0 else
              np.concatenate([[math.log(sp.rho / (1 - sp.rho)) + math.log(sp.N.sum() + sp.alpha)], np.log(sp.N), [math.log(sp.alpha)]]))
    else:
        lp = np.concatenate([np.log(sp.N), [math.log(sp.alpha)]])
    p = np.exp(lp - np.logaddexp.reduce(lp))
    p = p[p > 0]
    return float(-(p * np.log(p)).sum())


def mean(f, ps):
    """mean(f, ps) (src/filters.jl:17-18)."""
    return ps.mean(f)


def posterior_predictive(ps, xs):
    """pdf.(posterior_predictive(ps), xs) (src/filters.jl:22-24), evaluated on the device."""
    return ps.posterior_predictive(xs)


def assignment_similarity(ps):
    """assignment_similarity(ps) (src/filters.jl:39-42): 1 - Hamming distance between rows (device)."""
    return ps.assignment_similarity()


def randindex(ps, truth, kind="adjusted"):
    """randindex(ps, truth, type) (src/filters.jl:53-80) on the device-computed similarity counts."""
    return ps.randindex(truth, kind)


def save(ps, path):
    return ps.save(path)


def load(path, device=0):
    return ParticleFilter.load(path, device)


def select(w, N, u, order=PF_ORDER_INDEX, device=0):
    """Fearnhead-Clifford selection (cutoff_ascending + sample_stratified!, src/fearnhead.jl:59-114)
    on given weights, on the GPU.  Returns (survivors, resampled_flag, stats)."""
    L = _lib.load()
    w, pw = _d(w)
    surv = np.empty(max(int(N), 1) + 1, dtype=np.int64)
    flag = np.empty(max(int(N), 1) + 1, dtype=np.uint8)
    st = pf_step_stats()
    _lib.check(L.pf_select(device, pw, len(w), int(N), float(u), order, surv.ctypes.data_as(C.POINTER(C.c_int64)),
                           flag.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(st)))
    n = st.n_out
    return surv[:n].copy(), flag[:n].astype(bool), st.asdict()
