"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE, not the product).

Loads oracle/_build/libparticles_oracle.so (built from particles_oracle.c by
oracle/Makefile) and adds one thing the C file does not have: `select_exact`, the
selection step of Fearnhead's filter (src/fearnhead.jl:59-77, 97-114 of the reference)
evaluated in EXACT rational arithmetic on the FP64 weights (Python integers).  The C file
is the literal floating-point restatement; the exact variant is what a rounding-free
evaluation of the same formulas gives, and is the specification the CUDA fixed-point
selection is checked against bit-for-bit.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from fractions import Fraction

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libparticles_oracle.so")

NIX2, NIW = 0, 1
FEARNHEAD, CHENLIU = 0, 1
ORDER_SORTED, ORDER_INDEX = 0, 1
SP_CRP, SP_STICKY, SP_CHANGEPOINT, SP_NSTATE = 0, 1, 2, 3

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "particles_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL,
                              stderr=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_double, C.c_double,
                                 C.c_double, _dp, C.c_double, C.c_double, C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_fh_step.argtypes = [C.c_void_p, _dp, C.c_double]
        L.orc_fh_step.restype = C.c_int
        L.orc_fh_expand.argtypes = [C.c_void_p, _dp]
        L.orc_fh_expand.restype = C.c_int
        L.orc_fh_commit.argtypes = [C.c_void_p, _ip, _dp, C.c_int]
        L.orc_fh_step_labeled.argtypes = [C.c_void_p, _dp, C.c_double, C.c_int]
        L.orc_fh_step_labeled.restype = C.c_int
        L.orc_set_stateprior.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, _dp, C.c_int]
        L.orc_set_stateprior.restype = C.c_int
        L.orc_get_sticky.argtypes = [C.c_void_p, C.c_int, _dp]
        L.orc_get_sticky.restype = C.c_int
        L.orc_sp_log_prior.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_sp_log_prior.restype = C.c_double
        L.orc_sp_candidates.argtypes = [C.c_void_p, C.c_int, _ip]
        L.orc_sp_candidates.restype = C.c_int
        L.orc_cl_step.argtypes = [C.c_void_p, _dp, _dp, _dp]
        L.orc_cl_step.restype = C.c_int
        for name in ("orc_n_particles", "orc_n_steps", "orc_n_putatives", "orc_resampled",
                     "orc_n_kept", "orc_n_resamp_req", "orc_n_resamp_got"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = C.c_int
        for name in ("orc_total_w", "orc_w_resamp", "orc_cv"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = C.c_double
        L.orc_get_weights.argtypes = [C.c_void_p, _dp]
        L.orc_get_ncomponents.argtypes = [C.c_void_p, _ip]
        L.orc_get_last_ancestors.argtypes = [C.c_void_p, _ip]
        L.orc_get_last_assignments.argtypes = [C.c_void_p, _ip]
        L.orc_get_putative_weights.argtypes = [C.c_void_p, _dp]
        L.orc_get_putative_parents.argtypes = [C.c_void_p, _ip, _ip]
        L.orc_get_particle.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.orc_get_particle.restype = C.c_int
        L.orc_get_assignments.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.orc_get_assignments.restype = C.c_int
        L.orc_marginal_log_posterior.argtypes = [C.c_void_p, C.c_int]
        L.orc_marginal_log_posterior.restype = C.c_double
        L.orc_cutoff_ascending.argtypes = [_dp, C.c_long, C.c_long, _dp]
        L.orc_cutoff_ascending.restype = C.c_long
        L.orc_sample_stratified.argtypes = [_dp, C.c_long, C.c_long, C.c_double, _lp]
        L.orc_sample_stratified.restype = C.c_long
        L.orc_sum_pairwise.argtypes = [_dp, C.c_long]
        L.orc_sum_pairwise.restype = C.c_double
        L.orc_coefvar.argtypes = [_dp, C.c_long]
        L.orc_coefvar.restype = C.c_double
        L.orc_stats_add.argtypes = [C.c_int, _dp, _dp, _dp]
        L.orc_stats_sub.argtypes = [C.c_int, _dp, _dp, _dp]
        L.orc_prior_sizeof.restype = C.c_int
        L.orc_make_prior.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, C.c_double, C.c_double,
                                     C.c_double, _dp]
        L.orc_mll.argtypes = [C.c_void_p, _dp]
        L.orc_mll.restype = C.c_double
        L.orc_logpred.argtypes = [C.c_void_p, _dp, _dp]
        L.orc_logpred.restype = C.c_double
        L.orc_posterior_nix2.argtypes = [C.c_void_p, _dp, _dp]
        L.orc_posterior_niw.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp, _dp]
        L.orc_posterior_niw.restype = C.c_int
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def ss_len(d: int) -> int:
    return 1 + 2 * d + d * d


class Prior:
    """NormalInverseChisq(mu, s2, kappa, nu) or NormalInverseWishart(mu, kappa, Lam, nu)."""

    def __init__(self, kind, d, mu0, kappa0, nu0, s20=0.0, Lam0=None):
        self.kind, self.d = kind, d
        self.mu0 = np.atleast_1d(np.asarray(mu0, dtype=np.float64)).copy()
        self.kappa0, self.nu0, self.s20 = float(kappa0), float(nu0), float(s20)
        self.Lam0 = (np.asarray(Lam0, dtype=np.float64).reshape(d, d).copy()
                     if Lam0 is not None else np.zeros((d, d)))
        self._buf = C.create_string_buffer(lib().orc_prior_sizeof())
        lam_f = np.asfortranarray(self.Lam0)
        lib().orc_make_prior(self._buf, kind, d, self.mu0.ctypes.data_as(_dp), self.kappa0,
                             self.nu0, self.s20, lam_f.ctypes.data_as(_dp))

    @staticmethod
    def nix2(mu, s2, kappa, nu):
        return Prior(NIX2, 1, [mu], kappa, nu, s20=s2)

    @staticmethod
    def niw(mu, kappa, Lam, nu):
        mu = np.atleast_1d(np.asarray(mu, dtype=np.float64))
        return Prior(NIW, len(mu), mu, kappa, nu, Lam0=Lam)

    # --- unit-level arithmetic -------------------------------------------------------
    def empty(self):
        return np.zeros(ss_len(self.d))

    def add(self, ss, x):
        ss, pss = _d(ss)
        x, px = _d(np.atleast_1d(x))
        out = np.empty_like(ss)
        lib().orc_stats_add(self.d, pss, px, out.ctypes.data_as(_dp))
        return out

    def sub(self, ss, x):
        ss, pss = _d(ss)
        x, px = _d(np.atleast_1d(x))
        out = np.empty_like(ss)
        lib().orc_stats_sub(self.d, pss, px, out.ctypes.data_as(_dp))
        return out

    def mll(self, ss):
        ss, pss = _d(ss)
        return lib().orc_mll(self._buf, pss)

    def logpred(self, ss, y):
        ss, pss = _d(ss)
        y, py = _d(np.atleast_1d(y))
        return lib().orc_logpred(self._buf, pss, py)

    def posterior(self, ss):
        ss, pss = _d(ss)
        if self.kind == NIX2:
            out = np.empty(4)
            lib().orc_posterior_nix2(self._buf, pss, out.ctypes.data_as(_dp))
            return tuple(out)
        mu = np.empty(self.d)
        U = np.empty((self.d, self.d), order="F")
        kn, nn = C.c_double(), C.c_double()
        rc = lib().orc_posterior_niw(self._buf, pss, mu.ctypes.data_as(_dp), U.ctypes.data_as(_dp),
                                     C.byref(kn), C.byref(nn))
        assert rc == 0
        return mu, np.array(U), kn.value, nn.value


class OracleFilter:
    """FearnheadParticles / ChenLiuParticles restated (see particles_oracle.c)."""

    def __init__(self, filter_kind, N, prior: Prior, alpha, rejuv=50.0, order=ORDER_SORTED,
                 keep_history=True, stateprior=None):
        """stateprior: None = ChineseRestaurantProcess(alpha); ("sticky", alpha, rho) = StickyCRP;
        ("changepoint", p) = ChangePoint(p); ("nstate", counts) = NStatePrior(counts) (src/statepriors.jl)."""
        self.prior, self.N, self.kind = prior, N, filter_kind
        lam_f = np.asfortranarray(prior.Lam0)
        self.h = lib().orc_create(filter_kind, N, prior.kind, prior.d,
                                  prior.mu0.ctypes.data_as(_dp), prior.kappa0, prior.nu0,
                                  prior.s20, lam_f.ctypes.data_as(_dp), float(alpha),
                                  float(rejuv), order, int(keep_history))
        if not self.h:
            raise ValueError("orc_create failed")
        if stateprior is not None:
            kind = stateprior[0]
            if kind == "sticky":
                rc = lib().orc_set_stateprior(self.h, SP_STICKY, float(stateprior[1]), float(stateprior[2]), None, 0)
            elif kind == "changepoint":
                rc = lib().orc_set_stateprior(self.h, SP_CHANGEPOINT, 0.0, float(stateprior[1]), None, 0)
            elif kind == "nstate":
                c, pc = _d(stateprior[1])
                rc = lib().orc_set_stateprior(self.h, SP_NSTATE, 0.0, 0.0, pc, len(c))
            else:
                raise ValueError(f"unknown state prior {kind}")
            if rc != 0:
                raise ValueError("invalid state prior")          # ArgumentError, src/statepriors.jl:169

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    # Fearnhead
    def fh_step(self, y, u):
        y, py = _d(np.atleast_1d(y))
        return lib().orc_fh_step(self.h, py, float(u))

    def fh_step_labeled(self, y, u, label):
        """fit!(ps, Labeled(y, label)) (src/labeled.jl:10-11); label None / 0 = missing."""
        y, py = _d(np.atleast_1d(y))
        return lib().orc_fh_step_labeled(self.h, py, float(u), int(label or 0))

    def sp_log_prior(self, i, x):
        return lib().orc_sp_log_prior(self.h, i, int(x))

    def sp_candidates(self, i):
        first = C.c_int()
        n = lib().orc_sp_candidates(self.h, i, C.byref(first))
        return list(range(first.value, first.value + n))

    def sticky(self, i):
        K = int(self.ncomponents()[i])
        ns = np.zeros(max(K, 1))
        last = lib().orc_get_sticky(self.h, i, ns.ctypes.data_as(_dp))
        return ns[:K], last

    def fh_expand(self, y):
        y, py = _d(np.atleast_1d(y))
        return lib().orc_fh_expand(self.h, py)

    def fh_commit(self, idx, w):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        w, pw = _d(w)
        lib().orc_fh_commit(self.h, idx.ctypes.data_as(_ip), pw, len(idx))

    # Chen-Liu
    def cl_step(self, y, u_prop, u_rejuv):
        y, py = _d(np.atleast_1d(y))
        up, pup = _d(u_prop)
        ur, pur = _d(u_rejuv)
        assert len(up) >= self.n and len(ur) >= self.n
        return lib().orc_cl_step(self.h, py, pup, pur)

    # read-out
    @property
    def n(self):
        return lib().orc_n_particles(self.h)

    @property
    def M(self):
        return lib().orc_n_putatives(self.h)

    @property
    def T(self):
        return lib().orc_n_steps(self.h)

    def stat(self, name):
        return getattr(lib(), "orc_" + name)(self.h)

    def weights(self):
        out = np.empty(self.n)
        lib().orc_get_weights(self.h, out.ctypes.data_as(_dp))
        return out

    def ncomponents(self):
        out = np.empty(self.n, dtype=np.int32)
        lib().orc_get_ncomponents(self.h, out.ctypes.data_as(_ip))
        return out

    def last_ancestors(self):
        out = np.empty(self.n, dtype=np.int32)
        lib().orc_get_last_ancestors(self.h, out.ctypes.data_as(_ip))
        return out

    def last_assignments(self):
        out = np.empty(self.n, dtype=np.int32)
        lib().orc_get_last_assignments(self.h, out.ctypes.data_as(_ip))
        return out

    def putative_weights(self):
        out = np.empty(self.M)
        lib().orc_get_putative_weights(self.h, out.ctypes.data_as(_dp))
        return out

    def putative_parents(self):
        p = np.empty(self.M, dtype=np.int32)
        s = np.empty(self.M, dtype=np.int32)
        lib().orc_get_putative_parents(self.h, p.ctypes.data_as(_ip), s.ctypes.data_as(_ip))
        return p, s

    def particle(self, i):
        K = int(self.ncomponents()[i])
        st = np.empty(max(K, 1) * ss_len(self.prior.d))
        crp = np.empty(max(K, 1))
        lib().orc_get_particle(self.h, i, st.ctypes.data_as(_dp), crp.ctypes.data_as(_dp))
        return st[:K * ss_len(self.prior.d)].reshape(K, -1), crp[:K]

    def assignments(self):
        """T x n matrix of 1-based labels (src/filters.jl:26-34)."""
        out = np.empty((self.n, self.T), dtype=np.int64)
        rc = lib().orc_get_assignments(self.h, out.ctypes.data_as(C.POINTER(C.c_int64)))
        assert rc == 0
        return out.T.copy()

    def marginal_log_posterior(self, i):
        return lib().orc_marginal_log_posterior(self.h, i)


    # ---- filter-level summaries (src/filters.jl:17-42, 53-80), restated on the oracle's population
    def sp_weights(self, i):
        """weights(p) = exp.(log_prior(p.stateprior)) (src/particle.jl:160, src/statepriors.jl:24-29): the normalised
        prior over the candidate states of particle i."""
        lp = np.array([self.sp_log_prior(i, x) for x in self.sp_candidates(i)])
        m = lp.max()
        return np.exp(lp - (m + np.log(np.exp(lp - m).sum())))             # StatsFuns.logsumexp

    def state_entropy(self):
        """state_entropy(ps) = mean(state_entropy, ps) (src/filters.jl:17-20); state_entropy(p) = entropy of weights(p)
        (src/particle.jl:162, src/statepriors.jl:36; StatsBase.entropy = -sum(p log p), 0 log 0 = 0)."""
        w = self.weights()
        H = np.empty(self.n)
        for i in range(self.n):
            p = self.sp_weights(i)
            p = p[p > 0]
            H[i] = -(p * np.log(p)).sum()
        return float((H * w).sum() / w.sum())

    def mean_ncomponents(self):
        """mean(ncomponents, ps) (src/filters.jl:17-18)."""
        w = self.weights()
        return float((self.ncomponents() * w).sum() / w.sum())

    def ncomponents_dist(self):
        """ncomponents_dist(ps) = fit(Categorical, ncomponents, weights) (src/filters.jl:36-37): p[k-1] = weighted share
        of the particles with k components, k = 1..max."""
        K, w = self.ncomponents(), self.weights()
        return np.bincount(K, weights=w, minlength=1)[1:] / w.sum()

    def posterior_predictive(self, xs):
        """pdf(posterior_predictive(ps), x) (src/filters.jl:22-24): mixture over particles (normalised weights) of
        MixtureModel(posterior_predictive.(components(p)), weights(p)) (src/particle.jl:156-172), components(p) =
        [p.components..., p.prior]; the component's predictive is src/component.jl:12-23 (orc_logpred)."""
        xs = np.atleast_2d(np.asarray(xs, dtype=np.float64))
        if self.prior.d == 1 and xs.shape[0] == 1 and xs.shape[1] != 1:
            xs = xs.T
        w = self.weights()
        w = w / w.sum()
        out = np.zeros(len(xs))
        empty = self.prior.empty()
        for i in range(self.n):
            st, _ = self.particle(i)
            pi = self.sp_weights(i)
            assert len(pi) == len(st) + 1
            for q, x in enumerate(xs):
                dens = sum(pi[j] * np.exp(self.prior.logpred(st[j], x)) for j in range(len(st)))
                dens += pi[-1] * np.exp(self.prior.logpred(empty, x))
                out[q] += w[i] * dens
        return out

    def assignment_similarity(self):
        """assignment_similarity(ps) (src/filters.jl:39-42): 1 .- pairwise(Hamming(), as, dims=1) ./ size(as, 2)."""
        a = self.assignments()
        T, n = a.shape
        ham = np.array([[(a[s] != a[t]).sum() for t in range(T)] for s in range(T)], dtype=np.float64)
        return 1.0 - ham / n

    def randindex(self, truth, kind="adjusted"):
        """randindex(ps, truth, type) (src/filters.jl:53-80)."""
        truth = np.asarray(truth)
        ps_sim = self.assignment_similarity()
        truth_sim = (truth[:, None] == truth[None, :]).astype(np.float64)
        if truth_sim.shape != ps_sim.shape:
            raise ValueError("Ground truth and particles have different number of obs!")        # DimensionMismatch, :57-58
        N = ps_sim.size
        nis, njs = ps_sim.sum(), truth_sim.sum()
        both_same = (ps_sim * truth_sim).sum()
        both_diff = N - nis - njs + both_same
        if kind == "adjusted":
            n = len(truth)
            nc = (n * (n ** 2 + 1) - (n + 1) * nis - (n + 1) * njs + 2 * (nis * njs) / n) / (2 * (n - 1))
            return (both_same + both_diff - nc) / (N - nc)
        return (both_same + both_diff) / N


# ---------------------------------------------------------------------------------------
# stand-alone literal helpers

def cutoff_ascending(ws, N):
    """src/fearnhead.jl:59-77 -> (kept_i 1-based, w_resamp)."""
    ws, p = _d(ws)
    w = C.c_double()
    i = lib().orc_cutoff_ascending(p, len(ws), N, C.byref(w))
    if i < 0:
        raise ValueError("weights must be sorted in ascending order")
    return i, w.value


def sample_stratified(w, n, u):
    """src/fearnhead.jl:97-114 -> 0-based positions picked."""
    w, p = _d(w)
    picks = np.empty(max(len(w), 1), dtype=np.int64)
    k = lib().orc_sample_stratified(p, len(w), n, float(u), picks.ctypes.data_as(_lp))
    return picks[:k].copy()


def coefvar(x):
    x, p = _d(x)
    return lib().orc_coefvar(p, len(x))


def select_literal(w, N, u, order=ORDER_SORTED):
    """The selection part of fit!(::FearnheadParticles) (src/fearnhead.jl:146-179) on a
    given weight vector, literal floating point.  Returns (survivors, resampled_flag,
    w_resamp, n_kept): survivors are 0-based putative indices in output order."""
    w = np.asarray(w, dtype=np.float64)
    M = len(w)
    perm = np.argsort(w, kind="stable")
    ws = w[perm]
    kept_i, w_resamp = cutoff_ascending(ws, N)
    n_kept = M - (kept_i - 1)
    n_resamp = N - n_kept
    if order == ORDER_SORTED:
        picks = sample_stratified(ws[:kept_i - 1], n_resamp, u) if kept_i > 1 else np.empty(0, np.int64)
        surv = np.concatenate([perm[picks], perm[kept_i - 1:]])
        flag = np.concatenate([np.ones(len(picks), bool), np.zeros(n_kept, bool)])
    else:
        kappa = ws[kept_i - 1] if kept_i <= M else np.inf
        low = np.nonzero(w < kappa)[0]
        picks = sample_stratified(w[low], n_resamp, u) if len(low) else np.empty(0, np.int64)
        take = np.zeros(M, bool)
        take[low[picks]] = True
        keep = w >= kappa
        surv = np.nonzero(keep | take)[0]
        flag = take[surv]
    return surv.astype(np.int64), flag, w_resamp, n_kept


# ---------------------------------------------------------------------------------------
# exact rational selection

def _as_int(x: float) -> int:
    """x * 2**1074 as an exact Python integer (x finite, >= 0)."""
    m, e = math.frexp(x)
    return int(m * (1 << 53)) << (e - 53 + 1074) if e - 53 + 1074 >= 0 else int(m * (1 << 53)) >> -(e - 53 + 1074)


def select_exact(w, N, u, order=ORDER_INDEX):
    """Fearnhead-Clifford selection in exact arithmetic on the FP64 weights.

    kappa  = smallest non-zero weight w_i with  sum_{w_k < w_i} w_k <= (N - #{w_k >= w_i}) w_i
             (the predicate of cutoff_ascending, src/fearnhead.jl:69; ties decided jointly);
             no such weight -> everything is resampled (kept_i = M+1, src/fearnhead.jl:76).
    low    = {k : w_k < kappa},  n = N - #{w_k >= kappa},  T = sum(low)  (exact)
    comb   = sample_stratified! (src/fearnhead.jl:97-114) with K = T/n and U0 = u K in exact
             arithmetic: walking the low items in `order`, item i is picked iff the number
             of teeth {(u + k) T / n, k = 0,1,..} strictly below the running sum S_i grows.
    Returns (survivors, resampled_flag, (T_num, n), n_kept): survivors in output order
    (ORDER_SORTED: [picked ascending ; kept ascending]; ORDER_INDEX: index order);
    c = T/n is returned as the exact pair (T * 2**1074, n)."""
    w = np.asarray(w, dtype=np.float64)
    M = len(w)
    W = [_as_int(float(x)) for x in w]
    perm = sorted(range(M), key=lambda k: (W[k], k))
    # threshold search over distinct values
    tot = 0
    kept_start = M          # position in perm of the first kept item
    i = 0
    while i < M:
        j = i
        while j < M and W[perm[j]] == W[perm[i]]:
            j += 1
        wi = W[perm[i]]
        if wi != 0:
            n_geq = M - i
            if tot <= (N - n_geq) * wi:
                kept_start = i
                break
            tot += wi * (j - i)
        i = j
    n_kept = M - kept_start
    n = N - n_kept
    low_sorted = perm[:kept_start]
    kept_sorted = perm[kept_start:]
    T = sum(W[k] for k in low_sorted)
    if order == ORDER_SORTED:
        walk = low_sorted
    else:
        lowset = set(low_sorted)
        walk = [k for k in range(M) if k in lowset]
    picked = []
    if n > 0 and T > 0:
        fu = Fraction(u)                          # exact value of the double
        un, ud = fu.numerator, fu.denominator
        # tooth k is below S  <=>  (u + k) T / n < S  <=>  (un + k ud) T < S n ud
        S = 0
        k_next = 0                                 # index of the next tooth not yet passed
        for item in walk:
            S += W[item]
            lhs = S * n * ud
            if (un + k_next * ud) * T < lhs:
                picked.append(item)
                # pass every tooth now below S (at most one in non-degenerate cases)
                k_next = (lhs - un * T + ud * T - 1) // (ud * T)   # smallest k with (un+k ud)T >= lhs
    if order == ORDER_SORTED:
        surv = list(picked) + list(kept_sorted)
        flag = [True] * len(picked) + [False] * n_kept
    else:
        ps = set(picked)
        ks = set(kept_sorted)
        surv = [k for k in range(M) if k in ps or k in ks]
        flag = [k in ps for k in surv]
    return (np.asarray(surv, dtype=np.int64), np.asarray(flag, dtype=bool), (T, n), n_kept)


def exact_c_value(T_n):
    """c = T/n of select_exact as the correctly rounded double."""
    T, n = T_n
    if n <= 0:
        return float("nan")
    return float(Fraction(T, n) / (1 << 1074))


def select_exact_fast(w, N, u, order=ORDER_INDEX):
    """select_exact for large M (millions): the same exact-arithmetic specification, evaluated with a
    numpy sort, C-speed integer prefix sums (itertools.accumulate over Python integers scaled by the
    smallest exponent present) and a binary search over the distinct weights (the predicate of
    src/fearnhead.jl:69 is monotone along the ascending weights).  tests/test_oracle_selection.py pins
    it to select_exact on random and edge cases.  Returns the same tuple, except that the exact c = T/n
    is returned as (T * 2**-emin_shift, n, emin_shift) folded into (T_scaled_to_2**1074, n)."""
    from itertools import accumulate
    w = np.ascontiguousarray(w, dtype=np.float64)
    M = len(w)
    bits = w.view(np.uint64)
    mant, expo = np.frexp(w)
    m53 = np.ldexp(mant, 53).astype(np.int64)                  # exact: w = m53 * 2**(expo - 53)
    e = expo.astype(np.int64) - 53
    nz = m53 != 0
    emin = int(e[nz].min()) if nz.any() else 0
    sh = np.where(nz, e - emin, 0)
    W = [m << s for m, s in zip(m53.tolist(), sh.tolist())]     # w * 2**(-emin), exact integers
    perm = np.argsort(bits, kind="stable")                     # ascending weight, ties by index
    sb = bits[perm]
    Ws = [W[k] for k in perm.tolist()]
    pre = [0] + list(accumulate(Ws))                           # pre[i] = sum of the i smallest
    starts = np.flatnonzero(np.concatenate([[True], sb[1:] != sb[:-1]]))   # first position of each distinct value
    starts = starts[sb[starts] != 0]                           # exact zeros are skipped (src/fearnhead.jl:66)

    def pred(pos):                                             # predicate at the distinct value starting at `pos`
        return pre[pos] <= (N - (M - pos)) * Ws[pos]
    lo, hi = 0, len(starts)                                    # first group whose predicate holds
    while lo < hi:
        mid = (lo + hi) // 2
        if pred(int(starts[mid])):
            hi = mid
        else:
            lo = mid + 1
    kept_start = int(starts[lo]) if lo < len(starts) else M
    n_kept = M - kept_start
    n = N - n_kept
    T = pre[kept_start]
    lowmask = np.zeros(M, dtype=bool)
    lowmask[perm[:kept_start]] = True
    walk = perm[:kept_start] if order == ORDER_SORTED else np.flatnonzero(lowmask)
    picked = np.zeros(0, dtype=np.int64)
    if n > 0 and T > 0:
        fu = Fraction(u)
        un, ud = fu.numerator, fu.denominator
        den = ud * T
        unT = un * T
        nud = n * ud
        # teeth strictly below S: 0 if un T >= S n ud, else floor((S n ud - un T - 1) / (ud T)) + 1
        run = accumulate(W[k] for k in walk.tolist())
        teeth = [0 if (x := s * nud - unT) <= 0 else (x - 1) // den + 1 for s in run]
        t = np.asarray(teeth, dtype=np.int64)
        grow = np.diff(np.concatenate([[0], t])) > 0
        picked = walk[grow]
    if order == ORDER_SORTED:
        surv = np.concatenate([picked, perm[kept_start:]])
        flag = np.concatenate([np.ones(len(picked), bool), np.zeros(n_kept, bool)])
    else:
        take = np.zeros(M, dtype=bool)
        take[picked] = True
        keep = ~lowmask
        surv = np.flatnonzero(keep | take)
        flag = take[surv]
    shift = 1074 + emin                                        # W = w * 2**(-emin)  ->  w * 2**1074
    Tn = (T << shift if shift >= 0 else Fraction(T, 1 << -shift), n)
    return surv.astype(np.int64), flag, Tn, n_kept
