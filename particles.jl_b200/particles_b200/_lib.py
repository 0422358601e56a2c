"""ctypes binding of libparticles_b200.so (include/particles_b200.h).

The shared object is built in-tree by particles.jl_b200/build.py (nvcc, sm_100a).  There is
no fallback: if the library is missing the import fails, and every compute entry point
fails with PF_ERR_CUDA when no CUDA device is usable."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PARTICLES_B200_LIB") or os.path.join(os.path.dirname(_HERE), "libparticles_b200.so")

PF_OK, PF_ERR_ARG, PF_ERR_CUDA, PF_ERR_STATE, PF_ERR_HISTORY, PF_ERR_OVERFLOW, PF_ERR_NCCL = range(7)
PF_FEARNHEAD, PF_CHENLIU = 0, 1
PF_NIX2, PF_NIW = 0, 1
PF_F64, PF_F32 = 0, 1
PF_ORDER_INDEX, PF_ORDER_SORTED = 0, 1
PF_SP_CRP, PF_SP_STICKY, PF_SP_CHANGEPOINT, PF_SP_NSTATE = 0, 1, 2, 3
PF_MEAN_NCOMPONENTS, PF_MEAN_STATE_ENTROPY = 0, 1

_dp = C.POINTER(C.c_double)


class pf_config(C.Structure):
    _fields_ = [("filter_kind", C.c_int32), ("prior_kind", C.c_int32), ("d", C.c_int32), ("precision", C.c_int32),
                ("n_particles", C.c_int64), ("k_max", C.c_int32), ("order", C.c_int32), ("history", C.c_int64),
                ("device", C.c_int32), ("stateprior_kind", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32),
                ("alpha", C.c_double),
                ("rejuv_threshold", C.c_double), ("kappa0", C.c_double), ("nu0", C.c_double),
                ("sigma2_0", C.c_double), ("mu0", _dp), ("lambda0", _dp), ("seed", C.c_uint64),
                ("sp_param", C.c_double), ("sp_counts", _dp), ("sp_n", C.c_int32), ("reserved2", C.c_int32)]


class pf_step_stats(C.Structure):
    _fields_ = [("step", C.c_int64), ("n_in", C.c_int64), ("n_putative", C.c_int64), ("n_kept", C.c_int64),
                ("n_resamp_req", C.c_int64), ("n_resamp_got", C.c_int64), ("n_out", C.c_int64),
                ("overflow", C.c_int64), ("log_total_w", C.c_double), ("log_w_resamp", C.c_double),
                ("threshold", C.c_double), ("cv", C.c_double), ("resampled", C.c_int32),
                ("select_rounds", C.c_int32), ("n_candidates", C.c_int64),
                ("select_phase_us", C.c_double * 4), ("n_plateau", C.c_int64), ("select_status", C.c_int32),
                ("reserved1", C.c_int32)]

    def asdict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_}
        d["select_phase_us"] = list(self.select_phase_us)
        return d


PF_MG_NARRAYS, PF_MG_IPC_BYTES = 17, 64
PF_KERNEL_CLASSES = 11

# every symbol include/particles_b200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("pf_create", C.c_int32, [C.POINTER(pf_config), C.POINTER(C.c_void_p)]),
    ("pf_destroy", C.c_int32, [C.c_void_p]),
    ("pf_last_error", C.c_char_p, [C.c_void_p]),
    ("pf_step", C.c_int32, [C.c_void_p, _dp, _dp, C.c_int64]),
    ("pf_filter", C.c_int32, [C.c_void_p, _dp, C.c_int64, _dp, C.c_int64, _dp]),
    ("pf_filter_labeled", C.c_int32, [C.c_void_p, _dp, C.POINTER(C.c_int32), C.c_int64, _dp, C.c_int64, _dp]),
    ("pf_get_stateprior", C.c_int32, [C.c_void_p, C.c_int64, _dp, _dp, C.POINTER(C.c_int32)]),
    ("pf_n_particles", C.c_int64, [C.c_void_p]),
    ("pf_n_steps", C.c_int64, [C.c_void_p]),
    ("pf_get_logweights", C.c_int32, [C.c_void_p, _dp]),
    ("pf_get_ncomponents", C.c_int32, [C.c_void_p, C.POINTER(C.c_int32)]),
    ("pf_get_particle", C.c_int32, [C.c_void_p, C.c_int64, C.POINTER(C.c_int32), _dp, _dp, _dp, _dp]),
    ("pf_get_assignments", C.c_int32, [C.c_void_p, C.POINTER(C.c_int32)]),
    ("pf_get_particle_assignments", C.c_int32, [C.c_void_p, C.c_int64, C.POINTER(C.c_int32)]),
    ("pf_genealogy_info", C.c_int32, [C.c_void_p, C.POINTER(C.c_int64)]),
    ("pf_get_last_step", C.c_int32, [C.c_void_p, C.POINTER(pf_step_stats)]),
    ("pf_posterior_predictive", C.c_int32, [C.c_void_p, _dp, C.c_int64, _dp]),
    ("pf_weighted_mean", C.c_int32, [C.c_void_p, C.c_int32, _dp]),
    ("pf_ncomponents_dist", C.c_int32, [C.c_void_p, _dp]),
    ("pf_assignment_similarity", C.c_int32, [C.c_void_p, _dp]),
    ("pf_randindex", C.c_int32, [C.c_void_p, C.POINTER(C.c_int32), C.c_int64, C.c_int32, _dp]),
    ("pf_save", C.c_int32, [C.c_void_p, C.c_char_p]),
    ("pf_load", C.c_int32, [C.c_char_p, C.c_int32, C.POINTER(C.c_void_p)]),
    ("pf_get_config", C.c_int32, [C.c_void_p, C.POINTER(pf_config)]),
    ("pf_get_prior", C.c_int32, [C.c_void_p, _dp, _dp, _dp]),
    ("pf_get_putative_weights", C.c_int32, [C.c_void_p, _dp, C.c_int64, C.POINTER(C.c_int64)]),
    ("pf_get_last_ancestors", C.c_int32, [C.c_void_p, C.POINTER(C.c_int32)]),
    ("pf_get_last_assignments", C.c_int32, [C.c_void_p, C.POINTER(C.c_int32)]),
    ("pf_select", C.c_int32, [C.c_int32, _dp, C.c_int64, C.c_int64, C.c_double, C.c_int32, C.POINTER(C.c_int64),
                              C.POINTER(C.c_uint8), C.POINTER(pf_step_stats)]),
    ("pf_set_stream", C.c_int32, [C.c_void_p, C.c_void_p]),
    ("pf_mg_local_arrays", C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p)]),
    ("pf_mg_ipc_handles", C.c_int32, [C.c_void_p, C.c_void_p]),
    ("pf_mg_set_peer", C.c_int32, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.c_void_p]),
    ("pf_mg_connect", C.c_int32, [C.POINTER(C.c_void_p), C.c_int32]),
    ("pf_mg_filter", C.c_int32, [C.POINTER(C.c_void_p), C.c_int32, _dp, C.c_int64, _dp, C.c_int64, _dp]),
    ("pf_mg_n_global", C.c_int64, [C.c_void_p]),
    ("pf_mg_get_assignments", C.c_int32, [C.c_void_p, C.POINTER(C.c_int32)]),
    ("pf_mg_summary_partials", C.c_int32, [C.c_void_p, _dp, C.c_int64, _dp, _dp, C.POINTER(C.c_uint64)]),
    ("pf_mg_create_local", C.c_int32, [C.POINTER(pf_config), C.POINTER(C.c_int32), C.c_int32, C.POINTER(C.c_void_p)]),
    ("pf_debug_dump", C.c_int32, [C.c_void_p, C.POINTER(C.c_int64)]),
    ("pf_last_timing", C.c_int32, [C.c_void_p, _dp, C.POINTER(C.c_int64)]),
    ("pf_set_profiling", C.c_int32, [C.c_void_p, C.c_int32]),
    ("pf_get_kernel_times", C.c_int32, [C.c_void_p, _dp, C.POINTER(C.c_int64)]),
    ("pf_kernel_class_name", C.c_char_p, [C.c_int32]),
    ("pf_stream", C.c_void_p, [C.c_void_p]),
    ("pf_version", C.c_char_p, []),
]

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python particles.jl_b200/build.py` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class PFError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[pf error {code}] {msg}")
        self.code = code


def check(rc, handle=None):
    if rc != PF_OK:
        msg = load().pf_last_error(handle)
        msg = msg.decode() if msg else ""
        if rc == PF_ERR_ARG:
            raise ValueError(f"[pf error {rc}] {msg}")      # ArgumentError in the reference
        raise PFError(rc, msg)
