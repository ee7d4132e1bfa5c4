"""ctypes binding of include/mgvi_b200.h.

This is the same binding a Julia `ccall` makes (julia/MGVIB200.jl); plain pointers and sizes,
no torch types.  The product path loads `libmgvib200.so` (sm_100a CUDA) and raises when it is
missing -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmgvib200.so")
TRACE_MAX = 64

c_double_p = C.POINTER(C.c_double)
c_int64_p = C.POINTER(C.c_int64)


class ModelDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("ndim", C.c_int32), ("dims", C.c_int64 * 3),
        ("amplitude", c_double_p), ("scale", C.c_double), ("offset", C.c_double), ("link", C.c_int32),
        ("likelihood", C.c_int32), ("sigma", C.c_double), ("lambda_layout", C.c_int32),
        ("data", c_double_p), ("n_data", C.c_int64),
        ("n_groups", C.c_int32), ("group_sizes", c_int64_p), ("x", c_double_p), ("n_coef", C.c_int32),
        ("coef_scale", c_double_p), ("noise_scale", C.c_double),
        ("J", c_double_p), ("m", C.c_int64), ("k", C.c_int64),
        ("sigma_vec", c_double_p),
        ("kdist", c_double_p), ("hyper", C.c_double * 4), ("bin_start", C.c_int64), ("grain", C.c_int32),
        ("binsize", C.c_double),
    ]


class NewtonCGOpts(C.Structure):
    _fields_ = [("alpha", C.c_double), ("steps", C.c_int64), ("i0", C.c_int64), ("i1", C.c_int64)]


class NewtonCGTrace(C.Structure):
    _fields_ = [
        ("steps", C.c_int64), ("f_history", C.c_double * (TRACE_MAX + 1)), ("n_cg_iterations", C.c_int64),
        ("cg_iterations", C.c_int64 * (TRACE_MAX + 1)), ("cg_updates", C.c_int64 * TRACE_MAX),
        ("step_lengths", C.c_double * TRACE_MAX), ("f_calls", C.c_int64), ("g_calls", C.c_int64),
        ("initial_f", C.c_double), ("minimum", C.c_double),
    ]

    def as_dict(self):
        s = int(self.steps)
        return dict(f_history=list(self.f_history[: s + 1]),
                    cg_iterations=list(self.cg_iterations[: int(self.n_cg_iterations)]),
                    cg_updates=list(self.cg_updates[:s]), step_lengths=list(self.step_lengths[:s]),
                    f_calls=int(self.f_calls), g_calls=int(self.g_calls), initial_f=float(self.initial_f),
                    minimum=float(self.minimum))


# int (*mgvib200_allgather_fn)(void* user, const double* send, double* recv, int64_t count)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)


PROBE_LS_MAX = 96


class NewtonProbe(C.Structure):
    """mgvib200_newton_probe (include/mgvi_b200.h)."""
    _fields_ = [
        ("step", C.c_int64), ("f_prev", C.c_double), ("df_n", C.c_double), ("force_k", C.c_int64),
        ("max_rec", C.c_int64),
        ("grad", c_double_p), ("wbar", c_double_p), ("dx_iters", c_double_p), ("f_k", c_double_p),
        ("cg_resid", c_double_p), ("x_new", c_double_p),
        ("k_done", C.c_int64), ("k_own_exit", C.c_int64), ("dphi0", C.c_double), ("beta", C.c_double),
        ("f_new", C.c_double), ("n_ls", C.c_int64), ("ls_kind", C.c_int32 * PROBE_LS_MAX),
        ("ls_a", C.c_double * PROBE_LS_MAX), ("ls_val", C.c_double * PROBE_LS_MAX),
    ]


class CGOpts(C.Structure):
    _fields_ = [("abstol", C.c_double), ("reltol", C.c_double), ("maxiters", C.c_int64)]


EXPORTS = {
    # name: (restype, argtypes)
    "mgvib200_init": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "mgvib200_destroy": (None, [C.c_void_p]),
    "mgvib200_last_error": (C.c_char_p, [C.c_void_p]),
    "mgvib200_version": (C.c_char_p, []),
    "mgvib200_comm_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "mgvib200_comm_init": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_uint8)]),
    "mgvib200_comm_init_external": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mgvib200_model_create": (C.c_int, [C.c_void_p, C.POINTER(ModelDesc), C.POINTER(C.c_void_p)]),
    "mgvib200_model_destroy": (None, [C.c_void_p]),
    "mgvib200_model_n_theta": (C.c_int64, [C.c_void_p]),
    "mgvib200_model_n_lambda": (C.c_int64, [C.c_void_p]),
    "mgvib200_linearize": (C.c_int, [C.c_void_p, c_double_p]),
    "mgvib200_get_weights": (C.c_int, [C.c_void_p, c_double_p, c_double_p]),
    "mgvib200_metric_apply": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p]),
    "mgvib200_sample_residuals_cg": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int64, C.POINTER(CGOpts),
                                               c_double_p, c_int64_p, c_double_p]),
    "mgvib200_sample_residuals_dense": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p, c_double_p]),
    "mgvib200_kl_value_grad": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int64, c_double_p, c_double_p]),
    "mgvib200_mean_metric_apply": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int64, c_double_p, c_double_p]),
    "mgvib200_newtoncg": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int64, C.POINTER(NewtonCGOpts),
                                    c_double_p, c_double_p, C.POINTER(NewtonCGTrace)]),
    "mgvib200_objective_bind": (C.c_int, [C.c_void_p, c_double_p, C.c_int64]),
    "mgvib200_objective_eval": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p]),
    "mgvib200_newton_probe": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int64, C.POINTER(NewtonCGOpts),
                                        C.POINTER(NewtonProbe)]),
    "mgvib200_ncg_probe": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int64, C.c_int64, c_double_p, c_double_p,
                                     c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "mgvib200_build_samples": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int64, c_double_p]),
    "mgvib200_step": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, C.c_int64, C.c_int,
                                C.POINTER(CGOpts), C.POINTER(NewtonCGOpts), c_double_p, c_double_p, c_double_p,
                                c_int64_p, C.POINTER(NewtonCGTrace)]),
    "mgvib200_sample": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_double_p, C.c_int64, C.c_int,
                                  C.POINTER(CGOpts), c_double_p, c_int64_p]),
    "mgvib200_dev_alloc": (C.c_void_p, [C.c_void_p, C.c_int64]),
    "mgvib200_dev_free": (None, [C.c_void_p, C.c_void_p]),
    "mgvib200_memcpy_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "mgvib200_memcpy_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "mgvib200_linearize_dev": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mgvib200_metric_apply_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "mgvib200_sample_residuals_cg_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(CGOpts),
                                                   C.c_void_p, c_int64_p, c_double_p]),
    "mgvib200_step_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int,
                                    C.POINTER(CGOpts), C.POINTER(NewtonCGOpts), C.c_void_p, C.c_void_p, c_double_p,
                                    c_int64_p, C.POINTER(NewtonCGTrace)]),
    "mgvib200_launch_count": (C.c_int64, [C.c_void_p, C.c_int]),
    "mgvib200_last_timing": (C.c_int, [C.c_void_p, c_double_p, c_double_p, c_int64_p]),
    "mgvib200_profile_begin": (C.c_int, [C.c_void_p]),
    "mgvib200_profile_end": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int32), c_double_p, c_int64_p]),
    "mgvib200_stream": (C.c_void_p, [C.c_void_p]),
}


class MGVIB200Error(RuntimeError):
    pass


class Library:
    """A loaded libmgvib200 with typed entry points."""

    def __init__(self, path: str | None = None):
        path = path or LIB_PATH
        if not os.path.exists(path):
            raise MGVIB200Error(
                "CUDA library %s is missing: build it with `python __graft_entry__.py` "
                "(nvcc, sm_100a). There is no CPU fallback." % path)
        self.path = path
        self.dll = C.CDLL(path, mode=C.RTLD_GLOBAL)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(self.dll, name)   # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
            setattr(self, name[len("mgvib200_"):], fn)


_default_lib: Library | None = None


def default_library() -> Library:
    global _default_lib
    if _default_lib is None:
        # MGVIB200_LIB selects another build of the SAME CUDA sources (tuning experiments)
        _default_lib = Library(os.environ.get("MGVIB200_LIB") or None)
    return _default_lib


def as_f64(a, shape=None) -> np.ndarray:
    """Column-major float64 view/copy suitable for the ABI (leading dimension = rows)."""
    a = np.asarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return np.asfortranarray(a)


def ptr(a: np.ndarray):
    return a.ctypes.data_as(c_double_p)
