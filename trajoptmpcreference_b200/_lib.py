"""ctypes binding of the per-robot CUDA libraries (include/b2t.h).  There is NO CPU fallback: if the library cannot be
built or loaded, or no CUDA device is present, every entry point raises."""
import ctypes
import os
import re
from ctypes import POINTER, c_char_p, c_double, c_int, c_longlong, c_size_t, c_void_p

from . import build as _build
from .model import extract_model, builtin_urdf, model_digest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# ---- mirrors of include/b2t.h
METHOD_N, METHOD_S, METHOD_PCG_J, METHOD_PCG_BJ, METHOD_PCG_SS = 0, 1, 2, 3, 4
F64, F32 = 0, 1
COST_QUADRATIC, COST_URDF_EE = 0, 1
LIMIT_NONE, LIMIT_QUADRATIC_PENALTY, LIMIT_AUGMENTED_LAGRANGIAN, LIMIT_ACTIVE_SET = 0, 1, 2, 3
STATUS_FIELDS, SCALAR_FIELDS, TRACE_FIELDS, KERNEL_FAMILIES = 8, 4, 12, 9
KERNEL_FAMILY_NAMES = ["fd", "fd_grad", "kkt", "schur", "pcg", "recover", "trial_fd", "merit", "ctrl"]
ARR = {"x": 0, "u": 1, "xkp1": 2, "dqdd": 3, "Ghat": 4, "g": 5, "Sd": 6, "So": 7, "Pd": 8, "gamma": 9, "l": 10, "dz": 11, "xn": 12, "un": 13,
       "cost_value": 14, "cost_grad": 15, "cost_hess": 16, "cost_err": 17, "kkt_hess": 18, "AB": 19, "soft_value": 20, "soft_grad": 21, "nu_trace": 22,
       "cost_jtot": 23, "plant_terms": 24}
HOOK_LINSYS, HOOK_STEP = 1, 2
ITERATION_HOOK = ctypes.CFUNCTYPE(c_int, c_void_p, c_int, c_int)


class ProblemDesc(ctypes.Structure):
    _fields_ = [("batch", c_int), ("knots", c_int), ("integrator_type", c_int), ("dtype", c_int), ("dt", c_double), ("gravity", c_double),
                ("cost_kind", c_int), ("qf_start", c_int), ("Q", POINTER(c_double)), ("QF", POINTER(c_double)), ("R", POINTER(c_double)),
                ("limit_mode", c_int * 3), ("lower", POINTER(c_double)), ("upper", POINTER(c_double)),
                ("mu_init", c_double * 3), ("mu_factor", c_double * 3), ("mu_max", c_double * 3), ("phi_init", c_double * 3), ("phi_factor", c_double * 3),
                ("hess_mode", c_int)]


class Options(ctypes.Structure):
    _fields_ = [("exit_tolerance_linSys", c_double), ("max_iter_linSys", c_int), ("exit_tolerance_SQP", c_double), ("max_iter_SQP", c_int),
                ("alpha_factor", c_double), ("alpha_min", c_double), ("rho_factor", c_double), ("rho_min", c_double), ("rho_max", c_double),
                ("rho_init", c_double), ("expected_reduction_min", c_double), ("expected_reduction_max", c_double),
                ("exit_tolerance_soft", c_double), ("max_iter_soft", c_int), ("merit_mu", c_double)]


_DP = POINTER(c_double)
_IP = POINTER(c_int)
_SIGNATURES = {
    "b2t_abi_version": (c_int, []),
    "b2t_model_name": (c_char_p, []),
    "b2t_model_digest": (c_char_p, []),
    "b2t_model_dims": (c_int, [_IP, _IP, _IP]),
    "b2t_last_error": (c_char_p, []),
    "b2t_default_options": (None, [POINTER(Options)]),
    "b2t_solver_create": (c_int, [POINTER(ProblemDesc), c_int, POINTER(c_void_p)]),
    "b2t_solver_destroy": (c_int, [c_void_p]),
    "b2t_workspace_bytes": (c_size_t, [c_void_p]),
    "b2t_set_stream": (c_int, [c_void_p, c_void_p]),
    "b2t_set_trajectory": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "b2t_set_goals": (c_int, [c_void_p, c_void_p, c_int]),
    "b2t_set_initial_state": (c_int, [c_void_p, c_void_p]),
    "b2t_set_multipliers": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2t_reset_multipliers": (c_int, [c_void_p]),
    "b2t_sqp_solve": (c_int, [c_void_p, c_int, POINTER(Options)]),
    "b2t_set_iteration_hook": (c_int, [c_void_p, ITERATION_HOOK, c_void_p]),
    "b2t_ilqr_solve": (c_int, [c_void_p, POINTER(Options)]),
    "b2t_mpc_shift": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2t_get_trajectory": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "b2t_get_status": (c_int, [c_void_p, c_void_p]),
    "b2t_get_scalars": (c_int, [c_void_p, c_void_p]),
    "b2t_get_trace": (c_int, [c_void_p, c_void_p, c_int]),
    "b2t_get_multipliers": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2t_get_launch_stats": (c_int, [c_void_p, POINTER(c_longlong), _DP]),
    "b2t_get_pass_trace": (c_int, [c_void_p, POINTER(c_int), c_int, POINTER(c_int)]),
    "b2t_pcg_kernel_name": (ctypes.c_char_p, [c_void_p]),
    "b2t_set_profiling": (c_int, [c_void_p, c_int]),
    "b2t_get_kernel_times": (c_int, [c_void_p, _DP, POINTER(c_longlong)]),
    "b2t_sqp_solve_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, POINTER(Options), c_void_p, c_void_p, c_void_p]),
    "b2t_stage_dynamics": (c_int, [c_void_p]),
    "b2t_stage_kkt": (c_int, [c_void_p, c_double, c_int]),
    "b2t_stage_pcg": (c_int, [c_void_p, c_int, c_double, c_int, c_void_p]),
    "b2t_stage_recover": (c_int, [c_void_p]),
    "b2t_set_block_system": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2t_stage_precond": (c_int, [c_void_p, c_int]),
    "b2t_stage_merit": (c_int, [c_void_p, c_double, c_void_p, c_void_p, c_void_p]),
    "b2t_fetch": (c_int, [c_void_p, c_int, c_void_p]),
    "b2t_measure_fma_peak": (c_int, [c_int, c_int, _DP]),
}


def declared_symbols():
    """Every function name declared in include/b2t.h (used by the symbol-export test)."""
    with open(os.path.join(ROOT, "include", "b2t.h")) as f:
        txt = f.read()
    return sorted(set(re.findall(r"\b(b2t_[a-z_0-9]+)\s*\(", txt)) - {"b2t_solver"})


class B2TError(RuntimeError):
    pass


_LIBS = {}


def load_library(model: dict, tag: str):
    """Build (if stale) and load libb2t_<tag>.so for `model`; binds every symbol of include/b2t.h."""
    key = (tag, model_digest(model))
    if key in _LIBS:
        return _LIBS[key]
    path = _build.build_model(model, tag)
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.b2t_abi_version() != 2:
        raise B2TError("ABI version mismatch in %s" % path)
    if lib.b2t_model_digest().decode() != model_digest(model):
        raise B2TError("stale library %s: model digest mismatch" % path)
    lib._path = path
    _LIBS[key] = lib
    return lib


def load_builtin(name: str):
    return load_library(extract_model(builtin_urdf(name)), name)


def check(lib, code):
    if code != 0:
        raise B2TError("b2t error %d: %s" % (code, lib.b2t_last_error().decode()))
