"""ctypes loader of the C-ABI library (include/janus_b200.h).  No CPU fallback: if the CUDA
library is missing or there is no GPU, calls fail loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JB200_LIB", os.path.join(_HERE, "libjanus_b200.so"))   # override: developer A/B builds

NYMAX = 16
NSTATS = 8


class jb200_opts(C.Structure):
    """struct jb200_opts (include/janus_b200.h) == OptionsTe, src/cpp/radaute.hpp:52-96."""
    _fields_ = [("RelTol", C.c_double * NYMAX), ("AbsTol", C.c_double * NYMAX), ("n_tol", C.c_int),
                ("InitialStep", C.c_double), ("MaxStep", C.c_double), ("JacRecompute", C.c_double),
                ("Start_Newt", C.c_int), ("MaxNbrNewton", C.c_int), ("NbrStg", C.c_int),
                ("MinNbrStg", C.c_int), ("MaxNbrStg", C.c_int), ("MaxNbrStep", C.c_int64),
                ("Safe", C.c_double), ("Quot1", C.c_double), ("Quot2", C.c_double), ("FacL", C.c_double),
                ("FacR", C.c_double), ("Vitu", C.c_double), ("Vitd", C.c_double), ("hhou", C.c_double),
                ("hhod", C.c_double), ("Gustafsson", C.c_int), ("costate_sign", C.c_int),
                ("radaute_quirks", C.c_int)]


class JanusB200Error(RuntimeError):
    pass


_lib = None
_plugins = {}

# every symbol include/janus_b200.h declares
SYMBOLS = ("jb200_opts_default", "jb200_create", "jb200_destroy", "jb200_ny", "jb200_solve", "jb200_solve_dense",
           "jb200_solve_host", "jb200_solve_host_compact", "jb200_set_events", "jb200_solve_out", "jb200_state_create", "jb200_state_step", "jb200_state_read", "jb200_state_destroy",
           "jb200_last_launches", "jb200_eval_dyns", "jb200_eval_jac", "jb200_eval_jac_dual", "jb200_lu_solve", "jb200_dopri853_step", "jb200_dopri853_step_dual", "jb200_seulex_step", "jb200_seulex_step_dual", "jb200_seulex_init", "jb200_solve_sens", "jb200_solve_sens_par",
           "jb200_det_pow", "jb200_fp64_peak", "jb200_dmma_peak", "jb200_last_error", "jb200_version")


def lib(path=None):
    """the C-ABI library; `path` selects a plugin build around a user Hamiltonian (same ABI)."""
    global _lib
    if path is not None:
        if path not in _plugins:
            _plugins[path] = _load(path)
        return _plugins[path]
    if _lib is None:
        _lib = _load(LIB_PATH)
    return _lib


def _load(path):
    if True:
        if not os.path.exists(path):
            raise JanusB200Error(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(janus_ode_b200 has no CPU fallback)")
        L = C.CDLL(path)
        vp, dp, i64 = C.c_void_p, C.c_void_p, C.c_int64
        L.jb200_opts_default.argtypes = [C.POINTER(jb200_opts)]
        L.jb200_create.argtypes = [C.POINTER(jb200_opts), C.c_int, C.POINTER(C.c_double), C.c_int,
                                   C.POINTER(C.c_void_p)]
        L.jb200_destroy.argtypes = [vp]
        L.jb200_destroy.restype = None
        L.jb200_ny.argtypes = [vp]
        L.jb200_solve.argtypes = [vp, dp, dp, i64, dp, dp, dp, dp, dp, dp, dp, vp]
        L.jb200_solve_dense.argtypes = [vp, dp, dp, C.c_int, i64, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, vp]
        L.jb200_solve_host.argtypes = [vp, dp, dp, i64, dp, dp, dp, dp, dp, dp, dp]
        L.jb200_solve_host_compact.argtypes = [vp, dp, dp, i64, dp, dp, dp, dp, dp, dp, dp]
        L.jb200_set_events.argtypes = [vp, C.c_int, dp, dp, dp, dp]
        L.jb200_solve_out.argtypes = [vp, dp, dp, i64, dp, dp, dp, dp, dp, dp, dp, C.c_int, C.c_int, dp, dp, dp, C.c_int, dp, dp,
                                      dp, dp, vp]
        L.jb200_state_create.argtypes = [vp, dp, dp, i64, C.POINTER(C.c_void_p), vp]
        L.jb200_state_step.argtypes = [vp, dp, vp]
        L.jb200_state_read.argtypes = [vp, dp, dp, dp, dp, dp, dp, dp, dp, vp]
        L.jb200_state_destroy.argtypes = [vp]
        L.jb200_state_destroy.restype = None
        L.jb200_last_launches.argtypes = [vp]
        L.jb200_eval_dyns.argtypes = [vp, dp, i64, dp, vp]
        L.jb200_eval_jac.argtypes = [vp, dp, i64, dp, vp]
        L.jb200_lu_solve.argtypes = [C.c_int, i64, dp, dp, dp, dp, dp, dp, vp]
        L.jb200_solve_sens.argtypes = [vp, dp, dp, dp, i64, dp, dp, dp, dp, dp, dp, dp, dp, dp, vp]
        L.jb200_solve_sens_par.argtypes = [vp, dp, dp, dp, C.c_int32, vp, i64, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp, vp]
        L.jb200_dopri853_step.argtypes = [vp, i64, i64, C.c_double, C.c_double, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, vp]
        L.jb200_dopri853_step_dual.argtypes = [vp, i64, i64, C.c_double, C.c_double, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, dp, C.c_int32, vp]
        L.jb200_eval_jac_dual.argtypes = [vp, dp, dp, C.c_int, i64, dp, dp, dp, dp, dp, vp]
        L.jb200_seulex_step.argtypes = [vp, i64, C.c_double, C.c_double, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, dp,
                                        dp, dp, vp]
        L.jb200_seulex_step_dual.argtypes = [vp, i64, C.c_double, C.c_double, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp,
                                             dp, dp, dp, dp, dp, C.c_int32, vp]
        L.jb200_seulex_init.argtypes = [C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int32)]
        L.jb200_det_pow.argtypes = [dp, dp, i64, dp, vp]
        L.jb200_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.jb200_dmma_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
        L.jb200_last_error.restype = C.c_char_p
        L.jb200_version.restype = C.c_char_p
        return L


def check(rc, L=None):
    if rc != 0:
        msg = (L or lib()).jb200_last_error().decode()
        if rc == -1:
            raise ValueError(msg)
        raise JanusB200Error(f"janus_b200 error {rc}: {msg}")


def check_h(handle, rc):
    check(rc, handle.L)
