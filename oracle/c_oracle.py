"""ctypes binding of oracle/c/libradau_oracle.so (the C parity oracle).

TEST INFRASTRUCTURE ONLY -- see oracle/c/radau_oracle.c header.  Never imported by
janus_ode_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "c", "libradau_oracle.so")

NYMAX = 16
PB = dict(harmonic=0, vdp=1, vdp_pmp_smooth=2, vdp_pmp_bang=3, vdp_aug_bang=4, mlp=5)
STAT_NAMES = ("FcnNbr", "JacNbr", "DecompNbr", "SolveNbr", "StepNbr", "AccptNbr",
              "StepRejNbr", "NewtRejNbr")


class OracleOpts(C.Structure):
    _fields_ = [("RelTol", C.c_double * NYMAX), ("AbsTol", C.c_double * NYMAX),
                ("InitialStep", C.c_double), ("JacRecompute", C.c_double),
                ("Start_Newt", C.c_int), ("MaxNbrNewton", C.c_int), ("NbrStg", C.c_int),
                ("MinNbrStg", C.c_int), ("MaxNbrStg", C.c_int), ("MaxNbrStep", C.c_longlong),
                ("Safe", C.c_double), ("Quot1", C.c_double), ("Quot2", C.c_double),
                ("FacL", C.c_double), ("FacR", C.c_double), ("Vitu", C.c_double),
                ("Vitd", C.c_double), ("hhou", C.c_double), ("hhod", C.c_double),
                ("Gustafsson", C.c_int)]


def build(force=False):
    src = os.path.join(_HERE, "c", "radau_oracle.c")
    if force or not os.path.exists(_SO) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)):
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "c"), "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _opts(opts, Ny):
    """opts: oracle.radau_scalar.RadauOptions (or anything with the same attributes)."""
    o = OracleOpts()
    rt = np.broadcast_to(np.asarray(opts.RelTol, dtype=np.float64), (Ny,))
    at = np.broadcast_to(np.asarray(opts.AbsTol, dtype=np.float64), (Ny,))
    for i in range(Ny):
        o.RelTol[i] = rt[i]
        o.AbsTol[i] = at[i]
    for k in ("InitialStep", "JacRecompute", "Safe", "Quot1", "Quot2", "FacL", "FacR", "Vitu",
              "Vitd", "hhou", "hhod"):
        setattr(o, k, float(getattr(opts, k)))
    for k in ("MaxNbrNewton", "NbrStg", "MinNbrStg", "MaxNbrStg", "MaxNbrStep"):
        setattr(o, k, int(getattr(opts, k)))
    o.Start_Newt = int(bool(opts.Start_Newt))
    o.Gustafsson = int(bool(opts.Gustafsson))
    return o


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def set_arith(mode):
    """'reference' (no FMA, libm pow, closed-form derivatives) or 'mirror' (the CUDA path's
    arithmetic contract, bit-identical to the GPU)."""
    assert mode in ("reference", "mirror")
    lib().oracle_set_arith(C.c_int(1 if mode == "mirror" else 0))


Q8_GUS_UNGATED = 1


def set_quirks(q):
    """bit 0 (default on): RadauTe's Gustafsson limiter also on order-change steps (radaute_impl.hpp:1092-1100);
    0 = MATLAB radau.m:1066-1076 gating."""
    lib().oracle_set_quirks(C.c_int(int(q)))


def det_pow(x, y):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.empty_like(x)
    lib().oracle_det_pow(_dp(x), _dp(y), C.c_longlong(x.size), _dp(out))
    return out


def det_tanh(x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lib().oracle_det_tanh(_dp(x), C.c_longlong(x.size), _dp(out))
    return out


def solve(problem, par, sign, y0, tspan, opts, nthreads=0, arith="reference"):
    """y0 [M,Ny], tspan [M,2] -> dict(y, t, h, res, stats[M,8], nstg)."""
    set_arith(arith)
    set_quirks(getattr(opts, "RadauTeQuirks", 1))
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    tspan = np.ascontiguousarray(tspan, dtype=np.float64)
    par = np.ascontiguousarray(par, dtype=np.float64)
    M, Ny = y0.shape
    y = np.empty_like(y0)
    t = np.empty(M)
    h = np.empty(M)
    res = np.empty(M, dtype=np.int32)
    stats = np.empty((M, 8), dtype=np.int64)
    nstg = np.empty(M, dtype=np.int32)
    o = _opts(opts, Ny)
    rc = lib().oracle_radau_solve(
        C.c_int(PB[problem] if isinstance(problem, str) else problem), _dp(par), C.c_int(sign),
        C.c_int(Ny), C.c_longlong(M), _dp(y0), _dp(tspan), C.byref(o), _dp(y), _dp(t), _dp(h),
        res.ctypes.data_as(C.POINTER(C.c_int)), stats.ctypes.data_as(C.POINTER(C.c_longlong)),
        nstg.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(nthreads))
    assert rc == 0
    return dict(y=y, t=t, h=h, res=res, stats=stats, nstg=nstg)


def solve_dense(problem, par, sign, y0, tspan, opts, nthreads=0, arith="reference"):
    """tspan [M, ntspan] with interior output points -> additionally tout [M,ntspan], yout [M,ntspan,Ny], nout."""
    set_arith(arith)
    set_quirks(getattr(opts, "RadauTeQuirks", 1))
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    tspan = np.ascontiguousarray(tspan, dtype=np.float64)
    par = np.ascontiguousarray(par, dtype=np.float64)
    M, Ny = y0.shape
    nts = tspan.shape[1]
    y = np.empty_like(y0)
    t = np.empty(M)
    h = np.empty(M)
    res = np.empty(M, dtype=np.int32)
    stats = np.empty((M, 8), dtype=np.int64)
    nstg = np.empty(M, dtype=np.int32)
    tout = np.zeros((M, nts))
    yout = np.zeros((M, nts, Ny))
    nout = np.zeros(M, dtype=np.int32)
    o = _opts(opts, Ny)
    ip = C.POINTER(C.c_int)
    rc = lib().oracle_radau_solve_dense(
        C.c_int(PB[problem]), _dp(par), C.c_int(sign), C.c_int(Ny), C.c_longlong(M), _dp(y0), _dp(tspan),
        C.c_int(nts), C.byref(o), _dp(y), _dp(t), _dp(h), res.ctypes.data_as(ip),
        stats.ctypes.data_as(C.POINTER(C.c_longlong)), nstg.ctypes.data_as(ip), _dp(tout), _dp(yout),
        nout.ctypes.data_as(ip), C.c_int(nthreads))
    assert rc == 0
    return dict(y=y, t=t, h=h, res=res, stats=stats, nstg=nstg, tout=tout, yout=yout, nout=nout)


def solve_out(problem, par, sign, y0, tspan, opts, cap, refine=1, events=None, ecap=8, nthreads=0, arith="reference"):
    """every accepted step recorded (OutFlag 1 / 2) + located events.  events = (A [K,Ny], b [K], stop [K], slope [K])."""
    set_arith(arith)
    set_quirks(getattr(opts, "RadauTeQuirks", 1))
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    tspan = np.ascontiguousarray(tspan, dtype=np.float64)
    par = np.ascontiguousarray(par, dtype=np.float64)
    M, Ny = y0.shape
    y = np.empty_like(y0)
    t = np.empty(M)
    h = np.empty(M)
    res = np.empty(M, dtype=np.int32)
    stats = np.empty((M, 8), dtype=np.int64)
    nstg = np.empty(M, dtype=np.int32)
    tout = np.zeros((M, max(cap, 1)))
    yout = np.zeros((M, max(cap, 1), Ny))
    nout = np.zeros(M, dtype=np.int32)
    K = 0
    A = b = np.zeros(1)
    stop = slope = np.zeros(1, dtype=np.int32)
    if events is not None:
        A = np.ascontiguousarray(events[0], dtype=np.float64).reshape(-1, Ny)
        K = A.shape[0]
        b = np.ascontiguousarray(events[1], dtype=np.float64).reshape(K)
        stop = np.ascontiguousarray(events[2], dtype=np.int32).reshape(K)
        slope = np.ascontiguousarray(events[3], dtype=np.int32).reshape(K)
    te = np.zeros((M, max(ecap, 1)))
    ye = np.zeros((M, max(ecap, 1), Ny))
    ie = np.zeros((M, max(ecap, 1)), dtype=np.int32)
    nev = np.zeros(M, dtype=np.int32)
    o = _opts(opts, Ny)
    ip = C.POINTER(C.c_int)
    rc = lib().oracle_radau_solve_out(
        C.c_int(PB[problem]), _dp(par), C.c_int(sign), C.c_int(Ny), C.c_longlong(M), _dp(y0), _dp(tspan), C.byref(o), _dp(y),
        _dp(t), _dp(h), res.ctypes.data_as(ip), stats.ctypes.data_as(C.POINTER(C.c_longlong)), nstg.ctypes.data_as(ip),
        C.c_int(refine), C.c_int(cap), _dp(tout), _dp(yout), nout.ctypes.data_as(ip), C.c_int(K), _dp(A), _dp(b),
        stop.ctypes.data_as(ip), slope.ctypes.data_as(ip), C.c_int(ecap), _dp(te), _dp(ye), ie.ctypes.data_as(ip),
        nev.ctypes.data_as(ip), C.c_int(nthreads))
    assert rc == 0
    return dict(y=y, t=t, h=h, res=res, stats=stats, nstg=nstg, tout=tout, yout=yout, nout=nout, te=te, ye=ye, ie=ie, nev=nev)


def trace(problem, par, sign, y0, tspan, opts, cap=100000, arith="reference"):
    """one trajectory, with the accepted step sizes."""
    set_arith(arith)
    set_quirks(getattr(opts, "RadauTeQuirks", 1))
    y0 = np.ascontiguousarray(y0, dtype=np.float64).reshape(-1)
    tspan = np.ascontiguousarray(tspan, dtype=np.float64).reshape(2)
    par = np.ascontiguousarray(par, dtype=np.float64)
    Ny = y0.shape[0]
    y = np.empty(Ny)
    t = C.c_double()
    stats = np.empty(8, dtype=np.int64)
    rec = np.empty(cap)
    n = C.c_int()
    o = _opts(opts, Ny)
    res = lib().oracle_radau_trace(
        C.c_int(PB[problem] if isinstance(problem, str) else problem), _dp(par), C.c_int(sign),
        C.c_int(Ny), _dp(y0), _dp(tspan), C.byref(o), _dp(y), C.byref(t),
        stats.ctypes.data_as(C.POINTER(C.c_longlong)), _dp(rec), C.c_int(cap), C.byref(n))
    return dict(y=y, t=t.value, res=res, stats=stats, h_accept=rec[:n.value].copy())


def trace_passes(problem, par, sign, y0, tspan, opts, cap=100000, arith="reference"):
    """one trajectory; per attempted step [StepNbr, t, h, NbrStg, Newt, Theta, err (-1 = Newton failed), FacConv]."""
    set_arith(arith)
    set_quirks(getattr(opts, "RadauTeQuirks", 1))
    y0 = np.ascontiguousarray(y0, dtype=np.float64).reshape(-1)
    tspan = np.ascontiguousarray(tspan, dtype=np.float64).reshape(2)
    par = np.ascontiguousarray(par, dtype=np.float64)
    Ny = y0.shape[0]
    y = np.empty(Ny)
    stats = np.empty(8, dtype=np.int64)
    log = np.zeros((cap, 8))
    n = C.c_int()
    o = _opts(opts, Ny)
    res = lib().oracle_radau_trace2(
        C.c_int(PB[problem] if isinstance(problem, str) else problem), _dp(par), C.c_int(sign),
        C.c_int(Ny), _dp(y0), _dp(tspan), C.byref(o), _dp(y), stats.ctypes.data_as(C.POINTER(C.c_longlong)),
        _dp(log), C.c_int(cap), C.byref(n))
    return dict(y=y, res=res, stats=stats, log=log[:n.value].copy())


def eval_dyns(problem, par, sign, y, arith="reference"):
    set_arith(arith)
    y = np.ascontiguousarray(y, dtype=np.float64)
    par = np.ascontiguousarray(par, dtype=np.float64)
    M, Ny = y.shape
    f = np.empty_like(y)
    lib().oracle_eval_dyns(C.c_int(PB[problem]), _dp(par), C.c_int(sign), C.c_int(Ny),
                           C.c_longlong(M), _dp(y), _dp(f))
    return f


def eval_jac(problem, par, sign, y, arith="reference"):
    set_arith(arith)
    y = np.ascontiguousarray(y, dtype=np.float64)
    par = np.ascontiguousarray(par, dtype=np.float64)
    M, Ny = y.shape
    J = np.empty((M, Ny, Ny))
    lib().oracle_eval_jac(C.c_int(PB[problem]), _dp(par), C.c_int(sign), C.c_int(Ny),
                          C.c_longlong(M), _dp(y), _dp(J))
    return J


def lu_solve(A, b, arith="reference"):
    """A [M,n,n] real or complex, b [M,n] -> x (zgetf2-rule LU with reciprocal pivots)."""
    set_arith(arith)
    A = np.asarray(A)
    b = np.asarray(b)
    M, n, _ = A.shape
    if np.iscomplexobj(A) or np.iscomplexobj(b):
        Ar = np.ascontiguousarray(A.real, dtype=np.float64)
        Ai = np.ascontiguousarray(A.imag, dtype=np.float64)
        br = np.ascontiguousarray(b.real, dtype=np.float64)
        bi = np.ascontiguousarray(b.imag, dtype=np.float64)
        xr = np.empty_like(br)
        xi = np.empty_like(bi)
        lib().oracle_lu_solve(C.c_int(n), C.c_longlong(M), _dp(Ar), _dp(Ai), _dp(br), _dp(bi),
                              _dp(xr), _dp(xi))
        return xr + 1j * xi
    Ar = np.ascontiguousarray(A, dtype=np.float64)
    br = np.ascontiguousarray(b, dtype=np.float64)
    xr = np.empty_like(br)
    lib().oracle_lu_solve(C.c_int(n), C.c_longlong(M), _dp(Ar), None, _dp(br), None, _dp(xr), None)
    return xr


class SeulexOracleState:
    """controller state of one batch for seulex_step (mirrors the fields of the reference's SeulexState,
    /root/reference/src/python/seulext.py:8-70 that survive between step() calls)."""

    def __init__(self, y, x, atol, rtol, KMAXX=13):
        self.y = np.ascontiguousarray(y, dtype=np.float64).copy()
        self.M, self.D = self.y.shape
        self.x = np.ascontiguousarray(np.broadcast_to(np.asarray(x, dtype=np.float64), (self.M,))).copy()
        self.atol, self.rtol, self.KMAXX = float(atol), float(rtol), int(KMAXX)
        th, kt = C.c_double(), C.c_int()
        lib().oracle_seulex_init(C.c_double(self.atol), C.c_double(self.rtol), C.c_int(self.KMAXX), C.byref(th), C.byref(kt))
        self.theta = np.full(self.M, th.value)
        self.k_targ = np.full(self.M, kt.value, dtype=np.int32)
        self.hnext = np.full(self.M, -1.0e99)
        self.hdid = np.zeros(self.M)
        self.flags = np.ones(self.M, dtype=np.int32)          # first_step
        self.dfdy = np.zeros((self.M, self.D, self.D))
        self.k = np.zeros(self.M, dtype=np.int32)
        self.nrej = np.zeros(self.M, dtype=np.int32)


def seulex_step(problem, par, sign, htry, st, quirks=0, arith="reference"):
    """one SeulexT.step() per trajectory of `st` (in place)."""
    set_arith(arith)
    par = np.ascontiguousarray(par, dtype=np.float64)
    htry = np.ascontiguousarray(np.broadcast_to(np.asarray(htry, dtype=np.float64), (st.M,))).copy()
    ip = C.POINTER(C.c_int)
    rc = lib().oracle_seulex_step(
        C.c_int(PB[problem]), _dp(par), C.c_int(sign), C.c_int(st.D), C.c_longlong(st.M), C.c_double(st.atol),
        C.c_double(st.rtol), C.c_int(st.KMAXX), C.c_int(quirks), _dp(htry), _dp(st.y), _dp(st.x), _dp(st.theta),
        st.k_targ.ctypes.data_as(ip), _dp(st.hnext), _dp(st.hdid), st.flags.ctypes.data_as(ip), _dp(st.dfdy),
        st.k.ctypes.data_as(ip), st.nrej.ctypes.data_as(ip))
    assert rc == 0
    return st


def max_threads():
    return lib().oracle_max_threads()
