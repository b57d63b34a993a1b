"""ctypes driver of tests/hostsim/libhostsim.so -- TEST TOOL ONLY (CPU run of the CUDA core)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhostsim.so")
HAM = dict(harmonic=0, vdp=1, vdp_pmp_smooth=2, vdp_pmp_bang=3, vdp_aug_bang=4)


def build(force=False):
    src = os.path.join(_HERE, "hostsim.cpp")
    core = os.path.join(_HERE, "..", "..", "janus_ode_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(core, f)) for f in os.listdir(core) if f.endswith(".cuh"))
    newest = max(newest, os.path.getmtime(src))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-mfma", "-shared",
                               "-x", "c++", "-o", _SO, src])
    return _SO


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _gus(opts):
    """RadauParams::Gustafsson bit mask: bit 0 on, bit 1 limiter not gated by ChangeFlag (quirk Q8)."""
    return (1 | (2 if getattr(opts, "RadauTeQuirks", 1) & 1 else 0)) if opts.Gustafsson else 0


def _params(lib, opts, ny, sign, par):
    P = (C.c_char * lib.hostsim_params_size())()
    rt = np.ascontiguousarray(np.broadcast_to(np.asarray(opts.RelTol, dtype=np.float64), (ny,)))
    at = np.ascontiguousarray(np.broadcast_to(np.asarray(opts.AbsTol, dtype=np.float64), (ny,)))
    d = C.c_double
    lib.hostsim_fill_params(P, _dp(rt), _dp(at), C.c_int(ny), d(opts.InitialStep), d(opts.JacRecompute),
                            d(opts.Safe), d(opts.Quot1), d(opts.Quot2), d(opts.FacL), d(opts.FacR), d(opts.Vitu),
                            d(opts.Vitd), d(opts.hhou), d(opts.hhod), C.c_int(sign), C.c_longlong(opts.MaxNbrStep),
                            C.c_int(int(opts.Start_Newt)), C.c_int(opts.MaxNbrNewton), C.c_int(opts.NbrStg),
                            C.c_int(opts.MinNbrStg), C.c_int(opts.MaxNbrStg), C.c_int(_gus(opts)),
                            _dp(par[:8]), C.c_int(min(par.size, 8)))
    return P


def solve_mlp(par, sign, y0, tspan, opts, oracle_lib):
    """The MLP-Hamiltonian phase machine (csrc/mlp_phase.cuh) on the CPU; stage / Jacobian evaluations
    are the callbacks oracle_eval_dyns / oracle_eval_jac of `oracle_lib` (set it to mirror arithmetic)."""
    build()
    lib = C.CDLL(_SO)
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    tspan = np.ascontiguousarray(tspan, dtype=np.float64)
    par = np.ascontiguousarray(par, dtype=np.float64)
    M, ny = y0.shape
    assert ny == 16 and tspan.shape[1] == 2
    P = _params(lib, opts, ny, sign, par)
    y = np.empty_like(y0)
    t = np.empty(M)
    h = np.empty(M)
    res = np.empty(M, dtype=np.int32)
    stats = np.empty((M, 8), dtype=np.int64)
    nstg = np.empty(M, dtype=np.int32)
    work = np.empty(M, dtype=np.int64)
    fd = C.cast(oracle_lib.oracle_eval_dyns, C.c_void_p)
    fj = C.cast(oracle_lib.oracle_eval_jac, C.c_void_p)
    rc = lib.hostsim_mlp_solve(fd, fj, _dp(par), P, _dp(y0), _dp(tspan), C.c_longlong(M), _dp(y), _dp(t), _dp(h),
                               res.ctypes.data_as(C.POINTER(C.c_int)), stats.ctypes.data_as(C.POINTER(C.c_longlong)),
                               nstg.ctypes.data_as(C.POINTER(C.c_int)), work.ctypes.data_as(C.POINTER(C.c_longlong)))
    assert rc == 0
    return dict(y=y, t=t, h=h, res=res, stats=stats, nstg=nstg, work=work)


def dopri_dual(ham, par, sign, y0, x0, xf, atol, rtol, h0=0.01, dual=True):
    """csrc/dopri853.cuh on the host: integrate to xf with repeated accepted steps, optionally with the dual part d y / d y0"""
    build()
    lib = C.CDLL(_SO)
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    par = np.ascontiguousarray(par, dtype=np.float64)
    M, ny = y0.shape
    y = np.empty_like(y0)
    S = np.empty((M, ny, ny)) if dual else None
    steps = np.zeros(M, dtype=np.int32)
    d = C.c_double
    rc = lib.hostsim_dopri_dual(C.c_int(HAM[ham]), _dp(par), d(sign), _dp(y0), C.c_longlong(M), d(x0), d(xf), d(atol), d(rtol),
                                d(h0), _dp(y), _dp(S) if dual else None, steps.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0
    return dict(y=y, sens=S, steps=steps)


def seulex_dual(ham, par, sign, y0, x0, xf, atol, rtol, KMAXX=12, h0=0.01, dual=True):
    """csrc/seulex.cuh on the host: the reference's driver loop (htry = hnext, clipped to xf), optionally with d y / d y0"""
    build()
    lib = C.CDLL(_SO)
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    par = np.ascontiguousarray(par, dtype=np.float64)
    M, ny = y0.shape
    y = np.empty_like(y0)
    S = np.empty((M, ny, ny)) if dual else None
    steps = np.zeros(M, dtype=np.int32)
    d = C.c_double
    rc = lib.hostsim_seulex_dual(C.c_int(HAM[ham]), _dp(par), d(sign), _dp(y0), C.c_longlong(M), d(x0), d(xf), d(atol), d(rtol),
                                 C.c_int(KMAXX), d(h0), _dp(y), _dp(S) if dual else None, steps.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0
    return dict(y=y, sens=S, steps=steps)


def det_tanh(x):
    build()
    lib = C.CDLL(_SO)
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lib.hostsim_det_tanh(_dp(x), C.c_longlong(x.size), _dp(out))
    return out


def solve_sens(ham, par, sign, y0, tspan, opts, par_dirs=()):
    """forward sensitivities d y / d y0 (and d y / d par[k], k in par_dirs) by the CUDA core's sens_step, on the CPU"""
    build()
    lib = C.CDLL(_SO)
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    tspan = np.ascontiguousarray(tspan, dtype=np.float64)
    par = np.ascontiguousarray(par, dtype=np.float64)
    M, ny = y0.shape
    P = _params(lib, opts, ny, sign, par)
    y = np.empty_like(y0)
    S = np.empty((M, ny, ny))
    stats = np.empty((M, 8), dtype=np.int64)
    iters = np.empty(M, dtype=np.int32)
    if par_dirs:
        Sp = np.empty((M, len(par_dirs), ny))
        pidx = (C.c_int * len(par_dirs))(*par_dirs)
        rc = lib.hostsim_solve_sens_par(C.c_int(HAM[ham]), P, _dp(y0), _dp(tspan), C.c_longlong(M), C.c_int(len(par_dirs)),
                                        pidx, _dp(y), _dp(S), _dp(Sp), stats.ctypes.data_as(C.POINTER(C.c_longlong)),
                                        iters.ctypes.data_as(C.POINTER(C.c_int)))
        assert rc == 0
        return dict(y=y, sens=S, spar=Sp.transpose(0, 2, 1).copy(), stats=stats, iters=iters)
    rc = lib.hostsim_solve_sens(C.c_int(HAM[ham]), P, _dp(y0), _dp(tspan), C.c_longlong(M), _dp(y), _dp(S),
                                stats.ctypes.data_as(C.POINTER(C.c_longlong)), iters.ctypes.data_as(C.POINTER(C.c_int)))
    assert rc == 0
    return dict(y=y, sens=S, stats=stats, iters=iters)


def solve(ham, par, sign, y0, tspan, opts, epochs_K=0):
    """epochs_K > 0: emulate the variable-order epoch scheme (park / regroup / resume every K passes)."""
    build()
    lib = C.CDLL(_SO)
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    tspan = np.ascontiguousarray(tspan, dtype=np.float64)
    par = np.ascontiguousarray(par, dtype=np.float64)
    M, ny = y0.shape
    P = (C.c_char * lib.hostsim_params_size())()
    rt = np.ascontiguousarray(np.broadcast_to(np.asarray(opts.RelTol, dtype=np.float64), (ny,)))
    at = np.ascontiguousarray(np.broadcast_to(np.asarray(opts.AbsTol, dtype=np.float64), (ny,)))
    d = C.c_double
    lib.hostsim_fill_params(P, _dp(rt), _dp(at), C.c_int(ny), d(opts.InitialStep), d(opts.JacRecompute),
                            d(opts.Safe), d(opts.Quot1), d(opts.Quot2), d(opts.FacL), d(opts.FacR), d(opts.Vitu),
                            d(opts.Vitd), d(opts.hhou), d(opts.hhod), C.c_int(sign), C.c_longlong(opts.MaxNbrStep),
                            C.c_int(int(opts.Start_Newt)), C.c_int(opts.MaxNbrNewton), C.c_int(opts.NbrStg),
                            C.c_int(opts.MinNbrStg), C.c_int(opts.MaxNbrStg), C.c_int(_gus(opts)),
                            _dp(par), C.c_int(par.size))
    y = np.empty_like(y0)
    t = np.empty(M)
    h = np.empty(M)
    res = np.empty(M, dtype=np.int32)
    stats = np.empty((M, 8), dtype=np.int64)
    nstg = np.empty(M, dtype=np.int32)
    nts = tspan.shape[1]
    out = {}
    if nts > 2:
        assert epochs_K == 0
        out = dict(tout=np.zeros((M, nts)), yout=np.zeros((M, nts, ny)), nout=np.zeros(M, dtype=np.int32))
        lib.hostsim_set_dense(C.c_int(nts), _dp(out["tout"]), _dp(out["yout"]), out["nout"].ctypes.data_as(C.POINTER(C.c_int)))
    else:
        lib.hostsim_set_dense(C.c_int(2), None, None, None)
    args = [C.c_int(HAM[ham]), P, _dp(y0), _dp(tspan), C.c_longlong(M), _dp(y), _dp(t), _dp(h),
            res.ctypes.data_as(C.POINTER(C.c_int)), stats.ctypes.data_as(C.POINTER(C.c_longlong)),
            nstg.ctypes.data_as(C.POINTER(C.c_int))]
    rc = lib.hostsim_solve_epochs(*args, C.c_int(epochs_K)) if epochs_K != 0 else lib.hostsim_solve(*args)
    assert rc == 0
    return dict(y=y, t=t, h=h, res=res, stats=stats, nstg=nstg, **out)


def seulex_step(ham_id, par, sign, htry, st, quirks=0):
    """csrc/seulex.cuh compiled for the host: one step of every trajectory of an oracle.c_oracle.SeulexOracleState
    (in place; st.dfdy is not used -- the kept Jacobian is st.yjac here)."""
    lib = build()
    L = C.CDLL(lib) if isinstance(lib, str) else lib
    par = np.ascontiguousarray(par, dtype=np.float64)
    htry = np.ascontiguousarray(np.broadcast_to(np.asarray(htry, dtype=np.float64), (st.M,))).copy()
    if not hasattr(st, "yjac"):
        st.yjac = st.y.copy()
    ip = C.POINTER(C.c_int)
    rc = L.hostsim_seulex_step(C.c_int(ham_id), _dp(par), C.c_int(sign), C.c_longlong(st.M), C.c_double(st.atol),
                               C.c_double(st.rtol), C.c_int(st.KMAXX), C.c_int(quirks), _dp(htry), _dp(st.y), _dp(st.x),
                               _dp(st.theta), st.k_targ.ctypes.data_as(ip), _dp(st.hnext), _dp(st.hdid),
                               st.flags.ctypes.data_as(ip), _dp(st.yjac), st.k.ctypes.data_as(ip), st.nrej.ctypes.data_as(ip))
    assert rc == 0
    return st
