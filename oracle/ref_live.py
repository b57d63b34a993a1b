"""ctypes binding of oracle/_ref/libref_radau.so: the UNMODIFIED reference RadauTe run live.

TEST INFRASTRUCTURE ONLY -- never imported by janus_ode_b200/.  See oracle/ref_driver.cpp and
oracle/Makefile.ref.  The library only builds where /root/reference exists (this container); on the GPU
box the prebuilt .so is used if present (it needs nothing from /root/reference at run time), and
everything pinned by it is committed as fixtures under tests/golden/ (make_reference_live.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref_radau.so")
REF = os.environ.get("JANUS_REFERENCE", "/root/reference")

PB = dict(harmonic=0, vdp=1, vdp_pmp_smooth=2, vdp_pmp_bang=3)
STAT_NAMES = ("FcnNbr", "JacNbr", "DecompNbr", "SolveNbr", "StepNbr", "AccptNbr", "StepRejNbr", "NewtRejNbr")


def can_build():
    return os.path.exists(os.path.join(REF, "src", "cpp", "radaute_impl.hpp"))


def available():
    return os.path.exists(_SO) or can_build()


def build(force=False):
    srcs = [os.path.join(_HERE, "ref_driver.cpp")]
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs)
    if stale and can_build():
        subprocess.check_call(["make", "-s", "-C", _HERE, "-f", "Makefile.ref", f"REF={REF}"])
    return _SO if os.path.exists(_SO) else None


_lib = None


def lib():
    global _lib
    if _lib is None:
        import torch  # noqa: F401  (loads libtorch before the driver that links against it)
        if build() is None:
            raise RuntimeError("oracle/_ref/libref_radau.so is not built and /root/reference is absent")
        _lib = C.CDLL(_SO)
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def evaluate(problem, par, sign, y, adaptor=True):
    """OdeFcn / JacFn of the live reference on rows y [M,Ny] -> (ydot [M,Ny], jac [M,Ny,Ny])."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    M, Ny = y.shape
    f = np.empty_like(y)
    J = np.empty((M, Ny, Ny))
    rc = lib().ref_eval(C.c_int(PB[problem]), C.c_double(par), C.c_int(sign), C.c_int(int(adaptor)),
                        C.c_int64(M), C.c_int(Ny), _dp(y), _dp(f), _dp(J))
    assert rc == 0, rc
    return f, J


def evaluate_dual(problem, par, y, yd):
    """evalDynsDual / evalJacDual (+U) of the live reference: (f [M,Ny], fd [M,Ny,D], J [M,Ny,Ny], Jd [M,Ny,Ny,D])."""
    y = np.ascontiguousarray(y, dtype=np.float64)
    yd = np.ascontiguousarray(yd, dtype=np.float64)
    M, Ny = y.shape
    D = yd.shape[2]
    f, fd, J, Jd = np.empty((M, Ny)), np.empty((M, Ny, D)), np.empty((M, Ny, Ny)), np.empty((M, Ny, Ny, D))
    rc = lib().ref_eval_dual(C.c_int(PB[problem]), C.c_double(par), C.c_int64(M), C.c_int(Ny), C.c_int(D), _dp(y), _dp(yd),
                             _dp(f), _dp(fd), _dp(J), _dp(Jd))
    if rc != 0:
        lib().ref_last_error.restype = C.c_char_p
        raise RuntimeError(lib().ref_last_error().decode()[:400])
    return f, fd, J, Jd


def _solve_chunk(args):
    (problem, par, sign, adaptor, batch, y0, tspan, rtol, atol, h0, stg, max_steps) = args
    import time
    M, Ny = y0.shape
    y = np.empty_like(y0)
    t = np.empty(M)
    h = np.empty(M)
    res = np.empty(M)
    stats = np.empty((M, 8), dtype=np.int64)
    nstg = np.empty(M, dtype=np.int64)
    passes = np.empty(M, dtype=np.int64)
    lib()
    t_start = time.perf_counter()
    rc = lib().ref_radau_solve(
        C.c_int(PB[problem]), C.c_double(par), C.c_int(sign), C.c_int(int(adaptor)), C.c_int(int(batch)),
        C.c_int64(M), C.c_int(Ny), _dp(y0), _dp(tspan), C.c_double(rtol), C.c_double(atol), C.c_double(h0),
        C.c_int(stg[0]), C.c_int(stg[1]), C.c_int(stg[2]), C.c_int(max_steps),
        _dp(y), _dp(t), _dp(h), _dp(res), _ip(stats), _ip(nstg), _ip(passes))
    if rc != 0:
        lib().ref_last_error.restype = C.c_char_p
        raise RuntimeError(f"reference solver failed ({rc}): {lib().ref_last_error().decode()[:400]}")
    secs = np.full(M, (time.perf_counter() - t_start) / max(M, 1))   # solver seconds per row (no process start-up)
    return dict(y=y, t=t, h=h, res=res.astype(np.int32), stats=stats, nstg=nstg.astype(np.int32), passes=passes, secs=secs)


def trace(problem, par, sign, y0, tspan, rtol, atol, h0=1e-2, stages=(3, 3, 7), max_steps=100000, adaptor=False,
          cap=200000):
    """one trajectory; log [n,10] of every OdeFcn call (see ref_driver.cpp: ref_radau_trace)."""
    y0 = np.ascontiguousarray(y0, dtype=np.float64).reshape(-1)
    tspan = np.ascontiguousarray(tspan, dtype=np.float64).reshape(2)
    Ny = y0.shape[0]
    log = np.zeros((cap, 10))
    n = C.c_int()
    y = np.empty(Ny)
    stats = np.empty(8, dtype=np.int64)
    rc = lib().ref_radau_trace(
        C.c_int(PB[problem]), C.c_double(par), C.c_int(sign), C.c_int(int(adaptor)), C.c_int(Ny), _dp(y0), _dp(tspan),
        C.c_double(rtol), C.c_double(atol), C.c_double(h0), C.c_int(stages[0]), C.c_int(stages[1]), C.c_int(stages[2]),
        C.c_int(max_steps), _dp(log), C.c_int(cap), C.byref(n), _dp(y), _ip(stats))
    assert rc == 0
    return dict(log=log[:n.value].copy(), y=y, stats=stats)


def steps(problem, par, sign, y0, tspan, rtol, atol, h0=1e-2, stages=(3, 3, 7), max_steps=100000, adaptor=False, cap=4096):
    """one trajectory: the reference's own OutFlag-1 record (tout [n], yout [n,Ny]) of every accepted step."""
    y0 = np.ascontiguousarray(y0, dtype=np.float64).reshape(-1)
    tspan = np.ascontiguousarray(tspan, dtype=np.float64).reshape(2)
    Ny = y0.shape[0]
    tout = np.zeros(cap)
    yout = np.zeros((cap, Ny))
    n = C.c_int()
    y = np.empty(Ny)
    stats = np.empty(8, dtype=np.int64)
    rc = lib().ref_radau_steps(
        C.c_int(PB[problem]), C.c_double(par), C.c_int(sign), C.c_int(int(adaptor)), C.c_int(Ny), _dp(y0), _dp(tspan),
        C.c_double(rtol), C.c_double(atol), C.c_double(h0), C.c_int(stages[0]), C.c_int(stages[1]), C.c_int(stages[2]),
        C.c_int(max_steps), C.c_int(cap), _dp(tout), _dp(yout), C.byref(n), _dp(y), _ip(stats))
    assert rc == 0, rc
    k = min(n.value, cap)
    return dict(tout=tout[:k].copy(), yout=yout[:k].copy(), nout=n.value, y=y, stats=stats)


def solve(problem, par, sign, y0, tspan, rtol, atol, h0=1e-2, stages=(3, 3, 7), max_steps=100000,
          adaptor=False, batch=False, procs=1):
    """Run janus::RadauTe on every row (one solver object per row unless batch=True).
    stages = (NbrStg, MinNbrStg, MaxNbrStg).  procs > 1 forks worker processes over row chunks."""
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    tspan = np.ascontiguousarray(tspan, dtype=np.float64)
    M = y0.shape[0]
    base = (problem, float(par), int(sign), adaptor, batch)
    tail = (float(rtol), float(atol), float(h0), tuple(int(s) for s in stages), int(max_steps))
    if procs <= 1 or batch or M < 2 * procs:
        return _solve_chunk(base + (y0, tspan) + tail)
    # workers are separate interpreters (python -m oracle.ref_live <job.npz> <out.npz>): libtorch does not
    # survive fork(), and multiprocessing's spawn re-imports the caller's __main__
    import pickle
    import sys
    import tempfile
    lib()                                            # build before starting workers
    edges = np.linspace(0, M, procs * 2 + 1).astype(int)
    chunks = [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]
    parts = [None] * len(chunks)
    with tempfile.TemporaryDirectory() as td:
        running, nxt = {}, 0
        env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1")
        while nxt < len(chunks) or running:
            while nxt < len(chunks) and len(running) < procs:
                a, b = chunks[nxt]
                job, out = os.path.join(td, f"j{nxt}.pkl"), os.path.join(td, f"o{nxt}.pkl")
                with open(job, "wb") as fh:
                    pickle.dump(base + (y0[a:b].copy(), tspan[a:b].copy()) + tail, fh)
                running[nxt] = (subprocess.Popen([sys.executable, "-m", "oracle.ref_live", job, out],
                                                 cwd=os.path.dirname(_HERE), env=env), out)
                nxt += 1
            for k in list(running):
                pr, out = running[k]
                if pr.poll() is not None:
                    assert pr.returncode == 0, f"reference worker {k} failed ({pr.returncode})"
                    with open(out, "rb") as fh:
                        parts[k] = pickle.load(fh)
                    del running[k]
            import time
            time.sleep(0.05)
    return {k: np.concatenate([p[k] for p in parts]) for k in parts[0]}


if __name__ == "__main__":
    import pickle
    import sys
    with open(sys.argv[1], "rb") as _fh:
        _job = pickle.load(_fh)
    _r = _solve_chunk(_job)
    with open(sys.argv[2], "wb") as _fh:
        pickle.dump(_r, _fh)
