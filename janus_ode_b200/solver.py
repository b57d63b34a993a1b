"""Python host side of the accelerated path: the reference's solver API shape over the C ABI.

Mirrors
  RadauTe(OdeFcn, JacFn, tspan, y0, Op, params).solve() -> res, public fields y, t, h, stats
      (/root/reference/src/cpp/radaute.hpp:243-483, impl src/cpp/radaute_impl.hpp:5-1345)
  the `dyns(t, y, params)` / `dyns.jacobian(t, y, params)` protocol of the Python steppers
      (/root/reference/src/python/seulex_dual_example_vdp.py:26-67)
PyTorch is only the owner of device memory and streams here; all arithmetic happens in
libjanus_b200.so.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .hamiltonians import Hamiltonian
from .options import OptionsTe, StatsTe


def _stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


_HANDLE_CACHE = {}


def _get_handle(ham: Hamiltonian, options: OptionsTe, costate_sign: int, device):
    """Solver objects are one-shot like the reference's (construct per solve), but the native handle
    behind them -- device workspace of the host-buffer and variable-order paths -- is reused for the
    same (Hamiltonian, options, device)."""
    import threading

    import numpy as np

    from .options import _to_np
    try:
        # one native handle per (Hamiltonian, options, device, stream, thread): the C ABI allows one call at a time per
        # handle (include/janus_b200.h), so concurrent streams / threads never share one
        stream = torch.cuda.current_stream(device).cuda_stream if isinstance(device, torch.device) else 0
        key = (ham, costate_sign, str(device), stream, threading.get_ident(),
               tuple((k, tuple(np.atleast_1d(np.asarray(_to_np(v), dtype=float)).tolist()) if k in ("RelTol", "AbsTol") else v)
                     for k, v in sorted(vars(options).items())))
        h = _HANDLE_CACHE.get(key)
    except TypeError:
        return _Handle(ham, options, costate_sign)
    if h is None:
        if len(_HANDLE_CACHE) >= 8:
            # evict the oldest entry from the cache only: solver / state objects that still reference it keep it
            # alive, _Handle.__del__ destroys the native handle when the last of them goes away
            _HANDLE_CACHE.pop(next(iter(_HANDLE_CACHE)))
        h = _HANDLE_CACHE[key] = _Handle(ham, options, costate_sign)
    return h


class _Handle:
    def __init__(self, ham: Hamiltonian, options: OptionsTe, costate_sign: int):
        L = _lib.lib(ham.lib_path)
        o = options.to_c(ham.Ny, costate_sign)
        par = (C.c_double * max(1, len(ham.params)))(*ham.params)
        h = C.c_void_p()
        _lib.check(L.jb200_create(C.byref(o), ham.ham_id, par, len(ham.params), C.byref(h)), L)
        self.h = h
        self.L = L
        ev = options.EventsFcn
        if ev is not None:
            K = len(ev.b)
            A = (C.c_double * (K * ham.Ny))(*[v for row in ev.A for v in row])
            if len(ev.A[0]) != ham.Ny:
                raise ValueError(f"LinearEvents.A must be [K,{ham.Ny}]")
            b = (C.c_double * K)(*ev.b)
            stop = (C.c_int32 * K)(*ev.stop)
            slope = (C.c_int32 * K)(*ev.slope)
            _lib.check(L.jb200_set_events(h, K, A, b, stop, slope), L)

    def close(self):
        if self.h:
            self.L.jb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class HamiltonianSystem:
    """`dyns` object of the reference protocol: dyns(t, y[, params]) -> ydot [M,Ny] and
    dyns.jacobian(t, y[, params]) -> J [M,Ny,Ny], evaluated by the dual / hyper-dual kernels
    (evalDyns / evalJac, src/cpp/janus_ode_common.hpp:2161-2299).
    costate_sign +1 = reference (pdot = +dH/dx), -1 = canonical."""

    def __init__(self, hamiltonian: Hamiltonian, costate_sign: int = +1):
        if not isinstance(hamiltonian, Hamiltonian):
            raise TypeError("the accelerated path needs a janus_ode_b200.Hamiltonian descriptor; a Python "
                            "callable cannot run inside a CUDA kernel")
        if costate_sign not in (1, -1):
            raise ValueError("costate_sign must be +1 or -1")
        self.hamiltonian = hamiltonian
        self.costate_sign = costate_sign
        self._handle = None

    def _h(self):
        if self._handle is None:
            self._handle = _Handle(self.hamiltonian, OptionsTe(), self.costate_sign)
        return self._handle

    def _check(self, y):
        if y.dim() != 2 or y.shape[1] != self.hamiltonian.Ny:
            raise ValueError(f"y must be [M,{self.hamiltonian.Ny}]")
        if y.dtype != torch.float64:
            raise TypeError("y must be float64")
        if not y.is_cuda:
            raise _lib.JanusB200Error("stage evaluation runs on CUDA tensors only (no CPU fallback)")
        return y.contiguous()

    def __call__(self, t, y, params=None):
        y = self._check(y)
        out = torch.empty_like(y)
        H = self._h()
        _lib.check_h(H, H.L.jb200_eval_dyns(H.h, _ptr(y), y.shape[0], _ptr(out), _stream_ptr(y.device)))
        return out

    def jacobian(self, t, y, params=None):
        y = self._check(y)
        ny = self.hamiltonian.Ny
        out = torch.empty(y.shape[0], ny, ny, dtype=torch.float64, device=y.device)
        H = self._h()
        _lib.check_h(H, H.L.jb200_eval_jac(H.h, _ptr(y), y.shape[0], _ptr(out), _stream_ptr(y.device)))
        return out


def _hs_third(self, y):
    """T [M,Ny,Ny,Ny], T[m,r,c,k] = d J[r][c] / d y_k: the reference's third-order ladders in y-order."""
    return self.jacobian_dual(None, y)[2]


def _hs_jacobian_dual(self, t, y, yd=None):
    """evalJacDual (+ evalDynsDual) of the reference, src/cpp/janus_ode_common.hpp:2075-2158, 2302-2416:
    y [M,Ny], yd [M,Ny,D] (dual part of the TensorDual state) -> (f, J, T, fd, Jd): dynamics [M,Ny], Jacobian
    [M,Ny,Ny], third-order tensor [M,Ny,Ny,Ny], and their dual parts fd [M,Ny,D] = J yd, Jd [M,Ny,Ny,D] = T yd
    (None without yd).  One third-order jet of H per trajectory instead of N + N^2 backward passes per block."""
    y = self._check(y)
    M, ny = y.shape
    dev = y.device
    f = torch.empty(M, ny, dtype=torch.float64, device=dev)
    J = torch.empty(M, ny, ny, dtype=torch.float64, device=dev)
    T = torch.empty(M, ny, ny, ny, dtype=torch.float64, device=dev)
    fd = Jd = None
    D = 0
    if yd is not None:
        if yd.dim() != 3 or yd.shape[:2] != (M, ny) or yd.dtype != torch.float64 or not yd.is_cuda:
            raise ValueError(f"yd must be a CUDA float64 tensor [M,{ny},D]")
        yd = yd.contiguous()
        D = yd.shape[2]
        fd = torch.empty(M, ny, D, dtype=torch.float64, device=dev)
        Jd = torch.empty(M, ny, ny, D, dtype=torch.float64, device=dev)
    H = self._h()
    _lib.check_h(H, H.L.jb200_eval_jac_dual(H.h, _ptr(y), _ptr(yd), D, M, _ptr(f), _ptr(J), _ptr(T), _ptr(fd), _ptr(Jd),
                                            _stream_ptr(dev)))
    return f, J, T, fd, Jd


HamiltonianSystem.jacobian_dual = _hs_jacobian_dual
HamiltonianSystem.third_derivatives = _hs_third


class RadauTe:
    """Batched variable-order Radau IIA for a Hamiltonian system, one adaptive step per trajectory.

    RadauTe(dyns, jac, tspan[M,2], y0[M,Ny], options, params=None), like the reference; `dyns`
    must be a HamiltonianSystem (jac is implied by it and may be None or dyns.jacobian).
    solve() returns res[M] (float64 codes 0/1/2/3 like the reference) and fills the public
    fields y, t, h, stats, nstg, count.  CUDA tensors are integrated in place on their device
    and stream; CPU tensors go through jb200_solve_host (copy in, solve, copy out) -- the
    drop-in for a reference caller, who holds CPU tensors.
    """

    def __init__(self, dyns, jac, tspan, y0, options: OptionsTe | None = None, params=None):
        if not isinstance(dyns, HamiltonianSystem):
            raise TypeError("dyns must be a janus_ode_b200.HamiltonianSystem")
        if jac is not None and getattr(jac, "__self__", None) is not dyns and jac is not dyns:
            raise TypeError("jac must be None or dyns.jacobian (the Jacobian is the Hessian of H)")
        self.dyns = dyns
        self.options = options or OptionsTe()
        ham = dyns.hamiltonian
        if tspan.dim() != y0.dim():
            raise ValueError("tspan must be a tensor")          # radaute_impl.hpp:10-14
        if y0.dim() != 2:
            raise ValueError("Initial conditions argument must be a valid vector")
        if y0.shape[1] != ham.Ny:
            raise ValueError(f"y0 must be [M,{ham.Ny}] for Hamiltonian '{ham.name}'")
        if tspan.shape[0] != y0.shape[0] or tspan.shape[1] < 2:
            raise ValueError("tspan must be [M,>=2]")
        if tspan.shape[1] > 2 and not y0.is_cuda:
            raise NotImplementedError("dense output (tspan with interior points) takes CUDA tensors")
        if tspan.shape[1] > 2:
            d = tspan[:, 1:] - tspan[:, :-1]
            if not (bool((d > 0).all()) or bool((d < 0).all())):
                raise ValueError("Time vector must be strictly monotonic")
        if y0.dtype != torch.float64 or tspan.dtype != torch.float64:
            raise TypeError("y0 and tspan must be float64")
        if y0.device != tspan.device:
            raise ValueError("y0 and tspan must live on the same device")
        t0, tf = tspan[:, 0], tspan[:, -1]
        # radaute_impl.hpp:54-59; large host batches: jb200_solve_host_compact makes the same check on its way through the rows
        if (y0.is_cuda or y0.shape[0] <= 65536) and bool((tf == t0).any()):
            raise ValueError("Time vector must be strictly monotonic")
        dev_key = y0.device if y0.is_cuda else torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "none"
        if y0.is_cuda:
            with torch.cuda.device(y0.device):
                self._handle = _get_handle(ham, self.options, dyns.costate_sign, dev_key)   # validates the options
        else:
            self._handle = _get_handle(ham, self.options, dyns.costate_sign, dev_key)
        self.M, self.Ny = y0.shape
        self.device = y0.device
        self.tspan = tspan.contiguous()
        self.ntspan = tspan.shape[1]
        self.tout = self.yout = self.nout = None
        self.te = self.ye = self.ie = self.nev = None     # located events (teout, yeout, ieout of radau.m:1110-1121)
        self.y = y0.contiguous()          # never written: solve() replaces it with the result tensor
        self.t = t0                        # a view until solve() replaces it with the final times
        self.h = None                      # set by solve()
        self.stats = None
        self.nstg = None
        self._res = self._res32 = None
        self.want_work = False     # host path: also fetch the per-trajectory FLOP counts (roofline accounting)

    def solve(self):
        H = self._handle
        M, ny, dev = self.M, self.Ny, self.device
        # host path: page-locked result buffers so the D2H copies of jb200_solve_host are direct DMA
        kw = dict(device=dev) if self.y.is_cuda else dict(device=dev, pin_memory=True)
        host = not self.y.is_cuda
        launches = None
        y_out = torch.empty(M, ny, dtype=torch.float64, **kw)
        t_out = torch.empty(M, dtype=torch.float64, **kw)
        h_out = torch.empty(M, dtype=torch.float64, **kw)
        res = torch.empty(M, dtype=torch.int32, **kw)
        # host path: int32 counters on the PCIe link (jb200_solve_host_compact); StatsTe widens lazily
        stats = torch.empty(M, 8, dtype=torch.int32 if host else torch.int64, **kw)
        nstg = torch.empty(M, dtype=torch.int32, **kw)
        work = None if (host and not self.want_work) else torch.empty(M, dtype=torch.int64, **kw)
        if self.ntspan > 2 and (self.options.EventsFcn is not None or self.options.OutputCapacity > 0 or self.options.Refine > 1):
            raise NotImplementedError("events / per-step output need a two-column tspan (OutFlag 3 records the tspan points only)")
        if not self.y.is_cuda and (self.options.EventsFcn is not None or self.options.OutputCapacity > 0 or self.options.Refine > 1):
            raise NotImplementedError("events / per-step output run on CUDA tensors")
        if self.y.is_cuda and self.ntspan > 2:
            # output at the requested tspan points (radau.m OutFlag 3): tout, yout, nout like the reference
            self.tout = torch.zeros(M, self.ntspan, dtype=torch.float64, **kw)
            self.yout = torch.zeros(M, self.ntspan, ny, dtype=torch.float64, **kw)
            self.nout = torch.zeros(M, dtype=torch.int32, **kw)
            with torch.cuda.device(dev):
                _lib.check_h(H, H.L.jb200_solve_dense(H.h, _ptr(self.y), _ptr(self.tspan), self.ntspan, M, _ptr(y_out),
                                                     _ptr(t_out), _ptr(h_out), _ptr(res), _ptr(stats), _ptr(nstg),
                                                     _ptr(work), _ptr(self.tout), _ptr(self.yout), _ptr(self.nout),
                                                     _stream_ptr(dev)))
        elif self.y.is_cuda and (self.options.OutputCapacity > 0 or self.options.EventsFcn is not None or self.options.Refine > 1):
            # every computed point (OutFlag 1 / 2) and / or located events
            cap, ecap = int(self.options.OutputCapacity), int(self.options.EventsCapacity)
            self.tout = torch.zeros(M, max(cap, 1), dtype=torch.float64, **kw)
            self.yout = torch.zeros(M, max(cap, 1), ny, dtype=torch.float64, **kw)
            self.nout = torch.zeros(M, dtype=torch.int32, **kw)
            self.te = torch.zeros(M, max(ecap, 1), dtype=torch.float64, **kw)
            self.ye = torch.zeros(M, max(ecap, 1), ny, dtype=torch.float64, **kw)
            self.ie = torch.zeros(M, max(ecap, 1), dtype=torch.int32, **kw)
            self.nev = torch.zeros(M, dtype=torch.int32, **kw)
            with torch.cuda.device(dev):
                _lib.check_h(H, H.L.jb200_solve_out(H.h, _ptr(self.y), _ptr(self.tspan), M, _ptr(y_out), _ptr(t_out), _ptr(h_out),
                                                   _ptr(res), _ptr(stats), _ptr(nstg), _ptr(work), int(self.options.Refine), cap,
                                                   _ptr(self.tout), _ptr(self.yout), _ptr(self.nout), ecap, _ptr(self.te),
                                                   _ptr(self.ye), _ptr(self.ie), _ptr(self.nev), _stream_ptr(dev)))
        elif self.y.is_cuda:
            with torch.cuda.device(dev):
                _lib.check_h(H, H.L.jb200_solve(H.h, _ptr(self.y), _ptr(self.tspan), M, _ptr(y_out), _ptr(t_out),
                                           _ptr(h_out), _ptr(res), _ptr(stats), _ptr(nstg), _ptr(work), _stream_ptr(dev)))
        else:
            _lib.check_h(H, H.L.jb200_solve_host_compact(H.h, _ptr(self.y), _ptr(self.tspan), M, _ptr(y_out), _ptr(t_out),
                                                         _ptr(h_out), _ptr(res), _ptr(stats), _ptr(nstg), _ptr(work)))
        self.y, self.t, self.h = y_out, t_out, h_out
        self.stats = StatsTe.from_tensor(stats)
        self.nstg = nstg
        self.work = work           # algorithmic FLOPs per trajectory (roofline accounting)
        self._res32 = res
        self._res = None
        self.gpu_launches = launches if launches is not None else int(H.L.jb200_last_launches(H.h))
        return self.res

    @property
    def res(self):
        """[M] fp64 result codes like the reference's `Tensor solve()` (widened from int32 on first use)."""
        if self._res is None and self._res32 is not None:
            self._res = self._res32.to(torch.float64)
        return self._res

    @res.setter
    def res(self, v):
        self._res, self._res32 = v, None

    @property
    def count(self):
        """passes of the reference's outer loop (radaute.hpp:417): the largest StepNbr of the batch"""
        if self.stats is None or self.M == 0:
            return 0
        return int(self.stats.raw[:, 4].max().item())


class RadauState:
    """State object of the incremental protocol, shaped like the reference's SeulexState(y, x, atol, rtol,
    KMAXX, alpha, ft) (src/python/seulext.py:8-148): y [M,Ny], x [M] (time), ft (final time, scalar or [M]),
    atol / rtol.  After each step: y, x, hdid (step taken), hnext (proposed next step), done [M] bool,
    res [M], stats.  The integrator's internal memory (collocation polynomial, Jacobian, flags, counters)
    stays parked on the device."""

    def __init__(self, y, x, atol=1.0e-6, rtol=1.0e-3, ft=None):
        if ft is None:
            raise ValueError("ft (final time) is required")
        if not y.is_cuda:
            raise _lib.JanusB200Error("RadauState lives on a CUDA device (no CPU fallback)")
        self.y = y.contiguous()
        self.x = x.to(y).contiguous().reshape(-1)
        self.ft = torch.as_tensor(ft, dtype=torch.float64, device=y.device).expand(self.x.shape).contiguous()
        self.atol, self.rtol = atol, rtol
        self.M, self.D = y.shape
        self.hnext = torch.zeros_like(self.x)
        self.hdid = torch.zeros_like(self.x)
        self.done = torch.zeros(self.M, dtype=torch.bool, device=y.device)
        self.res = torch.zeros(self.M, dtype=torch.int32, device=y.device)
        self.stats = None
        self.nstg = None
        self._native = None
        self._owner = None

    def __del__(self):
        try:
            if self._native is not None:
                self._owner.L.jb200_state_destroy(self._native)
                self._native = None
        except Exception:
            pass


class RadauT:
    """Stepper object of the incremental protocol: `state = RadauT(options).step(htry, dyns, ft, state)` in a
    loop `htry = state.hnext` until `state.done.all()`, like SeulexT.step (src/python/seulext.py:158-581,
    driver src/python/seulex_dual_example_vdp.py:136-175).  Every call advances each unfinished trajectory
    by exactly one accepted Radau step; looping reproduces RadauTe.solve() bit for bit."""

    def __init__(self, options: OptionsTe | None = None):
        self.options = options

    def step(self, htry, dyns, ft, state: RadauState):
        if not isinstance(dyns, HamiltonianSystem):
            raise TypeError("dyns must be a janus_ode_b200.HamiltonianSystem")
        dev = state.y.device
        if state._native is None:
            op = self.options or OptionsTe(RelTol=state.rtol, AbsTol=state.atol)
            if htry is not None:
                import copy
                op = copy.copy(op)
                op.InitialStep = float(torch.as_tensor(htry, dtype=torch.float64).reshape(-1)[0])   # first trial step, like step(htry, ...)
            with torch.cuda.device(dev):
                H = _get_handle(dyns.hamiltonian, op, dyns.costate_sign, dev)
                tspan = torch.stack([state.x, torch.as_tensor(ft, dtype=torch.float64, device=dev).expand_as(state.x)], 1).contiguous()
                if bool((tspan[:, 1] == tspan[:, 0]).any()):
                    raise ValueError("Time vector must be strictly monotonic")
                nat = C.c_void_p()
                _lib.check_h(H, H.L.jb200_state_create(H.h, _ptr(state.y), _ptr(tspan), state.M, C.byref(nat), _stream_ptr(dev)))
            state._native, state._owner = nat, H
            htry = None                                   # the state starts with InitialStep = htry
        H = state._owner
        ht = None
        if htry is not None:
            ht = torch.as_tensor(htry, dtype=torch.float64, device=dev).expand_as(state.x).contiguous()
        with torch.cuda.device(dev):
            _lib.check_h(H, H.L.jb200_state_step(state._native, _ptr(ht), _stream_ptr(dev)))
            M, ny = state.M, state.D
            y = torch.empty(M, ny, dtype=torch.float64, device=dev)
            t = torch.empty(M, dtype=torch.float64, device=dev)
            hn, hd = torch.empty_like(t), torch.empty_like(t)
            res = torch.empty(M, dtype=torch.int32, device=dev)
            st = torch.empty(M, 8, dtype=torch.int64, device=dev)
            ns = torch.empty(M, dtype=torch.int32, device=dev)
            dn = torch.empty(M, dtype=torch.int32, device=dev)
            _lib.check_h(H, H.L.jb200_state_read(state._native, _ptr(y), _ptr(t), _ptr(hn), _ptr(hd), _ptr(res), _ptr(st),
                                                 _ptr(ns), _ptr(dn), _stream_ptr(dev)))
        state.y, state.x, state.hnext, state.hdid = y, t, hn, hd
        state.res, state.stats, state.nstg, state.done = res, StatsTe.from_tensor(st), ns, dn.bool()
        return state


def lu_solve(A, b):
    """Batched small dense LU factor + solve on CUDA tensors, A [M,n,n], b [M,n], n in {2,3,4} or 16;
    real float64 or complex128 (the torch::linalg_lu_factor / lu_solve role on the Radau path,
    src/cpp/radaute_impl.hpp:1558,1754)."""
    if not A.is_cuda:
        raise _lib.JanusB200Error("lu_solve runs on CUDA tensors only (no CPU fallback)")
    M, n, _ = A.shape
    L = _lib.lib()
    st = _stream_ptr(A.device)
    if A.is_complex() or b.is_complex():
        A = A.to(torch.complex128)
        b = b.to(torch.complex128)
        Ar, Ai = A.real.contiguous(), A.imag.contiguous()
        br, bi = b.real.contiguous(), b.imag.contiguous()
        xr, xi = torch.empty_like(br), torch.empty_like(bi)
        _lib.check(L.jb200_lu_solve(n, M, _ptr(Ar), _ptr(Ai), _ptr(br), _ptr(bi), _ptr(xr), _ptr(xi), st))
        return torch.complex(xr, xi)
    A = A.contiguous()
    b = b.contiguous()
    x = torch.empty_like(b)
    _lib.check(L.jb200_lu_solve(n, M, _ptr(A), None, _ptr(b), None, _ptr(x), None, st))
    return x


def det_pow(x, y):
    """deterministic pow of the controller (test hook)."""
    out = torch.empty_like(x)
    _lib.check(_lib.lib().jb200_det_pow(_ptr(x.contiguous()), _ptr(y.contiguous()), x.numel(), _ptr(out),
                                        _stream_ptr(x.device)))
    return out


def fp64_peak_tflops(iters=20000):
    v = C.c_double()
    _lib.check(_lib.lib().jb200_fp64_peak(int(iters), C.byref(v)))
    return v.value


def fp64_dmma_tflops(iters=20000):
    """FP64 tensor-core (mma.sync m8n8k4 f64) throughput probe, TFLOP/s"""
    v = C.c_double()
    _lib.check(_lib.lib().jb200_dmma_peak(int(iters), C.byref(v)))
    return v.value


class SeulexState:
    """State object of the reference's stiff extrapolation stepper, same constructor and fields as
    SeulexState(y, x, atoll, rtoll, KMAXX, alpha, ft) (/root/reference/src/python/seulext.py:8-70): y [M,D],
    x [M], hnext / hdid [M], k_targ [M], theta [M], first_step / last_step / reject / prev_reject [M] (bool views
    of one flag word).  `alpha` and `ft` are carried like the reference does (its step() does not read them).
    Everything lives on the CUDA device of `y`; `dfdy` (the kept Jacobian) is evaluated on demand from the state it
    was taken at."""

    def __init__(self, y, x, atoll, rtoll, KMAXX=13, alpha=0.0, ft=None):
        if not (isinstance(y, torch.Tensor) and y.is_cuda and y.dtype == torch.float64 and y.dim() == 2):
            raise _lib.JanusB200Error("SeulexState lives on CUDA float64 tensors [M,D] only (no CPU fallback)")
        dev = y.device
        self.M, self.D = y.shape
        self.y = y.contiguous().clone()
        self.x = torch.as_tensor(x, dtype=torch.float64, device=dev).reshape(-1).expand(self.M).contiguous().clone()
        self.atol, self.rtol = float(atoll), float(rtoll)
        self.alpha, self.ft = alpha, ft
        self.KMAXX, self.IMAXX = int(KMAXX), int(KMAXX) + 1
        th, kt = C.c_double(), C.c_int32()
        _lib.check(_lib.lib().jb200_seulex_init(self.atol, self.rtol, self.KMAXX, C.byref(th), C.byref(kt)))
        self.theta = torch.full((self.M,), th.value, dtype=torch.float64, device=dev)
        self.k_targ = torch.full((self.M,), kt.value, dtype=torch.int32, device=dev)
        self.hnext = torch.full((self.M,), -1.0e99, dtype=torch.float64, device=dev)
        self.hdid = torch.zeros(self.M, dtype=torch.float64, device=dev)
        self.flags = torch.ones(self.M, dtype=torch.int32, device=dev)       # first_step
        self.k = torch.zeros(self.M, dtype=torch.int32, device=dev)          # table rows of the last step
        self.nrej = torch.zeros(self.M, dtype=torch.int32, device=dev)       # rejected attempts of the last step
        self._yjac = self.y.clone()
        self.xold = self.x.clone()
        self.EPS = torch.finfo(torch.float64).eps
        self._dyns = None

    first_step = property(lambda self: (self.flags & 1) != 0)
    last_step = property(lambda self: (self.flags & 2) != 0)
    reject = property(lambda self: (self.flags & 4) != 0)
    prev_reject = property(lambda self: (self.flags & 8) != 0)

    @property
    def dfdy(self):
        """the Jacobian the stepper currently holds, [M,D,D] (state.dfdy of the reference)"""
        if self._dyns is None:
            raise _lib.JanusB200Error("no step has been taken yet")
        return self._dyns.jacobian(self.x, self._yjac)

    _FIELDS = ("y", "x", "theta", "k_targ", "hnext", "hdid", "flags", "k", "nrej", "_yjac", "xold")

    def __getitem__(self, mask):
        """sub-state of the selected trajectories (seulext.py:118-148), used as `state[mask] = step(..., state[mask])`"""
        import copy
        sub = copy.copy(self)
        for f in self._FIELDS:
            setattr(sub, f, getattr(self, f)[mask].contiguous())
        sub.M = sub.y.shape[0]
        return sub

    def __setitem__(self, mask, new):
        for f in self._FIELDS:
            getattr(self, f)[mask] = getattr(new, f)
        self._dyns = new._dyns or self._dyns


class SeulexT:
    """`state = SeulexT().step(htry, dyns, state)`: one step of the semi-implicit extrapolation method for every
    trajectory of `state`, like /root/reference/src/python/seulext.py:158-581 (driver loop:
    src/python/seulex_batch_example.py:44-77: htry = state.hnext, clipped to the final time).  `dyns` must be a
    HamiltonianSystem; the whole step runs in one kernel (`jb200_seulex_step`, csrc/seulex.cuh).
    reference_clock=True reproduces the reference's time bookkeeping after an in-step reduction (x advances by htry,
    seulext.py:483-485); the default advances x by the step actually taken."""

    def __init__(self, reference_clock: bool = False):
        self.quirks = 1 if reference_clock else 0

    def step(self, htry, dyns, state: SeulexState) -> SeulexState:
        if not isinstance(dyns, HamiltonianSystem):
            raise TypeError("dyns must be a janus_ode_b200.HamiltonianSystem")
        if state.D != dyns.hamiltonian.Ny:
            raise ValueError(f"state.y must be [M,{dyns.hamiltonian.Ny}]")
        dev = state.y.device
        ht = torch.as_tensor(htry, dtype=torch.float64, device=dev).reshape(-1).expand(state.M).contiguous()
        H = dyns._h()
        state._dyns = dyns
        state.xold = state.x.clone()
        with torch.cuda.device(dev):
            _lib.check_h(H, H.L.jb200_seulex_step(H.h, state.M, state.atol, state.rtol, state.KMAXX, self.quirks, _ptr(ht),
                                                  _ptr(state.y), _ptr(state.x), _ptr(state.theta), _ptr(state.k_targ),
                                                  _ptr(state.hnext), _ptr(state.hdid), _ptr(state.flags), _ptr(state._yjac),
                                                  _ptr(state.k), _ptr(state.nrej), _stream_ptr(dev)))
        return state


class SeulexTedState(SeulexState):
    """SeulexState with a dual part, the role of the reference's SeulexTedState(yin, xin, params, nparams_step, atoll, rtoll,
    KMAXX) (/root/reference/src/python/seulexted.py:12-95, whose y is a TensorDual seeded with the identity): `yd [M,D,Nd]`
    (default: identity, Nd = D <= 8) rides along; `state.y` stays the real part."""

    def __init__(self, yin, xin, atoll, rtoll, KMAXX=13, yd=None, ft=None):
        super().__init__(yin, xin, atoll, rtoll, KMAXX, 0.0, ft)
        if yd is None:
            yd = torch.eye(self.D, dtype=torch.float64, device=self.y.device).repeat(self.M, 1, 1)
        if yd.dim() != 3 or yd.shape[0] != self.M or yd.shape[1] != self.D or not (1 <= yd.shape[2] <= 8):
            raise ValueError(f"yd must be [M,{self.D},Nd] with 1 <= Nd <= 8")
        self.yd = yd.to(self.y.device, torch.float64).contiguous().clone()
        self._sjac = self.yd.clone()

    _FIELDS = SeulexState._FIELDS + ("yd", "_sjac")


class SeulexTed(SeulexT):
    """`state = SeulexTed().step(htry, dyns, state)` on a SeulexTedState: the step of SeulexT plus the dual part of y
    propagated across it (`jb200_seulex_step_dual`; /root/reference/src/python/seulexted.py:238-1396).  The real part --
    states, orders, step sizes -- is SeulexT's bit for bit."""

    def step(self, htry, dyns, state: SeulexTedState) -> SeulexTedState:
        if not isinstance(dyns, HamiltonianSystem):
            raise TypeError("dyns must be a janus_ode_b200.HamiltonianSystem")
        if state.D != dyns.hamiltonian.Ny:
            raise ValueError(f"state.y must be [M,{dyns.hamiltonian.Ny}]")
        dev = state.y.device
        ht = torch.as_tensor(htry, dtype=torch.float64, device=dev).reshape(-1).expand(state.M).contiguous()
        H = dyns._h()
        state._dyns = dyns
        state.xold = state.x.clone()
        with torch.cuda.device(dev):
            _lib.check_h(H, H.L.jb200_seulex_step_dual(H.h, state.M, state.atol, state.rtol, state.KMAXX, self.quirks, _ptr(ht),
                                                       _ptr(state.y), _ptr(state.x), _ptr(state.theta), _ptr(state.k_targ),
                                                       _ptr(state.hnext), _ptr(state.hdid), _ptr(state.flags),
                                                       _ptr(state._yjac), _ptr(state.k), _ptr(state.nrej), _ptr(state.yd),
                                                       _ptr(state._sjac), state.yd.shape[2], _stream_ptr(dev)))
        return state


class Dopri853Te:
    """Explicit Dormand-Prince 8(5,3) stepper for non-stiff batches with the reference's object shape
    (/root/reference/src/python/dopri853te.py:4-5, 290-315):

        s = Dopri853Te(yy, dydxx, xx, atoll, rtoll, dens, derivs);  s.step(htry)  ->  s.y, s.x, s.dydx, s.hnext, s.hdid

    `derivs` must be a HamiltonianSystem (the device evaluates H's dynamics by dual numbers); `dydxx` is used as
    given, like the reference (None: evaluated here).  One step() = one accepted step of every trajectory on the
    GPU (`jb200_dopri853_step`); the reference's float32 coefficients, batch-size-dependent error norm and
    non-growing hnext are reproduced (csrc/dopri853.cuh).  `norm_M`: batch size of the error norm (default M),
    for sharded runs.  Dense output (`dens`) is not built."""

    def __init__(self, yy, dydxx, xx, atoll, rtoll, dens=False, derivs=None, norm_M=None):
        if not isinstance(derivs, HamiltonianSystem):
            raise TypeError("derivs must be a janus_ode_b200.HamiltonianSystem")
        if dens:
            raise NotImplementedError("dense output of DOPRI853 is not built")
        if not (isinstance(yy, torch.Tensor) and yy.is_cuda and yy.dtype == torch.float64 and yy.dim() == 2):
            raise _lib.JanusB200Error("Dopri853Te runs on CUDA float64 tensors [M,Ny] only (no CPU fallback)")
        self.derivs = derivs
        self.M, self.D = yy.shape
        if self.D != derivs.hamiltonian.Ny:
            raise ValueError(f"y must be [M,{derivs.hamiltonian.Ny}]")
        dev = yy.device
        self.y = yy.contiguous().clone()
        self.x = torch.as_tensor(xx, dtype=torch.float64, device=dev).expand(self.M).contiguous().clone()
        self.dydx = (derivs(self.x, self.y) if dydxx is None else dydxx.to(dev, torch.float64)).contiguous().clone()
        self.atol, self.rtol = float(atoll), float(rtoll)
        self.norm_M = int(norm_M or self.M)
        self.errold = torch.full((self.M,), 1.0e-4, dtype=torch.float64, device=dev)
        self.hnext = torch.zeros(self.M, dtype=torch.float64, device=dev)
        self.hdid = torch.zeros(self.M, dtype=torch.float64, device=dev)
        self.xold = self.x.clone()
        self.nrej = torch.zeros(self.M, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            self._handle = _get_handle(derivs.hamiltonian, OptionsTe(), derivs.costate_sign, dev)
        self.gpu_launches = 0

    def step(self, htry):
        dev = self.y.device
        h = torch.as_tensor(htry, dtype=torch.float64, device=dev).expand(self.M).contiguous()
        self.xold = self.x.clone()
        H = self._handle
        with torch.cuda.device(dev):
            if getattr(self, "yd", None) is not None:
                _lib.check_h(H, H.L.jb200_dopri853_step_dual(H.h, self.M, self.norm_M, self.atol, self.rtol,
                                                            1 if self.gpu_launches else 0, _ptr(self.x), _ptr(self.y),
                                                            _ptr(self.dydx), _ptr(self.errold), _ptr(self.hnext), _ptr(h),
                                                            _ptr(self.hdid), _ptr(self.nrej), _ptr(self.yd),
                                                            self.yd.shape[2], _stream_ptr(dev)))
            else:
                _lib.check_h(H, H.L.jb200_dopri853_step(H.h, self.M, self.norm_M, self.atol, self.rtol,
                                                       1 if self.gpu_launches else 0, _ptr(self.x),
                                                       _ptr(self.y), _ptr(self.dydx), _ptr(self.errold), _ptr(self.hnext),
                                                       _ptr(h), _ptr(self.hdid), _ptr(self.nrej), _stream_ptr(dev)))
        self.gpu_launches += 1
        return self


class Dopri853TeD(Dopri853Te):
    """Dopri853Te with a dual part, shaped like the reference's C++ `Dopri853TeD` (TensorDual state,
    /root/reference/src/cpp/dopri853ted.hpp:26-734): `yd [M,Ny,Nd]` (default: identity, Nd = Ny <= 8) is propagated across every
    accepted step by differentiating the stages (`jb200_dopri853_step_dual`); `s.y` and all step sizes are those of Dopri853Te
    bit for bit."""

    def __init__(self, yy, dydxx, xx, atoll, rtoll, dens=False, derivs=None, norm_M=None, yd=None):
        super().__init__(yy, dydxx, xx, atoll, rtoll, dens, derivs, norm_M)
        if yd is None:
            yd = torch.eye(self.D, dtype=torch.float64, device=self.y.device).repeat(self.M, 1, 1)
        if yd.dim() != 3 or yd.shape[0] != self.M or yd.shape[1] != self.D or not (1 <= yd.shape[2] <= 8):
            raise ValueError(f"yd must be [M,{self.D},Nd] with 1 <= Nd <= 8")
        self.yd = yd.to(self.y.device, torch.float64).contiguous().clone()


class RadauTeD(RadauTe):
    """Radau IIA with forward sensitivities, shaped like the reference's dual-number solver
    (`RadauTeD r(dyns_dual, jac_dual, tspan, y, options, params); r.solve(); r.y.r, r.y.d`,
    /root/reference/src/cpp/radauted.hpp, tests/cpp/ode_solver_test.cpp:620-770):

        r = RadauTeD(dyns, dyns.jacobian, tspan, y0, options, y0d=None);  r.solve()  ->  r.y [M,Ny], r.yd [M,Ny,Ny]

    y0d [M,Ny,Ny] is the dual part of the initial state (None: identity, i.e. yd = d y(tfinal) / d y0).  The
    sensitivity of the final state to the final time is the dynamics there: `r.dydtf = dyns(t, r.y)`.
    `par_dirs=[k, ...]` (at most 4) adds parameter directions: `r.ypar [M,Ny,len(par_dirs)]` = d y(tfinal) / d par[k] for
    the parameters of the Hamiltonian -- the reference appends the parameter to the state with zero dynamics and seeds its
    dual part (ode_solver_test.cpp:630-640, y = [y1, y2, mu]).  `r.yd_full()` returns the reference test's layout
    [M, Ny, Ny + len(par_dirs) + 1]: state directions, parameter directions, final time.  CUDA tensors only;
    thread-resident Hamiltonians (built-in or user plugins with N <= 2)."""

    def __init__(self, dyns, jac, tspan, y0, options=None, y0d=None, par_dirs=()):
        super().__init__(dyns, jac, tspan, y0, options)
        self.par_dirs = tuple(int(k) for k in par_dirs)
        if len(self.par_dirs) > 4:
            raise ValueError("at most 4 parameter directions per solve")
        self.ypar = None
        if not y0.is_cuda:
            raise _lib.JanusB200Error("RadauTeD runs on CUDA tensors only (no CPU fallback)")
        if self.ntspan != 2:
            raise NotImplementedError("RadauTeD: tspan must be [M,2]")
        if y0d is not None and tuple(y0d.shape) != (self.M, self.Ny, self.Ny):
            raise ValueError(f"y0d must be [M,{self.Ny},{self.Ny}]")
        self.y0d = None if y0d is None else y0d.to(y0.device, torch.float64).contiguous()
        self.yd = None
        self.sens_iters = None

    def solve(self):
        H, M, ny, dev = self._handle, self.M, self.Ny, self.device
        y_out = torch.empty(M, ny, dtype=torch.float64, device=dev)
        S_out = torch.empty(M, ny, ny, dtype=torch.float64, device=dev)
        t_out = torch.empty(M, dtype=torch.float64, device=dev)
        h_out = torch.empty(M, dtype=torch.float64, device=dev)
        res = torch.empty(M, dtype=torch.int32, device=dev)
        stats = torch.empty(M, 8, dtype=torch.int64, device=dev)
        nstg = torch.empty(M, dtype=torch.int32, device=dev)
        work = torch.empty(M, dtype=torch.int64, device=dev)
        it = torch.empty(M, dtype=torch.int32, device=dev)
        npd = len(self.par_dirs)
        Sp = torch.empty(M, npd, ny, dtype=torch.float64, device=dev) if npd else None
        pidx = (C.c_int32 * max(npd, 1))(*self.par_dirs)
        with torch.cuda.device(dev):
            _lib.check_h(H, H.L.jb200_solve_sens_par(H.h, _ptr(self.y), _ptr(self.tspan), _ptr(self.y0d), npd, pidx, M,
                                                    _ptr(y_out), _ptr(S_out), _ptr(Sp), _ptr(t_out), _ptr(h_out), _ptr(res),
                                                    _ptr(stats), _ptr(nstg), _ptr(work), _ptr(it), _stream_ptr(dev)))
        self.y, self.yd, self.t, self.h, self.nstg, self.work, self.sens_iters = y_out, S_out, t_out, h_out, nstg, work, it
        self.ypar = Sp.transpose(1, 2).contiguous() if npd else None      # [M, Ny, n_par]
        self.stats = StatsTe.from_tensor(stats)
        self.gpu_launches = 1
        self.res = res.to(torch.float64)
        self.dydtf = self.dyns(self.t, self.y)
        return self.res

    def yd_full(self):
        """[M, Ny, Ny + n_par + 1]: d y(tfinal) / d (y0 directions, parameters, final time) -- the column layout of the
        reference's SensitivityTest (its N = D + 1 dual variables)."""
        cols = [self.yd] + ([self.ypar] if self.ypar is not None else []) + [self.dydtf.unsqueeze(2)]
        return torch.cat(cols, 2)
