"""OptionsTe / StatsTe mirrors (reference: /root/reference/src/cpp/radaute.hpp:52-108)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from ._lib import NYMAX, jb200_opts

STAT_NAMES = ("FcnNbr", "JacNbr", "DecompNbr", "SolveNbr", "StepNbr", "AccptNbr", "StepRejNbr",
              "NewtRejNbr")


class LinearEvents:
    """EventsFcn of the accelerated path (reference: EventFnType, src/cpp/radaute.hpp:44-47 -- value, isterminal,
    direction; zero location src/matlab/radau.m:1494-1618): K <= 4 event functions E_k(t, y) = A[k] . y - b[k].
    stop[k]: terminal event; slope[k]: +1 only rising zeros, -1 only falling, 0 both.  A kernel cannot call back into
    Python, so the family is closed; a component crossing a level (the reference's examples) is A = unit vector."""

    def __init__(self, A, b, stop=None, slope=None):
        A = np.atleast_2d(np.asarray(_to_np(A), dtype=np.float64))
        K = A.shape[0]
        if not 1 <= K <= 4:
            raise ValueError("1 to 4 event functions")
        self.A = tuple(map(tuple, A.tolist()))
        self.b = tuple(np.broadcast_to(np.asarray(_to_np(b), dtype=np.float64), (K,)).tolist())
        self.stop = tuple(int(bool(v)) for v in np.broadcast_to(np.asarray(0 if stop is None else stop), (K,)))
        self.slope = tuple(int(np.sign(v)) for v in np.broadcast_to(np.asarray(0 if slope is None else slope), (K,)))

    def _key(self):
        return (self.A, self.b, self.stop, self.slope)

    def __hash__(self):
        return hash(self._key())

    def __eq__(self, other):
        return isinstance(other, LinearEvents) and self._key() == other._key()


@dataclass
class OptionsTe:
    """Same field names and defaults as the reference's OptionsTe (solver knobs).

    MaxStep is accepted and ignored exactly like the reference (hmax = tfinal - t0,
    src/cpp/radaute_impl.hpp:227).  EventsFcn takes a LinearEvents object.  MassFcn / OutputFcn / Complex are
    outside the accelerated path and raise if set.
    OutputCapacity > 0 (two-column tspan): the computed points are recorded like the reference's OutFlag 1 (Refine = 1)
    / OutFlag 2 (Refine > 1) into tout [M,cap], yout [M,cap,Ny], nout [M] -- the reference grows these buffers step by
    step for the whole batch; here the caller sizes them (nout > cap: too small).  EventsCapacity: events kept per row.
    """
    RelTol: object = 1.0e-3
    AbsTol: object = 1.0e-6
    InitialStep: float = 1.0e-2
    MaxStep: float = 1.0
    JacRecompute: float = 1.0e-3
    Start_Newt: bool = False
    MaxNbrNewton: int = 7
    NbrStg: int = 3
    MinNbrStg: int = 3
    MaxNbrStg: int = 7
    MaxNbrStep: int = 10000
    Safe: float = 0.9
    Quot1: float = 1.0
    Quot2: float = 1.2
    FacL: float = 0.2
    FacR: float = 8.0
    Vitu: float = 0.002
    Vitd: float = 0.8
    hhou: float = 1.2
    hhod: float = 0.8
    Gustafsson: bool = True
    # RadauTe deviations from its MATLAB twin that are reproduced (bit mask; 0 = radau.m semantics).  Bit 0 (quirk Q8):
    # the Gustafsson limiter also acts on accepted steps whose order just changed, src/cpp/radaute_impl.hpp:1092-1100.
    RadauTeQuirks: int = 1
    Complex: bool = False
    Refine: int = 1
    MassFcn: object = None
    EventsFcn: object = None
    OutputFcn: object = None
    OutputCapacity: int = 0
    EventsCapacity: int = 8

    def to_c(self, ny: int, costate_sign: int) -> jb200_opts:
        for nm in ("MassFcn", "OutputFcn"):
            if getattr(self, nm) is not None:
                raise NotImplementedError(f"{nm} is outside the accelerated Radau path")
        if self.EventsFcn is not None and not isinstance(self.EventsFcn, LinearEvents):
            raise NotImplementedError("EventsFcn must be a janus_ode_b200.LinearEvents (a kernel cannot call a Python function)")
        if self.Complex:
            raise NotImplementedError("Complex=true tables are outside the accelerated path")
        if self.Refine < 1:
            raise ValueError("Refine must be >= 1")
        o = jb200_opts()
        rt = np.atleast_1d(np.asarray(_to_np(self.RelTol), dtype=np.float64)).ravel()
        at = np.atleast_1d(np.asarray(_to_np(self.AbsTol), dtype=np.float64)).ravel()
        for name, v in (("RelTol", rt), ("AbsTol", at)):
            if v.size not in (1, ny):
                raise ValueError(f"{name} vector of length 1 or {ny}")
        n_tol = max(rt.size, at.size)
        rt = np.broadcast_to(rt, (n_tol,))
        at = np.broadcast_to(at, (n_tol,))
        if n_tol > NYMAX:
            raise ValueError("too many tolerance components")
        for i in range(n_tol):
            o.RelTol[i] = rt[i]
            o.AbsTol[i] = at[i]
        o.n_tol = n_tol
        for k in ("InitialStep", "MaxStep", "JacRecompute", "Safe", "Quot1", "Quot2", "FacL", "FacR", "Vitu",
                  "Vitd", "hhou", "hhod"):
            setattr(o, k, float(_to_np(getattr(self, k))))
        for k in ("MaxNbrNewton", "NbrStg", "MinNbrStg", "MaxNbrStg", "MaxNbrStep"):
            setattr(o, k, int(_to_np(getattr(self, k))))
        o.Start_Newt = int(bool(self.Start_Newt))
        o.Gustafsson = int(bool(self.Gustafsson))
        o.costate_sign = int(costate_sign)
        o.radaute_quirks = int(self.RadauTeQuirks)
        return o


def _to_np(v):
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    a = np.asarray(v)
    return a if a.ndim else a.item()


@dataclass
class StatsTe:
    """Per-trajectory counters, one tensor per field (src/cpp/radaute.hpp:98-108): int64 like the reference for device
    solves; solves on host tensors carry them as int32 over the PCIe link (jb200_solve_host_compact)."""
    FcnNbr: object = None
    JacNbr: object = None
    DecompNbr: object = None
    SolveNbr: object = None
    StepNbr: object = None
    AccptNbr: object = None
    StepRejNbr: object = None
    NewtRejNbr: object = None
    raw: object = field(default=None, repr=False)   # [M, 8]

    @classmethod
    def from_tensor(cls, st):
        return cls(*[st[:, i] for i in range(8)], raw=st)
