"""Scalar (one trajectory at a time) CPU restatement of the reference's RADAU stepper.

TEST INFRASTRUCTURE ONLY.  Nothing in janus_ode_b200/ may import this module; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.

PARITY STATUS: "parity unpinned" for final states / step counts of a WHOLE Radau solve --
the reference's RadauTe / RadauTeD cannot be compiled or imported in this environment (they
need the un-vendored janus-tensor-dual / janus-linear headers, Eigen3, SUNDIALS, PETSc) and
the reference ships no golden vectors for a solve.  What IS pinned:
  * against LIVE reference code run in the build container (tests/golden/make_reference_pins.py,
    tests/test_reference_pins.py): solvrad() == the reference's Python solve_radau, estrad() ==
    its estrad incl. the First-step retry, the C / Dd literals and gamma, alpha +- i beta ==
    its radau_tables (src/python/odesolvers/src/odesolvers/{solvrad,estrad,radau_tables}.py);
  * the coefficient tables across radau.m and radaute_impl.hpp (oracle/tools/gen_tables.py), the
    closed-form Hamiltonian fixtures of the reference's GTest file (tests/test_oracle_ham.py),
    tolerance-level agreement with SciPy Radau on the reference's own cross-check scripts
    (tests/test_oracle_scipy.py);
  * sens_step(): the reference's own criterion for RadauTeD (finite differences of the solver,
    tests/cpp/ode_solver_test.cpp:620-770) and the exact harmonic-oscillator sensitivities.

What is restated, line by line:
  control flow      /root/reference/src/matlab/radau.m:630-1256 (the scalar truth the C++
                    was transcribed from), == masked form src/cpp/radauted_impl.hpp:620-1645
                    and src/cpp/radaute_impl.hpp:537-1345
  DecomRC           src/cpp/radaute_impl.hpp:1502-1661  (radau.m:1349-1393)
  Solvrad           src/cpp/radaute_impl.hpp:1721-1897  (radau.m:1395-1431)
  Estrad            src/cpp/radaute_impl.hpp:1902-1986  (radau.m:1432-1459)
  tables            src/cpp/radaute_impl.hpp:2187-2644  -> oracle/radau_tables.py
Arithmetic *forms* follow the batched C++ where it is specific (ValP * (1/h), Dd * (1/h),
err ** (1/NbrStg + 1), LAPACK zgetrf/zgetrs on complex128 for every block).

Deliberate deviations from RadauTe's batch-coupled behaviour:
  Q3  controller runs per trajectory (RadauTeD `any()` / MATLAB semantics), not `all()`
  Q4  Newt > Nit breaks out of the Newton loop (radau.m:915-925)
  Q6  no |h| <= hmin abort (C++ has none); "0.1|h| <= |t| eps" -> res = 1 is kept
  Q7  MaxStep ignored: hmax = tfinal - t0 (radaute_impl.hpp:227)
  Q8  Gustafsson: per trajectory; by default RadauTe's gate (limiter also on order-change steps,
      radaute_impl.hpp:1092-1100), RadauTeQuirks=0 -> radau.m:1066-1076
  Q9  Reject is cleared at the top of every Newton iteration (radau.m:913) -- reproduced,
      so Estrad's retry fires on First only and the Reject limiter on accept is dead code.
Counters (defined here, the reference is inconsistent between files): MATLAB semantics,
radau.m:684,733,757,829,942,953,1056,1195,1206,1250,917-918,981-982,992, + Estrad :1454.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla

from .radau_tables import TABLES

EPS = 2.220446049250313e-16


@dataclass
class RadauOptions:
    """Field-for-field mirror of OptionsTe (src/cpp/radaute.hpp:52-96) -- solver knobs only."""
    RelTol: float | np.ndarray = 1.0e-3
    AbsTol: float | np.ndarray = 1.0e-6
    InitialStep: float = 1.0e-2
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
    RadauTeQuirks: int = 1      # bit 0: quirk Q8, Gustafsson limiter not gated by ChangeFlag (radaute_impl.hpp:1092-1100)

    def validate(self):
        """Option checks of the constructor, src/cpp/radaute_impl.hpp:65-214 (raise, not exit)."""
        if np.any(np.asarray(self.AbsTol) < 0):
            raise ValueError("AbsTol must be a positive number")
        if np.any(np.asarray(self.RelTol) < 10 * EPS):
            raise ValueError("Relative tolerance are too small.")
        if not (0 <= self.JacRecompute < 1.0):
            raise ValueError("JacRecompute must be in [0,1)")
        if self.MaxNbrNewton < 4:
            raise ValueError("MaxNbrNewton integer >= 4")
        for nm in ("NbrStg", "MinNbrStg", "MaxNbrStg"):
            if getattr(self, nm) not in (1, 3, 5, 7):
                raise ValueError(f"{nm} must be 1, 3, 5 or 7")
        if self.NbrStg < self.MinNbrStg or self.NbrStg > self.MaxNbrStg:
            raise ValueError("NbrStg must be between MinNbrStg and MaxNbrStg")
        if not (0.001 < self.Safe < 1):
            raise ValueError("Curious input for Safe")
        if self.Quot1 > 1 or self.Quot2 < 1:
            raise ValueError("Curious input for Quot1/Quot2")
        if self.FacL > 1.0 or self.FacR < 1.0:
            raise ValueError("Curious input for FacL/FacR")
        if self.Vitu < 0 or self.Vitd < 0 or self.hhou < 1 or self.hhod < 0:
            raise ValueError("Curious input for Vitu/Vitd/hhou/hhod")


@dataclass
class RadauResult:
    y: np.ndarray
    t: float
    h: float
    res: int
    stats: dict
    nstg: int
    h_accept: list = field(default_factory=list)
    t_accept: list = field(default_factory=list)
    y_accept: list = field(default_factory=list)
    nstg_accept: list = field(default_factory=list)
    newt_accept: list = field(default_factory=list)
    sens: np.ndarray | None = None     # forward sensitivities d y / d (directions of sens0), [Ny, Nd]
    sens_iters: int = 0


STAT_NAMES = ("FcnNbr", "JacNbr", "DecompNbr", "SolveNbr", "StepNbr", "AccptNbr",
              "StepRejNbr", "NewtRejNbr")


def _tables(s):
    t = TABLES[s]
    return (np.array(t["T"]), np.array(t["TI"]), np.array(t["C"]),
            np.array(t["ValP"]), np.array(t["Dd"]))


def _nit(s, MaxNbrNewton):
    # radau.m:616-629
    return {1: MaxNbrNewton - 3, 3: MaxNbrNewton, 5: MaxNbrNewton + 5, 7: MaxNbrNewton + 10}[s]


def decomrc(h, ValP, Jac):
    """src/cpp/radaute_impl.hpp:1502-1625 (RealYN): LU of gamma/h I - J and (alpha+i beta)/h I - J."""
    n = Jac.shape[0]
    s = len(ValP)
    valp = ValP * (1.0 / h)                       # :1516  einsum(ValP, h.reciprocal())
    eye = np.eye(n)
    B0 = -Jac + eye * valp[0]                     # :1535
    lus = [sla.lu_factor(B0.astype(np.complex128), check_finite=False)]   # :1553-1558
    for q in range(1, (s - 1) // 2 + 1):          # :1580-1616
        q2, q3 = 2 * q, 2 * q + 1
        B_r = eye * valp[q2 - 1] - Jac
        B_i = eye * valp[q3 - 1]
        lus.append(sla.lu_factor(B_r + 1j * B_i, check_finite=False))
    return lus


def solvrad(z, w, ValP, h, lus):
    """src/cpp/radaute_impl.hpp:1721-1868.  z, w are [Ny, s]; returns new z."""
    s = len(ValP)
    valp = ValP * (1.0 / h)                       # :1731
    z = z.copy()
    z[:, 0] = z[:, 0] - valp[0] * w[:, 0]         # :1737-1738
    z[:, 0] = sla.lu_solve(lus[0], z[:, 0].astype(np.complex128), check_finite=False).real
    for q in range(1, (s - 1) // 2 + 1):          # :1792-1857
        q2, q3 = 2 * q, 2 * q + 1
        z2 = z[:, q2 - 1] - valp[q2 - 1] * w[:, q2 - 1] + valp[q3 - 1] * w[:, q3 - 1]
        z3 = z[:, q3 - 1] - valp[q3 - 1] * w[:, q2 - 1] - valp[q2 - 1] * w[:, q3 - 1]
        sol = sla.lu_solve(lus[q], z2 + 1j * z3, check_finite=False)
        z[:, q2 - 1] = sol.real
        z[:, q3 - 1] = sol.imag
    return z


def estrad(z, Dd, h, lus, Scal, f0, First, Reject, f, t, y, stats):
    """src/cpp/radaute_impl.hpp:1902-1986 == radau.m:1432-1459."""
    n = len(y)
    SqrtNy = math.sqrt(n)
    Ddoh = Dd * (1.0 / h)                         # :1908
    temp = z @ Ddoh                               # :1909
    err_v = sla.lu_solve(lus[0], (f0 + temp).astype(np.complex128), check_finite=False).real
    err = math.sqrt(float(np.sum((err_v / Scal) ** 2)))
    err = max(err / SqrtNy, 1e-10)                # :1945
    if err >= 1 and (First or Reject):            # :1948
        err_v = f(t, y + err)                     # :1951  (scalar err added to every component)
        stats["FcnNbr"] += 1
        err_v = sla.lu_solve(lus[0], (err_v + temp).astype(np.complex128), check_finite=False).real
        err = math.sqrt(float(np.sum((err_v / Scal) ** 2)))
        err = max(err / SqrtNy, 1e-10)
    return err


def sens_step(jac, t, y, z, h, C, T, TI, ValP, lus, Scal, Sm, Nit, FNewt):
    """Forward sensitivities across one accepted step (csrc/radau_core.cuh sens_step; the dual part of
    RadauTeD's stage equations, src/cpp/radauted_impl.hpp:1746-1936): the linear system
        Zs_i = h sum_j a_ij J(y + z_j) (S + Zs_j)
    solved by the state's simplified-Newton iteration with the same factors, J at the converged stage points.
    Sm [Ny, Nd] -> returns the updated Sm and the iterations spent."""
    Ny, s = z.shape
    Jq = [np.asarray(jac(t + C[q] * h, y + z[:, q]), dtype=np.float64) for q in range(s)]
    out = Sm.copy()
    iters = 0
    for d in range(Sm.shape[1]):
        Sd = Sm[:, d]
        rs = 1.0 / max(1.0, float(np.max(np.abs(Sd))))
        Zs = np.zeros((Ny, s))
        Ws = np.zeros((Ny, s))
        nrm_old = 0.0
        for it in range(Nit + 10):
            F = np.stack([Jq[q] @ (Sd + Zs[:, q]) for q in range(s)], axis=1)
            zz = F @ TI.T                                        # zz_j = sum_q TI[j,q] F_q
            dW = solvrad(zz, Ws, ValP, h, lus)
            Ws = Ws + dW
            Zs = Ws @ T.T
            dZ = dW @ T.T
            nrm = math.sqrt(float(np.sum((dZ / Scal[:, None] * rs) ** 2))) / math.sqrt(s * Ny)
            iters += 1
            if not nrm > 0.1 * FNewt:
                break
            if it >= 1:                      # linear iteration: theta / (1 - theta) * |increment| bounds the remainder
                theta = nrm / nrm_old
                if theta < 0.99 and theta / (1.0 - theta) * nrm <= 0.1 * FNewt:
                    break
            nrm_old = nrm
        out[:, d] = Sd + Zs[:, s - 1]
    return out, iters


def radau_solve(f, jac, t0, tfinal, y0, opts: RadauOptions | None = None, record=False, sens0=None):
    """Integrate y' = f(t,y) from t0 to tfinal for ONE trajectory.

    f(t, y) -> ndarray[Ny];  jac(t, y) -> ndarray[Ny, Ny].
    Returns RadauResult; res: 0 ok, 1 step too small, 2 NaN/inf, 3 MaxNbrStep exceeded.
    """
    op = opts or RadauOptions()
    getattr(op, "validate", lambda: None)()   # any object with OptionsTe's field names works
    y = np.array(y0, dtype=np.float64).copy()
    Ny = y.shape[0]
    t = float(t0)
    tfinal = float(tfinal)
    PosNeg = math.copysign(1.0, tfinal - t0)
    if not (PosNeg * (tfinal - t0) > 0):
        raise ValueError("Time vector must be strictly monotonic")
    RelTol = np.broadcast_to(np.asarray(op.RelTol, dtype=np.float64), (Ny,)).copy()
    AbsTol = np.broadcast_to(np.asarray(op.AbsTol, dtype=np.float64), (Ny,)).copy()
    stats = {k: 0 for k in STAT_NAMES}
    out = RadauResult(y=y, t=t, h=0.0, res=0, stats=stats, nstg=op.NbrStg)
    Sm = None if sens0 is None else np.array(sens0, dtype=np.float64).copy()   # forward sensitivities [Ny, Nd]
    out.sens, out.sens_iters = Sm, 0

    Thet = op.JacRecompute
    Safe, Quot1, Quot2 = op.Safe, op.Quot1, op.Quot2
    FacL = 1.0 / op.FacL                            # radaute_impl.hpp:185
    FacR = 1.0 / op.FacR                            # :191
    MinNbrStg, MaxNbrStg, NbrStg = op.MinNbrStg, op.MaxNbrStg, op.NbrStg
    Variab = (MaxNbrStg - MinNbrStg) != 0          # radau.m:638
    T, TI, C, ValP, Dd = _tables(NbrStg)
    Nit = _nit(NbrStg, op.MaxNbrNewton)

    # radau.m:631-696 / radaute_impl.hpp:406-441
    UnExpStepRej = UnExpNewtRej = Keep = ChangeFlag = False
    ChangeNbr = 0
    Theta = Thetat = 0.0
    hmax = tfinal - t0                              # :227 (Q7)
    hmaxn = min(abs(hmax), abs(tfinal - t0))
    h = float(op.InitialStep)
    if abs(h) <= 10 * EPS:
        h = 1e-6
    h = PosNeg * min(abs(h), hmaxn)
    h_old = h
    hopt = h
    if (t + h * 1.0001 - tfinal) * PosNeg >= 0:
        h = tfinal - t
        Last = True
    else:
        Last = False
    FacConv = 1.0
    Reject = False
    First = True
    ExpmNs = (NbrStg + 1.0) / (2.0 * NbrStg)
    QuotTol = AbsTol / RelTol
    RelTol1 = 0.1 * RelTol ** ExpmNs
    AbsTol1 = RelTol1 * QuotTol
    Scal = AbsTol1 + RelTol1 * np.abs(y)
    f0 = np.asarray(f(t, y), dtype=np.float64)
    stats["FcnNbr"] += 1
    if np.any(np.isnan(f0)):
        out.res = 2
        return out
    NeedNewJac = True
    NeedNewLU = True
    SqrtStgNy = math.sqrt(NbrStg * Ny)
    Newt = 0
    z = np.zeros((Ny, NbrStg))
    w = np.zeros((Ny, NbrStg))
    cont = np.zeros((Ny, NbrStg))
    Jac = None
    lus = None
    hacc = erracc = 0.0
    thqold = OldNrm = 0.0
    err = 0.0
    finished = False

    # ------- MAIN LOOP  radau.m:712
    while stats["StepNbr"] <= op.MaxNbrStep and PosNeg * t < PosNeg * tfinal:
        stats["StepNbr"] += 1
        FacConv = max(FacConv, EPS) ** 0.8          # :735
        if NeedNewJac:                              # :737-764
            Jac = np.asarray(jac(t, y), dtype=np.float64)
            if np.any(np.isnan(Jac)):
                out.res = 2
                break
            stats["JacNbr"] += 1
            NeedNewJac = False
            NeedNewLU = True
        if Variab and not Keep:                     # :766-824
            ChangeNbr += 1
            NbrStgNew = NbrStg
            hquot = h / h_old
            Thetat = min(10.0, max(Theta, Thetat * 0.5))
            if Newt > 1 and Thetat <= op.Vitu and hquot < op.hhou and hquot > op.hhod:
                NbrStgNew = min(MaxNbrStg, NbrStg + 2)
            if Thetat >= op.Vitd or UnExpStepRej:
                NbrStgNew = max(MinNbrStg, NbrStg - 2)
            if ChangeNbr >= 1 and UnExpNewtRej:
                NbrStgNew = max(MinNbrStg, NbrStg - 2)
            if ChangeNbr <= 10:
                NbrStgNew = min(NbrStg, NbrStgNew)
            ChangeFlag = NbrStg != NbrStgNew
            UnExpNewtRej = False
            UnExpStepRej = False
            if ChangeFlag:
                NbrStg = NbrStgNew
                ChangeNbr = 1
                ExpmNs = (NbrStg + 1.0) / (2.0 * NbrStg)
                RelTol1 = 0.1 * RelTol ** ExpmNs
                AbsTol1 = RelTol1 * QuotTol
                Scal = AbsTol1 + RelTol1 * np.abs(y)
                NeedNewLU = True
                SqrtStgNy = math.sqrt(NbrStg * Ny)
                T, TI, C, ValP, Dd = _tables(NbrStg)
                Nit = _nit(NbrStg, op.MaxNbrNewton)
        if NeedNewLU:                               # :827-844 (U_Sing never set, Q6)
            lus = decomrc(h, ValP, Jac)
            stats["DecompNbr"] += 1
            NeedNewLU = False
        if Variab and Keep:                         # :846-853 (the inner assignment there
            Keep = False                            #  writes a misspelt variable: no effect)
            ChangeNbr += 1
            if ChangeNbr >= 10 and NbrStg < MaxNbrStg:
                NeedNewJac = False                  # already false on this path
        if 0.1 * abs(h) <= abs(t) * EPS:            # :855 / radaute_impl.hpp:711-719
            out.res = 1
            break
        ExpmNs = (NbrStg + 1.0) / (2.0 * NbrStg)    # :858-862
        QuotTol = AbsTol / RelTol
        RelTol1 = 0.1 * RelTol ** ExpmNs
        AbsTol1 = RelTol1 * QuotTol
        Scal = AbsTol1 + RelTol1 * np.abs(y)
        if First or op.Start_Newt or ChangeFlag:    # :872-891
            z = np.zeros((Ny, NbrStg))
            w = np.zeros((Ny, NbrStg))
            cont = np.zeros((Ny, NbrStg))
        else:
            hquot = h / h_old
            cq = C * hquot
            for q in range(NbrStg):
                zq = (cq[q] - C[0] + 1.0) * cont[:, NbrStg - 1]
                for q1 in range(2, NbrStg + 1):
                    zq = (zq + cont[:, NbrStg - q1]) * (cq[q] - C[q1 - 1] + 1.0)
                z[:, q] = zq
            # w(n,:) = sum_q TI_ref(q,:) z(n,q)  -> w_j = sum_q TI[j][q] z_q, q ascending
            w = np.zeros((Ny, NbrStg))
            for j in range(NbrStg):
                acc = TI[j, 0] * z[:, 0]
                for q in range(1, NbrStg):
                    acc = acc + TI[j, q] * z[:, q]
                w[:, j] = acc
        # ------- simplified Newton  :897-1018
        mrt = float(np.min(RelTol1))
        FNewt = max(10 * EPS / mrt, min(0.03, mrt ** (1.0 / ExpmNs - 1.0)))
        if NbrStg == 1:
            FNewt = max(10 * EPS / mrt, 0.03)
        FacConv = max(FacConv, EPS) ** 0.8
        Theta = abs(Thet)
        Newt = 0
        newton_failed = False
        nan_fail = False
        while True:
            Reject = False                          # :913 (Q9)
            Newt += 1
            if Newt > Nit:                          # :915-925 (Q4: break)
                UnExpStepRej = True
                stats["StepRejNbr"] += 1
                stats["NewtRejNbr"] += 1
                h = h * 0.5
                Reject = True
                Last = False
                NeedNewLU = True
                newton_failed = True
                break
            fq = np.empty((Ny, NbrStg))
            for q in range(NbrStg):                 # :932-941
                fq[:, q] = f(t + C[q] * h, y + z[:, q])
            if np.any(np.isnan(fq)):
                nan_fail = True
                break
            stats["FcnNbr"] += NbrStg
            for j in range(NbrStg):                 # :943-948  z = TI f
                acc = TI[j, 0] * fq[:, 0]
                for q in range(1, NbrStg):
                    acc = acc + TI[j, q] * fq[:, q]
                z[:, j] = acc
            z = solvrad(z, w, ValP, h, lus)         # :952
            stats["SolveNbr"] += 1
            NewNrm = 0.0
            for q in range(NbrStg):                 # :955-959
                NewNrm += math.sqrt(float(np.sum((z[:, q] / Scal) ** 2)))
            NewNrm = NewNrm / SqrtStgNy
            if not math.isfinite(NewNrm):
                nan_fail = True
                break
            if 1 < Newt < Nit:                      # :962-996
                thq = NewNrm / OldNrm
                if Newt == 2:
                    Theta = thq
                else:
                    Theta = math.sqrt(thq * thqold)
                thqold = thq
                if Theta < 0.99:
                    FacConv = Theta / (1.0 - Theta)
                    dyth = FacConv * NewNrm * Theta ** (Nit - 1 - Newt) / FNewt
                    if dyth >= 1:
                        qNewt = max(1.0e-4, min(20.0, dyth))
                        hhfac = 0.8 * qNewt ** (-1.0 / (4.0 + Nit - 1 - Newt))
                        h = hhfac * h
                        Reject = True
                        Last = False
                        UnExpNewtRej = hhfac <= 0.5
                        stats["NewtRejNbr"] += 1
                        stats["StepRejNbr"] += 1
                        NeedNewLU = True
                        newton_failed = True
                        break
                else:
                    h = h * 0.5
                    Reject = True
                    Last = False
                    UnExpStepRej = True
                    stats["StepRejNbr"] += 1
                    NeedNewLU = True
                    newton_failed = True
                    break
            OldNrm = max(NewNrm, EPS)               # :1001
            w = w + z                               # :1002
            for j in range(NbrStg):                 # :1007-1012  z = T w
                acc = T[j, 0] * w[:, 0]
                for q in range(1, NbrStg):
                    acc = acc + T[j, q] * w[:, q]
                z[:, j] = acc
            if FacConv * NewNrm > FNewt:            # :1014
                continue
            break
        if nan_fail:
            out.res = 2
            break
        if newton_failed:
            continue                                # :1020-1022
        # ------- error estimate  :1033
        err = estrad(z, Dd, h, lus, Scal, f0, First, Reject, f, t, y, stats)
        if not math.isfinite(err):
            out.res = 2
            break
        fac = min(Safe, (2.0 * Nit + 1.0) / (2.0 * Nit + Newt))   # :1038
        exponent = 1.0 / NbrStg + 1.0               # :1041 (sic)
        quot = max(FacR, min(FacL, err ** exponent / fac))
        hnew = h / quot
        if err < 1:                                 # ------- accepted  :1054
            First = False
            stats["AccptNbr"] += 1
            gus_upd = op.Gustafsson and not ChangeFlag          # radau.m:1066-1076
            gus_lim = op.Gustafsson if (op.RadauTeQuirks & 1) else gus_upd   # radaute_impl.hpp:1092-1100 (Q8)
            if gus_lim:
                if stats["AccptNbr"] > 1:
                    facgus = (hacc / h) * (err * err / erracc) ** (1.0 / (NbrStg + 1.0)) / Safe
                    facgus = max(FacR, min(FacL, facgus))
                    quot = max(quot, facgus)
                    hnew = h / quot
            if gus_upd:
                hacc = h
                erracc = max(1.0e-2, err)
            if Sm is not None:
                Sm, its = sens_step(jac, t, y, z, h, C, T, TI, ValP, lus, Scal, Sm, Nit, FNewt)
                out.sens, out.sens_iters = Sm, out.sens_iters + its
            h_old = h
            t = t + h
            Scal = AbsTol1 + RelTol1 * np.abs(y)    # :1080 (y before the update; dead value)
            y = y + z[:, NbrStg - 1]                # :1091
            cont[:, NbrStg - 1] = z[:, 0] / C[0]    # :1093-1107
            for q in range(1, NbrStg):
                Fact = 1.0 / (C[NbrStg - q - 1] - C[NbrStg - q])
                cont[:, q - 1] = (z[:, NbrStg - q - 1] - z[:, NbrStg - q]) * Fact
            for jj in range(2, NbrStg + 1):
                for k in range(NbrStg, jj - 1, -1):
                    if NbrStg - k == 0:
                        Fact = 1.0 / (-C[jj - 1])
                    else:
                        Fact = 1.0 / (C[NbrStg - k - 1] - C[NbrStg - k + jj - 1])
                    cont[:, k - 1] = (cont[:, k - 1] - cont[:, k - 2]) * Fact
            if record:
                out.h_accept.append(h)
                out.t_accept.append(t)
                out.y_accept.append(y.copy())
                out.nstg_accept.append(NbrStg)
                out.newt_accept.append(Newt)
            NeedNewJac = True                       # :1192
            if Last:                                # :1193-1197
                h = hopt
                stats["StepRejNbr"] += 1
                finished = True
                break
            f0 = np.asarray(f(t, y), dtype=np.float64)      # :1198-1206
            stats["FcnNbr"] += 1
            if np.any(np.isnan(f0)):
                out.res = 2
                break
            hnew = PosNeg * min(abs(hnew), abs(hmaxn))
            hopt = PosNeg * min(abs(h), abs(hnew))
            if Reject:                              # dead (Q9)
                hnew = PosNeg * min(abs(hnew), abs(h))
            Reject = False
            if (t + hnew / Quot1 - tfinal) * PosNeg >= 0:   # :1214
                h = tfinal - t
                Last = True
            else:
                qt = hnew / h
                if Theta <= Thet and qt >= Quot1 and qt <= Quot2:
                    Keep = True
                    NeedNewJac = False
                    NeedNewLU = False
                    continue
                h = hnew
            NeedNewLU = True
            if Theta <= Thet:
                NeedNewJac = False
        else:                                       # ------- rejected  :1234-1253
            Reject = True
            Last = False
            if First:
                h = h / 10.0
            else:
                h = hnew
            if stats["AccptNbr"] >= 1:
                stats["StepRejNbr"] += 1
            NeedNewLU = True

    if out.res == 0 and not finished and PosNeg * t < PosNeg * tfinal:
        out.res = 3                                 # MaxNbrStep exceeded (radau.m:1287)
    out.y, out.t, out.h, out.nstg = y, t, h, NbrStg
    return out
