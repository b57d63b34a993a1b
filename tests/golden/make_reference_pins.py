#!/usr/bin/env python3
"""Golden vectors from the reference's OWN code, run here: the torch-only modules of its Python Radau package
(/root/reference/src/python/odesolvers/src/odesolvers/{solvrad,estrad,radau_tables}.py) are
imported and executed on seeded inputs; inputs and outputs go to tests/golden/ref_python_pins.npz.  The tests
(tests/test_reference_pins.py) then check oracle/radau_scalar.py against these vectors WITHOUT the reference
tree (it does not exist on the GPU box).  What this pins, against live reference code:
  * solve_radau   -- right-hand sides and real/complex block solves of Solvrad (sign and pairing of alpha, beta)
  * estrad        -- the error estimate incl. the First/Reject retry with the scalar err added to every component
  * radau_tables  -- C, Dd literals; the inverted eigenvalues of the Radau IIA matrix (gamma, alpha +- i beta)
Not pinned by any runnable reference code: whole solves (RadauTe needs janus-tensor-dual; radau.py needs petsc4py),
interp_radau.py, whose Horner order differs from radau.m:1291-1308 / radaute_impl.hpp:1413-1440 (the oracle
follows the latter two), and simplified_newton.py, whose TI product `(f.T @ TI.T).T` only type-checks for n == s.          Run from the repo root:  python tests/golden/make_reference_pins.py
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src/python/odesolvers/src")
from odesolvers import estrad as ref_estrad            # noqa: E402  (reference code)
from odesolvers import radau_tables as ref_tables      # noqa: E402
from odesolvers import solvrad as ref_solvrad          # noqa: E402

from oracle.radau_tables import TABLES                 # noqa: E402  (only to feed OUR tables to the reference functions)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_python_pins.npz")


def solvers(Jac, ValP, h):
    """LAPACK solve closures for (gamma/h I - J) and ((alpha + i beta)/h I - J), as solve_radau expects them."""
    n = Jac.shape[0]
    eye = torch.eye(n, dtype=torch.float64)
    fns = [lambda r, A=(ValP[0] / h) * eye - Jac: torch.linalg.solve(A.to(r.dtype), r)]
    for q in range(1, len(ValP), 2):
        A = ((ValP[q] + 1j * ValP[q + 1]) / h) * eye.to(torch.complex128) - Jac.to(torch.complex128)
        fns.append(lambda r, A=A: torch.linalg.solve(A, r.to(torch.complex128)))
    return fns


def vdp4(t, y):
    """stiff van der Pol in Hamiltonian form, reference sign (tests/cpp/ode_solver_test.cpp:204-223 shape), mu = 10"""
    mu = 10.0
    p1, p2, x1, x2 = y
    return torch.stack([p2 * (-2 * mu * x1 * x2 - 1), p1 + p2 * mu * (1 - x1 * x1), x2, mu * (1 - x1 * x1) * x2 - x1])


def main():
    g = torch.Generator().manual_seed(0)
    out = {}
    for s in (3, 5, 7):
        T = torch.tensor(TABLES[s]["T"], dtype=torch.float64)
        TI = torch.tensor(TABLES[s]["TI"], dtype=torch.float64)
        C = torch.tensor(TABLES[s]["C"], dtype=torch.float64)
        ValP = torch.tensor(TABLES[s]["ValP"], dtype=torch.float64)
        Dd = torch.tensor(TABLES[s]["Dd"], dtype=torch.float64)
        for n in (4, 16):
            k = f"s{s}_n{n}"
            Jac = torch.randn(n, n, generator=g, dtype=torch.float64) * 3.0
            h = float(10.0 ** (-2 + torch.rand(1, generator=g, dtype=torch.float64).item()))
            z = torch.randn(n, s, generator=g, dtype=torch.float64)
            w = torch.randn(n, s, generator=g, dtype=torch.float64)
            fns = solvers(Jac, ValP, h)
            out[k + "_Jac"], out[k + "_h"], out[k + "_z"], out[k + "_w"] = Jac.numpy(), h, z.numpy(), w.numpy()
            zz = z.clone()
            ref_solvrad.solve_radau(zz, w, ValP, h, fns, Mass=None, real_layout=True)
            out[k + "_solvrad"] = zz.numpy()
            # estrad: both branches
            y = torch.randn(n, generator=g, dtype=torch.float64)
            f0 = torch.randn(n, generator=g, dtype=torch.float64)
            Scal = 1e-3 * (1.0 + y.abs())
            odef = (lambda t, yy: torch.tanh(yy) * 2.0 - yy.flip(0))
            for tag, sc, first in (("small", Scal * 1e6, False), ("retry", Scal * 1e-3, True), ("noretry", Scal * 1e-3, False)):
                st = SimpleNamespace(FcnNbr=0)
                err, st = ref_estrad.estrad(zz, Dd, h, fns[0], None, sc, f0, First=first, Reject=False, Stat=st,
                                            OdeFcn=odef, t=0.3, y=y)
                out[f"{k}_estrad_{tag}"] = np.array([err, st.FcnNbr])
                out[f"{k}_estrad_{tag}_scal"] = sc.numpy()
            out[k + "_y"], out[k + "_f0"] = y.numpy(), f0.numpy()
        # tables of the reference's Python package
        Tr, TIr, Cr, ValPr, Ddr = getattr(ref_tables, f"coertv{s}")()
        out[f"s{s}_tab_C"], out[f"s{s}_tab_Dd"] = Cr.numpy(), Ddr.numpy()
        out[f"s{s}_tab_ValP"] = ValPr.numpy()           # complex: 1 / eig(A)
    np.savez(OUT, **out)
    print("written", OUT, len(out), "arrays")


if __name__ == "__main__":
    main()
