"""oracle/radau_scalar.py against golden vectors produced by the reference's OWN Python code
(tests/golden/make_reference_pins.py imports /root/reference/src/python/odesolvers/src/odesolvers/
{solvrad,estrad,radau_tables}.py in the build container and runs them on seeded inputs).  The reference tree is
not needed here.  Tolerances: the reference solves with LAPACK gesv on the assembled matrix, the oracle with
its own factorisation -> agreement to roundoff times the condition number."""
import os

import numpy as np
import pytest
import scipy.linalg as sla

from oracle import radau_scalar as rs
from oracle.radau_tables import TABLES

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_python_pins.npz"))


@pytest.mark.parametrize("s", [3, 5, 7])
@pytest.mark.parametrize("n", [4, 16])
def test_solvrad_matches_reference_solve_radau(s, n):
    k = f"s{s}_n{n}"
    ValP = np.array(TABLES[s]["ValP"])
    h = float(G[k + "_h"])
    lus = rs.decomrc(h, ValP, G[k + "_Jac"])
    z = rs.solvrad(G[k + "_z"], G[k + "_w"], ValP, h, lus)
    ref = G[k + "_solvrad"]
    assert np.allclose(z, ref, rtol=1e-10, atol=1e-12 * np.abs(ref).max())


@pytest.mark.parametrize("s", [3, 5, 7])
@pytest.mark.parametrize("n", [4, 16])
@pytest.mark.parametrize("tag,first", [("small", False), ("retry", True), ("noretry", False)])
def test_estrad_matches_reference_estrad(s, n, tag, first):
    k = f"s{s}_n{n}"
    ValP, Dd = np.array(TABLES[s]["ValP"]), np.array(TABLES[s]["Dd"])
    h = float(G[k + "_h"])
    lus = rs.decomrc(h, ValP, G[k + "_Jac"])
    stats = {"FcnNbr": 0}
    f = lambda t, y: np.tanh(y) * 2.0 - y[::-1]
    err = rs.estrad(G[k + "_solvrad"], Dd, h, lus, G[f"{k}_estrad_{tag}_scal"], G[k + "_f0"], first, False, f, 0.3,
                    G[k + "_y"], stats)
    ref_err, ref_fcn = G[f"{k}_estrad_{tag}"]
    assert err == pytest.approx(ref_err, rel=1e-9)
    assert stats["FcnNbr"] == int(ref_fcn)
    if tag == "retry":
        assert ref_fcn == 1          # the First-step retry branch (scalar err added to every component) was taken
    if tag == "small":
        assert ref_err < 1


@pytest.mark.parametrize("s", [3, 5, 7])
def test_tables_match_reference_python_tables(s):
    t = TABLES[s]
    assert np.allclose(G[f"s{s}_tab_C"], t["C"], rtol=0, atol=2e-16)
    assert np.allclose(G[f"s{s}_tab_Dd"], t["Dd"], rtol=1e-15, atol=0)
    # the reference computes 1/eig(A) of the Radau IIA matrix numerically; ours are Hairer's literals
    key = lambda v: sorted(v, key=lambda c: (round(c.real, 8), c.imag))     # conjugate pairs: equal real parts up to roundoff
    ref = np.array(key(np.asarray(G[f"s{s}_tab_ValP"]).astype(np.complex128)))
    ValP = np.array(t["ValP"])
    ours = [ValP[0] + 0j]
    for q in range(1, s, 2):
        ours += [ValP[q] + 1j * ValP[q + 1], ValP[q] - 1j * ValP[q + 1]]
    ours = np.array(key(np.array(ours)))
    assert np.allclose(ref, ours, rtol=1e-9)
