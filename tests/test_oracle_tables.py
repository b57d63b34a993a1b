"""Radau IIA tables: mathematical identities that pin the generated constants
(reference: src/cpp/radaute_impl.hpp:2187-2644 == src/matlab/radau.m:1621-2000)."""
import numpy as np
import pytest

from oracle.radau_tables import TABLES


def radau_A(C):
    s = len(C)
    CP = np.array([[c ** j for j in range(s)] for c in C])
    CQ = np.array([[c ** (j + 1) / (j + 1) for j in range(s)] for c in C])
    return CQ @ np.linalg.inv(CP)


@pytest.mark.parametrize("s", [1, 3, 5, 7])
def test_decomposition(s):
    t = TABLES[s]
    T, TI, V = np.array(t["T"]), np.array(t["TI"]), t["ValP"]
    L = np.zeros((s, s))
    L[0, 0] = V[0]
    for q in range((s - 1) // 2):
        i = 1 + 2 * q
        a, b = V[i], V[i + 1]
        L[i:i + 2, i:i + 2] = [[a, -b], [b, a]]
    Ainv = np.linalg.inv(radau_A(t["C"]))
    assert np.abs(T @ TI - np.eye(s)).max() < 1e-11
    assert np.abs(T @ L @ TI - Ainv).max() / np.abs(Ainv).max() < 1e-10


@pytest.mark.parametrize("s", [1, 3, 5, 7])
def test_nodes_and_error_coefficients(s):
    t = TABLES[s]
    C, Dd = np.array(t["C"]), np.array(t["Dd"])
    assert C[-1] == 1.0 and np.all(np.diff(C) > 0) and C[0] > 0
    # Radau IIA nodes: roots of d^{s-1}/dx^{s-1} (x^{s-1} (x-1)^s)
    p = np.poly1d([1.0])
    for _ in range(s - 1):
        p = p * np.poly1d([1.0, 0.0])
    for _ in range(s):
        p = p * np.poly1d([1.0, -1.0])
    for _ in range(s - 1):
        p = p.deriv()
    assert np.allclose(np.sort(p.roots.real), C, atol=1e-10)
    # Estrad on y' = const must give zero error: 1 + sum_q Dd_q c_q = 0
    assert abs(np.dot(Dd, C) + 1.0) < 1e-12


def test_generated_headers_match_python_tables():
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for path in ("oracle/c/radau_tables.h", "janus_ode_b200/csrc/radau_tables_t.cuh"):
        txt = open(os.path.join(root, path)).read()
        nums = set(float(x) for x in re.findall(r"-?\d+\.\d+(?:e[-+]?\d+)?", txt))
        for s in (3, 5, 7):
            for name in ("C", "Dd", "ValP", "T", "TI"):
                for v in np.array(TABLES[s][name]).ravel():
                    assert float(v) in nums, (path, s, name, v)
