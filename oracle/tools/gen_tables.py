#!/usr/bin/env python3
"""Generate the Radau IIA coefficient tables used by the oracle and by the CUDA kernels.

TEST/BUILD INFRASTRUCTURE -- run once in the build container, outputs are committed.

The constants are Hairer & Wanner's published RADAU tables (nodes C, error
coefficients Dd, transformation matrices T / TI and the inverted eigenvalues
ValP = [gamma, alpha1, beta1, ...]).  The reference keeps them as literals in
  /root/reference/src/matlab/radau.m:1621-2000       (Coertv1/3/5/7)
  /root/reference/src/cpp/radaute_impl.hpp:2187-2644 (same numbers)
This script reads the literals from the MATLAB twin, cross-checks every number
against the C++ file, verifies them mathematically (T * blockdiag(ValP) * TI ==
inv(A) for the Radau IIA Butcher matrix A built from C) and then writes

  oracle/radau_tables.py                     (python, for oracle/*.py)
  oracle/c/radau_tables.h                    (plain C, for oracle/c/radau_oracle.c)
  janus_ode_b200/csrc/radau_tables_t.cuh     (constexpr accessors RadauTab<S> for the kernels)

Storage convention in all outputs: Hairer's orientation, i.e.
  z_j = sum_q T[j][q] * w_q ,   w_j = sum_q TI[j][q] * z_q .
(The reference stores the transposes, T_ref[q][j] = T[j][q];
 radaute_impl.hpp:2217-2236,2388-2411; usage :881-890,969-977.)
"""
import math
import os
import re
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))


def _num(tok):
    return float(tok.replace("D", "e").replace("d", "e"))


def parse_matlab():
    src = open(os.path.join(REF, "src/matlab/radau.m")).read().splitlines()
    # restrict to the table functions
    start = next(i for i, l in enumerate(src) if l.startswith("function [T_3 TI_3 C3"))
    body = src[start:]
    pat = re.compile(r"^\s*(T|TI|C|Dd|ValP)(\d)\((\d)(?:,(\d))?\)\s*=\s*([-+0-9.eEdD]+)\s*;")
    out = {}
    for line in body:
        m = pat.match(line)
        if not m:
            continue
        name, s, i, j, val = m.groups()
        s = int(s)
        key = (name, s)
        out.setdefault(key, {})
        idx = (int(i) - 1,) if j is None else (int(i) - 1, int(j) - 1)
        # first occurrence wins (RealYN branch precedes the eig branch)
        out[key].setdefault(idx, (_num(val), val))
    return out


def build_tables():
    raw = parse_matlab()
    tabs = {}
    # --- s = 1 (implicit Euler)  radau.m:1621-1627
    tabs[1] = dict(C=[1.0], Dd=[-1.0], T=[[1.0]], TI=[[1.0]], ValP=[1.0])
    # --- s = 3: C, Dd, ValP are closed forms  radau.m:1641-1675
    sq6 = math.sqrt(6.0)
    C3 = [(4.0 - sq6) / 10.0, (4.0 + sq6) / 10.0, 1.0]
    Dd3 = [-(13.0 + 7.0 * sq6) / 3.0, (-13.0 + 7.0 * sq6) / 3.0, -1.0 / 3.0]
    st9 = 9.0 ** (1.0 / 3.0)
    v1 = (6.0 + st9 * (st9 - 1.0)) / 30.0
    v2 = (12.0 - st9 * (st9 - 1.0)) / 60.0
    v3 = st9 * (st9 + 1.0) * math.sqrt(3.0) / 60.0
    cno = v2 * v2 + v3 * v3
    ValP3 = [1.0 / v1, v2 / cno, v3 / cno]
    T3 = [[raw[("T", 3)][(i, j)][0] for j in range(3)] for i in range(3)]
    TI3 = [[raw[("TI", 3)][(i, j)][0] for j in range(3)] for i in range(3)]
    tabs[3] = dict(C=C3, Dd=Dd3, T=T3, TI=TI3, ValP=ValP3)
    for s in (5, 7):
        C = [raw[("C", s)][(i,)][0] for i in range(s)]
        Dd = [raw[("Dd", s)][(i,)][0] for i in range(s)]
        V = [raw[("ValP", s)][(i,)][0] for i in range(s)]
        T = [[raw[("T", s)][(i, j)][0] for j in range(s)] for i in range(s)]
        TI = [[raw[("TI", s)][(i, j)][0] for j in range(s)] for i in range(s)]
        tabs[s] = dict(C=C, Dd=Dd, T=T, TI=TI, ValP=V)
    return tabs


def crosscheck_cpp(tabs):
    """every literal of the tables must also appear in radaute_impl.hpp:2187-2644."""
    src = open(os.path.join(REF, "src/cpp/radaute_impl.hpp")).read().splitlines()[2186:2644]
    nums = set()
    for line in src:
        for tok in re.findall(r"[-+]?\d+\.\d*(?:[eE][-+]?\d+)?", line):
            nums.add(abs(float(tok)))
    missing = 0
    for s in (3, 5, 7):
        for name in ("T", "TI") + (("C", "Dd", "ValP") if s > 3 else ()):
            arr = np.array(tabs[s][name]).ravel()
            for v in arr:
                if v in (0.0, 1.0):
                    continue
                if abs(v) not in nums:
                    print("not found in C++ tables:", s, name, repr(v))
                    missing += 1
    assert missing == 0, "MATLAB and C++ tables disagree"


def radau_A(C):
    s = len(C)
    C = np.array(C, dtype=np.longdouble)
    CP = np.array([[c ** j for j in range(s)] for c in C], dtype=np.longdouble)
    CQ = np.array([[c ** (j + 1) / (j + 1) for j in range(s)] for c in C], dtype=np.longdouble)
    return np.array(CQ.astype(np.float64) @ np.linalg.inv(CP.astype(np.float64)))


def verify(tabs):
    for s, t in tabs.items():
        A = radau_A(t["C"])
        Ainv = np.linalg.inv(A)
        T = np.array(t["T"])
        TI = np.array(t["TI"])
        V = t["ValP"]
        L = np.zeros((s, s))
        L[0, 0] = V[0]
        for q in range((s - 1) // 2):
            a, b = V[1 + 2 * q], V[2 + 2 * q]
            i = 1 + 2 * q
            L[i, i] = a
            L[i, i + 1] = -b
            L[i + 1, i] = b
            L[i + 1, i + 1] = a
        err_inv = np.abs(T @ TI - np.eye(s)).max()
        err_dec = np.abs(T @ L @ TI - Ainv).max() / np.abs(Ainv).max()
        print(f"s={s}: |T TI - I|={err_inv:.2e}  |T L TI - inv(A)|/|inv(A)|={err_dec:.2e}")
        assert err_inv < 1e-11 and err_dec < 1e-10, (s, err_inv, err_dec)
        # error coefficients: Dd_q / gamma0 relation is checked indirectly (sum rule):
        # Estrad uses sum_q Dd_q z_q / h, for y' = const the estimate must vanish:
        # z_q = c_q h f  =>  f + sum Dd_q c_q f = 0  <=>  sum Dd_q c_q = -1
        sm = float(np.dot(t["Dd"], t["C"]))
        assert abs(sm + 1.0) < 1e-12, (s, sm)


def r17(x):
    return repr(float(x))


def emit_python(tabs, path):
    with open(path, "w") as f:
        f.write('"""Radau IIA tables -- GENERATED by oracle/tools/gen_tables.py, do not edit.\n\n')
        f.write("Source of the numbers: Hairer & Wanner RADAU constants as kept in\n")
        f.write("/root/reference/src/matlab/radau.m:1621-2000 == src/cpp/radaute_impl.hpp:2187-2644.\n")
        f.write("Orientation: z_j = sum_q T[j][q] w_q, w_j = sum_q TI[j][q] z_q (Hairer's).\n")
        f.write('ValP = [gamma, alpha1, beta1, alpha2, beta2, ...] (already inverted eigenvalues).\n"""\n')
        f.write("TABLES = {\n")
        for s in (1, 3, 5, 7):
            t = tabs[s]
            f.write(f"    {s}: {{\n")
            for name in ("C", "Dd", "ValP"):
                f.write(f"        '{name}': [" + ", ".join(r17(v) for v in t[name]) + "],\n")
            for name in ("T", "TI"):
                f.write(f"        '{name}': [\n")
                for row in t[name]:
                    f.write("            [" + ", ".join(r17(v) for v in row) + "],\n")
                f.write("        ],\n")
            f.write("    },\n")
        f.write("}\n")


def emit_c(tabs, path, device):
    qual = "__device__ __constant__ const double" if device else "static const double"
    guard = "JB200_RADAU_TABLES_CUH" if device else "JB200_ORACLE_RADAU_TABLES_H"
    with open(path, "w") as f:
        f.write("// Radau IIA tables -- GENERATED by oracle/tools/gen_tables.py, do not edit.\n")
        f.write("// Numbers: Hairer & Wanner RADAU constants as kept in the reference at\n")
        f.write("// src/matlab/radau.m:1621-2000 == src/cpp/radaute_impl.hpp:2187-2644.\n")
        f.write("// Orientation: z_j = sum_q T[j][q] w_q ; w_j = sum_q TI[j][q] z_q (Hairer's).\n")
        f.write("// ValP = [gamma, alpha1, beta1, ...] (inverted eigenvalues of A).\n")
        f.write(f"#ifndef {guard}\n#define {guard}\n\n")
        for s in (1, 3, 5, 7):
            t = tabs[s]
            for name in ("C", "Dd", "ValP"):
                f.write(f"{qual} RADAU_{name}{s}[{s}] = {{" + ", ".join(r17(v) for v in t[name]) + "};\n")
            for name in ("T", "TI"):
                f.write(f"{qual} RADAU_{name}{s}[{s}][{s}] = {{\n")
                for row in t[name]:
                    f.write("  {" + ", ".join(r17(v) for v in row) + "},\n")
                f.write("};\n")
            f.write("\n")
        f.write(f"#endif  // {guard}\n")


def emit_cuh(tabs, path):
    with open(path, "w") as f:
        f.write("// Radau IIA tables -- GENERATED by oracle/tools/gen_tables.py, do not edit.\n")
        f.write("// Numbers: Hairer & Wanner RADAU constants as kept in the reference at\n")
        f.write("// src/matlab/radau.m:1621-2000 == src/cpp/radaute_impl.hpp:2187-2644.\n")
        f.write("// Orientation: z_j = sum_q T(j,q) w_q ; w_j = sum_q TI(j,q) z_q (Hairer's).\n")
        f.write("// ValP = [gamma, alpha1, beta1, ...] (inverted eigenvalues of A).\n")
        f.write("// C, Dd, ValP: constexpr accessors (fold into immediates, and into compile-time quotients such\n")
        f.write("// as 1/(C_a - C_b)).  T, TI (2 s^2 numbers, used only as FMA multiplicands): on the device they\n")
        f.write("// sit in __constant__ memory so that each use is a constant-bank operand of the DFMA instead of\n")
        f.write("// a pair of UMOVs building a 64-bit immediate (9% of all issued instructions before).\n")
        f.write("#pragma once\n#include \"detmath.cuh\"\n\nnamespace jb {\n\ntemplate <int S>\nstruct RadauTab;\n\n")
        f.write("#if defined(__CUDACC__)\n")
        for s in (1, 3, 5, 7):
            t = tabs[s]
            for name in ("T", "TI"):
                f.write(f"static __device__ __constant__ double jb_c{name}{s}[{s}][{s}] = {{\n")
                for row in t[name]:
                    f.write("    {" + ", ".join(r17(v) for v in row) + "},\n")
                f.write("};\n")
        f.write("#endif\n\n")
        for s in (1, 3, 5, 7):
            t = tabs[s]
            f.write(f"template <>\nstruct RadauTab<{s}> {{\n")
            for name in ("C", "Dd", "ValP"):
                f.write(f"  JB_HD static constexpr double {name}(int i) {{\n")
                f.write(f"    constexpr double v[{s}] = {{" + ", ".join(r17(v) for v in t[name]) + "};\n")
                f.write("    return v[i];\n  }\n")
            for name in ("T", "TI"):
                f.write(f"  JB_HD static double {name}(int i, int j) {{\n")
                f.write("#if defined(__CUDA_ARCH__)\n")
                f.write(f"    return jb_c{name}{s}[i][j];\n")
                f.write("#else\n")
                f.write(f"    constexpr double v[{s}][{s}] = {{\n")
                for row in t[name]:
                    f.write("        {" + ", ".join(r17(v) for v in row) + "},\n")
                f.write("    };\n    return v[i][j];\n")
                f.write("#endif\n  }\n")
            f.write("};\n\n")
        f.write("}  // namespace jb\n")


def main():
    tabs = build_tables()
    crosscheck_cpp(tabs)
    verify(tabs)
    os.makedirs(os.path.join(ROOT, "oracle", "c"), exist_ok=True)
    emit_python(tabs, os.path.join(ROOT, "oracle", "radau_tables.py"))
    emit_c(tabs, os.path.join(ROOT, "oracle", "c", "radau_tables.h"), device=False)
    emit_cuh(tabs, os.path.join(ROOT, "janus_ode_b200", "csrc", "radau_tables_t.cuh"))
    print("tables written")


if __name__ == "__main__":
    sys.exit(main())
