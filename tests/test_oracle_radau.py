"""The Radau restatements against each other, SciPy, analytic solutions and the golden files."""
import math
import os

import numpy as np
import pytest
from scipy.integrate import solve_ivp

from janus_ode_b200 import configs
from oracle import c_oracle, problems
from oracle.radau_scalar import RadauOptions, radau_solve

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_scipy_radau_vdp_mu4():
    # examples/python/radau_test.py:13-33
    mu = 4.0
    f = lambda t, y: np.array([y[1], mu * (1 - y[0] ** 2) * y[1] - y[0]])
    jac = lambda t, y: np.array([[0, 1.0], [-2 * mu * y[0] * y[1] - 1, mu * (1 - y[0] ** 2)]])
    period = (3.0 - 2.0 * np.log(2.0)) * mu + 2.0 * 3.141592653589793 / 1000.0 / 3.0
    sol = solve_ivp(f, (0, period), [2.0, 0.0], method="Radau", rtol=1e-9, atol=1e-12, jac=jac)
    for kw in (dict(MinNbrStg=3, MaxNbrStg=3), dict()):
        r = radau_solve(f, jac, 0.0, period, [2.0, 0.0], RadauOptions(RelTol=1e-9, AbsTol=1e-12, MaxNbrStep=100000, **kw))
        assert r.res == 0 and r.t == period
        assert np.abs(r.y - sol.y[:, -1]).max() < 1e-7


def test_scipy_radau_aug_vdp_bangbang():
    # examples/python/aug_vdp_ref_test.py:6-86 == examples/cpp/radauted_vdp_aug_example.cpp
    def f(t, y):
        p1, p2, x1, x2 = y
        u = 100 if p2 * (1 - x1 ** 2) * x2 < 0 else 1
        return np.array([p2 * u * (-2.0 * x1) * x2 - p2, p1 + p2 * u * (1.0 - x1 ** 2), x2, u * (1.0 - x1 ** 2) * x2 - x1])
    y0 = [5.0, 5.0, 0.0, 1.5]
    sol = solve_ivp(f, (0, 10), y0, method="Radau", rtol=1e-10, atol=1e-13)
    for arith in ("reference", "mirror"):
        for kw in (dict(MinNbrStg=3, MaxNbrStg=3), dict()):
            op = RadauOptions(RelTol=1e-6, AbsTol=1e-9, MaxNbrStep=100000, **kw)
            o = c_oracle.solve("vdp_aug_bang", [1.0, 100.0], 1, np.array([y0]), np.array([[0.0, 10.0]]), op, arith=arith)
            assert o["res"][0] == 0
            assert np.allclose(o["y"][0], sol.y[:, -1], rtol=2e-3)


@pytest.mark.parametrize("kw", [dict(MinNbrStg=3, MaxNbrStg=3), dict(), dict(NbrStg=1, MinNbrStg=1, MaxNbrStg=1),
                                dict(NbrStg=5, MinNbrStg=5, MaxNbrStg=5), dict(NbrStg=7, MinNbrStg=7, MaxNbrStg=7)])
def test_python_vs_c_vdp_stiff(kw):
    mu = 1000.0
    tf = (3 - 2 * math.log(2)) * mu
    if kw.get("NbrStg") == 1:
        tf = 2.0      # implicit Euler needs ~1e5 steps for a whole period
    rng = np.random.default_rng(0)
    M = 3
    y0 = np.zeros((M, 4))
    y0[:, :2] = 1
    y0[:, 2] = 2 + rng.random(M)
    op = RadauOptions(RelTol=1e-6, AbsTol=1e-9, MaxNbrStep=1000000, **kw)
    c = c_oracle.solve("vdp", [mu], +1, y0, np.tile([0.0, tf], (M, 1)), op)
    f, j = problems.vdp(mu, +1)
    for m in range(M):
        r = radau_solve(f, j, 0.0, tf, y0[m], op)
        assert r.res == 0 and c["res"][m] == 0 and c["t"][m] == tf
        assert abs(r.stats["AccptNbr"] - c["stats"][m, 5]) <= max(4, 0.10 * r.stats["AccptNbr"])
        assert np.allclose(r.y, c["y"][m], rtol=1e-6, atol=1e-9 * np.abs(r.y).max())


@pytest.mark.parametrize("arith", ["reference", "mirror"])
def test_harmonic_analytic_forward_and_reverse(arith):
    cfg = configs.c1_harmonic(M=32)
    y0, op = cfg["y0"].numpy(), cfg["options"]
    for t0, tf in ((0.0, 10.0), (10.0, 0.0)):   # reverse-time integration (PosNeg = -1)
        ts = np.tile([t0, tf], (32, 1))
        o = c_oracle.solve("harmonic", [1.0], -1, y0, ts, op, arith=arith)
        dt = tf - t0
        exact = np.stack([y0[:, 0] * np.cos(dt) - y0[:, 1] * np.sin(dt), y0[:, 1] * np.cos(dt) + y0[:, 0] * np.sin(dt)], 1)
        assert (o["res"] == 0).all() and (o["t"] == tf).all()
        assert np.abs(o["y"] - exact).max() < 5e-9
    # reference sign makes the same H hyperbolic: x'' = +x
    o = c_oracle.solve("harmonic", [1.0], +1, y0, np.tile([0.0, 2.0], (32, 1)), op, arith=arith)
    exact = np.stack([y0[:, 0] * np.cosh(2) + y0[:, 1] * np.sinh(2), y0[:, 1] * np.cosh(2) + y0[:, 0] * np.sinh(2)], 1)
    assert np.abs(o["y"] - exact).max() < 1e-7


def test_edge_cases_result_codes():
    y0 = np.array([[1.0, 1.0, 2.5, 0.0]])
    # tiny interval: the initial step already overshoots -> one clipped step (Last at init)
    o = c_oracle.solve("vdp", [1000.0], 1, y0, np.array([[0.0, 1e-3]]), RadauOptions(RelTol=1e-6, AbsTol=1e-9))
    assert o["res"][0] == 0 and o["t"][0] == 1e-3 and o["stats"][0, 5] >= 1
    # step cap -> res 3, t < tfinal
    o = c_oracle.solve("vdp", [1000.0], 1, y0, np.array([[0.0, 3000.0]]), RadauOptions(RelTol=1e-6, AbsTol=1e-9, MaxNbrStep=20))
    assert o["res"][0] == 3 and o["t"][0] < 3000.0 and o["stats"][0, 4] == 21
    # NaN initial state -> res 2
    o = c_oracle.solve("vdp", [1000.0], 1, np.array([[1.0, 1.0, np.nan, 0.0]]), np.array([[0.0, 1.0]]), RadauOptions())
    assert o["res"][0] == 2
    # blow-up (canonical sign on stiff vdP) must end with a failure code, not hang
    o = c_oracle.solve("vdp", [1000.0], -1, y0, np.array([[0.0, 1614.0]]), RadauOptions(RelTol=1e-6, AbsTol=1e-9, MaxNbrStep=100000))
    assert o["res"][0] in (1, 2, 3)
    with pytest.raises(ValueError):
        RadauOptions(NbrStg=4).validate()
    with pytest.raises(ValueError):
        RadauOptions(RelTol=1e-17).validate()


@pytest.mark.parametrize("name,pb,par,sign", [
    ("c1_harmonic", "harmonic", [1.0], -1), ("c2_vdp", "vdp", [1000.0], 1),
    ("c3_pmp_smooth", "vdp_pmp_smooth", [1.0, 1.0, 100.0], 1), ("c3_pmp_bang", "vdp_pmp_bang", [1.0, 1.0, 100.0], 1)])
@pytest.mark.parametrize("tag", ["s3", "var"])
def test_golden_files_reproduce(name, pb, par, sign, tag):
    g = np.load(os.path.join(GOLD, f"{name}_{tag}.npz"))
    cfgf = dict(c1_harmonic=configs.c1_harmonic, c2_vdp=configs.c2_vdp,
                c3_pmp_smooth=lambda **k: configs.c3_pmp(control="smooth", **k),
                c3_pmp_bang=lambda **k: configs.c3_pmp(control="bangbang", **k))[name]
    cfg = cfgf(M=g["y0"].shape[0], fixed_order=(tag == "s3"))
    assert np.array_equal(cfg["y0"].numpy(), g["y0"])       # the seeded inputs are reproducible
    for arith, pre in (("reference", "ref"), ("mirror", "mir")):
        o = c_oracle.solve(pb, par, sign, g["y0"], g["tspan"], cfg["options"], arith=arith)
        assert np.array_equal(o["stats"], g[pre + "_stats"])
        assert np.array_equal(o["y"], g[pre + "_y"])
        assert np.array_equal(o["res"], g[pre + "_res"])


def test_det_pow_accuracy():
    rng = np.random.default_rng(0)
    x = 10 ** rng.uniform(-10, 3, 200000)
    y = rng.uniform(-2, 2, 200000)
    rel = np.abs(c_oracle.det_pow(x, y) / np.power(x, y) - 1)
    assert rel.max() < 2e-14
    assert c_oracle.det_pow(np.array([1.0]), np.array([0.8]))[0] == 1.0


def test_det_tanh_accuracy_and_device_code_parity():
    """csrc/detmath.cuh det_tanh_t (branch-free: shifter range reduction, 32-entry 2^(j/32) table, degree-6 polynomial,
    Newton reciprocal) -- the device code run on the host (tests/hostsim) equals the C oracle's copy bit for bit, both are
    within 2.3e-16 absolute / 2 ulp (|x| > 0.5) of libm, saturate exactly and propagate NaN"""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    import hostsim
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-25, 25, 1 << 18), 0.01 * rng.standard_normal(1 << 18),
                        10.0 ** rng.uniform(-300, -3, 1 << 14) * rng.choice([-1, 1], 1 << 14),
                        [0.0, 20.0, -20.0, 25.0, 1e300, -1e300, np.inf, -np.inf, np.nan]])
    a, b = c_oracle.det_tanh(x), hostsim.det_tanh(x)
    assert np.array_equal(a.view(np.int64), b.view(np.int64))
    ref = np.tanh(x)
    m = np.isfinite(ref)
    assert np.abs(a[m] - ref[m]).max() <= 2.3e-16
    big = m & (np.abs(x) > 0.5)
    assert (np.abs(a[big] - ref[big]) / np.abs(ref[big])).max() < 5e-16
    assert list(a[-8:-1]) == [1.0, -1.0, 1.0, 1.0, -1.0, 1.0, -1.0] and np.isnan(a[-1]) and a[-9] == 0.0
