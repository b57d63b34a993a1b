"""Parity against the LIVE reference: whole solves of the unmodified janus::RadauTe
(/root/reference/src/cpp/radaute.hpp + radaute_impl.hpp compiled by oracle/Makefile.ref, one solver object per
trajectory) stored as fixtures tests/golden/ref_live_*.npz by tests/golden/make_reference_live.py.

What the fixtures establish (numbers in profiles/r02_parity_report_live.txt):
  * fixed order: the oracle and the CUDA path reproduce the reference's accepted-step counts as well as the
    reference reproduces ITSELF when x1(0) moves by one ulp or when the closed-form callbacks are exchanged for
    its own autograd adaptor (`*_ulp`, `*_adaptor` fixtures) -- the roundoff floor of this stiff problem;
  * variable order: the same, PROVIDED RadauTe's ungated Gustafsson limiter is reproduced (quirk Q8,
    radaute_impl.hpp:1092-1100; OptionsTe.RadauTeQuirks bit 0, default on).  With the MATLAB gate
    (radau.m:1066-1076) agreement drops by ~15 points: that deviation was the whole variable-order gap.
BASELINE.json's ">= 99.9 % equal counts" therefore cannot be met by ANY second implementation of this
controller -- the reference misses it against itself -- and is asserted here as "at the reference's own floor".
"""
import os

import numpy as np
import pytest

from oracle import c_oracle
from oracle.radau_scalar import RadauOptions

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PB = {"c1_harmonic": ("harmonic", [1.0]), "c2_vdp": ("vdp", [1000.0]),
      "c3_pmp_bang": ("vdp_pmp_bang", [1.0, 1.0, 100.0]), "c3_pmp_smooth": ("vdp_pmp_smooth", [1.0, 1.0, 100.0])}


def load(name):
    p = os.path.join(GOLD, f"ref_live_{name}.npz")
    if not os.path.exists(p):
        pytest.skip(f"{p} not generated")
    b = np.load(p)
    key = [k for k in PB if name.startswith(k)][0]
    return b, PB[key]


def options_of(b, quirks=1):
    st = b["stages"]
    return RadauOptions(RelTol=float(b["rtol"]), AbsTol=float(b["atol"]), InitialStep=float(b["h0"]), NbrStg=int(st[0]),
                        MinNbrStg=int(st[1]), MaxNbrStg=int(st[2]), MaxNbrStep=100000, RadauTeQuirks=quirks)


def agreement(o, b):
    """fractions of rows with equal accepted / attempted counts, equal res, states within 1e-8 (BASELINE.json bar)."""
    ok = (b["res"] == 0) & (o["res"] == 0)
    rel = np.abs(o["y"] - b["y"]).max(1) / np.maximum(np.abs(b["y"]).max(1), 1e-300)
    return dict(res=(o["res"] == b["res"]).mean(), acc=(o["stats"][:, 5] == b["stats"][:, 5])[ok].mean(),
                att=(o["stats"][:, 4] == b["stats"][:, 4])[ok].mean(), y8=(rel[ok] < 1e-8).mean(),
                y6=(rel[ok] < 1e-6).mean(), ymax=rel[ok].max(), nstg=(o["nstg"] == b["nstg"])[ok].mean(),
                mean_acc=(o["stats"][:, 5][ok].mean(), b["stats"][:, 5][ok].mean()))


# (fixture, min equal-accepted-count fraction, min fraction of states within 1e-8): thresholds sit a few points
# below the measured live-vs-live floor of the same configuration (test_reference_roundoff_floor)
# The bang-bang control of C3 makes the right-hand side discontinuous: every trajectory crosses ~10 switching
# surfaces, each crossing costs a burst of rejected steps whose pattern is decided by roundoff, and the global error is
# set by how the switch is straddled (O(h_local * jump)), not by the tolerances -- the live reference differs from
# ITSELF (+1 ulp in x1(0)) by 1e-4..1e-3 there (test_reference_roundoff_floor).  Smooth control: 100 % / 1e-14.
CASES = [("c2_vdp_s3", 0.97, 0.999), ("c2_vdp_var", 0.93, 0.97), ("c3_pmp_bang_s3", 0.0, 0.0),
         ("c3_pmp_bang_var", 0.0, 0.0), ("c3_pmp_smooth_s3", 0.999, 0.999), ("c3_pmp_smooth_var", 0.999, 0.999)]
BANG_TOL = 5e-3


@pytest.mark.parametrize("arith", ["reference", "mirror"])
@pytest.mark.parametrize("name,min_acc,min_y", CASES)
def test_oracle_matches_live_reference(name, min_acc, min_y, arith):
    b, (pb, par) = load(name)
    o = c_oracle.solve(pb, par, int(b["sign"]), b["y0"], b["tspan"], options_of(b), arith=arith)
    a = agreement(o, b)
    assert a["res"] == 1.0
    assert a["acc"] >= min_acc and a["y8"] >= min_y, a
    assert (a["ymax"] < BANG_TOL) if "bang" in name else (a["y6"] >= 0.999), a
    # no systematic drift in the amount of work: mean accepted steps within 0.5 % (2 % across the switching bursts)
    assert abs(a["mean_acc"][0] - a["mean_acc"][1]) <= (0.02 if "bang" in name else 0.005) * a["mean_acc"][1], a


def test_variable_order_gap_was_the_gustafsson_gate():
    """quirk Q8: without RadauTe's ungated limiter the restatement leaves the reference on ~20 % of the rows."""
    b, (pb, par) = load("c2_vdp_var")
    with_q = agreement(c_oracle.solve(pb, par, 1, b["y0"], b["tspan"], options_of(b, 1)), b)
    without = agreement(c_oracle.solve(pb, par, 1, b["y0"], b["tspan"], options_of(b, 0)), b)
    assert with_q["acc"] >= 0.93 and without["acc"] <= 0.88, (with_q, without)
    assert with_q["nstg"] >= 0.95


@pytest.mark.parametrize("base,other,lo,hi", [("c2_vdp_s3", "c2_vdp_s3_ulp", 0.95, 0.998),
                                               ("c2_vdp_s3", "c2_vdp_s3_adaptor", 0.95, 1.0),
                                               ("c2_vdp_var", "c2_vdp_var_ulp", 0.88, 0.99),
                                               ("c2_vdp_var", "c2_vdp_var_adaptor", 0.88, 1.0),
                                               ("c1_harmonic_s3", "c1_harmonic_s3_ulp", 0.3, 0.9),
                                               ("c3_pmp_bang_var", "c3_pmp_bang_var_ulp", 0.0, 0.5)])
def test_reference_roundoff_floor(base, other, lo, hi):
    """the live reference against itself after a 1-ulp change of x1(0) / with its own autograd adaptor: equal
    accepted-step counts on well under 99.9 % of the rows -- the bar no reimplementation can clear."""
    a, _ = load(base)
    b, _ = load(other)
    n = len(b["res"])
    eq = (a["stats"][:n, 5] == b["stats"][:, 5]).mean()
    assert lo <= eq <= hi, eq


def test_harmonic_counts_are_roundoff_decided():
    """C1 (linear problem): Newton converges in one iteration up to roundoff, so whether a second iteration runs --
    and with it the step-size sequence -- is decided by the last bits; totals still agree closely and the states
    agree to 1e-9."""
    for name in ("c1_harmonic_s3", "c1_harmonic_var"):
        b, (pb, par) = load(name)
        o = c_oracle.solve(pb, par, int(b["sign"]), b["y0"], b["tspan"], options_of(b))
        a = agreement(o, b)
        assert a["res"] == 1.0 and a["y8"] == 1.0
        assert abs(a["mean_acc"][0] - a["mean_acc"][1]) <= 0.05 * a["mean_acc"][1], a
    # the reference against itself, x(0) + 1 ulp: far below 99.9 % too
    a, _ = load("c1_harmonic_var")
    b, _ = load("c1_harmonic_var_ulp")
    assert (a["stats"][:len(b["res"]), 5] == b["stats"][:, 5]).mean() < 0.9


def test_live_reference_reproduces_its_fixture():
    """only where oracle/_ref is built (needs /root/reference at build time): 4 rows, bit for bit."""
    from oracle import ref_live
    if not ref_live.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    b, (pb, par) = load("c2_vdp_var")
    r = ref_live.solve(pb, par[0], 1, b["y0"][:4], b["tspan"][:4], float(b["rtol"]), float(b["atol"]), h0=float(b["h0"]),
                       stages=tuple(int(s) for s in b["stages"]))
    assert np.array_equal(r["stats"], b["stats"][:4]) and np.array_equal(r["y"], b["y"][:4])


def test_live_reference_adaptor_is_the_closed_form():
    """the reference's evalDyns / evalJac (autograd) on the fixture Hamiltonian == the closed forms the
    fixtures were generated with, to roundoff (pins oracle/problems.py and the dual-number kernels' target)."""
    from oracle import ref_live
    if not ref_live.available():
        pytest.skip("oracle/_ref not built and /root/reference absent")
    rng = np.random.default_rng(0)
    y = rng.uniform(-2, 2, (64, 4))
    for pb, par in (("vdp", 1000.0), ("vdp_pmp_smooth", 1.0), ("vdp_pmp_bang", 1.0)):
        f1, J1 = ref_live.evaluate(pb, par, 1, y, adaptor=True)
        f0, J0 = ref_live.evaluate(pb, par, 1, y, adaptor=False)
        assert np.allclose(f1, f0, rtol=1e-13, atol=1e-11) and np.allclose(J1, J0, rtol=1e-13, atol=1e-10)
        pars = [par] if pb == "vdp" else [par, 1.0, 100.0]
        assert np.array_equal(f0, c_oracle.eval_dyns(pb, pars, 1, y)) and np.array_equal(J0, c_oracle.eval_jac(pb, pars, 1, y))


# ------------------------------------------------------------------ the CUDA path against the live reference
@pytest.mark.gpu
@pytest.mark.parametrize("name,min_acc,min_y", CASES + [("c1_harmonic_s3", 0.0, 1.0), ("c1_harmonic_var", 0.0, 1.0)])
def test_gpu_matches_live_reference(name, min_acc, min_y):
    import torch

    import janus_ode_b200 as jb
    b, (pb, par) = load(name)
    ham = {"harmonic": lambda: jb.Harmonic(1.0), "vdp": lambda: jb.VanDerPol(1000.0),
           "vdp_pmp_bang": lambda: jb.VdpPmp(1.0, "bangbang", 1.0, 100.0),
           "vdp_pmp_smooth": lambda: jb.VdpPmp(1.0, "smooth")}[pb]()
    st = b["stages"]
    op = jb.OptionsTe(RelTol=float(b["rtol"]), AbsTol=float(b["atol"]), InitialStep=float(b["h0"]), NbrStg=int(st[0]),
                      MinNbrStg=int(st[1]), MaxNbrStg=int(st[2]), MaxNbrStep=100000)
    dyn = jb.HamiltonianSystem(ham, int(b["sign"]))
    r = jb.RadauTe(dyn, dyn.jacobian, torch.from_numpy(b["tspan"]).cuda(), torch.from_numpy(b["y0"]).cuda(), op)
    res = r.solve()
    g = dict(y=r.y.cpu().numpy(), res=res.cpu().numpy().astype(np.int32), stats=r.stats.raw.cpu().numpy(),
             nstg=r.nstg.cpu().numpy())
    a = agreement(g, b)
    assert a["res"] == 1.0
    assert a["acc"] >= min_acc and a["y8"] >= min_y, a
    assert (a["ymax"] < BANG_TOL) if "bang" in name else (a["y6"] >= 0.999), a
    assert abs(a["mean_acc"][0] - a["mean_acc"][1]) <= (0.05 if name.startswith("c1") else 0.02) * a["mean_acc"][1], a
