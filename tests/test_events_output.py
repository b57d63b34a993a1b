"""Per-step output (OutFlag 1 / 2) and event location.
  * OutFlag 1 is pinned to the LIVE reference: tests/golden/ref_live_steps_c2.npz holds RadauTe's own tout / yout record
    (radaute_impl.hpp:1166-1183) of 16 stiff-vdP trajectories -- the whole accepted-step sequence, not just the end;
  * events follow src/matlab/radau.m:1494-1618 (the C++ port is incomplete): checked against the analytic zeros of the
    harmonic oscillator, the oracle in both arithmetics, and the CUDA path == mirror oracle bit for bit."""
import os

import numpy as np
import pytest

from janus_ode_b200 import configs
from oracle import c_oracle
from oracle.radau_scalar import RadauOptions

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_step_record_matches_the_live_reference():
    g = np.load(os.path.join(GOLD, "ref_live_steps_c2.npz"))
    n_same = 0
    for i in range(len(g["rows"])):
        st = (3, 3, 7) if g["var"][i] else (3, 3, 3)
        op = RadauOptions(RelTol=1e-6, AbsTol=1e-9, NbrStg=st[0], MinNbrStg=st[1], MaxNbrStg=st[2], MaxNbrStep=100000)
        o = c_oracle.solve_out("vdp", [1000.0], 1, g["y0"][i:i + 1], g["tspan"][i:i + 1], op, cap=512)
        n = int(g["nout"][i])
        assert abs(int(o["nout"][0]) - n) <= 2
        k = min(n, int(o["nout"][0])) - 1
        # the common prefix of the two step sequences: times agree to roundoff amplification (1e-9 -> 1e-6 over ~60
        # steps, see tests/test_reference_live.py), states to the tolerance
        m = min(k, 40)
        assert np.allclose(o["tout"][0, :m], g["tout"][i, :m], rtol=1e-7, atol=0)
        assert np.allclose(o["yout"][0, :m], g["yout"][i, :m], rtol=1e-6, atol=1e-9)
        if int(o["nout"][0]) == n:
            n_same += 1
            assert o["tout"][0, n - 1] == g["tout"][i, n - 1] == g["tspan"][i, 1]
            assert np.allclose(o["tout"][0, :n], g["tout"][i, :n], rtol=2e-3, atol=0)
    assert n_same >= len(g["rows"]) - 2


def _harmonic_zeros(p0, x0, tf):
    t0 = np.arctan2(-x0, p0)
    return np.array([t0 + k * np.pi for k in range(-2, 20) if 0 < t0 + k * np.pi < tf])


@pytest.mark.parametrize("arith", ["reference", "mirror"])
def test_events_of_the_harmonic_oscillator(arith):
    op = RadauOptions(RelTol=1e-10, AbsTol=1e-12, MaxNbrStep=100000)
    rng = np.random.default_rng(1)
    y0 = rng.uniform(-1, 1, (32, 2))
    ts = np.tile([0.0, 10.0], (32, 1))
    ev = (np.array([[0.0, 1.0]]), np.array([0.0]), np.array([0]), np.array([0]))
    r = c_oracle.solve_out("harmonic", [1.0], -1, y0, ts, op, cap=256, events=ev, ecap=8, arith=arith)
    for m in range(32):
        z = _harmonic_zeros(y0[m, 0], y0[m, 1], 10.0)
        assert r["nev"][m] == len(z) and np.allclose(r["te"][m, :len(z)], z, rtol=0, atol=2e-8)
        assert np.abs(r["ye"][m, :len(z), 1]).max() < 1e-8 and (r["ie"][m, :len(z)] == 1).all()
        assert r["t"][m] == 10.0
    # direction filter + terminal event: the first FALLING zero ends the trajectory there
    ev = (np.array([[0.0, 1.0]]), np.array([0.0]), np.array([1]), np.array([-1]))
    r = c_oracle.solve_out("harmonic", [1.0], -1, y0, ts, op, cap=256, events=ev, ecap=8, arith=arith)
    for m in range(32):
        z = _harmonic_zeros(y0[m, 0], y0[m, 1], 10.0)
        p_at = y0[m, 0] * np.cos(z) - y0[m, 1] * np.sin(z)          # xdot = p at the zeros
        zf = z[p_at < 0]
        assert r["nev"][m] == 1 and abs(r["t"][m] - zf[0]) < 2e-8 and r["res"][m] == 0
        n = r["nout"][m]
        assert r["tout"][m, n - 1] == r["t"][m] and np.array_equal(r["yout"][m, n - 1], r["y"][m])   # radau.m:1267-1269


def test_refined_output_interpolates_between_the_steps():
    op = RadauOptions(RelTol=1e-9, AbsTol=1e-12, MaxNbrStep=100000)
    y0 = np.array([[0.3, 1.0]])
    ts = np.array([[0.0, 5.0]])
    a = c_oracle.solve_out("harmonic", [1.0], -1, y0, ts, op, cap=1024, refine=1)
    b = c_oracle.solve_out("harmonic", [1.0], -1, y0, ts, op, cap=1024, refine=4)
    na, nb = a["nout"][0], b["nout"][0]
    assert nb == 4 * na and np.array_equal(b["tout"][0, 3:nb:4], a["tout"][0, :na])
    t = b["tout"][0, :nb]
    exact = np.stack([0.3 * np.cos(t) - 1.0 * np.sin(t), 1.0 * np.cos(t) + 0.3 * np.sin(t)], 1)
    assert np.abs(b["yout"][0, :nb] - exact).max() < 1e-7 and (np.diff(t) > 0).all()
    # a buffer that is too small: nout still counts every point
    c = c_oracle.solve_out("harmonic", [1.0], -1, y0, ts, op, cap=10, refine=1)
    assert c["nout"][0] == na and np.array_equal(c["tout"][0], a["tout"][0, :10])


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("var", [False, True])
def test_gpu_step_record_and_events_equal_the_mirror_oracle(var):
    import torch

    import janus_ode_b200 as jb
    cfg = configs.c2_vdp(M=1 << 20, fixed_order=not var, hi=512)
    A = [[0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]]           # x1 = 1 (both directions), x2 = 0 (rising)
    op = cfg["options"]
    op.EventsFcn = jb.LinearEvents(A, [1.0, 0.0], stop=[0, 0], slope=[0, 1])
    op.OutputCapacity, op.EventsCapacity = 400, 6
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    r = jb.RadauTe(dyn, dyn.jacobian, cfg["tspan"].cuda(), cfg["y0"].cuda(), op)
    r.solve()
    o = c_oracle.solve_out("vdp", [1000.0], 1, cfg["y0"].numpy(), cfg["tspan"].numpy(), op, cap=400,
                           events=(np.array(A), np.array([1.0, 0.0]), np.array([0, 0]), np.array([0, 1])), ecap=6, arith="mirror")
    for k, v in (("y", r.y), ("t", r.t), ("tout", r.tout), ("yout", r.yout), ("nout", r.nout), ("te", r.te), ("ye", r.ye),
                 ("ie", r.ie), ("nev", r.nev)):
        assert np.array_equal(v.cpu().numpy(), o[k]), k
    assert np.array_equal(r.stats.raw.cpu().numpy(), o["stats"]) and int(r.nev.sum()) > 0
    # the same solve without recording gives the same trajectory (recording is observation only)
    cfg2 = configs.c2_vdp(M=1 << 20, fixed_order=not var, hi=512)
    r2 = jb.RadauTe(dyn, dyn.jacobian, cfg2["tspan"].cuda(), cfg2["y0"].cuda(), cfg2["options"])
    r2.solve()
    assert torch.equal(r2.y, r.y) and torch.equal(r2.stats.raw, r.stats.raw)


@pytest.mark.gpu
def test_gpu_terminal_event_and_refine():
    import torch

    import janus_ode_b200 as jb
    cfg = configs.c1_harmonic(M=256)
    op = cfg["options"]
    op.EventsFcn = jb.LinearEvents([[0.0, 1.0]], [0.0], stop=1, slope=-1)
    op.OutputCapacity, op.Refine = 600, 3
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    r = jb.RadauTe(dyn, dyn.jacobian, cfg["tspan"].cuda(), cfg["y0"].cuda(), op)
    res = r.solve()
    o = c_oracle.solve_out("harmonic", [1.0], -1, cfg["y0"].numpy(), cfg["tspan"].numpy(), op, cap=600, refine=3,
                           events=(np.array([[0.0, 1.0]]), np.array([0.0]), np.array([1]), np.array([-1])), ecap=8, arith="mirror")
    for k, v in (("y", r.y), ("t", r.t), ("tout", r.tout), ("yout", r.yout), ("nout", r.nout), ("te", r.te), ("nev", r.nev)):
        assert np.array_equal(v.cpu().numpy(), o[k]), k
    assert bool((res == 0).all()) and bool((r.nev == 1).all()) and float(r.y[:, 1].abs().max()) < 1e-7
    y0 = cfg["y0"].numpy()
    for m in range(8):
        z = _harmonic_zeros(y0[m, 0], y0[m, 1], 10.0)
        zf = z[(y0[m, 0] * np.cos(z) - y0[m, 1] * np.sin(z)) < 0]
        assert abs(float(r.t[m]) - zf[0]) < 1e-7
