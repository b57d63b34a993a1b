"""SEULEX: the stiff extrapolation stepper of the reference's Python tree behind the same
boundary.  Parity chain:
  live reference  /root/reference/src/python/seulext.py (SeulexT.step, run with the janus-linear shim of
                  tests/golden/shims/, one state per trajectory)  ->  tests/golden/seulex_ref.npz
  C oracle        oracle/c/radau_oracle.c: seulex_one  == live reference step by step (same number of steps, same
                  order decisions, x / y / hnext to roundoff-level tolerances; the reference's LU is LAPACK through
                  the shim, the oracle's its own)
  CUDA / hostsim  csrc/seulex.cuh == C oracle in mirror arithmetic, bit for bit.
"""
import os

import numpy as np
import pytest

from oracle import c_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "seulex_ref.npz")
CASES = ["mu1000", "mu5"]


def drive(step_fn, y0, ft, atol, rtol, h0, KMAXX, nsteps, quirks):
    """the reference's driver loop (src/python/seulex_batch_example.py:44-77) for one trajectory"""
    st = c_oracle.SeulexOracleState(y0[None, :], 0.0, atol, rtol, KMAXX)
    htry, out = h0, []
    for _ in range(nsteps):
        step_fn(htry, st, quirks)
        out.append(np.concatenate([[st.x[0]], st.y[0], [st.hdid[0], st.hnext[0], float(st.k_targ[0])]]))
        htry = st.hnext[0]
        if htry > ft - st.x[0]:
            htry = ft - st.x[0]
    return np.array(out)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("arith", ["reference", "mirror"])
def test_oracle_follows_the_live_reference_stepper(name, arith):
    g = np.load(GOLD)
    mu, ft, atol, rtol, h0, KM = g[name + "_par"]
    for m in range(g[name + "_y0"].shape[0]):
        n = int(g[name + "_n"][m])
        rec = g[name + "_rec"][m, :n]
        got = drive(lambda h, st, q: c_oracle.seulex_step("vdp", [mu], 1, h, st, quirks=q, arith=arith),
                    g[name + "_y0"][m], ft, atol, rtol, h0, int(KM), n, quirks=1)
        assert got[-1, 0] == ft == rec[-1, 0]                                 # same number of steps to the final time
        assert np.array_equal(got[:, 7], rec[:, 7])                           # every order decision (k_targ)
        assert np.allclose(got[:, 0], rec[:, 0], rtol=1e-5, atol=0)           # clock
        ysc = np.abs(rec[:, 1:5]).max(1, keepdims=True)
        assert (np.abs(got[:, 1:5] - rec[:, 1:5]) / ysc).max() < 1e-6         # states
        assert np.allclose(got[:, 6], rec[:, 6], rtol=1e-4, atol=0)           # proposed next step


@pytest.mark.parametrize("name", CASES)
def test_cuda_source_equals_the_oracle_bit_for_bit_on_the_host(name):
    import hostsim
    g = np.load(GOLD)
    mu, ft, atol, rtol, h0, KM = g[name + "_par"]
    for quirks in (0, 1):
        for m in range(3):
            n = int(g[name + "_n"][m]) + 3
            a = drive(lambda h, st, q: c_oracle.seulex_step("vdp", [mu], 1, h, st, quirks=q, arith="mirror"),
                      g[name + "_y0"][m], ft, atol, rtol, h0, int(KM), n, quirks)
            b = drive(lambda h, st, q: hostsim.seulex_step(1, [mu], 1, h, st, quirks=q),
                      g[name + "_y0"][m], ft, atol, rtol, h0, int(KM), n, quirks)
            assert np.array_equal(a, b)


def test_clock_quirk_only_matters_after_an_in_step_reduction():
    """S1: with the reference's bookkeeping x advances by htry although y was integrated over a shorter step"""
    y0 = np.array([1.0, 1.0, 2.3, 2.0])                                      # off the slow manifold: a fast transient
    a = drive(lambda h, st, q: c_oracle.seulex_step("vdp", [1000.0], 1, h, st, quirks=q), y0, 1e9, 1e-9, 1e-6, 1.0, 13, 1, 0)
    b = drive(lambda h, st, q: c_oracle.seulex_step("vdp", [1000.0], 1, h, st, quirks=q), y0, 1e9, 1e-9, 1e-6, 1.0, 13, 1, 1)
    assert b[0, 0] == 1.0 and a[0, 0] < 1e-3 and a[0, 5] == a[0, 0]           # htry = 1 is cut to ~1e-4 inside the first step
    assert np.array_equal(a[0, 1:5], b[0, 1:5])                              # same state, different clock


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_seulex_equals_the_oracle_and_follows_the_reference(name):
    import torch

    import janus_ode_b200 as jb
    g = np.load(GOLD)
    mu, ft, atol, rtol, h0, KM = g[name + "_par"]
    y0 = g[name + "_y0"]
    M = y0.shape[0]
    dyn = jb.HamiltonianSystem(jb.VanDerPol(mu), +1)
    for quirks in (1, 0):
        st = jb.SeulexState(torch.from_numpy(y0).cuda(), torch.zeros(M).cuda(), atol, rtol, int(KM), 0.0, ft)
        solver = jb.SeulexT(reference_clock=bool(quirks))
        ost = c_oracle.SeulexOracleState(y0, 0.0, atol, rtol, int(KM))
        htry = torch.full((M,), h0, dtype=torch.float64).cuda()
        n = int(g[name + "_n"].min())
        for i in range(n):
            st = solver.step(htry, dyn, st)
            c_oracle.seulex_step("vdp", [mu], 1, htry.cpu().numpy(), ost, quirks=quirks, arith="mirror")
            for f in ("y", "x", "hnext", "hdid", "theta", "k_targ", "flags"):
                assert np.array_equal(getattr(st, f).cpu().numpy(), getattr(ost, f)), (f, i)
            if quirks:   # against the live reference's record
                rec = g[name + "_rec"][:, i]
                assert np.array_equal(st.k_targ.cpu().numpy(), rec[:, 7].astype(np.int32))
                assert np.abs(st.y.cpu().numpy() - rec[:, 1:5]).max() / np.abs(rec[:, 1:5]).max() < 1e-6
            htry = torch.minimum(st.hnext, ft - st.x)
        assert torch.equal(st.dfdy, dyn.jacobian(st.x, st._yjac))


@pytest.mark.gpu
def test_gpu_seulex_integrates_a_batch_to_the_final_time():
    """driver loop of src/python/seulex_batch_example.py:61-77 with sub-state selection; result vs the Radau solve"""
    import torch

    import janus_ode_b200 as jb
    from janus_ode_b200 import configs
    cfg = configs.c2_vdp(M=1 << 20, hi=2048)
    ft = 50.0
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    y0 = cfg["y0"].cuda()
    st = jb.SeulexState(y0, torch.zeros(2048).cuda(), 1e-9, 1e-6, 13, 0.0, ft)
    solver = jb.SeulexT()
    htry = torch.full((2048,), 0.01, dtype=torch.float64).cuda()
    for _ in range(400):
        live = (ft - st.x) > 0
        if not bool(live.any()):
            break
        sub = solver.step(torch.minimum(htry, ft - st.x)[live], dyn, st[live])
        st[live] = sub
        htry = st.hnext.clone()
    assert bool((st.x == ft).all())
    ts = torch.tensor([0.0, ft], dtype=torch.float64).repeat(2048, 1).cuda()
    r = jb.RadauTe(dyn, dyn.jacobian, ts, y0, jb.OptionsTe(RelTol=1e-9, AbsTol=1e-12, MaxNbrStep=100000))
    r.solve()
    rel = (st.y - r.y).abs().max(1).values / r.y.abs().max(1).values
    assert float(rel.max()) < 1e-4


# ------------------------------------------------------------------ dual part (the reference's SeulexTed / SeulexTedState)
def _hostsim():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    import hostsim
    return hostsim


@pytest.mark.parametrize("mu,tf", [(5.0, 3.0), (1000.0, 50.0)])
def test_cuda_core_dual_part_on_host_vs_finite_differences(mu, tf):
    """csrc/seulex.cuh seulex_dual_one (the accepted attempt replayed on [y, y'], A' from the third-order jet of H), run on the
    host through the reference's driver loop: d y(tf) / d y0 against central differences of the same stepper -- the reference's
    own check of its dual solvers (tests/cpp/ode_solver_test.cpp:620-770, rtol 1e-3 ... 1e-4) -- and against the Radau
    sensitivities of the same problem; the real part does not notice the dual part"""
    import janus_ode_b200 as jb
    hostsim = _hostsim()
    y0 = np.array([[1.0, 1.0, 2.0, 0.0], [0.5, -1.0, 1.7, 0.2]])
    kw = dict(atol=1e-9, rtol=1e-9, h0=0.01)
    r = hostsim.seulex_dual("vdp", [mu], +1, y0, 0.0, tf, **kw)
    r0 = hostsim.seulex_dual("vdp", [mu], +1, y0, 0.0, tf, dual=False, **kw)
    assert np.array_equal(r["y"], r0["y"]) and np.array_equal(r["steps"], r0["steps"]) and (r["steps"] > 5).all()
    S = np.zeros((2, 4, 4))
    for d in range(4):
        e = 1e-6
        yp, ym = y0.copy(), y0.copy()
        yp[:, d] += e
        ym[:, d] -= e
        S[:, :, d] = (hostsim.seulex_dual("vdp", [mu], +1, yp, 0.0, tf, dual=False, **kw)["y"]
                      - hostsim.seulex_dual("vdp", [mu], +1, ym, 0.0, tf, dual=False, **kw)["y"]) / (2 * e)
    assert np.allclose(r["sens"], S, rtol=1e-5, atol=1e-6 * np.abs(S).max())
    ts = np.tile(np.array([0.0, tf]), (2, 1))
    h = hostsim.solve_sens("vdp", [mu], +1, y0, ts, jb.OptionsTe(RelTol=1e-10, AbsTol=1e-13, MaxNbrStep=100000))
    assert np.allclose(r["sens"], h["sens"], rtol=1e-5, atol=1e-6 * np.abs(h["sens"]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("mu,tf", [(5.0, 3.0), (1000.0, 50.0)])
def test_gpu_dual_part(mu, tf):
    import torch
    import janus_ode_b200 as jb
    hostsim = _hostsim()
    M = 64
    g = torch.Generator().manual_seed(3)
    y0 = torch.tensor([1.0, 1.0, 2.0, 0.0], dtype=torch.float64).repeat(M, 1)
    y0[:, 2] += 0.5 * torch.rand(M, generator=g, dtype=torch.float64)
    dyn = jb.HamiltonianSystem(jb.VanDerPol(mu), +1)

    def run(state, solver):
        htry = torch.full((M,), 0.01, dtype=torch.float64, device="cuda")
        n = 0
        while bool(((tf - state.x) > 1e-14 * (1 + tf)).any()) and n < 10000:
            live = (tf - state.x) > 1e-14 * (1 + tf)
            state[live] = solver.step(torch.minimum(htry, tf - state.x)[live], dyn, state[live])
            htry = state.hnext.clone()
            n += 1
        return state

    x0 = torch.zeros(M, dtype=torch.float64, device="cuda")
    sd = run(jb.SeulexTedState(y0.cuda(), x0, 1e-9, 1e-9, KMAXX=12), jb.SeulexTed())
    sp = run(jb.SeulexState(y0.cuda(), x0, 1e-9, 1e-9, KMAXX=12), jb.SeulexT())
    assert torch.equal(sd.y, sp.y) and torch.equal(sd.x, sp.x) and torch.equal(sd.hnext, sp.hnext)      # real part untouched
    h = hostsim.seulex_dual("vdp", [mu], +1, y0.numpy(), 0.0, tf, 1e-9, 1e-9, KMAXX=12, h0=0.01)
    assert np.array_equal(sd.y.cpu().numpy(), h["y"])
    assert np.allclose(sd.yd.cpu().numpy(), h["sens"], rtol=1e-12, atol=1e-12 * np.abs(h["sens"]).max())
    # the implicit Radau path's sensitivities of the same rows
    ts = torch.tensor([0.0, tf], dtype=torch.float64).repeat(M, 1).cuda()
    r = jb.RadauTeD(dyn, dyn.jacobian, ts, y0.cuda(), jb.OptionsTe(RelTol=1e-10, AbsTol=1e-13, MaxNbrStep=100000))
    r.solve()
    assert torch.allclose(sd.yd, r.yd, rtol=1e-5, atol=1e-6 * float(r.yd.abs().max()))
