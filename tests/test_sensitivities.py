"""Forward sensitivities (the reference's RadauTeD, tests/cpp/ode_solver_test.cpp:620-770).
The reference's own check is central finite differences of the solver (rtol 1e-3 ... 1e-4): the same check is applied
to the oracle (CPU) and to the GPU path, plus the exact answer for the harmonic oscillator and GPU == oracle."""
import numpy as np
import pytest
import torch

import janus_ode_b200 as jb
from janus_ode_b200 import configs
from oracle import problems
from oracle.radau_scalar import RadauOptions, radau_solve


def fd_sens(f, j, t0, tf, y0, op, rel=1e-7):
    n = len(y0)
    S = np.zeros((n, n))
    for d in range(n):
        hh = rel * max(abs(y0[d]), 1.0)
        yp, ym = y0.copy(), y0.copy()
        yp[d] += hh
        ym[d] -= hh
        S[:, d] = (radau_solve(f, j, t0, tf, yp, op).y - radau_solve(f, j, t0, tf, ym, op).y) / (2 * hh)
    return S


def test_oracle_sensitivities_harmonic_exact():
    f, j = problems.harmonic(1.0, -1)
    op = RadauOptions(RelTol=1e-9, AbsTol=1e-12)
    y0 = np.array([0.3, -0.7])
    r = radau_solve(f, j, 0.0, 2.0, y0, op, sens0=np.eye(2))
    r0 = radau_solve(f, j, 0.0, 2.0, y0, op)
    assert np.array_equal(r.y, r0.y) and r.stats == r0.stats            # the real part is untouched
    c, s = np.cos(2.0), np.sin(2.0)
    assert np.allclose(r.sens, np.array([[c, -s], [s, c]]), atol=1e-8)   # y = [p, x]: p' = -x, x' = p


@pytest.mark.parametrize("mu,tf", [(1.0, 6.0), (30.0, 20.0)])
def test_oracle_sensitivities_vs_finite_differences(mu, tf):
    """the reference's test: vdP, tf = 6, sensitivities vs central differences at 1e-3 ... 1e-4 (here 2e-4)"""
    f, j = problems.vdp(mu, +1)
    op = RadauOptions(RelTol=1e-9, AbsTol=1e-12, MaxNbrStep=100000)
    y0 = np.array([1.0, 1.0, 2.0, 0.0])
    r = radau_solve(f, j, 0.0, tf, y0, op, sens0=np.eye(4))
    S = fd_sens(f, j, 0.0, tf, y0, op)
    assert r.res == 0 and r.sens_iters > 0
    assert np.allclose(r.sens, S, rtol=2e-4, atol=2e-4 * np.abs(S).max())


def test_cuda_core_sens_step_on_host_vs_oracle():
    """the device code itself (csrc/radau_core.cuh sens_step, compiled for the host by tests/hostsim) against the numpy
    oracle: same algorithm -> tolerance-level agreement; the real part equals the plain host-sim solve bit for bit"""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    import hostsim
    op = jb.OptionsTe(RelTol=1e-9, AbsTol=1e-12, MaxNbrStep=100000)
    y0 = np.array([[1.0, 1.0, 2.0, 0.0], [1.0, 1.0, 2.3, 0.0], [0.5, -1.0, 1.7, 0.2]])
    ts = np.tile(np.array([0.0, 6.0]), (3, 1))
    h = hostsim.solve_sens("vdp", [3.0], +1, y0, ts, op)
    p = hostsim.solve("vdp", [3.0], +1, y0, ts, op)
    assert np.array_equal(h["y"], p["y"]) and np.array_equal(h["stats"], p["stats"])
    f, j = problems.vdp(3.0, +1)
    for m in range(3):
        o = radau_solve(f, j, 0.0, 6.0, y0[m], op, sens0=np.eye(4))
        assert np.allclose(h["sens"][m], o.sens, rtol=1e-6, atol=1e-6 * np.abs(o.sens).max())
    hh = hostsim.solve_sens("harmonic", [1.0], -1, np.array([[0.3, -0.7]]), np.array([[0.0, 2.0]]), op)
    c, s = np.cos(2.0), np.sin(2.0)
    assert np.allclose(hh["sens"][0], np.array([[c, -s], [s, c]]), atol=1e-8)


def _hostsim():
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    import hostsim
    return hostsim


@pytest.mark.parametrize("ham,par,k,sign,mk,y0", [
    ("vdp", [3.0], 0, +1, lambda p: problems.vdp(p[0], +1), [[1.0, 1.0, 2.0, 0.0], [0.5, -1.0, 1.7, 0.2]]),
    ("harmonic", [1.3], 0, -1, lambda p: problems.harmonic(p[0], -1), [[0.3, -0.7], [1.0, 0.2]])])
def test_cuda_core_parameter_direction_on_host_vs_finite_differences(ham, par, k, sign, mk, y0):
    """d y(tf) / d par_k by the device code (sens_step + ham_dpar: one more hyper-dual variable of H), run on the host,
    against central differences IN THE PARAMETER of the oracle solver -- the reference's check of its mu direction
    (tests/cpp/ode_solver_test.cpp:630-700, rtol 1e-3); the state directions and the real part do not notice"""
    hostsim = _hostsim()
    op = jb.OptionsTe(RelTol=1e-9, AbsTol=1e-12, MaxNbrStep=100000)
    y0 = np.array(y0)
    tf = 4.0
    ts = np.tile(np.array([0.0, tf]), (2, 1))
    h = hostsim.solve_sens(ham, par, sign, y0, ts, op, par_dirs=[k])
    h0 = hostsim.solve_sens(ham, par, sign, y0, ts, op)
    assert np.array_equal(h["sens"], h0["sens"]) and np.array_equal(h["y"], h0["y"]) and np.array_equal(h["stats"], h0["stats"])
    oo = RadauOptions(RelTol=1e-10, AbsTol=1e-13, MaxNbrStep=100000)
    for m in range(2):
        e = 1e-6 * max(1.0, abs(par[k]))
        pp, pm = list(par), list(par)
        pp[k] += e
        pm[k] -= e
        fp, jp = mk(pp)
        fm, jm = mk(pm)
        fd = (radau_solve(fp, jp, 0.0, tf, y0[m].copy(), oo).y - radau_solve(fm, jm, 0.0, tf, y0[m].copy(), oo).y) / (2 * e)
        assert np.abs(fd).max() > 1e-3
        assert np.allclose(h["spar"][m, :, 0], fd, rtol=1e-4, atol=1e-5 * np.abs(fd).max())


# ------------------------------------------------------------------ GPU
DEV = "cuda:0"


def gpu_sens(cfg, y0d=None):
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    r = jb.RadauTeD(dyn, dyn.jacobian, cfg["tspan"].to(DEV), cfg["y0"].to(DEV), cfg["options"], y0d=y0d)
    r.solve()
    torch.cuda.synchronize()
    return r


@pytest.mark.gpu
@pytest.mark.parametrize("fixed", [True, False])
def test_gpu_real_part_is_bitwise_the_plain_solve(fixed):
    cfg = configs.c2_vdp(M=512, fixed_order=fixed)
    r = gpu_sens(cfg)
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    p = jb.RadauTe(dyn, dyn.jacobian, cfg["tspan"].to(DEV), cfg["y0"].to(DEV), cfg["options"])
    p.solve()
    assert torch.equal(r.y, p.y) and torch.equal(r.stats.raw, p.stats.raw) and torch.equal(r.h, p.h) and torch.equal(r.nstg, p.nstg)
    assert bool((r.sens_iters > 0).all()) and bool(torch.isfinite(r.yd).all())


@pytest.mark.gpu
def test_gpu_variable_order_epochs_equal_the_all_orders_core_bitwise():
    """variable order runs through the order-homogeneous epochs (epoch_sens_kernel) with S parked in the result matrix
    between epochs; the same core integrated in one go on the host (hostsim: integrate<Ham, 4, 7, true> + sens_step) must
    give the same bits: states, counters AND sensitivities"""
    hostsim = _hostsim()
    cfg = configs.c2_vdp(M=96)
    r = gpu_sens(cfg)
    h = hostsim.solve_sens("vdp", [1000.0], +1, cfg["y0"].numpy(), cfg["tspan"].numpy(), cfg["options"])
    assert np.array_equal(r.y.cpu().numpy(), h["y"])
    assert np.array_equal(r.stats.raw.cpu().numpy(), h["stats"])
    assert np.array_equal(r.yd.cpu().numpy().reshape(96, 16), h["sens"].reshape(96, 16))
    assert np.array_equal(r.sens_iters.cpu().numpy(), h["iters"])
    assert len(set(r.nstg.cpu().tolist())) >= 1 and int(r.stats.raw[:, 0].min()) > 50


@pytest.mark.gpu
@pytest.mark.parametrize("fixed", [True, False])
def test_gpu_reference_sensitivity_test_layout(fixed):
    """the reference's RadauTedTest.SensitivityTest (tests/cpp/ode_solver_test.cpp:620-770): vdP with mu = 1 + 0.001 i,
    y(0) = (2 + 1e-4 i, 0), tf = 6, RelTol 1e-6 / AbsTol 1e-9, dual directions = the initial state, mu and the final
    time, each checked against central differences of the solver at the reference's rtol (1e-3 ... 1e-4).  Here in
    Hamiltonian form (costates ride along), mu as a PARAMETER direction, differences taken with the GPU solver."""
    M = 8
    tf = 6.0
    op = jb.OptionsTe(RelTol=1e-6, AbsTol=1e-9)
    if fixed:
        op.MinNbrStg = op.MaxNbrStg = 3
    fdop = jb.OptionsTe(RelTol=1e-11, AbsTol=1e-13, MaxNbrStep=1000000)
    ts = torch.tensor([0.0, tf], dtype=torch.float64, device=DEV).repeat(M, 1)
    for i in (0, 5):
        mu = 1.0 + 0.001 * i
        dyn = jb.HamiltonianSystem(jb.VanDerPol(mu), +1)
        y0 = torch.tensor([1.0, 1.0, 2.0 + 1e-4 * i, 0.0], dtype=torch.float64, device=DEV).repeat(M, 1)
        r = jb.RadauTeD(dyn, dyn.jacobian, ts, y0, op, par_dirs=[0])
        assert bool((r.solve() == 0).all())
        full = r.yd_full()
        assert tuple(full.shape) == (M, 4, 4 + 1 + 1)
        # mu direction: central differences in the parameter
        e = 1e-6
        yp = jb.RadauTe(jb.HamiltonianSystem(jb.VanDerPol(mu + e), +1), None, ts, y0, fdop)
        ym = jb.RadauTe(jb.HamiltonianSystem(jb.VanDerPol(mu - e), +1), None, ts, y0, fdop)
        yp.solve()
        ym.solve()
        fd_mu = (yp.y - ym.y) / (2 * e)
        assert torch.allclose(full[:, :, 4], fd_mu, rtol=1e-3, atol=1e-3 * float(fd_mu.abs().max()))
        # final-time direction
        tp, tm = ts.clone(), ts.clone()
        tp[:, 1] += 1e-6
        tm[:, 1] -= 1e-6
        a = jb.RadauTe(dyn, None, tp, y0, fdop)
        b = jb.RadauTe(dyn, None, tm, y0, fdop)
        a.solve()
        b.solve()
        fd_t = (a.y - b.y) / 2e-6
        assert torch.allclose(full[:, :, 5], fd_t, rtol=1e-4, atol=1e-4 * float(fd_t.abs().max()))
        # x1(0) direction
        hh = 1e-7
        ypl, ymi = y0.clone(), y0.clone()
        ypl[:, 2] += hh
        ymi[:, 2] -= hh
        a = jb.RadauTe(dyn, None, ts, ypl, fdop)
        b = jb.RadauTe(dyn, None, ts, ymi, fdop)
        a.solve()
        b.solve()
        fd_x = (a.y - b.y) / (2 * hh)
        assert torch.allclose(full[:, :, 2], fd_x, rtol=1e-3, atol=1e-3 * float(fd_x.abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("fixed", [True, False])
def test_gpu_parameter_direction_equals_the_host_run_bitwise(fixed):
    hostsim = _hostsim()
    cfg = configs.c2_vdp(M=48, mu=30.0, fixed_order=fixed)
    cfg["tspan"][:, 1] = 20.0
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    r = jb.RadauTeD(dyn, dyn.jacobian, cfg["tspan"].to(DEV), cfg["y0"].to(DEV), cfg["options"], par_dirs=[0])
    r.solve()
    h = hostsim.solve_sens("vdp", [30.0], +1, cfg["y0"].numpy(), cfg["tspan"].numpy(), cfg["options"], par_dirs=[0])
    assert np.array_equal(r.y.cpu().numpy(), h["y"]) and np.array_equal(r.yd.cpu().numpy(), h["sens"])
    assert np.array_equal(r.ypar.cpu().numpy(), h["spar"])
    with pytest.raises(ValueError, match="out of range"):
        jb.RadauTeD(dyn, dyn.jacobian, cfg["tspan"].to(DEV), cfg["y0"].to(DEV), cfg["options"], par_dirs=[1]).solve()


@pytest.mark.gpu
def test_gpu_variable_order_large_batch_sorted_epochs():
    """>= 65,536 rows: the epoch lists are regime-sorted in H(y0) order; a sample must equal the small-batch solve"""
    cfg = configs.c2_vdp(M=1 << 16)
    r = gpu_sens(cfg)
    assert bool((r.res == 0).all())
    idx = torch.arange(0, 1 << 16, 997)
    sub = dict(cfg, y0=cfg["y0"][idx], tspan=cfg["tspan"][idx])
    q = gpu_sens(sub)
    assert torch.equal(r.y[idx.to(DEV)], q.y) and torch.equal(r.yd[idx.to(DEV)], q.yd)
    assert torch.equal(r.stats.raw[idx.to(DEV)], q.stats.raw)


@pytest.mark.gpu
def test_gpu_sensitivities_harmonic_exact_and_chain_rule():
    cfg = configs.c1_harmonic(M=256)
    r = gpu_sens(cfg)
    tf = 10.0
    c, s = np.cos(tf), np.sin(tf)
    exact = torch.tensor([[c, -s], [s, c]], dtype=torch.float64, device=DEV)
    assert torch.allclose(r.yd, exact.expand_as(r.yd), atol=1e-7)
    # a non-identity dual part of y0 is propagated linearly: yd(S0) = yd(I) S0
    S0 = torch.randn(256, 2, 2, dtype=torch.float64, device=DEV, generator=torch.Generator(DEV).manual_seed(0))
    r2 = gpu_sens(cfg, y0d=S0)
    assert torch.allclose(r2.yd, r.yd @ S0, rtol=1e-9, atol=1e-10)
    assert torch.allclose(r.dydtf, torch.stack([-r.y[:, 1], r.y[:, 0]], 1), rtol=1e-13, atol=1e-15)


@pytest.mark.gpu
@pytest.mark.parametrize("mu,tf", [(1.0, 6.0), (30.0, 20.0)])
def test_gpu_sensitivities_vs_oracle_and_finite_differences(mu, tf):
    M = 32
    g = torch.Generator().manual_seed(1)
    y0 = torch.tensor([1.0, 1.0, 2.0, 0.0], dtype=torch.float64).repeat(M, 1)
    y0[:, 2] += 0.5 * torch.rand(M, generator=g, dtype=torch.float64)
    op = jb.OptionsTe(RelTol=1e-9, AbsTol=1e-12, MaxNbrStep=100000)
    cfg = dict(ham=jb.VanDerPol(mu), sign=+1, y0=y0, tspan=torch.tensor([0.0, tf], dtype=torch.float64).repeat(M, 1), options=op)
    r = gpu_sens(cfg)
    assert bool((r.res == 0).all())
    # oracle (same algorithm, numpy): tolerance-level agreement
    f, j = problems.vdp(mu, +1)
    for m in (0, 7, 31):
        o = radau_solve(f, j, 0.0, tf, y0[m].numpy(), op, sens0=np.eye(4))
        assert np.allclose(r.yd[m].cpu().numpy(), o.sens, rtol=1e-6, atol=1e-6 * np.abs(o.sens).max())
    # finite differences of the GPU solver itself, all 2 * Ny * M perturbed solves in one batch
    rel = 1e-7
    yp = y0.repeat_interleave(8, 0).clone()
    hh = torch.zeros(M * 8, dtype=torch.float64)
    for d in range(4):
        step = rel * torch.clamp(y0[:, d].abs(), min=1.0)
        yp[2 * d::8, d] += step
        yp[2 * d + 1::8, d] -= step
        hh[2 * d::8] = step
    dyn = jb.HamiltonianSystem(cfg["ham"], +1)
    pb = jb.RadauTe(dyn, None, cfg["tspan"].repeat_interleave(8, 0).to(DEV), yp.to(DEV), op)
    pb.solve()
    Y = pb.y.cpu().reshape(M, 4, 2, 4)                      # [m, direction, +/-, component]
    fd = (Y[:, :, 0, :] - Y[:, :, 1, :]) / (2 * hh.reshape(M, 4, 2)[:, :, 0:1])
    fd = fd.permute(0, 2, 1)                                # [m, component, direction]
    S = r.yd.cpu()
    assert torch.allclose(S, fd, rtol=5e-4, atol=5e-4 * float(fd.abs().max()))


@pytest.mark.gpu
def test_gpu_sensitivities_unsupported():
    dyn = jb.HamiltonianSystem(jb.MlpHamiltonian(), -1)
    y0 = torch.zeros(2, 16, dtype=torch.float64, device=DEV)
    ts = torch.tensor([0.0, 1.0], dtype=torch.float64, device=DEV).repeat(2, 1)
    with pytest.raises(jb.JanusB200Error, match="thread-resident"):
        jb.RadauTeD(dyn, None, ts, y0, jb.OptionsTe()).solve()


# ------------------------------------------------------------------ against the LIVE reference solver
def _live_fd():
    import os
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_live_sens_fd.npz")
    return np.load(p)


@pytest.mark.parametrize("mu", [1, 30])
def test_oracle_sensitivities_vs_live_reference_finite_differences(mu):
    """d y(tf)/d y0 by central differences of the UNMODIFIED reference RadauTe (oracle/_ref, tests/golden/
    make_reference_sens_fd.py) -- the reference's own way of checking RadauTeD (ode_solver_test.cpp:620-770)."""
    g = _live_fd()
    mu_f, tf = g[f"par_mu{mu}"]
    f, j = problems.vdp(float(mu_f), +1)
    op = RadauOptions(RelTol=1e-10, AbsTol=1e-13, MaxNbrStep=1000000)
    for m, y0 in enumerate(g["y0"]):
        r = radau_solve(f, j, 0.0, float(tf), y0.copy(), op, sens0=np.eye(4))
        S = g[f"S_mu{mu}"][m]
        assert np.allclose(r.y, g[f"y_mu{mu}"][m], rtol=1e-7, atol=1e-9)
        assert np.allclose(r.sens, S, rtol=1e-4, atol=1e-4 * np.abs(S).max())


@pytest.mark.gpu
@pytest.mark.parametrize("mu", [1, 30])
@pytest.mark.parametrize("fixed", [True, False])
def test_gpu_sensitivities_vs_live_reference_finite_differences(mu, fixed):
    g = _live_fd()
    mu_f, tf = g[f"par_mu{mu}"]
    op = jb.OptionsTe(RelTol=1e-10, AbsTol=1e-13, MaxNbrStep=1000000)
    if fixed:
        op.MinNbrStg = op.MaxNbrStg = 3
    dyn = jb.HamiltonianSystem(jb.VanDerPol(float(mu_f)), +1)
    y0 = torch.from_numpy(g["y0"]).cuda()
    ts = torch.tensor([0.0, float(tf)], dtype=torch.float64).repeat(3, 1).cuda()
    r = jb.RadauTeD(dyn, dyn.jacobian, ts, y0, op)
    res = r.solve()
    assert bool((res == 0).all())
    S = g[f"S_mu{mu}"]
    assert np.allclose(r.yd.cpu().numpy(), S, rtol=1e-4, atol=1e-4 * np.abs(S).max())
    assert np.allclose(r.y.cpu().numpy(), g[f"y_mu{mu}"], rtol=1e-7, atol=1e-9)
