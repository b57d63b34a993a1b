"""Explicit DOPRI853 path.  CPU: the numpy oracle (oracle/dopri853.py) against golden vectors
of the LIVE reference class (tests/golden/make_dopri_golden.py imports /root/reference/src/python/dopri853te.py).
GPU: jb.Dopri853Te (jb200_dopri853_step) against the oracle on the same inputs."""
import os

import numpy as np
import pytest
import torch

import janus_ode_b200 as jb
from oracle.dopri853 import Dopri853Oracle
from oracle.dopri853_tables import EXACT, REF

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dopri853_ref.npz"))


def f_harmonic(t, y):
    return np.stack([-y[:, 1], y[:, 0]], 1)


def f_vdp(t, y, mu=2.0):
    p1, p2, x1, x2 = y[:, 0], y[:, 1], y[:, 2], y[:, 3]
    return np.stack([p2 * (-2.0 * mu * x1 * x2 - 1.0), p1 + p2 * (mu * (1.0 - x1 * x1)), x2,
                     mu * (1.0 - x1 * x1) * x2 - x1], 1)


CASES = {"harmonic": (f_harmonic, jb.Harmonic(1.0), -1), "vdp": (f_vdp, jb.VanDerPol(2.0), +1),
         "vdp_loose": (f_vdp, jb.VanDerPol(2.0), +1)}


def oracle_run(name, nsteps=None):
    f = CASES[name][0]
    y0 = G[name + "_y0"]
    atol, rtol = G[name + "_tol"]
    s = Dopri853Oracle(y0, f(None, y0), np.zeros(len(y0)), atol, rtol, f)
    h = np.full(len(y0), float(G[name + "_h0"]))
    xs, ys, hd, hn, calls = [], [], [], [], []
    for _ in range(nsteps or len(G[name + "_x"])):
        s.step(h)
        xs.append(s.x.copy()); ys.append(s.y.copy()); hd.append(s.hdid.copy()); hn.append(s.hnext.copy())
        calls.append(int(11 * (len(y0) + s.nrej.sum()) + len(y0)))     # 11 stage rows per attempt + f(x+h, yout)
        h = s.hnext.copy()
    return np.array(xs), np.array(ys), np.array(hd), np.array(hn), np.array(calls)


@pytest.mark.parametrize("name", ["harmonic", "vdp", "vdp_loose"])
def test_oracle_matches_live_reference(name):
    x, y, hd, hn, calls = oracle_run(name)
    assert np.array_equal(calls, G[name + "_calls"])                  # same accept / reject decisions in every call
    assert np.allclose(x, G[name + "_x"], rtol=1e-11, atol=0)
    assert np.allclose(y, G[name + "_y"], rtol=1e-11, atol=1e-13)
    assert np.allclose(hd, G[name + "_hdid"], rtol=1e-10) and np.allclose(hn, G[name + "_hnext"], rtol=1e-10)


def test_reference_quirks_are_visible():
    """hnext never exceeds hdid (control() starts from reject = all-true); the float32 tableau limits the accuracy
    to ~1e-8 relative per unit time whatever the tolerance, the exact DOP853 constants reach roundoff."""
    y0 = G["harmonic_y0"]
    p0, x0 = y0[:, 0], y0[:, 1]
    errs = {}
    for name, tab in (("ref", REF), ("exact", EXACT)):
        s = Dopri853Oracle(y0, f_harmonic(None, y0), np.zeros(len(y0)), 1e-9, 1e-9, f_harmonic, tables=tab)
        h = np.full(len(y0), 0.5)
        for _ in range(10):
            s.step(h)
            assert (s.hnext <= s.hdid).all()
            h = s.hnext.copy()
        errs[name] = np.abs(s.y[:, 1] - (x0 * np.cos(s.x) + p0 * np.sin(s.x))).max()
    assert errs["exact"] < 1e-14 and errs["ref"] > 1e-11


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["harmonic", "vdp", "vdp_loose"])
def test_gpu_matches_oracle_and_live_reference(name):
    f, ham, sign = CASES[name]
    dyn = jb.HamiltonianSystem(ham, sign)
    y0 = torch.from_numpy(G[name + "_y0"]).cuda()
    atol, rtol = G[name + "_tol"]
    s = jb.Dopri853Te(y0, None, torch.zeros(len(y0), dtype=torch.float64, device="cuda"), atol, rtol, False, dyn)
    assert np.allclose(s.dydx.cpu().numpy(), f(None, G[name + "_y0"]), rtol=1e-14, atol=1e-15)
    h = torch.full((len(y0),), float(G[name + "_h0"]), dtype=torch.float64, device="cuda")
    xo, yo, hdo, hno, _ = oracle_run(name)
    for k in range(len(xo)):
        s.step(h)
        h = s.hnext.clone()
        assert np.allclose(s.x.cpu().numpy(), xo[k], rtol=1e-12)
        assert np.allclose(s.y.cpu().numpy(), yo[k], rtol=1e-11, atol=1e-13)
        assert np.allclose(s.hdid.cpu().numpy(), hdo[k], rtol=1e-10) and np.allclose(s.hnext.cpu().numpy(), hno[k], rtol=1e-10)
    # ... and therefore the live reference
    assert np.allclose(s.y.cpu().numpy(), G[name + "_y"][-1], rtol=1e-10, atol=1e-12)
    assert np.allclose(s.x.cpu().numpy(), G[name + "_x"][-1], rtol=1e-10)


@pytest.mark.gpu
def test_gpu_large_batch_and_norm_M():
    """error norm divides by D * M: a sharded run with norm_M = global batch reproduces the unsharded steps."""
    torch.manual_seed(0)
    M = 4096
    y0 = (torch.rand(M, 2, dtype=torch.float64) * 2 - 1).cuda()
    dyn = jb.HamiltonianSystem(jb.Harmonic(1.0), -1)
    x0 = torch.zeros(M, dtype=torch.float64, device="cuda")
    full = jb.Dopri853Te(y0, None, x0, 1e-6, 1e-6, False, dyn)
    half = jb.Dopri853Te(y0[::2].contiguous(), None, x0[::2].contiguous(), 1e-6, 1e-6, False, dyn, norm_M=M)
    alone = jb.Dopri853Te(y0[::2].contiguous(), None, x0[::2].contiguous(), 1e-6, 1e-6, False, dyn)
    hf = hh = ha = 0.3
    for _ in range(5):
        full.step(hf); half.step(hh); alone.step(ha)
        hf, hh, ha = full.hnext.clone(), half.hnext.clone(), alone.hnext.clone()
    assert torch.equal(full.y[::2], half.y) and torch.equal(full.x[::2], half.x)
    assert not torch.equal(full.x[::2], alone.x)
    with pytest.raises(jb.JanusB200Error):
        jb.Dopri853Te(torch.zeros(4, 16, dtype=torch.float64, device="cuda"), None, 0.0, 1e-6, 1e-6, False,
                      jb.HamiltonianSystem(jb.MlpHamiltonian(), -1)).step(0.1)


# ------------------------------------------------------------------ dual part (the reference's Dopri853TeD)
def _hostsim():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
    import hostsim
    return hostsim


def test_cuda_core_dual_part_on_host_vs_finite_differences():
    """csrc/dopri853.cuh dopri_dual_one (the stages differentiated one by one, J from hyper-dual numbers), run on the host:
    d y(tf) / d y0 against central differences of the same stepper -- how the reference checks its dual solvers
    (tests/cpp/ode_solver_test.cpp:620-770) -- and the real part is untouched"""
    hostsim = _hostsim()
    y0 = np.array([[1.0, 1.0, 2.0, 0.0], [0.5, -1.0, 1.7, 0.2]])
    kw = dict(atol=1e-7, rtol=1e-7, h0=0.02)
    r = hostsim.dopri_dual("vdp", [1.0], +1, y0, 0.0, 3.0, **kw)
    r0 = hostsim.dopri_dual("vdp", [1.0], +1, y0, 0.0, 3.0, dual=False, **kw)
    assert np.array_equal(r["y"], r0["y"]) and np.array_equal(r["steps"], r0["steps"]) and (r["steps"] >= 150).all()
    S = np.zeros((2, 4, 4))
    for d in range(4):
        e = 1e-6
        yp, ym = y0.copy(), y0.copy()
        yp[:, d] += e
        ym[:, d] -= e
        S[:, :, d] = (hostsim.dopri_dual("vdp", [1.0], +1, yp, 0.0, 3.0, dual=False, **kw)["y"]
                      - hostsim.dopri_dual("vdp", [1.0], +1, ym, 0.0, 3.0, dual=False, **kw)["y"]) / (2 * e)
    assert np.allclose(r["sens"], S, rtol=1e-4, atol=1e-4 * np.abs(S).max())
    # ... and the implicit solver's sensitivities of the same problem (independent algorithm)
    ts = np.tile(np.array([0.0, 3.0]), (2, 1))
    h = hostsim.solve_sens("vdp", [1.0], +1, y0, ts, jb.OptionsTe(RelTol=1e-10, AbsTol=1e-13, MaxNbrStep=100000))
    assert np.allclose(r["sens"], h["sens"], rtol=1e-3, atol=1e-3 * np.abs(h["sens"]).max())


@pytest.mark.gpu
def test_gpu_dual_part():
    torch.manual_seed(1)
    M = 512
    y0 = torch.tensor([1.0, 1.0, 2.0, 0.0], dtype=torch.float64).repeat(M, 1)
    y0[:, 2] += 0.5 * torch.rand(M, dtype=torch.float64)
    y0 = y0.cuda()
    dyn = jb.HamiltonianSystem(jb.VanDerPol(1.0), +1)
    x0 = torch.zeros(M, dtype=torch.float64, device="cuda")
    mk = lambda y, cls, **kw: cls(y, None, x0, 1.0, 1.0, False, dyn, norm_M=1, **kw)     # error control off: no rejected attempts
    s, p = mk(y0, jb.Dopri853TeD), mk(y0, jb.Dopri853Te)
    fd = []
    for d in range(4):
        yp, ym = y0.clone(), y0.clone()
        yp[:, d] += 1e-6
        ym[:, d] -= 1e-6
        fd.append((mk(yp, jb.Dopri853Te), mk(ym, jb.Dopri853Te)))
    for _ in range(150):
        for o in [s, p] + [q for pair in fd for q in pair]:
            o.step(0.02)
    assert int(s.nrej.sum()) == 0
    assert torch.equal(s.y, p.y) and torch.equal(s.x, p.x) and torch.equal(s.dydx, p.dydx)      # real part untouched
    S = torch.stack([(a.y - b.y) / 2e-6 for a, b in fd], 2)
    assert torch.allclose(s.yd, S, rtol=1e-5, atol=1e-5 * float(S.abs().max()))
    # a non-identity dual part is propagated linearly
    S0 = torch.randn(M, 4, 3, dtype=torch.float64, device="cuda")
    t = mk(y0, jb.Dopri853TeD, yd=S0)
    for _ in range(150):
        t.step(0.02)
    assert torch.allclose(t.yd, s.yd @ S0, rtol=1e-9, atol=1e-9)
    with pytest.raises(ValueError):
        mk(y0, jb.Dopri853TeD, yd=torch.zeros(M, 4, 9, dtype=torch.float64, device="cuda"))
