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
