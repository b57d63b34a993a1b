"""Output at the interior points of tspan (OutFlag 3: src/matlab/radau.m:1150-1168, ntrprad :1291-1308 ==
src/cpp/radaute_impl.hpp:1413-1440; row a15 of SURVEY.md 8)."""
import numpy as np
import pytest
import torch

import hostsim
import janus_ode_b200 as jb
from janus_ode_b200 import configs
from oracle import c_oracle


def _harmonic_exact(y0, ts):
    return np.stack([y0[:, 0, None] * np.cos(ts) - y0[:, 1, None] * np.sin(ts),
                     y0[:, 1, None] * np.cos(ts) + y0[:, 0, None] * np.sin(ts)], 2)


@pytest.mark.parametrize("arith", ["reference", "mirror"])
def test_oracle_dense_output_matches_analytic_solution(arith):
    cfg = configs.c1_harmonic(M=16)
    ts = np.tile(np.linspace(0, 10, 41), (16, 1))
    o = c_oracle.solve_dense("harmonic", [1.0], -1, cfg["y0"].numpy(), ts, cfg["options"], arith=arith)
    assert (o["nout"] == 41).all() and np.array_equal(o["tout"], ts)
    assert np.abs(o["yout"] - _harmonic_exact(cfg["y0"].numpy(), ts)).max() < 2e-7
    assert np.array_equal(o["yout"][:, -1], o["y"])                     # last point = final state
    # requesting output does not perturb the integration
    o2 = c_oracle.solve("harmonic", [1.0], -1, cfg["y0"].numpy(), ts[:, [0, -1]], cfg["options"], arith=arith)
    assert np.array_equal(o["y"], o2["y"]) and np.array_equal(o["stats"], o2["stats"])


def test_core_dense_output_is_bit_identical_to_mirror_oracle():
    for ham, cfg, nts in (("harmonic", configs.c1_harmonic(M=16), 33), ("vdp", configs.c2_vdp(M=24), 150),
                          ("vdp", configs.c2_vdp(M=24, fixed_order=True), 7)):
        t0, tf = cfg["tspan"][0, 0].item(), cfg["tspan"][0, 1].item()
        ts = np.tile(np.linspace(t0, tf, nts), (cfg["y0"].shape[0], 1))
        par = list(cfg["ham"].params)
        a = hostsim.solve(ham, par, cfg["sign"], cfg["y0"].numpy(), ts, cfg["options"])
        b = c_oracle.solve_dense(ham, par, cfg["sign"], cfg["y0"].numpy(), ts, cfg["options"], arith="mirror")
        for k in ("y", "t", "stats", "tout", "yout", "nout"):
            assert np.array_equal(a[k], b[k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("fixed", [True, False])
def test_gpu_dense_output(fixed):
    cfg = configs.c2_vdp(M=2048, fixed_order=fixed)
    nts = 64
    tf = cfg["tspan"][0, 1].item()
    ts = torch.linspace(0, tf, nts, dtype=torch.float64).repeat(2048, 1)
    ts[1000:, 1:-1] *= 0.5                                             # ragged interior grids
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    r = jb.RadauTe(dyn, dyn.jacobian, ts.cuda(), cfg["y0"].cuda(), cfg["options"])
    assert (r.solve() == 0).all()
    o = c_oracle.solve_dense("vdp", [1000.0], 1, cfg["y0"].numpy(), ts.numpy(), cfg["options"], arith="mirror")
    assert np.array_equal(r.tout.cpu().numpy(), o["tout"]) and np.array_equal(r.yout.cpu().numpy(), o["yout"])
    assert np.array_equal(r.nout.cpu().numpy(), o["nout"]) and (o["nout"] == nts).all()
    assert np.array_equal(r.y.cpu().numpy(), o["y"]) and np.array_equal(r.stats.raw.cpu().numpy(), o["stats"])
    # reverse time + harmonic analytic check
    c1 = configs.c1_harmonic(M=64)
    tsr = torch.linspace(10, 0, 21, dtype=torch.float64).repeat(64, 1)
    dyn = jb.HamiltonianSystem(c1["ham"], c1["sign"])
    r = jb.RadauTe(dyn, None, tsr.cuda(), c1["y0"].cuda(), c1["options"])
    r.solve()
    y0 = c1["y0"].numpy()
    dt = tsr.numpy() - 10.0
    assert np.abs(r.yout.cpu().numpy() - _harmonic_exact(y0, dt)).max() < 2e-7
