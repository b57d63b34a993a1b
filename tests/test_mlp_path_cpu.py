"""MLP Hamiltonian (config 4) on the CPU: (a) the mirror arithmetic of the GEMM kernels
(oracle/c mlp_eval_mirror) against torch autograd, the engine the reference differentiates H with
(janus_ode_common.hpp:2161-2256); (b) the phase-split stepper the GPU kernels wrap
(csrc/mlp_phase.cuh, run by tests/hostsim) bit for bit against the oracle's scalar RADAU in mirror
arithmetic; (c) mirror vs reference arithmetic at tolerance level; (d) the committed golden vectors."""
import os
import sys

import numpy as np
import pytest
import torch

import janus_ode_b200 as jb
from janus_ode_b200 import configs
from oracle import c_oracle, ham

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
import hostsim  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_product_weights_equal_oracle_weights():
    """janus_ode_b200.MlpHamiltonian draws the same seeded weights as oracle/ham.py MLPHamiltonian."""
    H = jb.MlpHamiltonian(8, 128, 0)
    o = ham.MLPHamiltonian(N=8, width=128, seed=0)
    par = [8.0, 128.0]
    for W, b in zip(o.Ws, o.bs):
        par += W.reshape(-1).tolist() + b.reshape(-1).tolist()
    assert len(H.params) == 35331 and np.array_equal(np.array(H.params), np.array(par))


@pytest.mark.parametrize("sign", [-1, +1])
def test_mirror_mlp_vs_autograd(sign):
    o = ham.MLPHamiltonian(N=8, width=128, seed=0)
    par = np.array(jb.MlpHamiltonian(8, 128, 0).params)
    torch.manual_seed(5)
    y = 0.5 * torch.randn(6, 16, dtype=torch.float64)
    d = ham.eval_dyns(y, None, o, costate_sign=sign).numpy()
    J = ham.eval_jac(y, None, o, costate_sign=sign).numpy()
    for arith in ("reference", "mirror"):
        assert np.allclose(c_oracle.eval_dyns("mlp", par, sign, y.numpy(), arith=arith), d, rtol=1e-11, atol=1e-13)
        assert np.allclose(c_oracle.eval_jac("mlp", par, sign, y.numpy(), arith=arith), J, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("opts", [dict(MinNbrStg=3, MaxNbrStg=3), dict(), dict(RelTol=1e-9, AbsTol=1e-12),
                                  dict(NbrStg=1, MinNbrStg=1, MaxNbrStg=7, RelTol=1e-4, AbsTol=1e-6)])
def test_phase_machine_bitwise_vs_scalar_oracle(opts):
    cfg = configs.c4_mlp(M=12)
    op = cfg["options"]
    for k, v in opts.items():
        setattr(op, k, v)
    par = np.array(cfg["ham"].params)
    o = c_oracle.solve("mlp", par, -1, cfg["y0"].numpy(), cfg["tspan"].numpy(), op, arith="mirror")
    c_oracle.set_arith("mirror")
    h = hostsim.solve_mlp(par, -1, cfg["y0"].numpy(), cfg["tspan"].numpy(), op, c_oracle.lib())
    for k in ("y", "t", "h", "res", "stats", "nstg"):
        assert np.array_equal(o[k], h[k]), k
    assert (o["res"] == 0).all() and (h["work"] > 0).all()


def test_phase_machine_reverse_time_and_bias():
    H = jb.MlpHamiltonian(8, 128, 3, bias_scale=0.2)
    par = np.array(H.params)
    torch.manual_seed(2)
    y0 = (0.4 * torch.randn(6, 16, dtype=torch.float64)).numpy()
    ts = np.tile(np.array([1.0, 0.25]), (6, 1))
    op = jb.OptionsTe(RelTol=1e-6, AbsTol=1e-9)
    o = c_oracle.solve("mlp", par, -1, y0, ts, op, arith="mirror")
    c_oracle.set_arith("mirror")
    h = hostsim.solve_mlp(par, -1, y0, ts, op, c_oracle.lib())
    for k in ("y", "t", "h", "res", "stats", "nstg"):
        assert np.array_equal(o[k], h[k]), k
    assert np.all(o["t"] == 0.25)


def test_mirror_vs_reference_arithmetic():
    cfg = configs.c4_mlp(M=32)
    par = np.array(cfg["ham"].params)
    a = c_oracle.solve("mlp", par, -1, cfg["y0"].numpy(), cfg["tspan"].numpy(), cfg["options"], arith="mirror")
    b = c_oracle.solve("mlp", par, -1, cfg["y0"].numpy(), cfg["tspan"].numpy(), cfg["options"], arith="reference")
    rel = np.abs(a["y"] - b["y"]).max(1) / np.abs(b["y"]).max(1)
    assert rel.max() < 1e-8                                   # the 1e-8 bar of BASELINE.json
    assert (a["stats"][:, 5] == b["stats"][:, 5]).mean() >= 0.999   # smooth problem: the count bar holds across arithmetics


@pytest.mark.parametrize("tag", ["s3", "var"])
def test_golden_c4(tag):
    g = np.load(os.path.join(GOLD, f"c4_mlp_{tag}.npz"))
    cfg = configs.c4_mlp(M=g["y0"].shape[0], fixed_order=(tag == "s3"))
    assert np.array_equal(cfg["y0"].numpy(), g["y0"])
    par = np.array(cfg["ham"].params)
    m = c_oracle.solve("mlp", par, -1, g["y0"], g["tspan"], cfg["options"], arith="mirror")
    r = c_oracle.solve("mlp", par, -1, g["y0"], g["tspan"], cfg["options"], arith="reference")
    assert np.array_equal(m["y"], g["mir_y"]) and np.array_equal(m["stats"], g["mir_stats"])
    assert np.allclose(r["y"], g["ref_y"], rtol=1e-9, atol=1e-12) and np.array_equal(r["stats"], g["ref_stats"])
