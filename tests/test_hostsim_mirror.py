"""Arithmetic contract: the CUDA integrator core (janus_ode_b200/csrc/radau_core.cuh, compiled
for the host by tests/hostsim -- a test tool, never shipped) is bit-identical to the C oracle in
mirror mode: states, times, last step sizes, result codes, all 8 counters, final orders."""
import numpy as np
import pytest

import hostsim
from janus_ode_b200 import configs
from oracle import c_oracle


def _cmp(a, b):
    for k in ("y", "t", "h", "res", "stats", "nstg"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k


CFG = {
    "harmonic": lambda **k: configs.c1_harmonic(M=48, **k),
    "vdp": lambda **k: configs.c2_vdp(M=96, **k),
    "vdp_pmp_smooth": lambda **k: configs.c3_pmp(M=24, control="smooth", **k),
    "vdp_pmp_bang": lambda **k: configs.c3_pmp(M=24, control="bangbang", **k),
}


@pytest.mark.parametrize("ham", list(CFG))
@pytest.mark.parametrize("fixed", [True, False])
def test_core_equals_mirror_oracle(ham, fixed):
    cfg = CFG[ham](fixed_order=fixed)
    par = list(cfg["ham"].params)
    a = hostsim.solve(ham, par, cfg["sign"], cfg["y0"].numpy(), cfg["tspan"].numpy(), cfg["options"])
    b = c_oracle.solve(ham, par, cfg["sign"], cfg["y0"].numpy(), cfg["tspan"].numpy(), cfg["options"], arith="mirror")
    assert (a["res"] == 0).all()
    _cmp(a, b)


def test_core_equals_mirror_other_orders_and_failures():
    cfg = configs.c2_vdp(M=16)
    y0, ts = cfg["y0"].numpy(), cfg["tspan"].numpy().copy()
    ts[:, 1] = np.linspace(0.5, 30.0, 16)        # ragged horizons
    ts[8:, 0] = 5.0
    for kw in (dict(NbrStg=1, MinNbrStg=1, MaxNbrStg=1), dict(NbrStg=1, MinNbrStg=1, MaxNbrStg=3),
               dict(NbrStg=5, MinNbrStg=5, MaxNbrStg=5), dict(NbrStg=7, MinNbrStg=3, MaxNbrStg=7),
               dict(MaxNbrStep=25), dict(Gustafsson=False), dict(Start_Newt=True), dict(RelTol=[1e-5, 1e-6, 1e-7, 1e-8], AbsTol=1e-10)):
        op = configs.OptionsTe(RelTol=1e-6, AbsTol=1e-9, MaxNbrStep=200000)
        for k, v in kw.items():
            setattr(op, k, v)
        a = hostsim.solve("vdp", [1000.0], 1, y0, ts, op)
        b = c_oracle.solve("vdp", [1000.0], 1, y0, ts, op, arith="mirror")
        _cmp(a, b)
    # finite-time blow-up of the smooth feedback law: "step too small" (res 1) on both sides
    cfgb = configs.c3_pmp(M=8, control="smooth", tfinal=10.0)
    a = hostsim.solve("vdp_pmp_smooth", [1.0, 1.0, 100.0], 1, cfgb["y0"].numpy(), cfgb["tspan"].numpy(), cfgb["options"])
    b = c_oracle.solve("vdp_pmp_smooth", [1.0, 1.0, 100.0], 1, cfgb["y0"].numpy(), cfgb["tspan"].numpy(), cfgb["options"], arith="mirror")
    assert (a["res"] == 1).all()
    _cmp(a, b)
    # blow-up / NaN paths give the same failure codes
    op = configs.OptionsTe(RelTol=1e-6, AbsTol=1e-9, MaxNbrStep=100000)
    ts2 = cfg["tspan"].numpy()
    y1 = y0.copy()
    y1[3, 2] = np.nan
    a = hostsim.solve("vdp", [1000.0], -1, y1, ts2, op)
    b = c_oracle.solve("vdp", [1000.0], -1, y1, ts2, op, arith="mirror")
    assert (a["res"] != 0).any()
    assert np.array_equal(a["res"], b["res"]) and np.array_equal(a["stats"], b["stats"])


@pytest.mark.parametrize("K", [1, 5, 24, -1])   # -1: step() protocol, stop after every accepted step
def test_epoch_scheme_is_bit_identical(K):
    """variable-order path: park / regroup by stage count / resume (LU recomputed) every K attempted
    steps gives the same bits as the uninterrupted integration."""
    for ham, mk in (("vdp", lambda: configs.c2_vdp(M=192)), ("harmonic", lambda: configs.c1_harmonic(M=32)),
                    ("vdp_pmp_bang", lambda: configs.c3_pmp(M=16))):
        cfg = mk()
        par = list(cfg["ham"].params)
        a = hostsim.solve(ham, par, cfg["sign"], cfg["y0"].numpy(), cfg["tspan"].numpy(), cfg["options"], epochs_K=K)
        b = c_oracle.solve(ham, par, cfg["sign"], cfg["y0"].numpy(), cfg["tspan"].numpy(), cfg["options"], arith="mirror")
        _cmp(a, b)
