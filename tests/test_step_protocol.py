"""Incremental state / step(htry, dyns, ft, state) protocol (reference: src/python/seulext.py:8-148,158-581,
driver loop src/python/seulex_dual_example_vdp.py:136-175) carried by the Radau stepper."""
import numpy as np
import pytest
import torch

import janus_ode_b200 as jb
from janus_ode_b200 import configs
from oracle import c_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fixed", [True, False])
def test_step_loop_reproduces_solve_bit_for_bit(fixed):
    cfg = configs.c2_vdp(M=1024, fixed_order=fixed)
    cfg["tspan"][:, 1] = 900.0
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    ft = cfg["tspan"][:, 1].cuda()
    state = jb.RadauState(cfg["y0"].cuda(), cfg["tspan"][:, 0].cuda(), atol=1e-9, rtol=1e-6, ft=ft)
    solver = jb.RadauT(cfg["options"])
    htry = cfg["options"].InitialStep
    n = 0
    x_prev = state.x.clone()
    while not bool(state.done.all()):
        state = solver.step(htry, dyn, ft, state)
        htry = state.hnext
        live = ~state.done
        # every unfinished trajectory advanced by exactly the step it reports
        assert torch.equal((x_prev + state.hdid)[live], state.x[live])
        x_prev = state.x.clone()
        n += 1
        assert n < 5000
    o = c_oracle.solve("vdp", [1000.0], 1, cfg["y0"].numpy(), cfg["tspan"].numpy(), cfg["options"], arith="mirror")
    assert np.array_equal(state.y.cpu().numpy(), o["y"]) and np.array_equal(state.x.cpu().numpy(), o["t"])
    assert np.array_equal(state.stats.raw.cpu().numpy(), o["stats"]) and (state.res == 0).all()
    assert n == int(o["stats"][:, 5].max())          # one accepted step per call for the longest trajectory


def test_step_with_user_chosen_trial_step():
    cfg = configs.c1_harmonic(M=64)
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    state = jb.RadauState(cfg["y0"].cuda(), cfg["tspan"][:, 0].cuda(), atol=1e-12, rtol=1e-9, ft=10.0)
    solver = jb.RadauT(cfg["options"])
    state = solver.step(1e-3, dyn, 10.0, state)
    assert (state.hdid > 0).all() and (state.x == state.hdid).all()
    while not bool(state.done.all()):
        state = solver.step(torch.minimum(state.hnext, torch.full_like(state.hnext, 0.05)), dyn, 10.0, state)  # cap h
    y0 = cfg["y0"].numpy()
    exact = np.stack([y0[:, 0] * np.cos(10) - y0[:, 1] * np.sin(10), y0[:, 1] * np.cos(10) + y0[:, 0] * np.sin(10)], 1)
    assert np.abs(state.y.cpu().numpy() - exact).max() < 5e-9 and (state.x == 10.0).all()
    assert (state.hdid <= 0.05 + 1e-15).all()


def test_step_loop_on_a_large_batch_uses_the_ordered_lists():
    """>= 65,536 rows: the epoch lists are sorted by regime key, then H(y0) (csrc/solver_impl.cuh build_order); the step loop
    must still be the same integration as solve(), bit for bit"""
    cfg = configs.c2_vdp(M=70000)
    cfg["tspan"][:, 1] = 40.0
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    ft = cfg["tspan"][:, 1].cuda()
    state = jb.RadauState(cfg["y0"].cuda(), cfg["tspan"][:, 0].cuda(), atol=1e-9, rtol=1e-6, ft=ft)
    solver = jb.RadauT(cfg["options"])
    htry, n = cfg["options"].InitialStep, 0
    while not bool(state.done.all()):
        state = solver.step(htry, dyn, ft, state)
        htry = state.hnext
        n += 1
        assert n < 500
    r = jb.RadauTe(dyn, None, cfg["tspan"].cuda(), cfg["y0"].cuda(), cfg["options"])
    r.solve()
    assert torch.equal(state.y, r.y) and torch.equal(state.stats.raw, r.stats.raw) and torch.equal(state.x, r.t)
    idx = torch.arange(0, 70000, 97)
    o = c_oracle.solve("vdp", [1000.0], 1, cfg["y0"][idx].numpy(), cfg["tspan"][idx].numpy(), cfg["options"], arith="mirror")
    assert np.array_equal(r.y[idx.cuda()].cpu().numpy(), o["y"])
