"""User-supplied Hamiltonians: the library recompiled around a C++ expression (janus_ode_b200/user.py)."""
import numpy as np
import pytest
import torch

import janus_ode_b200 as jb
from janus_ode_b200 import _lib, configs


def test_plugin_builds_and_exports_the_abi():
    ham = jb.compile_hamiltonian("pendulum", 1, "0.5*p1*p1 - cos(x1)")
    assert ham.ham_id == 100 and ham.Ny == 2
    L = _lib.lib(ham.lib_path)
    for s in _lib.SYMBOLS:
        assert getattr(L, s) is not None
    with pytest.raises(RuntimeError, match="nvcc failed"):
        jb.compile_hamiltonian("broken", 1, "p1 +* x1")
    with pytest.raises(ValueError):
        jb.compile_hamiltonian("big", 5, "p1")


@pytest.mark.gpu
def test_user_vdp_is_bit_identical_to_builtin():
    user = jb.compile_hamiltonian("myvdp", 2, "p1 * x2 + p2 * (par[0] * (1.0 - x1 * x1) * x2 - x1)", params=(1000.0,))
    cfg = configs.c2_vdp(M=2048)
    out = []
    for ham in (user, cfg["ham"]):
        dyn = jb.HamiltonianSystem(ham, +1)
        r = jb.RadauTe(dyn, dyn.jacobian, cfg["tspan"].cuda(), cfg["y0"].cuda(), cfg["options"])
        r.solve()
        out.append((r.y.cpu(), r.stats.raw.cpu(), r.t.cpu()))
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1]) and torch.equal(out[0][2], out[1][2])


@pytest.mark.gpu
def test_user_pendulum_vs_autograd_oracle_and_energy():
    from oracle import ham as oham, problems
    from oracle.radau_scalar import radau_solve
    ham = jb.compile_hamiltonian("pendulum", 1, "0.5*p1*p1 - cos(x1)")
    dyn = jb.HamiltonianSystem(ham, costate_sign=-1)
    torch.manual_seed(0)
    y0 = torch.rand(256, 2, dtype=torch.float64) * 2 - 1
    ts = torch.tensor([0.0, 20.0], dtype=torch.float64).repeat(256, 1)
    op = jb.OptionsTe(RelTol=1e-9, AbsTol=1e-12, MaxNbrStep=100000)
    # stage evaluation against the reference's autograd engine
    Ht = lambda x, p, W: 0.5 * p[:, 0].pow(2) - torch.cos(x[:, 0])
    f = dyn(None, y0.cuda()).cpu()
    J = dyn.jacobian(None, y0.cuda()).cpu()
    assert torch.allclose(f, oham.eval_dyns(y0, None, Ht, costate_sign=-1), rtol=1e-14, atol=1e-15)
    assert torch.allclose(J, oham.eval_jac(y0, None, Ht, costate_sign=-1), rtol=1e-14, atol=1e-15)
    r = jb.RadauTe(dyn, dyn.jacobian, ts.cuda(), y0.cuda(), op)
    assert (r.solve() == 0).all()
    y = r.y.cpu()
    E0 = 0.5 * y0[:, 0] ** 2 - torch.cos(y0[:, 1])
    E1 = 0.5 * y[:, 0] ** 2 - torch.cos(y[:, 1])
    assert (E1 - E0).abs().max() < 1e-7                      # energy conserved to tolerance level
    fo, jo = problems.from_hamiltonian(Ht, sign=-1)
    for m in range(3):                                       # scalar oracle driven by autograd
        o = radau_solve(fo, jo, 0.0, 20.0, y0[m].numpy(), op)
        assert np.allclose(o.y, y[m].numpy(), rtol=1e-6, atol=1e-8)


CHAIN3 = ("0.5*(p1*p1+p2*p2+p3*p3) + 0.5*(x1*x1 + square(x2-x1) + square(x3-x2) + x3*x3) "
          "+ 0.25*par[0]*(square(square(x2-x1)) + square(square(x3-x2)))")


@pytest.mark.gpu
def test_user_three_dof_chain_ny6():
    """N = 3 (Ny = 6) Fermi-Pasta-Ulam-type chain: generic sizes beyond the built-ins; checked against
    SciPy DOP853 and energy conservation, fixed and variable order."""
    from scipy.integrate import solve_ivp
    ham = jb.compile_hamiltonian("chain3", 3, CHAIN3, params=(1.0,))
    dyn = jb.HamiltonianSystem(ham, costate_sign=-1)
    torch.manual_seed(1)
    y0 = torch.rand(128, 6, dtype=torch.float64) - 0.5
    ts = torch.tensor([0.0, 5.0], dtype=torch.float64).repeat(128, 1)

    def rhs(t, y):
        p, x = y[:3], y[3:]
        d1, d2 = x[1] - x[0], x[2] - x[1]
        dV = np.array([x[0] - d1 - d1 ** 3, d1 + d1 ** 3 - d2 - d2 ** 3, d2 + d2 ** 3 + x[2]])
        return np.concatenate([-dV, p])

    def energy(y):
        p, x = y[:, :3], y[:, 3:]
        d1, d2 = x[:, 1] - x[:, 0], x[:, 2] - x[:, 1]
        return 0.5 * (p ** 2).sum(1) + 0.5 * (x[:, 0] ** 2 + d1 ** 2 + d2 ** 2 + x[:, 2] ** 2) + 0.25 * (d1 ** 4 + d2 ** 4)
    f = dyn(None, y0.cuda()).cpu().numpy()
    assert np.allclose(f, np.array([rhs(0, r) for r in y0.numpy()]), rtol=1e-13, atol=1e-14)
    for kw in (dict(), dict(MinNbrStg=3, MaxNbrStg=3)):
        op = jb.OptionsTe(RelTol=1e-9, AbsTol=1e-12, MaxNbrStep=100000, **kw)
        r = jb.RadauTe(dyn, dyn.jacobian, ts.cuda(), y0.cuda(), op)
        assert (r.solve() == 0).all()
        y = r.y.cpu()
        assert (energy(y) - energy(y0)).abs().max() < 1e-7
        for m in range(3):
            sol = solve_ivp(rhs, (0, 5), y0[m].numpy(), method="DOP853", rtol=1e-12, atol=1e-14)
            assert np.allclose(sol.y[:, -1], y[m].numpy(), rtol=1e-6, atol=1e-8)
