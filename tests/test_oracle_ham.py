"""The reference's own GTest assertions (tests/cpp/ode_solver_test.cpp) restated against
oracle/ham.py, plus closed-form / C-oracle / golden cross-checks of the Hamiltonian adaptor."""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle, ham, problems

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fd(fun, x, p, W, wrt):
    """central FD exactly as ode_solver_test.cpp:52-66 (h = 1e-8 max(|x|,1))."""
    base = x if wrt == "x" else p
    out = torch.zeros_like(base)
    for i in range(base.shape[1]):
        h = 1e-8 * base[:, i].abs()
        h[base[:, i] < 1.0] = 1e-8
        a, b = base.clone(), base.clone()
        a[:, i] += h
        b[:, i] -= h
        if wrt == "x":
            out[:, i] = (fun(a, p, W) - fun(b, p, W)) / (2 * h)
        else:
            out[:, i] = (fun(x, a, W) - fun(x, b, W)) / (2 * h)
    return out


def test_pxH_ppH_vs_finite_differences():
    # HamiltonianTest.pxHTest / ppHTest, ode_solver_test.cpp:46-94
    torch.manual_seed(1)
    x = torch.rand(2, 2, dtype=torch.float64)
    p = torch.rand(2, 2, dtype=torch.float64)
    assert torch.allclose(ham.pxH(x, p, 1.0, ham.H_test_cubic), _fd(ham.H_test_cubic, x, p, 1.0, "x"), rtol=1e-5, atol=1e-8)
    assert torch.allclose(ham.ppH(x, p, 1.0, ham.H_test_cubic), _fd(ham.H_test_cubic, x, p, 1.0, "p"), rtol=1e-5, atol=1e-8)


def _vdpdyns_closed(y, W):
    # ode_solver_test.cpp:204-223 (closed-form dynamics of H_test_sq with the + sign)
    p1, p2, x1, x2 = y.T
    return torch.stack([-2 * p2 * p2 * ((1 - x1 * x1) * x2 - x1) * (-2 * x1 * x2 - 1) / W,
                        p1 - 2 * p2 * p2 * ((1 - x1 * x1) * x2 - x1) * (1 - x1 * x1) / W,
                        x2,
                        -2 * p2 * (x2 * (1 - x1 * x1) - x1).pow(2) / W], dim=1)


def test_dyns_explicit_vs_implicit_pins_plus_sign():
    # HamiltonianTest.DynsExplVsImplTest, ode_solver_test.cpp:372-433 (RK4 warm-up shortened)
    W = 0.1
    y = torch.tensor([[-1.0, -1.0, 2.0, 0.0]], dtype=torch.float64)
    h = 1e-6
    for _ in range(50):
        k1 = _vdpdyns_closed(y, W)
        k2 = _vdpdyns_closed(y + h / 2 * k1, W)
        k3 = _vdpdyns_closed(y + h / 2 * k2, W)
        k4 = _vdpdyns_closed(y + h * k3, W)
        y = y + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    d = ham.eval_dyns(y, W, ham.H_test_sq, costate_sign=+1)
    assert torch.allclose(d, _vdpdyns_closed(y, W), rtol=1e-12, atol=1e-12)
    # canonical sign flips the pdot rows only
    dm = ham.eval_dyns(y, W, ham.H_test_sq, costate_sign=-1)
    assert torch.allclose(dm[:, :2], -d[:, :2]) and torch.allclose(dm[:, 2:], d[:, 2:])


def test_Hu_dynamics_and_jacobian_vs_reference_fixtures():
    # HamiltonianUTest.DynsExplVsImplTestU / JacExplVsImplTest, ode_solver_test.cpp:456-573
    torch.manual_seed(2)
    M, W = 7, 0.1
    y = torch.rand(M, 4, dtype=torch.float64) * 2 - 1
    u = torch.rand(M, dtype=torch.float64)
    p1, p2, x1, x2 = y.T
    dyn = torch.stack([p2 * u * (-2 * x1 * x2 - 1), p1 + p2 * u * (1 - x1 * x1), x2, u * ((1 - x1 * x1) * x2 - x1)], 1)
    J = torch.zeros(M, 4, 4, dtype=torch.float64)
    J[:, 0, 1] = u * (-2 * x1 * x2 - 1)
    J[:, 0, 2] = p2 * u * (-2 * x2)
    J[:, 0, 3] = p2 * u * (-2 * x1)
    J[:, 1, 0] = 1
    J[:, 1, 1] = u * (1 - x1 * x1)
    J[:, 1, 2] = p2 * u * (-2 * x1)
    J[:, 2, 3] = 1
    J[:, 3, 2] = u * (-2 * x1 * x2 - 1)
    J[:, 3, 3] = u * (1 - x1 * x1)
    assert torch.allclose(ham.eval_dyns(y, W, ham.Hu_vdp_pmp, u), dyn, rtol=1e-13, atol=1e-14)
    assert torch.allclose(ham.eval_jac(y, W, ham.Hu_vdp_pmp, u), J, rtol=1e-13, atol=1e-14)


CASES = [("vdp", [1000.0], lambda: ham.H_vdp(1000.0), None, None),
         ("vdp_pmp_smooth", [1.0], lambda: ham.Hu_vdp_pmp, ham.control_smooth, 1.0),
         ("vdp_pmp_bang", [1.0, 1.0, 100.0], lambda: ham.Hu_vdp_pmp, ham.control_bangbang(1.0, 100.0), 1.0)]


@pytest.mark.parametrize("pb,par,Hf,ctrl,W", CASES)
@pytest.mark.parametrize("sign", [1, -1])
@pytest.mark.parametrize("arith", ["reference", "mirror"])
def test_c_oracle_derivatives_vs_autograd(pb, par, Hf, ctrl, W, sign, arith):
    torch.manual_seed(3)
    y = torch.rand(64, 4, dtype=torch.float64) * 4 - 2
    u = None if ctrl is None else ctrl(y[:, 2:], y[:, :2], W)
    d = ham.eval_dyns(y, W, Hf(), u, costate_sign=sign).numpy()
    J = ham.eval_jac(y, W, Hf(), u, costate_sign=sign).numpy()
    dc = c_oracle.eval_dyns(pb, par, sign, y.numpy(), arith)
    Jc = c_oracle.eval_jac(pb, par, sign, y.numpy(), arith)
    assert np.allclose(dc, d, rtol=1e-13, atol=1e-12)
    assert np.allclose(Jc, J, rtol=1e-13, atol=1e-12)


def test_harmonic_adaptor_and_q1_gate():
    torch.manual_seed(4)
    y = torch.rand(16, 2, dtype=torch.float64)
    J = ham.eval_jac(y, None, ham.H_harmonic(1.3), costate_sign=-1)
    ref = torch.tensor([[0.0, -1.69], [1.0, 0.0]], dtype=torch.float64).expand(16, 2, 2)
    assert torch.allclose(J, ref, atol=1e-14)
    # written with pow(2) the gated (quirk Q1) and the correct Hessians agree
    Jg = ham.eval_jac(y, None, ham.H_harmonic(1.3), costate_sign=-1, gate_constant_rows=True)
    assert torch.equal(J, Jg)
    # ... whereas the p*p form loses its constant rows under the reference's gate
    Hpp = lambda x, p, W: (p[:, 0] * p[:, 0] + x[:, 0] * x[:, 0]) / 2
    Jq = ham.eval_jac(y, None, Hpp, costate_sign=-1, gate_constant_rows=True)
    assert Jq.abs().max() == 0
    for arith in ("reference", "mirror"):
        assert np.allclose(c_oracle.eval_jac("harmonic", [1.3], -1, y.numpy(), arith), ref.numpy(), atol=1e-14)
        f, _ = problems.harmonic(1.3, -1)
        d = c_oracle.eval_dyns("harmonic", [1.3], -1, y.numpy(), arith)
        assert np.allclose(d, np.array([f(0, r) for r in y.numpy()]), rtol=1e-15)


def test_golden_autograd_vectors():
    g = np.load(os.path.join(GOLD, "ham_autograd.npz"))
    y = g["y"]
    for arith in ("reference", "mirror"):
        assert np.allclose(c_oracle.eval_dyns("vdp", [1000.0], 1, y, arith), g["vdp_dyn"], rtol=1e-13, atol=1e-12)
        assert np.allclose(c_oracle.eval_jac("vdp", [1000.0], 1, y, arith), g["vdp_jac"], rtol=1e-13, atol=1e-12)
        assert np.allclose(c_oracle.eval_dyns("vdp_pmp_smooth", [1.0], 1, y, arith), g["pmp_smooth_dyn"], rtol=1e-13, atol=1e-12)
        assert np.allclose(c_oracle.eval_jac("vdp_pmp_smooth", [1.0], 1, y, arith), g["pmp_smooth_jac"], rtol=1e-13, atol=1e-12)
        assert np.allclose(c_oracle.eval_dyns("vdp_pmp_bang", [1.0, 1.0, 100.0], 1, y, arith), g["pmp_bang_dyn"], rtol=1e-13, atol=1e-12)
        assert np.allclose(c_oracle.eval_jac("vdp_pmp_bang", [1.0, 1.0, 100.0], 1, y, arith), g["pmp_bang_jac"], rtol=1e-13, atol=1e-12)
        assert np.allclose(c_oracle.eval_dyns("harmonic", [1.0], -1, g["y2"], arith), g["harm_dyn"], rtol=1e-14)
        assert np.allclose(c_oracle.eval_jac("harmonic", [1.0], -1, g["y2"], arith), g["harm_jac"], rtol=1e-14)


def test_mlp_hamiltonian_c_vs_autograd():
    mlp = ham.MLPHamiltonian(N=2, width=16, seed=0)
    par = [2.0, 16.0]
    for W, b in zip(mlp.Ws, mlp.bs):
        par += W.numpy().ravel().tolist() + b.numpy().ravel().tolist()
    torch.manual_seed(5)
    y = torch.randn(8, 4, dtype=torch.float64) * 0.5
    d = ham.eval_dyns(y, None, mlp, costate_sign=-1).numpy()
    J = ham.eval_jac(y, None, mlp, costate_sign=-1).numpy()
    assert np.allclose(c_oracle.eval_dyns("mlp", par, -1, y.numpy()), d, rtol=1e-11, atol=1e-13)
    assert np.allclose(c_oracle.eval_jac("mlp", par, -1, y.numpy()), J, rtol=1e-10, atol=1e-12)
