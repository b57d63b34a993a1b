"""evalDynsDual / evalJacDual and the third-order ladders: one third-order jet of H
per trajectory (csrc/dual3.cuh) against
  * the LIVE reference: janus::evalDynsUDual / evalJacDualU of janus_ode_common.hpp run through oracle/_ref
    (tests/golden/ref_live_jac_dual.npz) on the PMP Hamiltonian with both controls.  For the uncontrolled vdP and the
    harmonic Hamiltonian the reference's own ladders throw ("tensor does not have a device"): its safe_jac returns an
    undefined tensor for a structurally-zero block (janus_ode_common.hpp:116-134 under libtorch 2.x);
  * the restatement oracle/ham.py: eval_jac_dual (same autograd engine, correct zeros) on every built-in Hamiltonian."""
import os

import numpy as np
import pytest
import torch

from oracle import ham

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_live_jac_dual.npz")


def _oracle(pb, par, y, yd, sign=+1):
    yt, ydt = torch.from_numpy(y), torch.from_numpy(yd)
    N = y.shape[1] // 2
    if pb == "vdp":
        return ham.eval_jac_dual(yt, ydt, None, ham.H_vdp(par), None, sign)
    if pb == "harmonic":
        return ham.eval_jac_dual(yt, ydt, None, ham.H_harmonic(par), None, sign)
    ctrl = ham.control_smooth if pb == "vdp_pmp_smooth" else ham.control_bangbang(1.0, 100.0)
    u = ctrl(yt[:, N:], yt[:, :N], par).detach()
    return ham.eval_jac_dual(yt, ydt, par, ham.Hu_vdp_pmp, u, sign)


def _ladders(pb, par, y, yd, gate):
    yt, ydt = torch.from_numpy(y), torch.from_numpy(yd)
    ctrl = ham.control_smooth if pb == "vdp_pmp_smooth" else ham.control_bangbang(1.0, 100.0)
    u = ctrl(yt[:, 2:], yt[:, :2], par).detach()
    return ham.eval_jac_dual_ladders(yt, ydt, par, ham.Hu_vdp_pmp, u, +1, gate=gate)


@pytest.mark.parametrize("pb", ["vdp_pmp_smooth", "vdp_pmp_bang"])
def test_restatement_equals_the_live_reference(pb):
    """f, fd, J equal the live reference to roundoff.  Jd equals it once the reference's gate is reproduced: its
    third-order ladders store a row only if the autograd result requires grad (quirk Q1), which drops the constant block
    d^3H/dx^3 = -2 p2 u of this Hamiltonian; the ungated restatement (= the CUDA jets) is the mathematically complete
    tensor and differs from the reference exactly there."""
    g = np.load(GOLD)
    y, yd, par = g[pb + "_y"], g[pb + "_yd"], float(g[pb + "_par"])
    f, fd, J, Jd = _ladders(pb, par, y, yd, gate=True)
    for a, k in ((f, "f"), (fd, "fd"), (J, "J"), (Jd, "Jd")):
        ref = g[pb + "_" + k]
        assert np.allclose(a.numpy(), ref, rtol=1e-12, atol=1e-12 * max(1.0, np.abs(ref).max())), k
    # ungated ladders == the jet restatement; the difference to the reference sits in the (pdot, x) block only
    Jd_full = _ladders(pb, par, y, yd, gate=False)[3]
    Jd_jet = _oracle(pb, par, y, yd)[3]
    assert np.allclose(Jd_full.numpy(), Jd_jet.numpy(), rtol=1e-12, atol=1e-11)
    diff = np.abs(Jd_full.numpy() - g[pb + "_Jd"]) > 1e-9
    assert diff[:, :2, 2:].any() and not diff[:, 2:].any() and not diff[:, :2, :2].any()


def test_third_order_tensor_is_symmetric_and_matches_finite_differences():
    rng = np.random.default_rng(0)
    y = rng.uniform(-1.5, 1.5, (8, 4))
    yd = rng.normal(size=(8, 4, 2))
    f, fd, J, Jd, T = _oracle("vdp", 7.0, y, yd)
    # d J / d y_k by central differences of the Jacobian
    e = 1e-5
    for k in range(4):
        yp, ym = y.copy(), y.copy()
        yp[:, k] += e
        ym[:, k] -= e
        Jp = _oracle("vdp", 7.0, yp, yd)[2].numpy()
        Jm = _oracle("vdp", 7.0, ym, yd)[2].numpy()
        assert np.allclose(T[:, :, :, k].numpy(), (Jp - Jm) / (2 * e), rtol=1e-6, atol=1e-6)
    # x-rows of the tensor are symmetric in (c, k) (mixed partials of one scalar H)
    assert np.allclose(T.numpy(), T.numpy().transpose(0, 1, 3, 2), atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("pb,sign", [("vdp_pmp_smooth", 1), ("vdp_pmp_bang", 1), ("vdp", 1), ("vdp", -1), ("harmonic", -1)])
def test_gpu_jets_against_the_reference_and_the_restatement(pb, sign):
    import janus_ode_b200 as jb
    g = np.load(GOLD)
    rng = np.random.default_rng(5)
    hamd = {"vdp": lambda: jb.VanDerPol(1000.0), "harmonic": lambda: jb.Harmonic(1.3), "vdp_pmp_smooth": lambda: jb.VdpPmp(1.0, "smooth"),
            "vdp_pmp_bang": lambda: jb.VdpPmp(1.0, "bangbang", 1.0, 100.0)}[pb]()
    par = {"vdp": 1000.0, "harmonic": 1.3}.get(pb, 1.0)
    if pb.startswith("vdp_pmp"):
        y, yd = g[pb + "_y"], g[pb + "_yd"]
    else:
        y, yd = rng.uniform(-2, 2, (257, hamd.Ny)), rng.normal(size=(257, hamd.Ny, 3))
    dyn = jb.HamiltonianSystem(hamd, sign)
    f, J, T, fd, Jd = dyn.jacobian_dual(None, torch.from_numpy(y).cuda(), torch.from_numpy(yd).cuda())
    of, ofd, oJ, oJd, oT = _oracle(pb, par, y, yd, sign)
    for a, b, k in ((f, of, "f"), (fd, ofd, "fd"), (J, oJ, "J"), (Jd, oJd, "Jd"), (T, oT, "T")):
        b = b.numpy()
        assert np.allclose(a.cpu().numpy(), b, rtol=1e-12, atol=1e-12 * max(1.0, np.abs(b).max())), k
    if pb.startswith("vdp_pmp"):   # ... and the live reference itself (Jd: see test_restatement_equals_the_live_reference)
        for a, k in ((f, "f"), (fd, "fd"), (J, "J")):
            ref = g[pb + "_" + k]
            assert np.allclose(a.cpu().numpy(), ref, rtol=1e-12, atol=1e-12 * max(1.0, np.abs(ref).max())), k
    # consistent with the first- / second-order kernels of the hot path, bit for bit in the values they share
    yt = torch.from_numpy(y).cuda()
    assert torch.allclose(f, dyn(None, yt), rtol=1e-14, atol=1e-12) and torch.allclose(J, dyn.jacobian(None, yt), rtol=1e-14, atol=1e-10)
    assert torch.equal(dyn.third_derivatives(yt), T)
