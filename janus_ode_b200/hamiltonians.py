"""Hamiltonian descriptors for the built-in device Hamiltonians (csrc/hamiltonians.cuh).

The reference takes any differentiable callable H(x, p, W) (janus_ode_common.hpp:2162-2164); a
CUDA kernel cannot call Python, so the accelerated path takes a descriptor of a Hamiltonian
compiled into the library.  State layout y = [p, x] (janus_ode_common.hpp:2167-2169).
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Hamiltonian:
    name: str
    ham_id: int
    N: int
    params: tuple
    lib_path: str | None = None     # plugin library of a user Hamiltonian (janus_ode_b200/user.py)

    @property
    def Ny(self):
        return 2 * self.N


def Harmonic(omega=1.0):
    """H = (p^2 + omega^2 x^2)/2."""
    return Hamiltonian("harmonic", 0, 1, (float(omega),))


def VanDerPol(mu=1000.0):
    """H = p1 x2 + p2 (mu (1 - x1^2) x2 - x1)  (vdP in Hamiltonian / costate form)."""
    return Hamiltonian("vdp", 1, 2, (float(mu),))


def VdpPmp(W=1.0, control="smooth", umin=1.0, umax=100.0):
    """Hu = p1 x2 + p2 u ((1-x1^2) x2 - x1) + W/2 u^2 + 1 with a supplied control law held constant
    under differentiation (tests/cpp/ode_solver_test.cpp:185-197, evalDynsU semantics).
    control: 'smooth'  u = -p2 ((1-x1^2) x2 - x1)/W          (ode_solver_test.cpp:17-24)
             'bangbang' u = umax if p2 (1-x1^2) x2 < 0 else umin (radauted_vdp_aug_example.cpp:30-33)"""
    if control == "smooth":
        return Hamiltonian("vdp_pmp_smooth", 2, 2, (float(W), float(umin), float(umax)))
    if control == "bangbang":
        return Hamiltonian("vdp_pmp_bang", 3, 2, (float(W), float(umin), float(umax)))
    raise ValueError("control must be 'smooth' or 'bangbang'")


def VdpAug(umin=1.0, umax=100.0):
    """H = p1 x2 + p2 (u (1-x1^2) x2 - x1), bang-bang u: the augmented vdP system of
    examples/cpp/radauted_vdp_aug_example.cpp:23-46 == examples/python/aug_vdp_ref_test.py."""
    return Hamiltonian("vdp_aug_bang", 4, 2, (float(umin), float(umax)))


def MlpHamiltonian(N=8, width=128, seed=0, bias_scale=0.0):
    """C4: H_theta([x;p]) = 4 linear layers 2N -> width -> width -> width -> 1, tanh, fp64,
    weights N(0, 1/fan_in) from torch.Generator(seed); biases 0 (bias_scale > 0: N(0, bias_scale^2), for
    tests).  The device kernels (csrc/mlp_kernels.cuh) are built for N = 8, width = 128.
    params = [N, width, W1, b1, W2, b2, W3, b3, W4, b4] (row-major, W_l is [out, in])."""
    import torch
    g = torch.Generator().manual_seed(seed)
    dims = [2 * N, width, width, width, 1]
    par = [float(N), float(width)]
    Ws, bs = [], []
    for i in range(4):
        fan_in = dims[i]
        Ws.append(torch.randn(dims[i + 1], fan_in, generator=g, dtype=torch.float64) / fan_in ** 0.5)
    for i in range(4):
        bs.append(bias_scale * torch.randn(dims[i + 1], generator=g, dtype=torch.float64) if bias_scale else
                  torch.zeros(dims[i + 1], dtype=torch.float64))
    for W, b in zip(Ws, bs):
        par += W.reshape(-1).tolist()
        par += b.reshape(-1).tolist()
    return Hamiltonian("mlp", 5, N, tuple(par))
