"""Closed-form dynamics / Jacobians of the benchmark Hamiltonians (numpy, scalar trajectories).

TEST INFRASTRUCTURE ONLY.  Hand-derived derivatives -- an independent check of both the
autograd restatement (oracle/ham.py) and the CUDA dual-number kernels.

State layout y = [p (N), x (N)], ydot = [sigma * dH/dx, dH/dp]
(/root/reference/src/cpp/janus_ode_common.hpp:2167-2184; sigma=+1 reference, -1 canonical).
Closed forms for the controlled vdP system are the reference's own fixtures:
tests/cpp/ode_solver_test.cpp:228-248 (vdpHudyns), :299-334 (vdpHujac).
"""
from __future__ import annotations

import numpy as np


def harmonic(omega=1.0, sign=-1):
    """H = (p^2 + omega^2 x^2)/2."""
    w2 = omega * omega

    def f(t, y):
        p, x = y
        return np.array([sign * w2 * x, p])

    def jac(t, y):
        return np.array([[0.0, sign * w2], [1.0, 0.0]])
    return f, jac


def vdp(mu, sign=+1):
    """H = p1 x2 + p2 (mu (1 - x1^2) x2 - x1)."""
    def f(t, y):
        p1, p2, x1, x2 = y
        return np.array([sign * (p2 * (-2.0 * mu * x1 * x2 - 1.0)),
                         sign * (p1 + p2 * (mu * (1.0 - x1 * x1))),
                         x2,
                         mu * (1.0 - x1 * x1) * x2 - x1])

    def jac(t, y):
        p1, p2, x1, x2 = y
        J = np.zeros((4, 4))
        J[0, 1] = sign * (-2.0 * mu * x1 * x2 - 1.0)
        J[0, 2] = sign * (p2 * (-2.0 * mu * x2))
        J[0, 3] = sign * (p2 * (-2.0 * mu * x1))
        J[1, 0] = sign * 1.0
        J[1, 1] = sign * (mu * (1.0 - x1 * x1))
        J[1, 2] = sign * (p2 * mu * (-2.0 * x1))
        J[2, 3] = 1.0
        J[3, 2] = -2.0 * mu * x1 * x2 - 1.0
        J[3, 3] = mu * (1.0 - x1 * x1)
        return J
    return f, jac


def ctrl_smooth(W):
    def u(y):
        p1, p2, x1, x2 = y
        return -p2 * ((1.0 - x1 * x1) * x2 - x1) / W
    return u


def ctrl_bangbang(umin=1.0, umax=100.0):
    def u(y):
        p1, p2, x1, x2 = y
        return umax if p2 * (1.0 - x1 * x1) * x2 < 0 else umin
    return u


def vdp_pmp(ctrl, W=1.0, sign=+1):
    """Hu = p1 x2 + p2 u ((1-x1^2) x2 - x1) + W/2 u^2 + 1, u = ctrl(y) held constant under d/dy."""
    def f(t, y):
        p1, p2, x1, x2 = y
        u = ctrl(y)
        return np.array([sign * (p2 * u * (-2.0 * x1 * x2 - 1.0)),
                         sign * (p1 + p2 * u * (1.0 - x1 * x1)),
                         x2,
                         u * ((1.0 - x1 * x1) * x2 - x1)])

    def jac(t, y):
        p1, p2, x1, x2 = y
        u = ctrl(y)
        J = np.zeros((4, 4))
        J[0, 1] = sign * (u * (-2.0 * x1 * x2 - 1.0))
        J[0, 2] = sign * (p2 * u * (-2.0 * x2))
        J[0, 3] = sign * (p2 * u * (-2.0 * x1))
        J[1, 0] = sign * 1.0
        J[1, 1] = sign * (u * (1.0 - x1 * x1))
        J[1, 2] = sign * (p2 * u * (-2.0 * x1))
        J[2, 3] = 1.0
        J[3, 2] = u * (-2.0 * x1 * x2 - 1.0)
        J[3, 3] = u * (1.0 - x1 * x1)
        return J
    return f, jac


def from_hamiltonian(H, W=None, control=None, sign=+1):
    """Adapt an oracle/ham.py torch Hamiltonian into scalar f/jac callables (autograd path)."""
    import torch

    from . import ham

    def _u(yt):
        if control is None:
            return None
        N = yt.shape[1] // 2
        return control(yt[:, N:], yt[:, :N], W).detach()

    def f(t, y):
        yt = torch.as_tensor(np.asarray(y, dtype=np.float64)).reshape(1, -1)
        return ham.eval_dyns(yt, W, H, _u(yt), costate_sign=sign)[0].numpy()

    def jac(t, y):
        yt = torch.as_tensor(np.asarray(y, dtype=np.float64)).reshape(1, -1)
        return ham.eval_jac(yt, W, H, _u(yt), costate_sign=sign)[0].numpy()
    return f, jac
