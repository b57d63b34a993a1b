"""CPU restatement of the reference's Hamiltonian adaptor (torch autograd, fp64).

TEST INFRASTRUCTURE ONLY (see oracle/radau_scalar.py header).  Uses the same autograd
engine the reference uses, so it is bit-comparable to what the reference would compute.

Follows /root/reference/src/cpp/janus_ode_common.hpp:
  safe_jac                      :116-134
  pxH / ppH (+ Hu twins)        :177-279
  ppppH, pxpxH, pxppH, pppxH    :282-718      (+ Hu twins)
  evalDyns / evalDynsU          :2161-2217    y = [p, x],  ydot = [dH/dx ; dH/dp]  (+ sign!)
  evalJac  / evalJacU           :2220-2299    J = [[pppxH, pxpxH], [ppppH, pxppH]]

Switches:
  costate_sign  +1 = reference (pdot = +dH/dx, pinned by tests/cpp/ode_solver_test.cpp:416-419)
                -1 = canonical Hamilton equations (pdot = -dH/dx), as the north star words it
  gate_constant_rows  False = mathematically correct Hessians (default)
                      True  = emulate quirk Q1: a Hessian row is stored only if its autograd
                              result `requires_grad()` (janus_ode_common.hpp:326,454,577,706)
H callables take (x[M,N], p[M,N], W) -> [M];  Hu callables take (x, p, u, W) -> [M].
"""
from __future__ import annotations

import torch


def _safe_jac(y, x):
    """janus_ode_common.hpp:116-134: d(sum_m y_m)/dx, zeros if y does not depend on x."""
    if not y.requires_grad:
        return torch.zeros_like(x)
    g = torch.autograd.grad([y], [x], [torch.ones_like(y)], retain_graph=True,
                            create_graph=True, allow_unused=True)[0]
    if g is None:
        return torch.zeros_like(x)
    return g


def _call(H, x, p, u, W):
    return H(x, p, W) if u is None else H(x, p, u, W)


def pxH(x, p, W, H, u=None):
    """dH/dx  [M,N]   (janus_ode_common.hpp:177-229)"""
    xt = x.clone().detach().requires_grad_(True)
    pt = p.clone().detach()
    return _safe_jac(_call(H, xt, pt, u, W), xt).detach()


def ppH(x, p, W, H, u=None):
    """dH/dp  [M,N]   (janus_ode_common.hpp:231-279)"""
    xt = x.clone().detach()
    pt = p.clone().detach().requires_grad_(True)
    return _safe_jac(_call(H, xt, pt, u, W), pt).detach()


def _hess_block(x, p, W, H, u, first, second, gate):
    """rows i = d/d(second_j) of (dH/d(first_i)).  first/second in {'x','p'}."""
    xt = x.clone().detach().requires_grad_(True)
    pt = p.clone().detach().requires_grad_(True)
    Hv = _call(H, xt, pt, u, W)
    a = xt if first == "x" else pt
    b = xt if second == "x" else pt
    g1 = _safe_jac(Hv, a)
    M, N = x.shape
    hess = torch.zeros(M, N, N, dtype=x.dtype)
    for i in range(N):
        gi = g1[:, i]
        row = _safe_jac(gi, b)
        if gate and not row.requires_grad:
            continue                                  # quirk Q1
        hess[:, i, :] = row.detach()
    return hess


def ppppH(x, p, W, H, u=None, gate=False):
    """[i,j] = d2H/dp_j dp_i   (:282-339)"""
    return _hess_block(x, p, W, H, u, "p", "p", gate)


def pxpxH(x, p, W, H, u=None, gate=False):
    """[i,j] = d2H/dx_j dx_i   (:411-466)"""
    return _hess_block(x, p, W, H, u, "x", "x", gate)


def pxppH(x, p, W, H, u=None, gate=False):
    """[i(p),j(x)] = d/dx_j (dH/dp_i)   (:530-590)"""
    return _hess_block(x, p, W, H, u, "p", "x", gate)


def pppxH(x, p, W, H, u=None, gate=False):
    """[i(x),j(p)] = d/dp_j (dH/dx_i)   (:660-718)"""
    return _hess_block(x, p, W, H, u, "x", "p", gate)


def eval_dyns(y, W, H, u=None, costate_sign=+1):
    """evalDyns / evalDynsU (:2161-2217).  y = [p, x] -> ydot = [sign*dH/dx, dH/dp]."""
    N = y.shape[1] // 2
    p, x = y[:, :N], y[:, N:]
    dotp = pxH(x, p, W, H, u)
    dotx = ppH(x, p, W, H, u)
    return torch.cat([costate_sign * dotp, dotx], dim=1)


def eval_jac(y, W, H, u=None, costate_sign=+1, gate_constant_rows=False):
    """evalJac / evalJacU (:2220-2299): d ydot / d y, rows [pdot; xdot], cols [p, x]."""
    N = y.shape[1] // 2
    p, x = y[:, :N], y[:, N:]
    g = gate_constant_rows
    pp_dotp = pppxH(x, p, W, H, u, g)
    px_dotp = pxpxH(x, p, W, H, u, g)
    pp_dotx = ppppH(x, p, W, H, u, g)
    px_dotx = pxppH(x, p, W, H, u, g)
    jac1 = costate_sign * torch.cat([pp_dotp, px_dotp], dim=2)
    jac2 = torch.cat([pp_dotx, px_dotx], dim=2)
    return torch.cat([jac1, jac2], dim=1)


def _third_block(x, p, W, H, u, a, b, c, gate):
    """[m,i,j,k] = d/dc_k d/db_j d/da_i H, a, b, c in {'x','p'}: the ladder `p{c}p{b}p{a}H` of janus_ode_common.hpp:787-2072.
    Like there, the pure blocks track gradients of their own class only (ppppppH :795-796, pxpxpxH :959-960) and a row is
    stored only if its autograd result requires grad (quirk Q1, e.g. :851-856) when gate=True."""
    need_x = "x" in (a, b, c)
    need_p = "p" in (a, b, c)
    xt = x.clone().detach().requires_grad_(need_x)
    pt = p.clone().detach().requires_grad_(need_p)
    Hv = _call(H, xt, pt, u, W)
    var = {"x": xt, "p": pt}
    g1 = _safe_jac(Hv, var[a])
    M, N = x.shape
    out = torch.zeros(M, N, N, N, dtype=x.dtype)
    for i in range(N):
        gi = _safe_jac(g1[:, i], var[b])
        for j in range(N):
            gij = _safe_jac(gi[:, j], var[c])
            if gate and not gij.requires_grad:
                continue
            out[:, i, j, :] = gij.detach()
    return out


def eval_jac_dual_ladders(y, yd, W, H, u=None, costate_sign=+1, gate=True):
    """evalDynsDual + evalJacDual assembled block by block from the eight third-order ladders exactly like
    janus_ode_common.hpp:2302-2416 (gate=True reproduces the reference, which DROPS third-order rows that are constant:
    e.g. d^3H/dx^3 of the PMP Hamiltonian, constant in x, comes back zero).  Returns (f, fd, J, Jd)."""
    N = y.shape[1] // 2
    p, x = y[:, :N], y[:, N:]
    pd, xd = yd[:, :N], yd[:, N:]
    f = eval_dyns(y, W, H, u, costate_sign)
    J = eval_jac(y, W, H, u, costate_sign)
    fd = torch.einsum("mrk,mkl->mrl", J, yd)

    def dual(a, b):   # J block rows d/da_i, cols d/db_j: its dual part
        return (torch.einsum("mijk,mkl->mijl", _third_block(x, p, W, H, u, a, b, "x", gate), xd) +
                torch.einsum("mijk,mkl->mijl", _third_block(x, p, W, H, u, a, b, "p", gate), pd))
    jac1 = costate_sign * torch.cat([dual("x", "p"), dual("x", "x")], dim=2)
    jac2 = torch.cat([dual("p", "p"), dual("p", "x")], dim=2)
    return f, fd, J, torch.cat([jac1, jac2], dim=1)


def eval_jac_dual(y, yd, W, H, u=None, costate_sign=+1):
    """evalDynsDual + evalJacDual (+ U twins), janus_ode_common.hpp:2075-2158, :2302-2416, restated with the same autograd
    engine: returns (f [M,Ny], fd [M,Ny,D], J [M,Ny,Ny], Jd [M,Ny,Ny,D], T [M,Ny,Ny,Ny]) with T[m,r,c,k] = d J[r][c] / d y_k
    (the eight third-order ladders :787-2072 assembled in y-order), fd = J yd, Jd = T yd.  Mathematically correct
    derivatives (structurally-zero blocks are zero; the reference's safe_jac throws there under libtorch 2.x)."""
    M, Ny = y.shape
    N = Ny // 2
    yt = y.clone().detach().requires_grad_(True)

    def dyn(yy):
        p, x = yy[:, :N], yy[:, N:]
        Hv = _call(H, x, p, u, W)
        g = torch.autograd.grad(Hv.sum(), yy, create_graph=True)[0]
        return torch.cat([costate_sign * g[:, N:], g[:, :N]], dim=1)

    f = dyn(yt)
    J = torch.zeros(M, Ny, Ny, dtype=y.dtype)
    T = torch.zeros(M, Ny, Ny, Ny, dtype=y.dtype)
    for r in range(Ny):
        if not f[:, r].requires_grad:
            continue
        Jr = torch.autograd.grad(f[:, r].sum(), yt, create_graph=True, allow_unused=True)[0]
        if Jr is None:
            continue
        J[:, r] = Jr.detach()
        for c in range(Ny):
            if not Jr[:, c].requires_grad:
                continue
            Trc = torch.autograd.grad(Jr[:, c].sum(), yt, retain_graph=True, allow_unused=True)[0]
            if Trc is not None:
                T[:, r, c] = Trc.detach()
    fd = torch.einsum("mrk,mkl->mrl", J, yd)
    Jd = torch.einsum("mrck,mkl->mrcl", T, yd)
    return f.detach(), fd, J, Jd, T


# ---------------------------------------------------------------------------------------
# Hamiltonians of the benchmark configs in torch form (for autograd) --
# written in forms for which quirk Q1 and the mathematically correct Hessian agree.
# ---------------------------------------------------------------------------------------

def H_harmonic(omega=1.0):
    """C1: H = (p^2 + omega^2 x^2)/2, N = 1 (run with canonical sign)."""
    def H(x, p, W):
        return (p[:, 0].pow(2) + (omega * omega) * x[:, 0].pow(2)) / 2
    return H


def H_vdp(mu):
    """C2/C5: H = p1 x2 + p2 (mu (1 - x1^2) x2 - x1), N = 2 (reference sign)."""
    def H(x, p, W):
        p1, p2, x1, x2 = p[:, 0], p[:, 1], x[:, 0], x[:, 1]
        return p1 * x2 + p2 * (mu * (1 - x1 * x1) * x2 - x1)
    return H


def Hu_vdp_pmp(x, p, u, W):
    """C3: tests/cpp/ode_solver_test.cpp:185-197  Hu = p1 x2 + p2 u ((1-x1^2) x2 - x1) + W/2 u^2 + 1."""
    p1, p2, x1, x2 = p[:, 0], p[:, 1], x[:, 0], x[:, 1]
    return p1 * x2 + p2 * u * ((1 - x1 * x1) * x2 - x1) + (W / 2) * u * u + 1


def control_smooth(x, p, W):
    """tests/cpp/ode_solver_test.cpp:17-24: u = -p2 ((1-x1^2) x2 - x1)/W."""
    p2, x1, x2 = p[:, 1], x[:, 0], x[:, 1]
    return -p2 * ((1 - x1 * x1) * x2 - x1) / W


def control_bangbang(umin=1.0, umax=100.0):
    """examples/cpp/radauted_vdp_aug_example.cpp:30-33: u = umax if p2 (1-x1^2) x2 < 0 else umin."""
    def ctrl(x, p, W):
        p2, x1, x2 = p[:, 1], x[:, 0], x[:, 1]
        m = p2 * (1 - x1 * x1) * x2 < 0
        return torch.where(m, torch.full_like(p2, umax), torch.full_like(p2, umin))
    return ctrl


def H_test_cubic(x, p, W):
    """tests/cpp/ode_solver_test.cpp:26-37 (control folded in and differentiated)."""
    p1, p2, x1, x2 = p[:, 0], p[:, 1], x[:, 0], x[:, 1]
    u = -p2 * ((1 - x1 * x1) * x2 - x1) / W
    return p1 * p1 * p1 * x2 + p2 * p2 * p2 * u * ((1 - x1 * x1) * x2 - x1) + W * u * u / 2 + 1


def H_test_sq(x, p, W):
    """tests/cpp/ode_solver_test.cpp:170-181: H = p1 x2 - (p2 ((1-x1^2) x2 - x1))^2 / W."""
    p1, p2, x1, x2 = p[:, 0], p[:, 1], x[:, 0], x[:, 1]
    return p1 * x2 - (p2 * ((1 - x1 * x1) * x2 - x1)).pow(2) / W


class MLPHamiltonian:
    """C4: H_theta([x;p]) = 4 linear layers 2N -> width -> width -> width -> 1, tanh, fp64.

    Weights N(0, 1/fan_in), seed 0.  Input order is [x, p].
    """

    def __init__(self, N=8, width=128, seed=0):
        g = torch.Generator().manual_seed(seed)
        dims = [2 * N, width, width, width, 1]
        self.N, self.width = N, width
        self.Ws, self.bs = [], []
        for i in range(4):
            fan_in = dims[i]
            self.Ws.append(torch.randn(dims[i + 1], fan_in, generator=g, dtype=torch.float64)
                           / fan_in ** 0.5)
            self.bs.append(torch.zeros(dims[i + 1], dtype=torch.float64))

    def __call__(self, x, p, W=None):
        a = torch.cat([x, p], dim=1)
        for i in range(3):
            a = torch.tanh(a @ self.Ws[i].T + self.bs[i])
        return (a @ self.Ws[3].T + self.bs[3])[:, 0]
