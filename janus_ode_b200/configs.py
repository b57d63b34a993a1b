"""Synthetic workloads of BASELINE.json / SURVEY.md 8d (C1..C5): initial states, time spans,
options.  Pure host code shared by tests and bench.py."""
from __future__ import annotations

import math

import torch

from .hamiltonians import Harmonic, MlpHamiltonian, VanDerPol, VdpPmp
from .options import OptionsTe


def vdp_tfinal(mu):
    return (3.0 - 2.0 * math.log(2.0)) * mu      # examples/cpp/radaute_example.cpp:83


def c1_harmonic(M=1024, seed=0, fixed_order=False):
    """C1: H=(p^2+x^2)/2, canonical sign, (p0,x0)~U(-1,1)^2, t in [0,10], rtol 1e-9, atol 1e-12."""
    g = torch.Generator().manual_seed(seed)
    y0 = torch.rand(M, 2, generator=g, dtype=torch.float64) * 2 - 1
    tspan = torch.tensor([0.0, 10.0], dtype=torch.float64).repeat(M, 1)
    op = OptionsTe(RelTol=1e-9, AbsTol=1e-12, MaxNbrStep=1000000)
    if fixed_order:
        op.MinNbrStg = op.MaxNbrStg = 3
    return dict(ham=Harmonic(1.0), sign=-1, y0=y0, tspan=tspan, options=op)


def c2_vdp(M=1 << 20, mu=1000.0, seed=0, fixed_order=False, lo=0, hi=None):
    """C2/C5: stiff vdP in Hamiltonian form, reference sign, x1_0 = 2 + U(0,1), x2_0 = 0, p0 = (1,1),
    t in [0,(3-2ln2)mu], rtol 1e-6, atol 1e-9.  Rows lo:hi of the seeded global batch."""
    g = torch.Generator().manual_seed(seed)
    u = torch.rand(M, generator=g, dtype=torch.float64)
    hi = M if hi is None else hi
    n = hi - lo
    y0 = torch.zeros(n, 4, dtype=torch.float64)
    y0[:, 0] = 1.0
    y0[:, 1] = 1.0
    y0[:, 2] = 2.0 + u[lo:hi]
    tspan = torch.tensor([0.0, vdp_tfinal(mu)], dtype=torch.float64).repeat(n, 1)
    op = OptionsTe(RelTol=1e-6, AbsTol=1e-9, MaxNbrStep=1000000)
    if fixed_order:
        op.MinNbrStg = op.MaxNbrStg = 3
    return dict(ham=VanDerPol(mu), sign=+1, y0=y0, tspan=tspan, options=op)


def c2_vdp_cloud(M=1 << 20, mu=1000.0, seed=0, fixed_order=False, lo=0, hi=None):
    """Config 2 with a genuinely 4-dimensional cloud of initial states (the stock C2 batch is a 1-parameter family in
    x1(0), which is what makes an ordering by H(y0) ideal): p(0) = (1,1) + 0.3 N(0,1), x1(0) = 2 + U(0,1),
    x2(0) = 0.2 N(0,1).  Same Hamiltonian, horizon and tolerances."""
    g = torch.Generator().manual_seed(seed + 7)
    hi = M if hi is None else hi
    u = torch.rand(M, generator=g, dtype=torch.float64)
    n = torch.randn(M, 3, generator=g, dtype=torch.float64)
    y0 = torch.stack([1.0 + 0.3 * n[:, 0], 1.0 + 0.3 * n[:, 1], 2.0 + u, 0.2 * n[:, 2]], 1)[lo:hi].contiguous()
    tspan = torch.tensor([0.0, vdp_tfinal(mu)], dtype=torch.float64).repeat(hi - lo, 1)
    op = OptionsTe(RelTol=1e-6, AbsTol=1e-9, MaxNbrStep=1000000)
    if fixed_order:
        op.MinNbrStg = op.MaxNbrStg = 3
    return dict(ham=VanDerPol(mu), sign=+1, y0=y0, tspan=tspan, options=op)


def c4_mlp(M=1 << 20, seed=0, fixed_order=False, lo=0, hi=None):
    """C4: MLP Hamiltonian 16 -> 128 -> 128 -> 128 -> 1 (tanh), N = 8, canonical sign, y0 ~ 0.5 N(0,1),
    t in [0,1], rtol 1e-6, atol 1e-9.  Rows lo:hi of the seeded global batch."""
    g = torch.Generator().manual_seed(seed + 1)
    y0 = 0.5 * torch.randn(M, 16, generator=g, dtype=torch.float64)
    hi = M if hi is None else hi
    y0 = y0[lo:hi].contiguous()
    tspan = torch.tensor([0.0, 1.0], dtype=torch.float64).repeat(hi - lo, 1)
    op = OptionsTe(RelTol=1e-6, AbsTol=1e-9, MaxNbrStep=1000000)
    if fixed_order:
        op.MinNbrStg = op.MaxNbrStg = 3
    return dict(ham=MlpHamiltonian(8, 128, seed), sign=-1, y0=y0, tspan=tspan, options=op)


def c3_pmp(M=1 << 18, control="bangbang", seed=0, fixed_order=False, tfinal=None, lo=0, hi=None):
    """C3: PMP Hamiltonian with supplied control, y0=[-5,-5,0,1.5]+0.1 N(0,1), reference sign.
    'bangbang': t in [0,10] (examples/cpp/radauted_vdp_aug_example.cpp:122-144).
    'smooth'  : the feedback law u=-p2(...)/W blows up in finite time (t~0.07) from these initial
                states, so the horizon is 0.05; [0,10] is kept as a failure-code (res=1) case."""
    g = torch.Generator().manual_seed(seed)
    y0 = torch.tensor([-5.0, -5.0, 0.0, 1.5], dtype=torch.float64).repeat(M, 1)
    y0 = y0 + 0.1 * torch.randn(M, 4, generator=g, dtype=torch.float64)
    if tfinal is None:
        tfinal = 10.0 if control == "bangbang" else 0.05
    hi = M if hi is None else hi
    y0 = y0[lo:hi].contiguous()
    tspan = torch.tensor([0.0, tfinal], dtype=torch.float64).repeat(hi - lo, 1)
    op = OptionsTe(RelTol=1e-6, AbsTol=1e-9, MaxNbrStep=1000000)
    if fixed_order:
        op.MinNbrStg = op.MaxNbrStg = 3
    return dict(ham=VdpPmp(1.0, control), sign=+1, y0=y0, tspan=tspan, options=op)


def shard_rows(M, rank, world):
    """interleaved shard i mod G: indices of the global batch owned by `rank`."""
    return torch.arange(rank, M, world)
