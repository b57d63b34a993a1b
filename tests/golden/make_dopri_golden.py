#!/usr/bin/env python3
"""Golden vectors of the explicit DOPRI853 path from the LIVE reference: /root/reference/src/python/dopri853te.py is
imported unmodified and stepped on seeded batches; every accepted step (x, y, hdid, hnext) and the number of
derivative evaluations go to tests/golden/dopri853_ref.npz.  Run from the repo root in the build container:
    python tests/golden/make_dopri_golden.py
Cases (closed-form dynamics = what a user's `derivs` would be for these Hamiltonians):
  harmonic  H = (p^2 + x^2)/2, canonical sign:  pdot = -x, xdot = p            y = [p, x]
  vdp       H = p1 x2 + p2 (mu (1 - x1^2) x2 - x1), mu = 2, reference sign       y = [p1, p2, x1, x2]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference/src/python")
from dopri853te import Dopri853Te as RefDopri          # noqa: E402  (reference code)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dopri853_ref.npz")


def harmonic(t, y):
    return torch.stack([-y[:, 1], y[:, 0]], 1)


def vdp(t, y, mu=2.0):
    p1, p2, x1, x2 = y[:, 0], y[:, 1], y[:, 2], y[:, 3]
    return torch.stack([p2 * (-2.0 * mu * x1 * x2 - 1.0), p1 + p2 * (mu * (1.0 - x1 * x1)), x2,
                        mu * (1.0 - x1 * x1) * x2 - x1], 1)


def run(name, f, y0, h0, atol, rtol, nsteps, out):
    calls = [0]

    def derivs(t, y):
        calls[0] += y.shape[0]
        return f(t, y)
    M = y0.shape[0]
    x0 = torch.zeros(M, dtype=torch.float64)
    s = RefDopri(y0.clone(), f(x0, y0), x0.clone(), atol, rtol, False, derivs)
    h = torch.full((M,), h0, dtype=torch.float64)
    xs, ys, hd, hn, nc = [], [], [], [], []
    for _ in range(nsteps):
        calls[0] = 0
        s.step(h.clone())
        xs.append(s.x.numpy().copy()); ys.append(s.y.numpy().copy()); hd.append(s.hdid.numpy().copy())
        hn.append(s.hnext.numpy().copy()); nc.append(calls[0])
        h = s.hnext.clone()
    out[name + "_y0"], out[name + "_h0"] = y0.numpy(), h0
    out[name + "_tol"] = np.array([atol, rtol])
    out[name + "_x"], out[name + "_y"] = np.array(xs), np.array(ys)
    out[name + "_hdid"], out[name + "_hnext"], out[name + "_calls"] = np.array(hd), np.array(hn), np.array(nc)
    print(name, "steps", nsteps, "x range", float(s.x.min()), float(s.x.max()), "derivative rows", int(sum(nc)))


def main():
    g = torch.Generator().manual_seed(0)
    out = {}
    run("harmonic", harmonic, torch.rand(8, 2, generator=g, dtype=torch.float64) * 2 - 1, 0.5, 1e-9, 1e-9, 40, out)
    y0 = torch.tensor([1.0, 1.0, 2.0, 0.0], dtype=torch.float64).repeat(6, 1)
    y0[:, 2] += torch.rand(6, generator=g, dtype=torch.float64)
    run("vdp", vdp, y0, 0.5, 1e-8, 1e-8, 40, out)
    run("vdp_loose", vdp, y0, 2.0, 1e-3, 1e-3, 25, out)          # large first step: several rejections per call
    np.savez(OUT, **out)
    print("written", OUT)


if __name__ == "__main__":
    main()
