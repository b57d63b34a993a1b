"""Generates tests/golden/seulex_ref.npz by running the reference's own Python stepper live:
/root/reference/src/python/seulext.py (SeulexState, SeulexT.step) imported unmodified, with tests/golden/shims/
standing in for the un-vendored janus-linear modules it imports (LUBatch, LU).  One state object per trajectory
(M = 1), driven like /root/reference/src/python/seulex_batch_example.py:44-77 (htry = hnext, clipped to the final
time).  Dynamics = the Hamiltonian van der Pol system of configs C2 in closed form (oracle/problems.py order).

Recorded per trajectory and step: x, y, hdid, hnext, k_targ after the call -- the quantities a step-by-step
restatement must reproduce.      python tests/golden/make_seulex_golden.py
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, "/root/reference/src/python")
sys.path.insert(0, ROOT)
from seulext import SeulexState, SeulexT  # noqa: E402  (the live reference)


class HamVdp:
    """ydot = [dH/dx ; dH/dp] of H = p1 x2 + p2 (mu (1 - x1^2) x2 - x1), y = [p1, p2, x1, x2] (reference sign)."""

    def __init__(self, mu):
        self.mu = mu

    def __call__(self, t, y):
        mu = self.mu
        p1, p2, x1, x2 = y[:, 0], y[:, 1], y[:, 2], y[:, 3]
        return torch.stack([p2 * (-2.0 * mu * x1 * x2 - 1.0), p1 + p2 * (mu * (1.0 - x1 * x1)), x2.clone(),
                            mu * (1.0 - x1 * x1) * x2 - x1], 1)

    def jacobian(self, t, y):
        mu = self.mu
        p1, p2, x1, x2 = y[:, 0], y[:, 1], y[:, 2], y[:, 3]
        J = torch.zeros(y.shape[0], 4, 4, dtype=y.dtype)
        J[:, 0, 1] = -2.0 * mu * x1 * x2 - 1.0
        J[:, 0, 2] = p2 * (-2.0 * mu * x2)
        J[:, 0, 3] = p2 * (-2.0 * mu * x1)
        J[:, 1, 0] = 1.0
        J[:, 1, 1] = mu * (1.0 - x1 * x1)
        J[:, 1, 2] = p2 * mu * (-2.0 * x1)
        J[:, 2, 3] = 1.0
        J[:, 3, 2] = -2.0 * mu * x1 * x2 - 1.0
        J[:, 3, 3] = mu * (1.0 - x1 * x1)
        return J


def run(mu, y0, ft, atol, rtol, htry0, KMAXX, max_steps):
    dyn = HamVdp(mu)
    rows = []
    for m in range(y0.shape[0]):
        x = torch.zeros(1, dtype=torch.float64)
        st = SeulexState(y0[m:m + 1].clone(), x, atol, rtol, KMAXX, 0.0, ft)
        solver = SeulexT()
        htry = torch.full((1,), htry0, dtype=torch.float64)
        rec = []
        for _ in range(max_steps):
            with contextlib.redirect_stdout(io.StringIO()):
                st = solver.step(htry, dyn, st)
            rec.append(np.concatenate([[st.x.item()], st.y[0].numpy(), [st.hdid.item(), st.hnext.item(), float(st.k_targ.item()),
                                                                     float(htry.item())]]))
            if not (ft - st.x.item() > 0.0):
                break
            htry = st.hnext.clone()
            if htry.item() > ft - st.x.item():
                htry = torch.full((1,), ft - st.x.item(), dtype=torch.float64)
            if not htry.item() > float(st.EPS):
                break
        rows.append(np.array(rec))
    n = max(len(r) for r in rows)
    out = np.full((len(rows), n, rows[0].shape[1]), np.nan)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out, np.array([len(r) for r in rows])


if __name__ == "__main__":
    torch.manual_seed(0)
    cases = {}
    # stiff (mu = 1000, a slice of the C2 horizon) and mildly stiff (mu = 5, several periods); 6 trajectories each
    for name, mu, ft, atol, rtol in (("mu1000", 1000.0, 200.0, 1e-9, 1e-6), ("mu5", 5.0, 8.0, 1e-8, 1e-6)):
        y0 = torch.zeros(6, 4, dtype=torch.float64)
        y0[:, 0] = 1.0
        y0[:, 1] = 1.0
        y0[:, 2] = 2.0 + torch.rand(6, dtype=torch.float64)
        rec, n = run(mu, y0, ft, atol, rtol, 0.01, 13, 400)
        cases[name + "_y0"] = y0.numpy()
        cases[name + "_rec"] = rec
        cases[name + "_n"] = n
        cases[name + "_par"] = np.array([mu, ft, atol, rtol, 0.01, 13])
        print(name, "steps per trajectory", n, "final x", rec[np.arange(6), n - 1, 0])
    np.savez_compressed(os.path.join(HERE, "seulex_ref.npz"), **cases)
