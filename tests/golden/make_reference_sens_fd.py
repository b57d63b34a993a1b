"""tests/golden/ref_live_sens_fd.npz: d y(tf) / d y0 of the LIVE reference solver by central finite differences -- the
check the reference applies to its own RadauTeD (tests/cpp/ode_solver_test.cpp:620-770: sensitivities vs central differences
of the solver, rtol 1e-3 ... 1e-4).  RadauTeD itself needs the whole un-vendored janus-tensor-dual library; the unmodified
RadauTe (oracle/_ref) run on perturbed initial states gives the same numbers to O(eps^2).
    python tests/golden/make_reference_sens_fd.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_live  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
Y0 = np.array([[1.0, 1.0, 2.0, 0.0], [1.0, 1.0, 2.3, 0.0], [0.5, -1.0, 1.7, 0.2]])
out = {"y0": Y0}
for mu, tf in ((1.0, 6.0), (30.0, 20.0)):
    rows, base = [], []
    for y0 in Y0:
        for d in range(4):
            hh = 1e-6 * max(abs(y0[d]), 1.0)
            for sgn in (+1, -1):
                yp = y0.copy()
                yp[d] += sgn * hh
                rows.append(yp)
        base.append(y0)
    allrows = np.array(rows + base)
    ts = np.tile([0.0, tf], (len(allrows), 1))
    r = ref_live.solve("vdp", mu, 1, allrows, ts, 1e-11, 1e-13, stages=(3, 3, 7), max_steps=1000000, procs=6)
    assert (r["res"] == 0).all()
    y = r["y"]
    S = np.zeros((3, 4, 4))
    for m in range(3):
        for d in range(4):
            hh = 1e-6 * max(abs(Y0[m, d]), 1.0)
            S[m, :, d] = (y[m * 8 + 2 * d] - y[m * 8 + 2 * d + 1]) / (2 * hh)
    out[f"S_mu{int(mu)}"] = S
    out[f"y_mu{int(mu)}"] = y[24:]
    out[f"par_mu{int(mu)}"] = np.array([mu, tf])
    print(mu, tf, np.abs(S).max())
np.savez_compressed(os.path.join(OUT, "ref_live_sens_fd.npz"), **out)
