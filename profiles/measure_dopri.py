#!/usr/bin/env python3
"""Throughput of the explicit DOPRI853 path (jb.Dopri853Te.step) on one GPU next to the numpy oracle on the host:
    python profiles/measure_dopri.py [M] [nsteps]      -> one JSON line (trajectory-steps/s, both arms)
Workload: van der Pol in Hamiltonian form, mu = 2 (non-stiff), reference sign, x1_0 = 2 + U(0,1), atol = rtol = 1e-6,
first htry = 0.1, then htry = hnext (the reference's driver loop)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import janus_ode_b200 as jb          # noqa: E402
from oracle.dopri853 import Dopri853Oracle   # noqa: E402


def f_vdp(t, y, mu=2.0):
    p1, p2, x1, x2 = y[:, 0], y[:, 1], y[:, 2], y[:, 3]
    return np.stack([p2 * (-2.0 * mu * x1 * x2 - 1.0), p1 + p2 * (mu * (1.0 - x1 * x1)), x2,
                     mu * (1.0 - x1 * x1) * x2 - x1], 1)


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    g = torch.Generator().manual_seed(0)
    y0 = torch.zeros(M, 4, dtype=torch.float64)
    y0[:, 0] = y0[:, 1] = 1.0
    y0[:, 2] = 2.0 + torch.rand(M, generator=g, dtype=torch.float64)
    dyn = jb.HamiltonianSystem(jb.VanDerPol(2.0), +1)
    x0 = torch.zeros(M, dtype=torch.float64, device="cuda")

    def run():
        s = jb.Dopri853Te(y0.cuda(), None, x0, 1e-6, 1e-6, False, dyn, norm_M=4096)
        h = torch.full((M,), 0.1, dtype=torch.float64, device="cuda")
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        att = 0
        for _ in range(nsteps):
            s.step(h)
            h = s.hnext
            att += M + int(s.nrej.sum())
        e1.record()
        torch.cuda.synchronize()
        return s, e0.elapsed_time(e1) * 1e-3, att
    run()
    s, dt, att = run()
    rows = 16384
    o = Dopri853Oracle(y0[:rows].numpy(), f_vdp(None, y0[:rows].numpy()), np.zeros(rows), 1e-6, 1e-6, f_vdp, norm_M=4096)
    h = np.full(rows, 0.1)
    t0 = time.perf_counter()
    for _ in range(nsteps):
        o.step(h)
        h = o.hnext.copy()
    dtc = time.perf_counter() - t0
    same = np.allclose(s.y[:rows].cpu().numpy(), o.y, rtol=1e-10, atol=1e-12)
    F = 14                                    # FP64 instructions of one vdP dynamics evaluation (SASS-counted)
    flops = att * (11 * F + 4 * (2 * 59 + 30)) + nsteps * M * F      # 11 stage evaluations + tableau sums + error per attempt
    print(json.dumps({"metric": "trajectory-steps/sec (fp64), explicit DOPRI853", "value": nsteps * M / dt,
                      "unit": "trajectory-steps/s", "M": M, "accepted_steps_per_trajectory": nsteps, "attempts": att,
                      "ms_per_step_call": 1e3 * dt / nsteps, "algorithmic_tflops": flops / dt / 1e12,
                      "cpu_baseline": {"value": nsteps * rows / dtc, "unit": "trajectory-steps/s", "kind": "port",
                                       "cores": 1, "sample": f"numpy oracle (vectorised over {rows} rows), {dtc:.1f} s"},
                      "gpu_equals_oracle_on_sample": bool(same)}))


if __name__ == "__main__":
    main()
