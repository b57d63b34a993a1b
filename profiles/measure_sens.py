#!/usr/bin/env python3
"""Throughput of the forward-sensitivity solve (jb.RadauTeD -> jb200_solve_sens) on config 2 rows:
    python profiles/measure_sens.py [M]  -> one JSON line.  The plain solve of the same rows is timed beside it."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import janus_ode_b200 as jb          # noqa: E402
from janus_ode_b200 import configs   # noqa: E402


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return r, best


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
    out = {"metric": "trajectory-steps/sec (fp64), Radau IIA with forward sensitivities d y / d y0 (Ny x Ny)", "M": M}
    for fixed in (True, False):
        cfg = configs.c2_vdp(M=M, fixed_order=fixed)
        dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
        y0, ts = cfg["y0"].cuda(), cfg["tspan"].cuda()

        def sens():
            r = jb.RadauTeD(dyn, None, ts, y0, cfg["options"])
            r.solve()
            return r

        def plain():
            r = jb.RadauTe(dyn, None, ts, y0, cfg["options"])
            r.solve()
            return r
        rs, ts_ = timed(sens)
        rp, tp = timed(plain)
        acc = int(rs.stats.AccptNbr.sum())
        assert torch.equal(rs.y, rp.y)
        out["fixed s=3" if fixed else "variable order"] = {
            "value": acc / ts_, "unit": "trajectory-steps/s", "ms": 1e3 * ts_, "plain_solve_ms": 1e3 * tp,
            "sens_newton_iterations_per_accepted_step": float(rs.sens_iters.sum()) / acc,
            "algorithmic_tflops": float(rs.work.sum()) / ts_ / 1e12, "real_part_bitwise_equal_to_plain_solve": True}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
