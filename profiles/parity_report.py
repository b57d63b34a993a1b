#!/usr/bin/env python3
"""Parity evidence at BASELINE.json's full sizes: GPU vs the C oracle (mirror / reference arithmetic).
    python profiles/parity_report.py > profiles/r01_parity_report_full.txt      (on a GPU box; ~5 min of host time)
config 2: all 2^20 rows, fixed s=3 and variable order;  config 4: the first 2^17 rows of the 2^20 batch."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import janus_ode_b200 as jb          # noqa: E402
from janus_ode_b200 import configs   # noqa: E402
from oracle import c_oracle          # noqa: E402


def report(tag, pb, cfg):
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    r = jb.RadauTe(dyn, None, cfg["tspan"].cuda(), cfg["y0"].cuda(), cfg["options"])
    res = r.solve()
    torch.cuda.synchronize()
    y, st, ns = r.y.cpu().numpy(), r.stats.raw.cpu().numpy(), r.nstg.cpu().numpy()
    outs = {}
    for arith in ("mirror", "reference"):
        o = c_oracle.solve(pb, np.array(cfg["ham"].params), cfg["sign"], cfg["y0"].numpy(), cfg["tspan"].numpy(), cfg["options"],
                           arith=arith)
        outs[arith] = o
        rel = np.abs(y - o["y"]).max(1) / np.maximum(np.abs(o["y"]).max(1), 1e-300)
        print(f"{tag} | GPU vs oracle[{arith:9s}] on {len(y)} rows: bit-identical states {np.mean((y == o['y']).all(1)):.6f}, "
              f"all 8 counters equal {np.mean((st == o['stats']).all(1)):.6f}, final order equal {np.mean(ns == o['nstg']):.6f}, "
              f"accepted-step count equal {np.mean(st[:, 5] == o['stats'][:, 5]):.6f}, rel.err median {np.median(rel):.2e} "
              f"max {rel.max():.2e}, within 1e-8: {np.mean(rel <= 1e-8):.6f}, res==0: {bool((res == 0).all())}", flush=True)
    a, b = outs["mirror"], outs["reference"]
    rel = np.abs(a["y"] - b["y"]).max(1) / np.maximum(np.abs(b["y"]).max(1), 1e-300)
    print(f"{tag} | CPU oracle[mirror] vs CPU oracle[reference] (two roundoff realisations of one algorithm): accepted-step "
          f"count equal {np.mean(a['stats'][:, 5] == b['stats'][:, 5]):.6f}, within 1e-8: {np.mean(rel <= 1e-8):.6f}, total "
          f"accepted steps {int(a['stats'][:, 5].sum())} vs {int(b['stats'][:, 5].sum())}", flush=True)


def main():
    print("# parity report at full sizes, one B200 vs the C oracle on the box's host cores")
    for fixed in (True, False):
        report(f"config 2, {'fixed s=3' if fixed else 'variable 3/5/7'}", "vdp", configs.c2_vdp(M=1 << 20, fixed_order=fixed))
    for fixed in (True, False):
        report(f"config 4, {'fixed s=3' if fixed else 'variable 3/5/7'}", "mlp",
               configs.c4_mlp(M=1 << 20, fixed_order=fixed, lo=0, hi=1 << 17))
    c3 = configs.c3_pmp(M=1 << 18, control="bangbang")
    report("config 3 (bang-bang control), variable 3/5/7", "vdp_pmp_bang", c3)


if __name__ == "__main__":
    main()
