"""Generates tests/golden/ref_live_*.npz: whole solves of the UNMODIFIED reference janus::RadauTe
(/root/reference/src/cpp/radaute.hpp + radaute_impl.hpp, compiled by oracle/Makefile.ref, one solver
object per trajectory -- see oracle/ref_driver.cpp) on the seeded rows of BASELINE.json's configs.

Run in the build container (needs /root/reference):   python tests/golden/make_reference_live.py [names...]
Each fixture holds the inputs (y0, tspan, tolerances, stage options) and the reference's outputs
y, t, h, res, stats [M,8] (StatsTe order; the reference never increments JacNbr/DecompNbr/SolveNbr/NewtRejNbr),
nstg (final NbrStg) and passes (the solver's `count`).

`*_adaptor` fixtures run the same rows through the reference's own evalDyns/evalJac autograd adaptor
instead of closed-form callbacks, `*_ulp` perturb x1(0) by one ulp: two more realisations of the SAME
reference algorithm whose disagreement with the base fixture is the reference's own roundoff floor.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from janus_ode_b200 import configs  # noqa: E402
from oracle import ref_live  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
PROCS = int(os.environ.get("REF_PROCS", "6"))


def run(name, problem, par, cfg, rows, var, adaptor=False, ulp=False):
    y0 = cfg["y0"][:rows].numpy().copy()
    ts = cfg["tspan"][:rows].numpy().copy()
    if ulp:
        y0[:, -2] = np.nextafter(y0[:, -2], np.inf)          # x1(0) + 1 ulp
    op = cfg["options"]
    stages = (3, 3, 7) if var else (3, 3, 3)
    t0 = time.time()
    r = ref_live.solve(problem, par, cfg["sign"], y0, ts, float(op.RelTol), float(op.AbsTol), h0=float(op.InitialStep),
                       stages=stages, max_steps=100000, adaptor=adaptor, procs=PROCS)
    dt = time.time() - t0
    np.savez_compressed(os.path.join(OUT, f"ref_live_{name}.npz"), y0=y0, tspan=ts, rtol=float(op.RelTol),
                        atol=float(op.AbsTol), h0=float(op.InitialStep), stages=np.array(stages), sign=cfg["sign"],
                        par=par, adaptor=int(adaptor), **r)
    print(f"{name}: {rows} rows in {dt:.0f} s; res {np.bincount(r['res'])}; accepted {r['stats'][:, 5].sum()};"
          f" nstg {np.bincount(r['nstg'])}", flush=True)


JOBS = {
    "c1_harmonic_var": lambda: run("c1_harmonic_var", "harmonic", 1.0, configs.c1_harmonic(1024), 1024, True),
    "c1_harmonic_s3": lambda: run("c1_harmonic_s3", "harmonic", 1.0, configs.c1_harmonic(1024, fixed_order=True), 1024, False),
    "c2_vdp_s3": lambda: run("c2_vdp_s3", "vdp", 1000.0, configs.c2_vdp(fixed_order=True), 1024, False),
    "c2_vdp_var": lambda: run("c2_vdp_var", "vdp", 1000.0, configs.c2_vdp(), 1024, True),
    "c2_vdp_s3_adaptor": lambda: run("c2_vdp_s3_adaptor", "vdp", 1000.0, configs.c2_vdp(fixed_order=True), 256, False, adaptor=True),
    "c2_vdp_var_adaptor": lambda: run("c2_vdp_var_adaptor", "vdp", 1000.0, configs.c2_vdp(), 256, True, adaptor=True),
    "c2_vdp_s3_ulp": lambda: run("c2_vdp_s3_ulp", "vdp", 1000.0, configs.c2_vdp(fixed_order=True), 256, False, ulp=True),
    "c2_vdp_var_ulp": lambda: run("c2_vdp_var_ulp", "vdp", 1000.0, configs.c2_vdp(), 256, True, ulp=True),
    "c3_pmp_bang_var": lambda: run("c3_pmp_bang_var", "vdp_pmp_bang", 1.0, configs.c3_pmp(1024, "bangbang"), 256, True),
    "c3_pmp_bang_s3": lambda: run("c3_pmp_bang_s3", "vdp_pmp_bang", 1.0, configs.c3_pmp(1024, "bangbang", fixed_order=True), 256, False),
    "c3_pmp_smooth_var": lambda: run("c3_pmp_smooth_var", "vdp_pmp_smooth", 1.0, configs.c3_pmp(1024, "smooth"), 1024, True),
    "c3_pmp_smooth_s3": lambda: run("c3_pmp_smooth_s3", "vdp_pmp_smooth", 1.0, configs.c3_pmp(1024, "smooth", fixed_order=True), 1024, False),
    "c1_harmonic_var_ulp": lambda: run("c1_harmonic_var_ulp", "harmonic", 1.0, configs.c1_harmonic(1024), 256, True, ulp=True),
    "c1_harmonic_s3_ulp": lambda: run("c1_harmonic_s3_ulp", "harmonic", 1.0, configs.c1_harmonic(1024, fixed_order=True), 256, False, ulp=True),
    "c3_pmp_bang_var_ulp": lambda: run("c3_pmp_bang_var_ulp", "vdp_pmp_bang", 1.0, configs.c3_pmp(1024, "bangbang"), 128, True, ulp=True),
}

if __name__ == "__main__":
    for n in (sys.argv[1:] or list(JOBS)):
        JOBS[n]()
