#!/usr/bin/env python3
"""Generate tests/golden/*.npz -- small golden vectors of the hot path.

The reference cannot be imported or compiled here, so these vectors come from
the repo's own restatements, cross-checked against each other before being written:
  * oracle/radau_scalar.py (Python, LAPACK LU)  ~  oracle/c (C, reference arithmetic): states to
    1e-6, accepted-step counts within 10 % (they differ in roundoff -- LAPACK vs hand LU -- and the
    reference's controller is sensitive to the last bit, see tests/test_oracle_sensitivity.py)
  * oracle/ham.py (torch autograd, the engine the reference uses) == closed forms, 1e-13
They pin (a) the reference-arithmetic results and (b) the mirror-arithmetic results that the
CUDA path must reproduce bit for bit.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from janus_ode_b200 import configs  # noqa: E402
from oracle import c_oracle, ham, problems  # noqa: E402
from oracle.radau_scalar import radau_solve  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def solve_case(name, pb, cfg, py_fj, n_py=4, rtol=1e-6, dacc=0.10):
    par = list(cfg["ham"].params)
    y0, ts, op = cfg["y0"].numpy(), cfg["tspan"].numpy(), cfg["options"]
    ref = c_oracle.solve(pb, par, cfg["sign"], y0, ts, op, arith="reference")
    mir = c_oracle.solve(pb, par, cfg["sign"], y0, ts, op, arith="mirror")
    f, j = py_fj
    for m in range(n_py):   # independent Python restatement on the first rows
        r = radau_solve(f, j, ts[m, 0], ts[m, 1], y0[m], op)
        st = np.array([r.stats[k] for k in c_oracle.STAT_NAMES])
        if dacc is not None:
            assert abs(st[5] - ref["stats"][m][5]) <= max(3, dacc * st[5]), (name, m, st, ref["stats"][m])
        assert np.allclose(r.y, ref["y"][m], rtol=rtol, atol=rtol * 1e-3 * np.abs(r.y).max()), (name, m, r.y, ref["y"][m])
    np.savez(os.path.join(OUT, name + ".npz"), y0=y0, tspan=ts,
             ref_y=ref["y"], ref_t=ref["t"], ref_res=ref["res"], ref_stats=ref["stats"], ref_nstg=ref["nstg"],
             mir_y=mir["y"], mir_t=mir["t"], mir_h=mir["h"], mir_res=mir["res"], mir_stats=mir["stats"],
             mir_nstg=mir["nstg"])
    print(name, "written; accepted steps", int(ref["stats"][:, 5].sum()))


def main():
    for fixed in (True, False):
        tag = "s3" if fixed else "var"
        # variable order at rtol 1e-9: the order-switch thresholds make the counts chaotic in
        # roundoff (35 vs 85 accepted steps between two arithmetics, both 1e-10 accurate) -> states only
        solve_case(f"c1_harmonic_{tag}", "harmonic", configs.c1_harmonic(M=64, fixed_order=fixed),
                   problems.harmonic(1.0, -1), dacc=0.10 if fixed else None)
        solve_case(f"c2_vdp_{tag}", "vdp", configs.c2_vdp(M=64, fixed_order=fixed), problems.vdp(1000.0, +1))
        solve_case(f"c3_pmp_smooth_{tag}", "vdp_pmp_smooth", configs.c3_pmp(M=32, control="smooth", fixed_order=fixed),
                   problems.vdp_pmp(problems.ctrl_smooth(1.0), 1.0, +1), n_py=2)
        solve_case(f"c3_pmp_bang_{tag}", "vdp_pmp_bang", configs.c3_pmp(M=32, control="bangbang", fixed_order=fixed),
                   problems.vdp_pmp(problems.ctrl_bangbang(1.0, 100.0), 1.0, +1), n_py=2, rtol=2e-3, dacc=0.1)
        # (discontinuous control: the switching times are tolerance-level quantities)
    # C4 (MLP Hamiltonian, Ny = 16): the Python restatement integrates the autograd dynamics (the engine the
    # reference differentiates H with) on the first rows
    mlp = ham.MLPHamiltonian(N=8, width=128, seed=0)
    for fixed in (True, False):
        solve_case(f"c4_mlp_{'s3' if fixed else 'var'}", "mlp", configs.c4_mlp(M=16, fixed_order=fixed),
                   problems.from_hamiltonian(mlp, sign=-1), n_py=2)
    # Hamiltonian adaptor golden: autograd (reference engine) on seeded points
    torch.manual_seed(0)
    y = torch.rand(32, 4, dtype=torch.float64) * 4 - 2
    out = dict(y=y.numpy())
    out["vdp_dyn"] = ham.eval_dyns(y, None, ham.H_vdp(1000.0)).numpy()
    out["vdp_jac"] = ham.eval_jac(y, None, ham.H_vdp(1000.0)).numpy()
    u = ham.control_smooth(y[:, 2:], y[:, :2], 1.0)
    out["pmp_smooth_dyn"] = ham.eval_dyns(y, 1.0, ham.Hu_vdp_pmp, u).numpy()
    out["pmp_smooth_jac"] = ham.eval_jac(y, 1.0, ham.Hu_vdp_pmp, u).numpy()
    ub = ham.control_bangbang()(y[:, 2:], y[:, :2], 1.0)
    out["pmp_bang_dyn"] = ham.eval_dyns(y, 1.0, ham.Hu_vdp_pmp, ub).numpy()
    out["pmp_bang_jac"] = ham.eval_jac(y, 1.0, ham.Hu_vdp_pmp, ub).numpy()
    y2 = y[:, :2].contiguous()
    out["y2"] = y2.numpy()
    out["harm_dyn"] = ham.eval_dyns(y2, None, ham.H_harmonic(1.0), costate_sign=-1).numpy()
    out["harm_jac"] = ham.eval_jac(y2, None, ham.H_harmonic(1.0), costate_sign=-1).numpy()
    np.savez(os.path.join(OUT, "ham_autograd.npz"), **out)
    print("ham_autograd written")


if __name__ == "__main__":
    main()
