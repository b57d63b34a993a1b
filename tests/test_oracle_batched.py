"""The masked-tensor torch restatement (the reference's op structure; CPU-baseline stand-in of BASELINE.md
section 3) against the scalar C restatement: same algorithm, independently written."""
import numpy as np

from janus_ode_b200 import configs
from oracle import c_oracle, ham, radau_batched


def _run(cfg, pb, H, ctrl=None, W=None):
    ode, jac = radau_batched.hamiltonian_callbacks(H, W, ctrl, cfg["sign"])
    r = radau_batched.RadauBatched(ode, jac, cfg["tspan"], cfg["y0"], cfg["options"])
    res = r.solve()
    o = c_oracle.solve(pb, list(cfg["ham"].params), cfg["sign"], cfg["y0"].numpy(), cfg["tspan"].numpy(), cfg["options"])
    return r, res.numpy(), o


def test_batched_vdp_fixed_and_variable_order():
    for fixed in (True, False):
        cfg = configs.c2_vdp(M=48, fixed_order=fixed)
        cfg["tspan"][:, 1] = 600.0
        r, res, o = _run(cfg, "vdp", ham.H_vdp(1000.0))
        assert (res == 0).all() and (o["res"] == 0).all()
        rel = np.abs(r.y.numpy() - o["y"]).max(1) / np.abs(o["y"]).max(1)
        assert np.median(rel) < 1e-9 and rel.max() < 1e-6
        st = r.stats.numpy()
        assert abs(st[:, 5].sum() / o["stats"][:, 5].sum() - 1) < 0.02         # same total work
        assert np.mean(st[:, 5] == o["stats"][:, 5]) > 0.8                      # counts equal up to roundoff flips
        assert np.array_equal(r.t.numpy(), o["t"])


def test_batched_harmonic_analytic():
    cfg = configs.c1_harmonic(M=16)
    r, res, o = _run(cfg, "harmonic", ham.H_harmonic(1.0))
    y0 = cfg["y0"].numpy()
    exact = np.stack([y0[:, 0] * np.cos(10) - y0[:, 1] * np.sin(10), y0[:, 1] * np.cos(10) + y0[:, 0] * np.sin(10)], 1)
    assert (res == 0).all() and np.abs(r.y.numpy() - exact).max() < 5e-9
