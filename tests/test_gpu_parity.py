"""GPU parity tests (run on the B200): every call goes through the public API -> C ABI -> CUDA.

Bar (BASELINE.json): final states within 1e-8 relative, accepted-step counts equal for >= 99.9 %
of the trajectories.  Because the reference's controller is sensitive to the last bit
(tests/test_oracle_sensitivity.py), the count bar is checked against the oracle in MIRROR
arithmetic, where the expected agreement is exact; against the oracle in REFERENCE arithmetic
(no FMA, libm pow, closed-form derivatives -- what libtorch CPU ops do) the states must agree
to 1e-8 and the counts as well as two CPU roundoff realisations agree with each other."""
import os

import numpy as np
import pytest
import torch

import janus_ode_b200 as jb
from janus_ode_b200 import configs
from oracle import c_oracle

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DEV = "cuda:0"

PB = {0: "harmonic", 1: "vdp", 2: "vdp_pmp_smooth", 3: "vdp_pmp_bang", 4: "vdp_aug_bang"}


def gpu_solve(cfg, host=False, options=None):
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    y0, ts = cfg["y0"], cfg["tspan"]
    if not host:
        y0, ts = y0.to(DEV), ts.to(DEV)
    r = jb.RadauTe(dyn, dyn.jacobian, ts, y0, options or cfg["options"])
    res = r.solve()
    torch.cuda.synchronize()
    return dict(y=r.y.cpu().numpy(), t=r.t.cpu().numpy(), h=r.h.cpu().numpy(), res=res.cpu().numpy().astype(np.int32),
                stats=r.stats.raw.cpu().numpy(), nstg=r.nstg.cpu().numpy(), launches=r.gpu_launches)


def oracle_solve(cfg, arith, options=None):
    return c_oracle.solve(PB[cfg["ham"].ham_id], list(cfg["ham"].params), cfg["sign"], cfg["y0"].numpy(),
                          cfg["tspan"].numpy(), options or cfg["options"], arith=arith)


def assert_bitwise(a, b):
    for k in ("y", "t", "h", "res", "stats", "nstg"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def relerr(a, b):
    return np.abs(a - b).max(1) / np.maximum(np.abs(b).max(1), 1e-300)


# ---------------------------------------------------------------- building blocks
@pytest.mark.parametrize("ham,sign", [(jb.Harmonic(1.3), -1), (jb.Harmonic(1.3), 1), (jb.VanDerPol(1000.0), 1),
                                      (jb.VanDerPol(7.0), -1), (jb.VdpPmp(1.0, "smooth"), 1),
                                      (jb.VdpPmp(0.1, "bangbang", 1.0, 100.0), 1), (jb.VdpAug(1.0, 100.0), 1)])
def test_stage_eval_dual_numbers(ham, sign):
    torch.manual_seed(0)
    y = (torch.rand(4099, ham.Ny, dtype=torch.float64) * 4 - 2)
    dyn = jb.HamiltonianSystem(ham, sign)
    f = dyn(None, y.to(DEV)).cpu().numpy()
    J = dyn.jacobian(None, y.to(DEV)).cpu().numpy()
    pb, par = PB[ham.ham_id], list(ham.params)
    # mirror arithmetic (dense duals on the CPU): bit-identical
    assert np.array_equal(f, c_oracle.eval_dyns(pb, par, sign, y.numpy(), "mirror"))
    assert np.array_equal(J, c_oracle.eval_jac(pb, par, sign, y.numpy(), "mirror"))
    # closed-form derivatives: roundoff
    assert np.allclose(f, c_oracle.eval_dyns(pb, par, sign, y.numpy(), "reference"), rtol=1e-13, atol=1e-12)
    assert np.allclose(J, c_oracle.eval_jac(pb, par, sign, y.numpy(), "reference"), rtol=1e-13, atol=1e-12)


def test_stage_eval_vs_reference_autograd_golden():
    g = np.load(os.path.join(GOLD, "ham_autograd.npz"))
    y = torch.from_numpy(g["y"]).to(DEV)
    for ham, key in ((jb.VanDerPol(1000.0), "vdp"), (jb.VdpPmp(1.0, "smooth"), "pmp_smooth"),
                     (jb.VdpPmp(1.0, "bangbang", 1.0, 100.0), "pmp_bang")):
        dyn = jb.HamiltonianSystem(ham, +1)
        assert np.allclose(dyn(None, y).cpu().numpy(), g[key + "_dyn"], rtol=1e-13, atol=1e-12)
        assert np.allclose(dyn.jacobian(None, y).cpu().numpy(), g[key + "_jac"], rtol=1e-13, atol=1e-12)
    dyn = jb.HamiltonianSystem(jb.Harmonic(1.0), -1)
    y2 = torch.from_numpy(g["y2"]).to(DEV)
    assert np.allclose(dyn(None, y2).cpu().numpy(), g["harm_dyn"], rtol=1e-14)
    assert np.allclose(dyn.jacobian(None, y2).cpu().numpy(), g["harm_jac"], rtol=1e-14)


@pytest.mark.parametrize("n", [2, 3, 4, 6, 8])
@pytest.mark.parametrize("cplx", [False, True])
def test_batched_lu(n, cplx):
    rng = np.random.default_rng(n)
    M = 5003
    # Radau-shaped matrices gamma/h I - J with badly scaled J (entries O(mu) next to O(1))
    J = rng.standard_normal((M, n, n)) * 10.0 ** rng.integers(-2, 4, (M, n, n))
    h = 10.0 ** rng.uniform(-6, 2, M)
    A = (3.6378 / h)[:, None, None] * np.eye(n) - J
    b = rng.standard_normal((M, n))
    if cplx:
        A = A + 1j * (3.05 / h)[:, None, None] * np.eye(n)
        b = b + 1j * rng.standard_normal((M, n))
    x = jb.lu_solve(torch.from_numpy(A).to(DEV), torch.from_numpy(b).to(DEV)).cpu().numpy()
    assert np.array_equal(x, c_oracle.lu_solve(A, b, "mirror"))          # same pivots, same bits
    xt = torch.linalg.solve(torch.from_numpy(A), torch.from_numpy(b)).numpy()     # LAPACK
    err = np.abs(x - xt).max(1) / np.abs(xt).max(1)
    cond = np.linalg.cond(A)
    assert (err < 1e-13 * np.maximum(cond, 10)).all()
    res = np.abs(np.einsum("mij,mj->mi", A, x) - b).max(1) / (np.abs(A).max((1, 2)) * np.abs(x).max(1))
    assert res.max() < 1e-14


def test_det_pow_bitwise():
    rng = np.random.default_rng(0)
    x = 10 ** rng.uniform(-10, 3, 1 << 18)
    y = rng.choice([0.8, 4.0 / 3.0, 1.0 / 5 + 1.0, 1.0 / 7 + 1.0, 2.0, 1.0 / 6.0, -0.2, -1.0 / 9.0], 1 << 18)
    g = jb.det_pow(torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV)).cpu().numpy()
    assert np.array_equal(g, c_oracle.det_pow(x, y))
    assert np.abs(g / np.power(x, y) - 1).max() < 2e-14


# ---------------------------------------------------------------- the integrator
@pytest.mark.parametrize("name,mk", [
    ("c1_harmonic", configs.c1_harmonic), ("c2_vdp", configs.c2_vdp),
    ("c3_pmp_smooth", lambda **k: configs.c3_pmp(control="smooth", **k)),
    ("c3_pmp_bang", lambda **k: configs.c3_pmp(control="bangbang", **k))])
@pytest.mark.parametrize("tag", ["s3", "var"])
def test_golden_files(name, mk, tag):
    g = np.load(os.path.join(GOLD, f"{name}_{tag}.npz"))
    cfg = mk(M=g["y0"].shape[0], fixed_order=(tag == "s3"))
    a = gpu_solve(cfg)
    assert a["launches"] == 1 if tag == "s3" else a["launches"] > 1   # fused kernel vs epochs
    assert np.array_equal(a["y"], g["mir_y"]) and np.array_equal(a["stats"], g["mir_stats"])
    assert np.array_equal(a["res"], g["mir_res"]) and np.array_equal(a["t"], g["mir_t"])
    assert np.array_equal(a["nstg"], g["mir_nstg"]) and np.array_equal(a["h"], g["mir_h"])
    if name != "c3_pmp_bang":      # switching times of a discontinuous control are tolerance-level
        tol = 1e-8 if name != "c1_harmonic" else 2e-9 / 1e-1
        assert (relerr(a["y"], g["ref_y"]) < tol).mean() >= 0.95


@pytest.mark.parametrize("fixed", [True, False])
def test_c2_vdp_parity_16k(fixed):
    cfg = configs.c2_vdp(M=16384, fixed_order=fixed)
    a = gpu_solve(cfg)
    m = oracle_solve(cfg, "mirror")
    assert (a["res"] == 0).all()
    # the bar: counts equal for >= 99.9 %, states within 1e-8  (here: exactly)
    assert (a["stats"][:, 5] == m["stats"][:, 5]).mean() >= 0.999
    assert (relerr(a["y"], m["y"]) <= 1e-8).mean() >= 0.999
    assert_bitwise(a, m)
    # against reference arithmetic: same answers, counts as close as CPU-vs-CPU roundoff allows
    r = oracle_solve(cfg, "reference")
    rel = relerr(a["y"], r["y"])
    assert np.median(rel) < 1e-10 and (rel < 1e-8).mean() > (0.99 if fixed else 0.96)
    gpu_vs_ref = (a["stats"][:, 5] == r["stats"][:, 5]).mean()
    cpu_vs_ref = (m["stats"][:, 5] == r["stats"][:, 5]).mean()
    assert gpu_vs_ref == cpu_vs_ref
    assert abs(a["stats"][:, 5].sum() / r["stats"][:, 5].sum() - 1) < 0.01


def test_c1_harmonic_full_size_and_analytic():
    cfg = configs.c1_harmonic(M=1024)
    a = gpu_solve(cfg)
    assert_bitwise(a, oracle_solve(cfg, "mirror"))
    y0 = cfg["y0"].numpy()
    exact = np.stack([y0[:, 0] * np.cos(10) - y0[:, 1] * np.sin(10), y0[:, 1] * np.cos(10) + y0[:, 0] * np.sin(10)], 1)
    assert np.abs(a["y"] - exact).max() < 5e-9
    assert (a["t"] == 10.0).all() and (a["res"] == 0).all()


@pytest.mark.parametrize("control", ["smooth", "bangbang"])
def test_c3_pmp_parity(control):
    cfg = configs.c3_pmp(M=4096, control=control)
    a = gpu_solve(cfg)
    assert (a["res"] == 0).all()
    assert_bitwise(a, oracle_solve(cfg, "mirror"))


def test_aug_vdp_vs_scipy_point():
    # examples/python/aug_vdp_ref_test.py (SciPy Radau, rtol 1e-10): tolerance-level agreement
    from scipy.integrate import solve_ivp

    def f(t, y):
        p1, p2, x1, x2 = y
        u = 100 if p2 * (1 - x1 ** 2) * x2 < 0 else 1
        return [p2 * u * (-2.0 * x1) * x2 - p2, p1 + p2 * u * (1.0 - x1 ** 2), x2, u * (1.0 - x1 ** 2) * x2 - x1]
    sol = solve_ivp(f, (0, 10), [5.0, 5.0, 0.0, 1.5], method="Radau", rtol=1e-10, atol=1e-13)
    cfg = dict(ham=jb.VdpAug(1.0, 100.0), sign=1, y0=torch.tensor([[5.0, 5.0, 0.0, 1.5]], dtype=torch.float64),
               tspan=torch.tensor([[0.0, 10.0]], dtype=torch.float64),
               options=jb.OptionsTe(RelTol=1e-6, AbsTol=1e-9, MaxNbrStep=100000))
    a = gpu_solve(cfg)
    assert a["res"][0] == 0 and np.allclose(a["y"][0], sol.y[:, -1], rtol=2e-3)


def test_host_buffer_path_equals_device_path():
    cfg = configs.c2_vdp(M=3000, fixed_order=True)
    a = gpu_solve(cfg, host=False)
    b = gpu_solve(cfg, host=True)        # jb200_solve_host: copies inside the C ABI
    assert_bitwise(a, b)


def test_edge_cases():
    # empty batch
    dyn = jb.HamiltonianSystem(jb.VanDerPol(1000.0))
    r = jb.RadauTe(dyn, None, torch.zeros(0, 2, dtype=torch.float64, device=DEV) + torch.tensor([0.0, 1.0], dtype=torch.float64, device=DEV),
                   torch.zeros(0, 4, dtype=torch.float64, device=DEV))
    assert r.solve().numel() == 0 and r.y.shape == (0, 4)
    # single trajectory, ragged horizons, reverse time, mixed failures
    cfg = configs.c2_vdp(M=257)
    ts = cfg["tspan"].clone()
    ts[:, 1] = torch.linspace(1e-4, 40.0, 257, dtype=torch.float64)
    ts[100:, 0] = 3.0
    ts[5, 1] = 1e-9                                  # first step already overshoots -> clipped
    cfg["tspan"] = ts
    y0 = cfg["y0"].clone()
    y0[7, 2] = float("nan")                          # NaN state -> res 2
    cfg["y0"] = y0
    a = gpu_solve(cfg)
    assert_bitwise(a, oracle_solve(cfg, "mirror"))
    assert a["res"][7] == 2 and (np.delete(a["res"], 7) == 0).all()
    one = dict(cfg, y0=cfg["y0"][:1], tspan=cfg["tspan"][:1])
    assert_bitwise(gpu_solve(one), oracle_solve(one, "mirror"))
    # step cap -> res 3 ; blow-up (canonical sign on stiff vdP) -> failure code, same as the oracle
    op = jb.OptionsTe(RelTol=1e-6, AbsTol=1e-9, MaxNbrStep=30)
    c2 = configs.c2_vdp(M=64)
    a = gpu_solve(c2, options=op)
    assert_bitwise(a, oracle_solve(c2, "mirror", options=op))
    assert (a["res"] == 3).any()
    c2["sign"] = -1
    op = jb.OptionsTe(RelTol=1e-6, AbsTol=1e-9, MaxNbrStep=50000)
    a = gpu_solve(c2, options=op)
    assert_bitwise(a, oracle_solve(c2, "mirror", options=op))
    # reverse-time harmonic oscillator
    c1 = configs.c1_harmonic(M=64)
    c1["tspan"] = c1["tspan"].flip(1).contiguous()
    a = gpu_solve(c1)
    assert_bitwise(a, oracle_solve(c1, "mirror"))
    assert (a["t"] == 0.0).all()


@pytest.mark.parametrize("kw", [dict(NbrStg=1, MinNbrStg=1, MaxNbrStg=1), dict(NbrStg=1, MinNbrStg=1, MaxNbrStg=3),
                                dict(NbrStg=5, MinNbrStg=5, MaxNbrStg=5), dict(NbrStg=7, MinNbrStg=7, MaxNbrStg=7),
                                dict(Gustafsson=False), dict(Start_Newt=True), dict(MaxNbrNewton=4),
                                dict(RelTol=[1e-5, 1e-6, 1e-7, 1e-8], AbsTol=1e-10), dict(JacRecompute=0.1, Quot2=2.0)])
def test_option_coverage(kw):
    cfg = configs.c2_vdp(M=512)
    cfg["tspan"][:, 1] = 25.0 if kw.get("MaxNbrStg") == 1 else 900.0
    op = jb.OptionsTe(RelTol=1e-6, AbsTol=1e-9, MaxNbrStep=1000000)
    for k, v in kw.items():
        setattr(op, k, v)
    a = gpu_solve(cfg, options=op)
    assert_bitwise(a, oracle_solve(cfg, "mirror", options=op))
    if "Start_Newt" not in kw:        # zero starting values make some Newton iterations fail for good
        assert (a["res"] == 0).all()


def test_full_size_1m_properties():
    """BASELINE config 2 at full size: size-independent properties + a sampled oracle check."""
    M = 1 << 20
    cfg = configs.c2_vdp(M=M)
    a = gpu_solve(cfg)
    assert (a["res"] == 0).all()
    assert np.allclose(a["t"], cfg["tspan"][:, 1].numpy(), rtol=1e-15)
    st = a["stats"]
    assert (st[:, 5] >= 1).all() and (st[:, 4] >= st[:, 5]).all() and (st[:, 3] >= st[:, 4] - st[:, 6]).all()
    assert np.isfinite(a["y"]).all()
    # determinism: a second run gives identical bits although lanes pick up work in a different order
    b = gpu_solve(cfg)
    assert_bitwise(a, b)
    # sampled rows against the oracle
    idx = np.random.default_rng(0).choice(M, 4096, replace=False)
    sub = dict(cfg, y0=cfg["y0"][idx], tspan=cfg["tspan"][idx])
    m = oracle_solve(sub, "mirror")
    assert np.array_equal(a["y"][idx], m["y"]) and np.array_equal(st[idx], m["stats"])
    # vdP limit cycle: |x1| <= ~2.1 for every trajectory that has completed a jump
    assert np.abs(a["y"][:, 2]).max() < 3.01


@pytest.mark.gpu
def test_epoch_list_ordering_is_scheduling_only(monkeypatch):
    """the per-batch ordering decision (JB200_EPOCH_HORDER=4: H(y0) order for a one-parameter family, step-size order for a
    cloud; solver_impl.cuh ws_init / coherence_kernel) and both fixed policies give the same bits on both kinds of batch"""
    from janus_ode_b200 import configs
    for k, mk in enumerate((configs.c2_vdp, configs.c2_vdp_cloud)):
        cfg = mk(M=1 << 16)
        cfg["tspan"][:, 1] = 40.0
        outs = []
        for j, h in enumerate(("4", "2", "3")):
            monkeypatch.setenv("JB200_EPOCH_HORDER", h)
            op = jb.OptionsTe(RelTol=1e-6, AbsTol=1e-9, MaxNbrStep=1000000 + 10 * k + j)    # a fresh native handle per policy
            dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
            r = jb.RadauTe(dyn, dyn.jacobian, cfg["tspan"].to(DEV), cfg["y0"].to(DEV), op)
            r.solve()
            outs.append((r.y.clone(), r.stats.raw.clone(), r.h.clone(), r.nstg.clone()))
        for o in outs[1:]:
            assert all(torch.equal(a, b) for a, b in zip(outs[0], o))
        assert int(outs[0][3].max()) >= 3 and bool((outs[0][1][:, 5] > 10).all())
