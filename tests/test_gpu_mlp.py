"""GPU parity of the MLP-Hamiltonian path (config 4): public API -> C ABI -> tile-GEMM kernels + phase
kernels, bit for bit against the oracle in mirror arithmetic; 1e-8 against reference arithmetic."""
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


def gpu_solve(cfg, host=False):
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    y0, ts = cfg["y0"], cfg["tspan"]
    if not host:
        y0, ts = y0.to(DEV), ts.to(DEV)
    r = jb.RadauTe(dyn, dyn.jacobian, ts, y0, cfg["options"])
    res = r.solve()
    torch.cuda.synchronize()
    return dict(y=r.y.cpu().numpy(), t=r.t.cpu().numpy(), h=r.h.cpu().numpy(), res=res.cpu().numpy().astype(np.int32),
                stats=r.stats.raw.cpu().numpy(), nstg=r.nstg.cpu().numpy(), launches=r.gpu_launches,
                work=r.work.cpu().numpy() if r.work is not None else None)


def oracle_solve(cfg, arith):
    return c_oracle.solve("mlp", np.array(cfg["ham"].params), cfg["sign"], cfg["y0"].numpy(), cfg["tspan"].numpy(),
                          cfg["options"], arith=arith)


def assert_bitwise(a, b):
    for k in ("y", "t", "h", "res", "stats", "nstg"):
        assert np.array_equal(a[k], b[k], equal_nan=True), k


@pytest.mark.parametrize("M", [1, 3, 63, 64, 65, 200])
@pytest.mark.parametrize("bias", [0.0, 0.25])
def test_stage_evaluation_bitwise(M, bias):
    """mlp_grad_kernel / mlp_hess_kernel (ragged tiles included) == scalar mirror, every bit."""
    H = jb.MlpHamiltonian(8, 128, 0, bias_scale=bias)
    par = np.array(H.params)
    torch.manual_seed(M)
    y = 0.7 * torch.randn(M, 16, dtype=torch.float64)
    for sign in (-1, +1):
        dyn = jb.HamiltonianSystem(H, sign)
        f = dyn(None, y.to(DEV)).cpu().numpy()
        J = dyn.jacobian(None, y.to(DEV)).cpu().numpy()
        assert np.array_equal(f, c_oracle.eval_dyns("mlp", par, sign, y.numpy(), arith="mirror"))
        assert np.array_equal(J, c_oracle.eval_jac("mlp", par, sign, y.numpy(), arith="mirror"))


def test_stage_evaluation_vs_reference_arithmetic_and_saturation():
    H = jb.MlpHamiltonian(8, 128, 0)
    par = np.array(H.params)
    torch.manual_seed(1)
    y = torch.cat([0.5 * torch.randn(40, 16, dtype=torch.float64), 50.0 * torch.randn(8, 16, dtype=torch.float64)])
    dyn = jb.HamiltonianSystem(H, -1)
    f = dyn(None, y.to(DEV)).cpu().numpy()
    J = dyn.jacobian(None, y.to(DEV)).cpu().numpy()
    assert np.allclose(f, c_oracle.eval_dyns("mlp", par, -1, y.numpy()), rtol=1e-11, atol=1e-13)
    assert np.allclose(J, c_oracle.eval_jac("mlp", par, -1, y.numpy()), rtol=1e-10, atol=1e-12)
    assert np.array_equal(f, c_oracle.eval_dyns("mlp", par, -1, y.numpy(), arith="mirror"))   # saturated tanh rows too


@pytest.mark.parametrize("fixed", [True, False])
@pytest.mark.parametrize("M", [1, 65, 300])
def test_solve_bitwise_vs_mirror(fixed, M):
    cfg = configs.c4_mlp(M=M, fixed_order=fixed)
    g = gpu_solve(cfg)
    assert_bitwise(g, oracle_solve(cfg, "mirror"))
    assert (g["res"] == 0).all() and g["launches"] > 10 and (g["work"] > 0).all()


@pytest.mark.parametrize("opts", [dict(RelTol=1e-9, AbsTol=1e-12), dict(NbrStg=1, MinNbrStg=1, MaxNbrStg=7, RelTol=1e-4, AbsTol=1e-6),
                                  dict(Start_Newt=True), dict(MaxNbrStep=5), dict(Gustafsson=False)])
def test_solve_options_bitwise(opts):
    cfg = configs.c4_mlp(M=96)
    for k, v in opts.items():
        setattr(cfg["options"], k, v)
    g = gpu_solve(cfg)
    assert_bitwise(g, oracle_solve(cfg, "mirror"))
    if "MaxNbrStep" in opts:
        assert (g["res"] == 3).any()


def test_solve_host_buffers_and_reverse_time():
    cfg = configs.c4_mlp(M=130)
    cfg["tspan"] = torch.tensor([1.0, 0.5], dtype=torch.float64).repeat(130, 1)
    g = gpu_solve(cfg, host=True)
    assert_bitwise(g, oracle_solve(cfg, "mirror"))
    assert np.all(g["t"] == 0.5)


def test_solve_vs_reference_arithmetic():
    cfg = configs.c4_mlp(M=256)
    g, o = gpu_solve(cfg), oracle_solve(cfg, "reference")
    rel = np.abs(g["y"] - o["y"]).max(1) / np.abs(o["y"]).max(1)
    assert rel.max() < 1e-8                                               # BASELINE.json bar: 1e-8 relative
    assert (g["stats"][:, 5] == o["stats"][:, 5]).mean() >= 0.999         # ... and >= 99.9 % equal accepted-step counts


@pytest.mark.parametrize("tag", ["s3", "var"])
def test_golden_c4(tag):
    gd = np.load(os.path.join(GOLD, f"c4_mlp_{tag}.npz"))
    cfg = configs.c4_mlp(M=gd["y0"].shape[0], fixed_order=(tag == "s3"))
    g = gpu_solve(cfg)
    assert np.array_equal(g["y"], gd["mir_y"]) and np.array_equal(g["stats"], gd["mir_stats"])
    assert np.allclose(g["y"], gd["ref_y"], rtol=1e-8, atol=1e-12)


@pytest.mark.parametrize("fixed", [True, False])
@pytest.mark.parametrize("reverse", [False, True])
def test_output_at_tspan_points_bitwise(fixed, reverse):
    """OutFlag 3 (radau.m:1150-1168): tout / yout / nout at interior tspan points, forward and reverse time."""
    cfg = configs.c4_mlp(M=70, fixed_order=fixed)
    grid = torch.linspace(0.0, 1.0, 9, dtype=torch.float64)
    if reverse:
        grid = torch.flip(grid, [0]) * 0.7 + 0.1
    ts = grid.repeat(70, 1)
    ts[:, 1:-1] += 0.01 * torch.rand(70, 7, dtype=torch.float64, generator=torch.Generator().manual_seed(3))   # per-row grids
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    r = jb.RadauTe(dyn, None, ts.to(DEV), cfg["y0"].to(DEV), cfg["options"])
    r.solve()
    torch.cuda.synchronize()
    o = c_oracle.solve_dense("mlp", np.array(cfg["ham"].params), cfg["sign"], cfg["y0"].numpy(), ts.numpy(), cfg["options"],
                             arith="mirror")
    assert (o["nout"] >= 8).all()      # (a final point missed by one ulp is overwritten with (tfinal, y): radau.m:1267-1269)
    assert np.array_equal(r.nout.cpu().numpy(), o["nout"]) and np.array_equal(r.tout.cpu().numpy(), o["tout"])
    assert np.array_equal(r.yout.cpu().numpy(), o["yout"]) and np.array_equal(r.y.cpu().numpy(), o["y"])
    assert np.array_equal(r.stats.raw.cpu().numpy(), o["stats"])


def test_full_size_properties():
    """config 4 at BASELINE.json's full size (2^20 trajectories): size-independent properties -- every trajectory ends at
    tfinal with res = 0, H is conserved along the flow (canonical sign), integrating back returns to y0 -- plus bit
    parity with the oracle on a strided sample."""
    from oracle import ham as oham
    M = 1 << 20
    cfg = configs.c4_mlp(M=M)
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    y0, ts = cfg["y0"].to(DEV), cfg["tspan"].to(DEV)
    r = jb.RadauTe(dyn, None, ts, y0, cfg["options"])
    res = r.solve()
    assert bool((res == 0).all()) and bool((r.t == 1.0).all())
    idx = torch.arange(0, M, 257)[:2048]
    o = c_oracle.solve("mlp", np.array(cfg["ham"].params), -1, cfg["y0"][idx].numpy(), cfg["tspan"][idx].numpy(),
                       cfg["options"], arith="mirror")
    assert np.array_equal(r.y[idx.to(DEV)].cpu().numpy(), o["y"]) and np.array_equal(r.stats.raw[idx.to(DEV)].cpu().numpy(), o["stats"])
    H = oham.MLPHamiltonian(N=8, width=128, seed=0)
    ya, yb = cfg["y0"][idx], r.y[idx.to(DEV)].cpu()
    Ha, Hb = H(ya[:, 8:], ya[:, :8]), H(yb[:, 8:], yb[:, :8])          # y = [p, x]
    assert float((Hb - Ha).abs().max()) < 1e-6
    back = jb.RadauTe(dyn, None, torch.flip(ts, [1]).contiguous(), r.y, cfg["options"])
    back.solve()
    assert bool((back.res == 0).all())
    assert float((back.y - y0).abs().max()) < 1e-5


@pytest.mark.parametrize("fixed", [True, False])
def test_step_protocol_reproduces_solve_bit_for_bit(fixed):
    """state = RadauState(...); state = RadauT(opts).step(htry, dyns, ft, state) -- one accepted step per trajectory and
    call (src/python/seulext.py:158-581); looping it is the same integration as solve(), bit for bit"""
    cfg = configs.c4_mlp(M=100, fixed_order=fixed)
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    ft = cfg["tspan"][:, 1].to(DEV)
    state = jb.RadauState(cfg["y0"].to(DEV), cfg["tspan"][:, 0].to(DEV), atol=1e-9, rtol=1e-6, ft=ft)
    solver = jb.RadauT(cfg["options"])
    htry, n = cfg["options"].InitialStep, 0
    x_prev = state.x.clone()
    while not bool(state.done.all()):
        state = solver.step(htry, dyn, ft, state)
        htry = state.hnext
        live = ~state.done
        assert torch.equal((x_prev + state.hdid)[live], state.x[live])
        x_prev = state.x.clone()
        n += 1
        assert n < 200
    o = oracle_solve(cfg, "mirror")
    assert np.array_equal(state.y.cpu().numpy(), o["y"]) and np.array_equal(state.x.cpu().numpy(), o["t"])
    assert np.array_equal(state.stats.raw.cpu().numpy(), o["stats"]) and bool((state.res == 0).all())
    assert n == int(o["stats"][:, 5].max())


def test_failure_codes_rowwise():
    """a NaN initial state fails only its own row (res = 2, radaute_impl.hpp:873-877); MaxNbrStep exhaustion gives 3"""
    cfg = configs.c4_mlp(M=40)
    cfg["y0"][7, 3] = float("nan")
    cfg["y0"][21] = float("inf")
    g = gpu_solve(cfg)
    o = oracle_solve(cfg, "mirror")
    assert np.array_equal(g["res"], o["res"]) and g["res"][7] == 2 and g["res"][21] == 2
    ok = g["res"] == 0
    assert ok.sum() == 38 and np.array_equal(g["y"][ok], o["y"][ok]) and np.array_equal(g["stats"], o["stats"])


def test_unsupported_and_invalid():
    H = jb.MlpHamiltonian(8, 128, 0)
    dyn = jb.HamiltonianSystem(H, -1)
    y0 = torch.zeros(4, 16, dtype=torch.float64, device=DEV)
    tgrid = torch.tensor([0.0, 0.5, 1.0], dtype=torch.float64, device=DEV).repeat(4, 1)
    with pytest.raises(jb.JanusB200Error, match="thread-resident"):
        jb.Dopri853Te(y0, None, 0.0, 1e-6, 1e-6, False, dyn).step(0.1)
    bad = jb.Hamiltonian("mlp", 5, 8, H.params[:100])
    with pytest.raises(ValueError, match="parameters"):
        jb.RadauTe(jb.HamiltonianSystem(bad, -1), None, tgrid[:, [0, 2]].contiguous(), y0, jb.OptionsTe()).solve()
    small = jb.Hamiltonian("mlp", 5, 8, (4.0, 128.0) + H.params[2:])
    with pytest.raises(jb.JanusB200Error, match="N = 8"):
        jb.RadauTe(jb.HamiltonianSystem(small, -1), None, tgrid[:, [0, 2]].contiguous(), y0, jb.OptionsTe()).solve()


@pytest.mark.parametrize("cplx", [False, True])
@pytest.mark.parametrize("kind", ["radau", "random"])
def test_lu16_bitwise(cplx, kind):
    """16-lane LU factor + solve (csrc/lu16.cuh) == the scalar mirror, every bit; 'random' matrices are
    not diagonally dominant, so the pivot search (first maximum), the row exchanges and the pivot
    bookkeeping are all exercised."""
    n, M = 16, 3001
    rng = np.random.default_rng(7)
    if kind == "radau":
        J = rng.standard_normal((M, n, n)) * 10.0 ** rng.integers(-2, 3, (M, n, n))
        h = 10.0 ** rng.uniform(-5, 1, M)
        A = (3.6378 / h)[:, None, None] * np.eye(n) - J
        if cplx:
            A = A + 1j * (3.05 / h)[:, None, None] * np.eye(n)
    else:
        A = rng.standard_normal((M, n, n))
        A[::7, 3, 0] = 5.0                          # exact tie for the first pivot: rows 3 and 9 ...
        A[::7, 9, 0] = 5.0                          # ... the first maximum (row 3) must win
        A[1::7, 12, 0] = -6.0                       # and a negative entry of larger modulus
        if cplx:
            A = A + 1j * rng.standard_normal((M, n, n))
    b = rng.standard_normal((M, n)) + (1j * rng.standard_normal((M, n)) if cplx else 0)
    x = jb.lu_solve(torch.from_numpy(A).to(DEV), torch.from_numpy(b).to(DEV)).cpu().numpy()
    xo = c_oracle.lu_solve(A, b, "mirror")
    assert np.array_equal(x, xo, equal_nan=True)
    xt = torch.linalg.solve(torch.from_numpy(A), torch.from_numpy(b)).numpy()
    err = np.abs(x - xt).max(1) / np.abs(xt).max(1)
    assert (err < 1e-12 * np.maximum(np.linalg.cond(A), 10)).all()
