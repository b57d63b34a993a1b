#!/usr/bin/env python3
"""Benchmark of the accelerated path: trajectory-steps/s (fp64) of the batched Hamiltonian Radau
solve on the BASELINE.json headline workload (config 2: stiff van der Pol, mu = 1000, 1M
trajectories per GPU, reference-default variable order 3/5/7).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--M rows-per-gpu]
                  [--fixed-order]
  N > 1: launched by torchrun, one rank per GPU; trajectories are independent, so each rank
  integrates its interleaved shard (rows r, r+N, ... of the seeded global batch of N*M rows) with
  no data-path collective; NCCL only reduces the counters / times ("scaling": "weak").

A "step" = one solve of the whole per-GPU batch.  `value` = accepted trajectory-steps of all
ranks / max-over-ranks device time with inputs resident in HBM; `e2e` = the same through the
public API with pinned HOST tensors (jb200_solve_host: H2D + solve + D2H inside the timed call).
`roofline` = algorithmic FLOPs (SURVEY.md 8d model, counted per trajectory inside the kernel) over
the CUDA-event kernel time, against the FP64 DFMA peak measured on this GPU in the same run
(MEASURED_PEAKS.json has no FP64 figure).  `cpu_baseline` / `--impl reference` = the UNMODIFIED reference
solver janus::RadauTe (oracle/_ref/libref_radau.so, built in the authoring container from /root/reference by
oracle/Makefile.ref; "kind": "reference"), one solver object per trajectory -- the only way it runs: its Solvrad
throws for any masked subset of more than one row (radaute_impl.hpp:1746-1749) -- one worker process per host
core, on a bounded sample of the same seeded rows; the same sample size in both arms.  If the prebuilt library
is missing the masked-tensor torch restatement (oracle/radau_batched.py, "kind": "port") stands in.
`cpu_baseline_strong` = the scalar C restatement with OpenMP over trajectories (oracle/c), a much faster CPU
design than the reference's, on a sample large enough that its rate no longer depends on the sample size.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "trajectory-steps/sec (fp64)"
UNIT = "trajectory-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--M", type=int, default=1 << 20, help="trajectories per GPU")
    ap.add_argument("--workload", default="c2", choices=["c2", "c2cloud", "c3", "c4"],
                    help="c2 = the headline (stiff van der Pol, BASELINE.json configs[1]); c2cloud = the same problem from a "
                         "4-dimensional cloud of initial states; c3 = PMP Hamiltonian with supplied control (configs[2], "
                         "--control); c4 = MLP Hamiltonian (configs[3]: 4-layer width-128 MLP, Ny = 16)")
    ap.add_argument("--control", default="bangbang", choices=["bangbang", "smooth"], help="control law of workload c3")
    ap.add_argument("--no-secondary", action="store_true", help="skip the short config-4 run appended to the default line")
    ap.add_argument("--fixed-order", action="store_true", help="fixed 3 stages instead of the default 3/5/7")
    ap.add_argument("--cpu-sample", type=int, default=0, help="rows of the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-kind", default="live", choices=["live", "c", "batched"],
                    help="reference arm: 'live' (default) = the unmodified reference RadauTe (oracle/_ref), one process per "
                         "host core; 'batched' = masked-tensor torch restatement (stand-in when oracle/_ref is missing, "
                         "always used for workload c4 which the reference cannot express); 'c' = scalar C restatement, "
                         "OpenMP over trajectories (always reported as cpu_baseline_strong)")
    return ap.parse_args()


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries (NCCL with NCCL_DEBUG=VERSION/INFO, nvcc, ...) write to fd 1; the contract is ONE JSON line
    on stdout.  Everything else goes to stderr, the JSON line to the original stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(out):
    line = (json.dumps(out) + "\n").encode()
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, line)


def make_cfg(a, M, lo=0, hi=None):
    from janus_ode_b200 import configs
    if a.workload == "c4":
        return configs.c4_mlp(M=M, fixed_order=a.fixed_order, lo=lo, hi=hi), "mlp"
    if a.workload == "c3":
        return (configs.c3_pmp(M=M, control=a.control, fixed_order=a.fixed_order, lo=lo, hi=hi),
                "vdp_pmp_bang" if a.control == "bangbang" else "vdp_pmp_smooth")
    if a.workload == "c2cloud":
        return configs.c2_vdp_cloud(M=M, fixed_order=a.fixed_order, lo=lo, hi=hi), "vdp"
    return configs.c2_vdp(M=M, fixed_order=a.fixed_order, lo=lo, hi=hi), "vdp"


def workload_name(a):
    order = "fixed s=3" if a.fixed_order else "variable order 3/5/7 (reference default)"
    if a.workload == "c4":
        return (f"config 4: MLP Hamiltonian 16->128->128->128->1 tanh (N=8, Ny=16), {a.M} trajectories/GPU fp64, "
                f"RelTol 1e-6 AbsTol 1e-9, t in [0,1], canonical sign, {order}")
    if a.workload == "c3":
        return (f"config 3: PMP Hamiltonian with supplied {a.control} control (costate system), {a.M} trajectories/GPU fp64, "
                f"y0=[-5,-5,0,1.5]+0.1N(0,1), RelTol 1e-6 AbsTol 1e-9, t in [0,{10 if a.control == 'bangbang' else 0.05}], {order}")
    cloud = " from a 4-D cloud of initial states (p0=(1,1)+0.3N, x1=2+U, x2=0.2N)" if a.workload == "c2cloud" else ""
    return (f"config 2: stiff van der Pol Hamiltonian form mu=1000{cloud}, {a.M} trajectories/GPU fp64, "
            f"RelTol 1e-6 AbsTol 1e-9, t in [0,(3-2ln2)mu], {order}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.rows = []
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(a, rows, threads=0):
    """C restatement of the reference algorithm on `rows` rows of the workload (reference arithmetic)."""
    import numpy as np
    from janus_ode_b200 import configs
    from oracle import c_oracle
    cfg, pb = make_cfg(a, a.M * a.gpus, 0, rows)
    y0, ts = cfg["y0"].numpy(), cfg["tspan"].numpy()
    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1: do not inherit that)
    cores = len(os.sched_getaffinity(0)) if threads <= 0 else threads
    t0 = time.perf_counter()
    o = c_oracle.solve(pb, list(cfg["ham"].params), cfg["sign"], y0, ts, cfg["options"], nthreads=cores,
                       arith="reference")
    dt = time.perf_counter() - t0
    acc = int(o["stats"][:, 5].sum())
    assert (o["res"] == 0).all()
    return acc / dt, dt, acc, cores


def batched_reference_run(a, rows):
    """masked-tensor torch restatement (oracle/radau_batched.py) with the autograd Hamiltonian adaptor: the op
    structure of the reference's RadauTe + evalDyns/evalJac, all host threads."""
    import torch
    from janus_ode_b200 import configs
    from oracle import ham, radau_batched
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    cfg, pb = make_cfg(a, a.M * a.gpus, 0, rows)
    H = ham.MLPHamiltonian(N=8, width=128, seed=0) if pb == "mlp" else ham.H_vdp(cfg["ham"].params[0])
    ode, jac = radau_batched.hamiltonian_callbacks(H, None, None, cfg["sign"])
    t0 = time.perf_counter()
    r = radau_batched.RadauBatched(ode, jac, cfg["tspan"], cfg["y0"], cfg["options"])
    res = r.solve()
    dt = time.perf_counter() - t0
    assert bool((res == 0).all())
    acc = int(r.stats[:, 5].sum())
    return acc / dt, dt, acc, cores


def live_reference_available(a):
    """the prebuilt oracle/_ref library (never built here: /root/reference does not exist on the GPU box); the
    reference has no MLP Hamiltonian instance, so config 4 always uses the port."""
    if a.workload == "c4":
        return False
    try:
        from oracle import ref_live
        return os.path.exists(ref_live._SO)
    except Exception:
        return False


def live_reference_rows(a):
    return a.cpu_sample or 16 * len(os.sched_getaffinity(0))


def live_reference_run(a, rows):
    """UNMODIFIED janus::RadauTe + its own evalDyns / evalJac autograd adaptor, one solver object per trajectory,
    one worker process per host core.  Rate = accepted steps / (solver seconds summed over rows / cores): process
    start-up (importing torch) is not charged to the reference."""
    from oracle import ref_live
    cores = len(os.sched_getaffinity(0))
    cfg, pb = make_cfg(a, a.M * a.gpus, 0, rows)
    op = cfg["options"]
    t0 = time.perf_counter()
    r = ref_live.solve(pb, cfg["ham"].params[0], cfg["sign"], cfg["y0"].numpy(), cfg["tspan"].numpy(), float(op.RelTol),
                       float(op.AbsTol), h0=float(op.InitialStep), stages=(op.NbrStg, op.MinNbrStg, op.MaxNbrStg),
                       max_steps=int(min(op.MaxNbrStep, 2000000000)), adaptor=True, procs=cores)
    wall = time.perf_counter() - t0
    assert (r["res"] == 0).all(), "the reference failed on some rows"
    acc = int(r["stats"][:, 5].sum())
    dt = float(r["secs"].sum()) / cores
    return acc / dt, dt, acc, cores, wall


LIVE_SAMPLE = ("first {rows} rows of the workload; UNMODIFIED reference janus::RadauTe + evalDyns/evalJac (autograd adaptor) "
               "from oracle/_ref/libref_radau.so, one solver object per trajectory (its Solvrad throws for masked subsets "
               "of more than one row), {cores} worker processes; {dt:.1f} s of solver time per core")


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if a.ref_kind == "live" and not live_reference_available(a):
        a.ref_kind = "batched"
    if a.ref_kind == "live":
        rows = live_reference_rows(a)
        rates, times = [], []
        for _ in range(max(a.steps, 1)):
            r, dt, acc, cores, wall = live_reference_run(a, rows)
            rates.append(r)
            times.append(dt)
        value = sum(rates) / len(rates)
        out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
               "warmup": 0, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload_name(a), "sample": f"first {rows} rows of the same seeded batch"},
               "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                                "sample": LIVE_SAMPLE.format(rows=rows, cores=cores, dt=sum(times) / len(times))},
               "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        emit(out)
        return
    if a.ref_kind == "batched":
        rows = a.cpu_sample or (16384 if a.workload == "c2" else 256)
        srows = 65536 if a.workload == "c2" else 4096
        rates, times = [], []
        for _ in range(max(a.steps, 1)):
            r, dt, acc, cores = batched_reference_run(a, rows)
            rates.append(r)
            times.append(dt)
        value = sum(rates) / len(rates)
        from oracle import c_oracle
        c_oracle.build()
        s_rate, s_dt, _, s_cores = cpu_reference_run(a, srows)
        out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
               "warmup": 0, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
               "cpu_baseline_strong": {"value": s_rate, "unit": UNIT, "cores": s_cores, "kind": "port",
                                       "sample": f"first {srows} rows, {s_dt:.1f} s; scalar C restatement, one trajectory per "
                                                 f"OpenMP thread (oracle/c) -- a stronger bar than the reference's design"},
               "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload_name(a), "sample": f"first {rows} rows of the same seeded batch"},
               "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"first {rows} rows; masked-tensor torch restatement of RadauTe with the autograd "
                                          f"Hamiltonian adaptor (oracle/radau_batched.py + oracle/ham.py): the reference's op "
                                          f"structure; the reference itself cannot be built here (SURVEY.md 8c)"},
               "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gpu_launches": 0}
        emit(out)
        return
    from oracle import c_oracle
    c_oracle.build()
    rows = a.cpu_sample or 65536
    r0 = 8192 if a.workload == "c2" else 1024
    rate0, dt0, _, cores = cpu_reference_run(a, r0)                 # size the sample: ~10 s per step
    rows = a.cpu_sample or int(min(a.M * a.gpus, max(r0, r0 * 10.0 / max(dt0, 1e-3))))
    for _ in range(a.warmup and 1):
        cpu_reference_run(a, min(rows, r0))
    rates, times = [], []
    for _ in range(a.steps):
        r, dt, acc, cores = cpu_reference_run(a, rows)
        rates.append(r)
        times.append(dt)
    value = sum(rates) / len(rates)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(a), "sample": f"first {rows} rows of the same seeded batch"},
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                            "sample": f"first {rows} rows of the workload, C restatement of the reference algorithm "
                                      f"(oracle/c, no-FMA reference arithmetic, OpenMP over trajectories); the "
                                      f"reference itself cannot be built here (SURVEY.md 8c)"},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


def run_ours(a):
    import torch
    import torch.distributed as dist

    import janus_ode_b200 as jb
    from janus_ode_b200 import configs

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torchrun --nproc-per-node N")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_bound = False
    if world > 1:
        from janus_ode_b200 import distributed
        numa_bound = distributed.bind_to_gpu_numa(local)   # before any page-locked allocation
        dist.init_process_group("nccl", device_id=dev)

    # shard: rows rank, rank+world, ... of the seeded global batch
    Mg = a.M * world
    cfg, _ = make_cfg(a, Mg)
    idx = configs.shard_rows(Mg, rank, world)
    y0_h = cfg["y0"][idx].contiguous().pin_memory()
    ts_h = cfg["tspan"][idx].contiguous().pin_memory()
    y0_d, ts_d = y0_h.to(dev), ts_h.to(dev)
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_device_step():
        r = jb.RadauTe(dyn, dyn.jacobian, ts_d, y0_d, cfg["options"])
        flush.zero_()                                   # evict L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r.solve()
        e1.record()
        return r, e0, e1

    for _ in range(max(a.warmup, 3)):
        r, _, _ = one_device_step()
    peak = jb.fp64_peak_tflops(20000) if rank == 0 else None       # FP64 DFMA probe (roofline denominator)
    dmma = jb.fp64_dmma_tflops(20000) if rank == 0 else None       # FP64 tensor-core (mma.sync f64) probe, for the MLP GEMMs
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    evs, launches = [], 0
    barrier()
    for _ in range(a.steps):
        r, e0, e1 = one_device_step()
        evs.append((e0, e1))
        launches += r.gpu_launches
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    t_dev = sum(ms) * 1e-3
    # res != 0 is a per-trajectory RESULT, not an error (the reference keeps integrating the others): config 3 with the smooth
    # control law has initial states whose costates blow up in finite time (3 of 262,144 rows end with res = 1, "step too
    # small", in the CPU oracle too).  Anything beyond a handful of rows would mean the path itself is broken.
    n_failed = int((r.res != 0).sum().item())
    assert n_failed <= max(0, int(1e-4 * r.res.numel())), f"{n_failed} trajectories failed"
    acc = int(r.stats.AccptNbr.sum().item())
    att = int(r.stats.StepNbr.sum().item())
    sol = int(r.stats.SolveNbr.sum().item())
    dec = int(r.stats.DecompNbr.sum().item())
    jac = int(r.stats.JacNbr.sum().item())
    flops = int(r.work.sum().item())

    # ---- e2e: public API with pinned host tensors, H2D + D2H inside the timed call
    ny = cfg["ham"].Ny
    h2d = a.M * (ny + 2) * 8
    gather_check = None
    if world == 1:
        # RadauTe on CPU tensors -> jb200_solve_host_compact: H2D, solve, D2H (y, t, h, res, int32 counters, nstg)
        rh = None
        for _ in range(3):
            del rh                                          # results of the previous call go back to torch's
            rh = jb.RadauTe(dyn, dyn.jacobian, ts_h, y0_h, cfg["options"])   # pinned-memory cache
            rh.solve()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            del rh
            rh = jb.RadauTe(dyn, dyn.jacobian, ts_h, y0_h, cfg["options"])
            rh.solve()
        t_e2e = time.perf_counter() - t0
        acc_e2e = int(rh.stats.AccptNbr.sum().item())
        d2h = a.M * (ny * 8 + 8 + 8 + 4 + 32 + 4)
        e2e_api = "RadauTe(dyn, dyn.jacobian, tspan_cpu_pinned, y0_cpu_pinned, opts).solve() -> jb200_solve_host_compact"
    else:
        # sharded solve (janus_ode_b200.distributed): each rank copies its shard in, solves it, the final states and
        # counters of ALL ranks are all-gathered over NCCL / NVLink (one packed all_gather_into_tensor), and each rank
        # reads its own shard's results back to page-locked host memory
        from janus_ode_b200 import distributed
        n = y0_h.shape[0]
        y_host = torch.empty(n, ny, dtype=torch.float64).pin_memory()
        r_host = torch.empty(n, dtype=torch.int32).pin_memory()
        s_host = torch.empty(n, 8, dtype=torch.int32).pin_memory()
        side = torch.cuda.Stream(dev)

        def one_e2e_step():
            yd, td = y0_h.to(dev, non_blocking=True), ts_h.to(dev, non_blocking=True)
            rr = jb.RadauTe(dyn, dyn.jacobian, td, yd, cfg["options"])
            rr.solve()
            st32 = rr.stats.raw.to(torch.int32)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                    # D2H of the own shard overlaps the NVLink gather
                y_host.copy_(rr.y, non_blocking=True)
                r_host.copy_(rr._res32, non_blocking=True)
                s_host.copy_(st32, non_blocking=True)
            g = distributed.gather_packed([rr.y, rr._res32, st32], Mg)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            return rr, g
        for _ in range(3):
            rr, g = one_e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            rr, g = one_e2e_step()
        t_e2e = time.perf_counter() - t0
        acc_e2e = int(s_host[:, 5].sum().item())
        d2h = a.M * (ny * 8 + 4 + 32)
        e2e_api = ("distributed: y0/tspan shard (pinned host) -> device, RadauTe(...).solve(), gather_packed([y, res, stats32]) = one "
                   "NCCL all_gather_into_tensor of the global result, own shard -> pinned host")
        # NCCL gather check: rank 0 re-solves a sample of GLOBAL rows alone; the gathered rows must equal it bit for bit
        if rank == 0:
            ns = min(Mg, 4096 * world)
            rs = jb.RadauTe(dyn, dyn.jacobian, cfg["tspan"][:ns].to(dev), cfg["y0"][:ns].to(dev), cfg["options"])
            rs.solve()
            gather_check = {"rows": ns, "states_bit_identical": bool(torch.equal(g[0][:ns], rs.y)),
                            "counters_identical": bool(torch.equal(g[2][:ns].to(torch.int64), rs.stats.raw)),
                            "nccl_bytes_per_step": int(Mg * (ny * 8 + 4 + 32))}

    tens = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device=dev)
    sums = torch.tensor([acc, att, sol, dec, jac, flops, acc_e2e, launches], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tens, op=dist.ReduceOp.MAX)     # time = max over ranks
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)     # work = sum over ranks
    t_dev, t_e2e = tens.tolist()
    acc, att, sol, dec, jac, flops, acc_e2e, launches = [int(v) for v in sums.tolist()]

    if rank == 0:
        value = acc * a.steps / t_dev
        kernel_s = t_dev / a.steps                       # CUDA-event time of one whole solve (all of its kernel launches)
        achieved = (flops / world) / kernel_s / 1e12     # per GPU
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
               "ms_per_step": 1e3 * t_dev / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload_name(a), "rows_total": Mg, "parallelism": f"batch-sharded x{world}",
                          "l2": "256 MiB buffer written between timed iterations",
                          "accepted_steps": acc, "attempted_steps": att, "newton_iterations": sol,
                          "lu_factorisations": dec, "jacobians": jac, "rows_res_nonzero_rank0": n_failed},
               "e2e": {"value": acc_e2e * a.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d * world,
                       "d2h_bytes_per_step": d2h * world, "ms_per_step": 1e3 * t_e2e / a.steps, "api": e2e_api,
                       "nccl_gather_check": gather_check, "numa_bound": numa_bound},
               "gpu_launches": launches,
               "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                            "frac": achieved / peak if peak else None, "traffic": _traffic(a),
                            "peak_source": "DFMA probe jb200_fp64_peak on this GPU in this run (MEASURED_PEAKS.json has no FP64 entry)",
                            "fp64_dmma_probe_tflops": dmma,
                            "algorithmic_flops_per_launch": flops / world,
                            "flops_per_accepted_step": flops / max(acc, 1),
                            "hbm_bytes_per_launch": h2d + d2h,
                            "hbm_frac_of_measured": ((h2d + d2h) / kernel_s / 1e9) / _hbm_peak()},
               "clocks": clocks}
        if not a.no_secondary and world == 1 and a.workload == "c2" and not a.fixed_order and a.M == 1 << 20:
            out["secondary"] = secondary_c4(a, dev, peak)
        if not a.no_cpu_baseline and world == 1:
            from oracle import c_oracle
            c_oracle.build()
            rows = a.cpu_sample or (65536 if a.workload != "c4" else 2048)
            rate, dt, _, cores = cpu_reference_run(a, rows)
            out["cpu_baseline_strong"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                          "sample": f"first {rows} rows of the workload, {dt:.1f} s; scalar C restatement, one "
                                                    f"trajectory per OpenMP thread (oracle/c, reference arithmetic) -- a "
                                                    f"stronger bar than the reference's masked-tensor design"}
            if live_reference_available(a):
                lrows = live_reference_rows(a)
                lrate, ldt, _, lcores, _ = live_reference_run(a, lrows)
                out["cpu_baseline"] = {"value": lrate, "unit": UNIT, "cores": lcores, "kind": "reference",
                                       "sample": LIVE_SAMPLE.format(rows=lrows, cores=lcores, dt=ldt)}
            else:
                brows = 8192 if a.workload != "c4" else 256
                brate, bdt, _, bcores = batched_reference_run(a, brows)
                out["cpu_baseline"] = {"value": brate, "unit": UNIT, "cores": bcores, "kind": "port",
                                       "sample": f"first {brows} rows of the workload, {bdt:.1f} s; masked-tensor torch restatement "
                                                 f"of RadauTe with the autograd Hamiltonian adaptor (oracle/radau_batched.py): "
                                                 f"the reference's op structure (oracle/_ref is not built on this box)"}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def secondary_c4(a, dev, peak):
    """short run of config 4 (MLP Hamiltonian, tile-GEMM stage evaluations) appended to the default line so that the
    driver's record carries it: 262,144 trajectories, two warm-ups + one timed solve, device-resident inputs."""
    import copy

    import torch

    import janus_ode_b200 as jb
    from janus_ode_b200 import configs
    M4 = 1 << 18
    c4 = configs.c4_mlp(M=1 << 20, lo=0, hi=M4)
    dyn = jb.HamiltonianSystem(c4["ham"], c4["sign"])
    y0, ts = c4["y0"].to(dev), c4["tspan"].to(dev)
    ms = None
    for _ in range(3):
        r = jb.RadauTe(dyn, dyn.jacobian, ts, y0, c4["options"])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r.solve()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    acc = int(r.stats.AccptNbr.sum().item())
    flops = int(r.work.sum().item())
    a4 = copy.copy(a)
    a4.workload, a4.M = "c4", M4
    ach = flops / (ms * 1e-3) / 1e12
    return {"c4": {"metric": METRIC, "value": acc / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": 1, "warmup": 2,
                   "config": {"workload": workload_name(a4)}, "gpu_launches": r.gpu_launches,
                   "all_rows_ok": bool((r.res == 0).all()),
                   "roofline": {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                                "frac": ach / peak if peak else None,
                                "tensor_core_path": "layer GEMMs as mma.sync.m8n8k4.f64 (DMMA); same peak as the DFMA pipe"}}}


def _traffic(a):
    """DRAM bytes of one solve from the committed ncu launch list (profiles/), only for the profiled size."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        if a.M != 1 << 20 or a.workload not in ("c2", "c4"):
            return None
        if a.workload == "c4":
            return None if a.fixed_order else t["c4_bytes_per_solve"]
        return t["fixed_bytes_per_solve"] if a.fixed_order else t["default_bytes_per_solve"]
    except Exception:
        return None


def _hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0      # fallback of /opt/skills/guides/B200_PROFILING.md


if __name__ == "__main__":
    args = parse()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
