"""N > 1 host logic on CPU: world_size-2 gloo processes shard the seeded batch (i mod G), solve
their shard with the CPU oracle standing in for the GPU solver, and all_gather the results;
the gathered global result must equal the single-process result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _OracleSolver:
    """same surface as janus_ode_b200.RadauTe, computed by the C oracle (mirror arithmetic)."""

    def __init__(self, cfg, ts, y0):
        self.cfg, self.ts, self.y0 = cfg, ts, y0

    def solve(self):
        from janus_ode_b200.options import StatsTe
        from oracle import c_oracle
        o = c_oracle.solve("vdp", list(self.cfg["ham"].params), self.cfg["sign"], self.y0.numpy(), self.ts.numpy(),
                           self.cfg["options"], nthreads=1, arith="mirror")
        self.y = torch.from_numpy(o["y"])
        self.t = torch.from_numpy(o["t"])
        self.stats = StatsTe.from_tensor(torch.from_numpy(o["stats"]))
        return torch.from_numpy(o["res"].astype(np.float64))


def _worker(rank, world, port, M, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from janus_ode_b200 import configs, distributed
    cfg = configs.c2_vdp(M=M, fixed_order=True)
    cfg["tspan"][:, 1] = 200.0
    y, t, r, st = distributed.solve_sharded(lambda ts, y0: _OracleSolver(cfg, ts, y0), cfg["y0"], cfg["tspan"])
    idx = distributed.shard_rows(M, rank, world)
    assert idx.tolist() == list(range(rank, M, world))
    if rank == 0:
        torch.save(dict(y=y, t=t, r=r, st=st), out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("M", [64, 37])      # even and ragged shards
def test_sharded_solve_equals_single_process(tmp_path, M):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, port, M, out), nprocs=2, join=True)
    g = torch.load(out)
    from janus_ode_b200 import configs
    from oracle import c_oracle
    cfg = configs.c2_vdp(M=M, fixed_order=True)
    cfg["tspan"][:, 1] = 200.0
    o = c_oracle.solve("vdp", [1000.0], 1, cfg["y0"].numpy(), cfg["tspan"].numpy(), cfg["options"], arith="mirror")
    assert np.array_equal(g["y"].numpy(), o["y"]) and np.array_equal(g["st"].numpy(), o["stats"])
    assert np.array_equal(g["t"].numpy(), o["t"]) and (g["r"] == 0).all()
