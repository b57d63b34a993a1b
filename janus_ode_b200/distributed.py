"""Batch sharding over the GPUs of one node.

Trajectories are independent (every reference op is row-wise over M; the only cross-sample
reductions are any()-style loop guards), so the global batch is split with NO data-path
collective: rank r owns rows r, r+G, r+2G, ... (interleaved, so monotone initial conditions do
not create per-GPU load skew).  Collectives are used once, after the solve, to gather final
states / counters to every rank (all_gather) -- NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_rows(M: int, rank: int, world: int) -> torch.Tensor:
    return torch.arange(rank, M, world)


def shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    return t[rank::world].contiguous()


def gather_rows(local: torch.Tensor, M: int, group=None) -> torch.Tensor:
    """inverse of shard(): all_gather the per-rank rows and re-interleave them into [M, ...]."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    n_max = (M + world - 1) // world
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = torch.empty((M,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        n_r = len(range(r, M, world))
        out[r::world] = parts[r][:n_r]
    del rank
    return out


def solve_sharded(make_solver, y0: torch.Tensor, tspan: torch.Tensor, group=None):
    """y0/tspan are the GLOBAL batch (identical on every rank, e.g. seeded).  `make_solver(ts, y)` builds
    a solver object with .solve() and fields y, t, stats.raw, res (janus_ode_b200.RadauTe on GPUs).
    Returns the gathered global (y, t, res, stats)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    M = y0.shape[0]
    s = make_solver(shard(tspan, rank, world), shard(y0, rank, world))
    res = s.solve()
    y = gather_rows(s.y, M, group)
    t = gather_rows(s.t, M, group)
    r = gather_rows(res, M, group)
    st = gather_rows(s.stats.raw, M, group)
    return y, t, r, st
