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


def gather_packed(tensors, M: int, group=None):
    """ONE collective for the whole result set: every rank packs the rows of its tensors ([n, ...] each, any dtypes)
    into a byte matrix [n_max, B], `all_gather_into_tensor` moves it (NCCL all-gather over NVLink on GPUs), and the
    interleaved shard order (row i lives on rank i mod G at position i // G) is undone by one strided view."""
    world = dist.get_world_size(group)
    n = tensors[0].shape[0]
    n_max = (M + world - 1) // world
    cols = [t.contiguous().view(torch.uint8).reshape(n, -1) if n else
            torch.empty(0, t[0:1].contiguous().view(torch.uint8).numel() if t.numel() else 0, dtype=torch.uint8, device=t.device)
            for t in tensors]
    widths = [int(t.element_size() * (t.numel() // max(n, 1))) for t in tensors]
    B = sum(widths)
    pad = torch.zeros(n_max, B, dtype=torch.uint8, device=tensors[0].device)
    if n:
        pad[:n] = torch.cat(cols, 1)
    out = torch.empty(world * n_max, B, dtype=torch.uint8, device=pad.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    glob = out.view(world, n_max, B).transpose(0, 1).reshape(world * n_max, B)[:M].contiguous()
    res, off = [], 0
    for t, w in zip(tensors, widths):
        res.append(glob[:, off:off + w].contiguous().view(t.dtype).reshape((M,) + tuple(t.shape[1:])))
        off += w
    return res


def gpu_numa_cpus(index: int):
    """host CPUs of the NUMA node the GPU hangs off (sysfs), or None when that cannot be determined."""
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return cpus or None
    except Exception:
        return None


def bind_to_gpu_numa(index: int) -> bool:
    """Pin this process to the CPUs next to GPU `index` BEFORE it allocates page-locked buffers: with one rank per GPU
    and no binding every rank's staging memory lands on NUMA node 0 and the host<->device copies of the far GPUs
    cross the socket interconnect (measured round 1: 35 -> 7 GB/s per rank at 8 ranks)."""
    import os
    cpus = gpu_numa_cpus(index)
    if not cpus:
        return False
    try:
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return True
    except Exception:
        pass
    return False


def solve_sharded(make_solver, y0: torch.Tensor, tspan: torch.Tensor, group=None):
    """y0/tspan are the GLOBAL batch (identical on every rank, e.g. seeded).  `make_solver(ts, y)` builds
    a solver object with .solve() and fields y, t, stats.raw, res (janus_ode_b200.RadauTe on GPUs).
    Returns the gathered global (y, t, res, stats)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    M = y0.shape[0]
    s = make_solver(shard(tspan, rank, world), shard(y0, rank, world))
    res = s.solve()
    y, t, r, st = gather_packed([s.y, s.t, res, s.stats.raw], M, group)   # one all-gather for the whole result set
    return y, t, r, st
