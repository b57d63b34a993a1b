#!/usr/bin/env python3
"""Per-kernel GPU time of ONE solve from the CUPTI activity trace (torch.profiler sees every kernel of the process, also the
library's own): python profiles/kernel_timeline.py [c2|c4] [M]  -> table on stdout.  Unlike an ncu launch list the kernels
run back to back with warm caches, so the shares are those of the real solve."""
import collections
import os
import re
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import janus_ode_b200 as jb          # noqa: E402
from janus_ode_b200 import configs   # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
    cfg = configs.c4_mlp(M=M) if wl == "c4" else configs.c2_vdp(M=M)
    dyn = jb.HamiltonianSystem(cfg["ham"], cfg["sign"])
    y0, ts = cfg["y0"].cuda(), cfg["tspan"].cuda()
    for _ in range(2):
        r = jb.RadauTe(dyn, dyn.jacobian, ts, y0, cfg["options"])
        r.solve()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        e0.record()
        r = jb.RadauTe(dyn, dyn.jacobian, ts, y0, cfg["options"])
        r.solve()
        e1.record()
        torch.cuda.synchronize()
    tot = collections.Counter()
    cnt = collections.Counter()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = re.sub(r"\(.*", "", ev.name)
            name = re.sub(r"^void ", "", name)[:70]
            tot[name] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
            cnt[name] += 1
    wall = e0.elapsed_time(e1) * 1e3
    busy = sum(tot.values())
    print(f"{wl} M={M}: solve {wall / 1e3:.2f} ms (CUDA events), kernels + copies {busy / 1e3:.2f} ms = {100 * busy / wall:.1f} % busy")
    for k, v in tot.most_common(16):
        print(f"  {k:70s} n={cnt[k]:5d}  {v / 1e3:9.3f} ms  {100 * v / wall:5.1f} % of the solve")


if __name__ == "__main__":
    main()
