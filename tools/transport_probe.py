#!/usr/bin/env python
"""Timing breakdown of the packed result transport (VHP_TRANSPORT_TRACE) on a bench workload.

    python tools/transport_probe.py c2s [pairs] [gpu_share] [host_threads]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["VHP_TRANSPORT_TRACE"] = "1"
if len(sys.argv) > 4:
    os.environ["VHP_HOST_THREADS"] = sys.argv[4]

import numpy as np
import torch

import visibility_heuristic_path_planner_b200 as vhp
from bench import workload

wl = sys.argv[1] if len(sys.argv) > 1 else "c2s"
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
share = int(sys.argv[3]) if len(sys.argv) > 3 else 0
maps, src, smap, desc = workload(wl, 0)
src = np.ascontiguousarray(src[:pairs])
n, (ny, nx) = len(src), maps.shape[1:]
out = torch.empty((n, ny, nx), dtype=torch.float32, pin_memory=True)
ctx = vhp.Context(0)
ctx.set_result_gpu_share(share)
for rep in range(3):
    t0 = time.perf_counter()
    ctx.visibility_batch(maps, src, dtype=vhp.F32, out=out.numpy())
    dt = time.perf_counter() - t0
    print(f"{wl} share {share}: {dt * 1e3:.1f} ms, {n * nx * ny / dt / 1e9:.1f} Gcells/s", flush=True)
