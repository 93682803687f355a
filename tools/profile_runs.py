"""Launch list of the binary / row-runs host calls on the c2 workload (run under ncu):
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/profile_runs.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import visibility_heuristic_path_planner_b200 as vhp  # noqa: E402

n = int(os.environ.get("N", 4096))
nx = ny = 1000
occ = np.ones((1, ny, nx), np.uint8)
rng = np.random.default_rng(0)
src = np.stack([rng.integers(0, nx, n), rng.integers(0, ny, n)], 1).astype(np.int32)
c = vhp.Context(0)
for k in range(3):
    t0 = time.perf_counter(); c.visibility_batch_bin(occ, src, 0.5); t1 = time.perf_counter()
    rc, pp, tr = c.visibility_batch_runs(occ, src, 0.5, trans_cap=16 * n * ny); t2 = time.perf_counter()
    print(f"bin {1e3 * (t1 - t0):.2f} ms  runs {1e3 * (t2 - t1):.2f} ms", flush=True)
