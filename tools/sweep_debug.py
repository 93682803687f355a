"""Compare the tile kernel with the CPU oracle on a few small cases and print where they differ."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import visibility_heuristic_path_planner_b200 as vhp
from oracle_py import Oracle
from conftest import rect_map
ora = Oracle(); ctx = vhp.Context(0)
cases = [(3, 5, 0), (40, 517, 3), (96, 80, 10), (200, 96, 14)]
for nx, ny, nobs in cases:
    occ = rect_map(nx, ny, nobs, nx + ny, 1, 9)
    srcs = [(0, 0), (nx - 1, ny - 1), (nx // 2, ny // 2), (nx - 1, 0), (0, ny - 1)]
    out = ctx.visibility_batch(occ, srcs, dtype=vhp.F64)
    for s, o in zip(srcs, out):
        ref = ora.compute_visibility(occ, *s)
        bad = np.argwhere(o != ref)
        if len(bad):
            ys, xs = bad[:, 0], bad[:, 1]
            print(f"{nx}x{ny} src {s}: {len(bad)} bad; x in [{xs.min()},{xs.max()}], y in [{ys.min()},{ys.max()}]; first {bad[:4].tolist()}",
                  [(float(o[y, x]), float(ref[y, x])) for y, x in bad[:3]])
        else:
            print(f"{nx}x{ny} src {s}: ok")
