"""Time ONE planner problem per map size on both routes: the persistent single-CTA planner
kernel (grid_sweep 0) and the many-CTA grid route (grid_sweep 2); host-buffer call without
field export, so the figure is the solve itself.  Prints ms per solve and per iteration."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import visibility_heuristic_path_planner_b200 as vhp

sizes = [int(a) for a in sys.argv[1:]] or [256, 512, 1000, 2048, 4096]
for n in sizes:
    g = np.random.default_rng(n)
    occ = np.ones((n, n), np.uint8)
    for _ in range(max(4, int(15 * (n / 1000) ** 2))):
        x, y = int(g.integers(1, n)), int(g.integers(1, n))
        w, h = int(g.integers(n // 10, n // 5 + 1)), int(g.integers(n // 10, n // 5 + 1))
        occ[y:y + h, x:x + w] = 0
    free = np.argwhere(occ != 0)
    a, b = free[len(free) // 50], free[-len(free) // 50]
    se = np.array([[a[1], a[0], b[1], b[0]]], np.int32)
    line = f"{n:5d}^2:"
    res = []
    for mode in (0, 2):
        ctx = vhp.Context(0)
        ctx.set_grid_sweep(mode)
        ctx.planner_batch(occ, se, threshold=0.25, max_iter=250, fields=False)
        ctx.synchronize()
        t0 = time.perf_counter()
        r = ctx.planner_batch(occ, se, threshold=0.25, max_iter=250, fields=False)
        ctx.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        nb = int(r["nb_sources"][0])
        res.append((int(r["status"][0]), nb, float(r["path_len"][0])))
        line += f"  {'cta' if mode == 0 else 'grid'} {ms:9.2f} ms ({ms / max(nb, 1):8.3f} ms/iter, nb {nb})"
        ctx.close()
    assert res[0] == res[1], res
    print(line, flush=True)
