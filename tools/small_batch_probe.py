"""Small-map batches at the sizes where the launcher picks 4-warp and 1-warp CTAs (A/B of launch bounds):
   VHP_LIB_VARIANT=<name> python tools/small_batch_probe.py"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import visibility_heuristic_path_planner_b200 as vhp


def rect_map(nx, ny, nobs, seed, lo, hi):
    g = np.random.default_rng(seed)
    m = np.ones((ny, nx), np.uint8)
    for _ in range(nobs):
        x, y = int(g.integers(1, nx)), int(g.integers(1, ny))
        m[y:y + int(g.integers(lo, hi)), x:x + int(g.integers(lo, hi))] = 0
    return m


dev = torch.device("cuda", 0)
c = vhp.Context(0)
for nx, n in ((101, 600), (101, 2000), (200, 1000), (400, 600), (101, 40000), (64, 40000)):
    occ = np.stack([rect_map(nx, nx, 8, k, 3, 14) for k in range(8)])
    rng = np.random.default_rng(1)
    src = np.stack([rng.integers(0, nx, n), rng.integers(0, nx, n)], 1).astype(np.int32)
    smap = rng.integers(0, 8, n).astype(np.int32)
    occ_t, src_t, smap_t = torch.from_numpy(occ).to(dev), torch.from_numpy(src).to(dev), torch.from_numpy(smap).to(dev)
    out = torch.empty((n, nx, nx), dtype=torch.float32, device=dev)
    c.prepare_maps_dev(occ_t)
    for _ in range(3):
        c.visibility_batch_dev(occ_t, src_t, out, smap_t)
    c.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(20):
        c.visibility_batch_dev(occ_t, src_t, out, smap_t)
    c.synchronize()
    t = (time.perf_counter() - t0) / 20
    print(os.environ.get("VHP_LIB_VARIANT", "base"), nx, n, f"{t*1e6:.1f} us  {n*nx*nx/t/1e9:.1f} Gcells/s", flush=True)
