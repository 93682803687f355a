"""Quick GPU check of the octant kernel against the oracle on small cases."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["VHP_SWEEP_IMPL"] = "octant"
import visibility_heuristic_path_planner_b200 as vhp
from oracle_py import Oracle
from conftest import rect_map
ora = Oracle()
ctx = vhp.Context(0)
print("ratio selftest", ctx.selftest_ratio(4096))
nbad = 0
for (nx, ny) in [(9, 9), (16, 12), (40, 36), (64, 64), (130, 61), (200, 256), (253, 253), (256, 256), (300, 200), (500, 400), (1000, 1000)]:
    occ = rect_map(nx, ny, max(1, (nx * ny) // 400), nx + ny, 1, 9)
    g = np.random.default_rng(nx * 7 + ny)
    srcs = [(0, 0), (nx - 1, 0), (0, ny - 1), (nx - 1, ny - 1), (nx // 2, ny // 2)]
    srcs += [(int(g.integers(0, nx)), int(g.integers(0, ny))) for _ in range(5 if nx < 1000 else 2)]
    for dt, name in ((vhp.F64, "f64"), (vhp.F32, "f32")):
        out = ctx.visibility_batch(occ, srcs, dtype=dt)
        for s, o in zip(srcs, out):
            ref = ora.compute_visibility(occ, *s)
            if dt == vhp.F32:
                ref = ref.astype(np.float32)
            if not np.array_equal(o, ref):
                bad = np.argwhere(o != ref)
                nbad += 1
                print(f"MISMATCH {nx}x{ny} {name} src={s}: {len(bad)} cells, first (y,x)={bad[:6].tolist()} got={[float(o[tuple(b)]) for b in bad[:3]]} want={[float(ref[tuple(b)]) for b in bad[:3]]}")
    print(f"{nx}x{ny} done", flush=True)
print("mismatching cases:", nbad)
