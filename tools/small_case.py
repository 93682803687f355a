"""Tiny sweep + raycast invocation (used under compute-sanitizer on the GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import visibility_heuristic_path_planner_b200 as vhp
from oracle_py import Oracle

o = Oracle()
ctx = vhp.Context(0)
rng = np.random.default_rng(0)
for nx, ny in [(70, 50), (257, 131), (300, 300)]:
    occ = (rng.random((ny, nx)) > 0.05).astype(np.uint8)
    srcs = [(0, 0), (nx - 1, ny - 1), (nx // 2, ny // 3), (1, ny - 2)]
    for dt in (vhp.F64, vhp.F32):
        out = ctx.visibility_batch(occ, srcs, dtype=dt)
        ref = np.stack([o.compute_visibility(occ, *s) for s in srcs])
        if dt == vhp.F32:
            ref = ref.astype(np.float32)
        print(nx, ny, "f64" if dt else "f32", "sweep equal:", np.array_equal(out, ref),
              "mismatches:", int((out != ref).sum()))
    ray = ctx.raycast_batch(occ, srcs[:2])
    print(nx, ny, "ray equal:", all(np.array_equal(r, o.raycast_all(occ, *s)) for r, s in zip(ray, srcs)))
print("launches", ctx.launches)
