"""C1 (101 x 101, one source): host-call and device latency; a few launches for ncu's launch list."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import visibility_heuristic_path_planner_b200 as vhp  # noqa: E402
from oracle_py import Oracle  # noqa: E402

occ = (Oracle().generate_environment(101, 101, 10, 10, 20, 10, 20, 2) != 0).astype(np.uint8)[None]
src = np.array([[5, 5]], np.int32)
ctx = vhp.Context(0)
for k in range(3):
    t0 = time.perf_counter()
    ctx.visibility_batch(occ, src, dtype=vhp.F64)
    print("host call", k, (time.perf_counter() - t0) * 1e6, "us")
t0 = time.perf_counter()
for _ in range(200):
    ctx.visibility_batch(occ, src, dtype=vhp.F64)
print("host call mean", (time.perf_counter() - t0) / 200 * 1e6, "us")
dev = torch.device("cuda", 0)
occ_t, src_t = torch.from_numpy(occ).to(dev), torch.from_numpy(src).to(dev)
out_t = torch.empty((1, 101, 101), dtype=torch.float64, device=dev)
ctx.prepare_maps_dev(occ_t)
for _ in range(5):
    ctx.visibility_batch_dev(occ_t, src_t, out_t)
ctx.synchronize()
for s in ((50, 50), (5, 5)):
    src_t.copy_(torch.tensor([s], dtype=torch.int32))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(1000):
        ctx.visibility_batch_dev(occ_t, src_t, out_t)
    ctx.synchronize()
    print("dev back-to-back, source", s, (time.perf_counter() - t0) / 1000 * 1e6, "us per sweep")
ctx.close()
