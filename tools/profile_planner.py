"""One planner batch on the device for ncu: python tools/profile_planner.py [nprob]
(592 problems = two waves of 2 CTAs x 148 SMs)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import visibility_heuristic_path_planner_b200 as vhp  # noqa: E402
from bench import planner_workload  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
maps, se, pmap = planner_workload(0)
se, pmap = np.ascontiguousarray(se[:n]), np.ascontiguousarray(pmap[:n])
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev)
ctx = vhp.torch_context(0, stream)
cap = 102
occ_t, se_t, pm_t = (torch.from_numpy(a).to(dev) for a in (maps, se, pmap))
o = [torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.int32, device=dev),
     torch.zeros((n, cap, 2), dtype=torch.int32, device=dev), torch.zeros(n, dtype=torch.float64, device=dev),
     torch.zeros(n, dtype=torch.int32, device=dev), torch.zeros((n, cap, 2), dtype=torch.int32, device=dev)]
po = vhp.PlannerOut(*[t.data_ptr() for t in o], None, None, None)
ctx.prepare_maps_dev(occ_t)
nmaps, ny, nx = maps.shape
with torch.cuda.stream(stream):
    for _ in range(2):
        st = ctx.lib.vhp_planner_batch_dev(ctx.h, occ_t.data_ptr(), nmaps, nx, ny, se_t.data_ptr(), pm_t.data_ptr(),
                                           n, 0.5, 100, cap, vhp.F64, C.byref(po))
        assert st == 0
stream.synchronize()
print("problems", n, "light sources", int(o[1].sum().item()))
