"""One batch of bit-writing sweeps (vhp_visibility_batch_bin_dev) on a bench workload, for ncu:
   ncu --set full --clock-control none --import-source on -k regex:sweep_tile_kernel -c 1 -f -o out \
       python tools/profile_bits.py c2d 1184"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import visibility_heuristic_path_planner_b200 as vhp  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
maps, src, smap, desc = bench.workload(wl, 0, 1)
src = src[:n]
smap = None if smap is None else smap[:n]
dev = torch.device("cuda", 0)
ny, nx = maps.shape[1:]
c = vhp.Context(0)
occ_t, src_t = torch.from_numpy(maps).to(dev), torch.from_numpy(src).to(dev)
smap_t = None if smap is None else torch.from_numpy(smap).to(dev)
out = torch.empty((len(src), ny, (nx + 31) // 32), dtype=torch.int32, device=dev)
for _ in range(4):
    c.visibility_batch_bin_dev(occ_t, src_t, 0.5, out, smap_t)
c.synchronize()
print(desc, len(src), int(out.ne(0).sum()))
