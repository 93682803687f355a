"""Time the planner kernel alone (CUDA events via torch) on the bench's planner batch and on maze-like single problems."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import visibility_heuristic_path_planner_b200 as vhp
from bench import planner_workload
maps, se, pmap = planner_workload(0)
ctx = vhp.Context(0)
for fields in (False, True):
    ctx.planner_batch(maps, se, prob_map=pmap, threshold=0.5, max_iter=100, fields=fields)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = ctx.planner_batch(maps, se, prob_map=pmap, threshold=0.5, max_iter=100, fields=fields)
    torch.cuda.synchronize()
    print("fields", fields, "host call ms", (time.perf_counter() - t0) * 1e3, "sweeps", int(r["nb_sources"].sum()))
# one problem at a time
t0 = time.perf_counter()
r1 = ctx.planner_batch(maps[:1], se[:1], threshold=0.5, max_iter=100, fields=False)
torch.cuda.synchronize()
print("single problem host call ms", (time.perf_counter() - t0) * 1e3, int(r1["nb_sources"][0]))
ctx.close()
