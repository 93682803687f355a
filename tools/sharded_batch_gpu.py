#!/usr/bin/env python
"""Sharded batch on N GPUs == the same batch on one GPU (SURVEY 4 item iv, on hardware): every rank
runs its contiguous block of a (map, source) batch and of a planner batch through the single-GPU
entry points (sharding.py: no data-path collective), the per-item SHA-256 of the results are gathered
over gloo, and rank 0 compares them with the whole batch computed on its own GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29551 tools/sharded_batch_gpu.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def digests(arrs):
    """One digest per item over the item's slice of every array."""
    n = len(arrs[0])
    out = []
    for i in range(n):
        h = hashlib.sha256()
        for a in arrs:
            h.update(np.ascontiguousarray(a[i]).tobytes())
        out.append(h.hexdigest()[:16])
    return np.array(out)


def main():
    import torch
    import torch.distributed as dist
    import visibility_heuristic_path_planner_b200 as vhp
    from visibility_heuristic_path_planner_b200.sharding import gather_blocks, shard_batch
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")
    g = np.random.default_rng(99)  # the same global batch on every rank
    nmaps, ny, nx, per = 24, 160, 200, 12
    maps = np.ones((nmaps, ny, nx), np.uint8)
    for m in range(nmaps):
        for _ in range(10):
            x, y = int(g.integers(1, nx)), int(g.integers(1, ny))
            maps[m, y:y + int(g.integers(6, 30)), x:x + int(g.integers(6, 30))] = 0
    smap = np.repeat(np.arange(nmaps, dtype=np.int32), per)
    n = len(smap)
    pts = []
    for m in smap:
        free = np.argwhere(maps[m] != 0)
        a, b = free[g.integers(0, len(free), 2)]
        pts.append((a[1], a[0], b[1], b[0]))
    se = np.array(pts, np.int32)
    src = np.ascontiguousarray(se[:, :2])
    ctx = vhp.Context(local)

    def run(maps_, src_, se_, smap_):
        vis = ctx.visibility_batch(maps_, src_, src_map=smap_, dtype=vhp.F64)
        bits = ctx.visibility_batch_bin(maps_, src_, 0.5, src_map=smap_)
        pl = ctx.planner_batch(maps_, se_, prob_map=smap_, threshold=0.3, max_iter=30)
        return digests([vis, bits]), digests([pl[k] for k in ("status", "nb_sources", "light_sources", "path_len",
                                                               "path_n", "path", "vg", "came", "vis")])

    lm, ls, lmap, (lo, hi) = shard_batch(maps, src, smap, rank, world)
    _, lse, _, _ = shard_batch(maps, se, smap, rank, world)
    d_sweep, d_plan = run(lm, ls, lse, lmap)
    all_sweep = gather_blocks(d_sweep, n, dist if world > 1 else None)
    all_plan = gather_blocks(d_plan, n, dist if world > 1 else None)
    if rank == 0:
        w_sweep, w_plan = run(maps, src, se, smap)
        print(json.dumps({"world": world, "items": n, "sweeps_equal": bool(np.array_equal(all_sweep, w_sweep)),
                          "planner_equal": bool(np.array_equal(all_plan, w_plan)),
                          "block_of_rank0": [lo, hi]}))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
