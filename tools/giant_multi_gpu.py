#!/usr/bin/env python
"""Strip-partitioned planner on N GPUs (one rank per GPU, NCCL): checks the result against
the single-GPU planner kernel and reports timings.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/giant_multi_gpu.py [--size 2048] [--queries 3]

BASELINE configs[4] shape: --size 8192 (dense random obstacles, multi-source planner)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--queries", type=int, default=3)
    ap.add_argument("--thr", type=float, default=0.3)
    ap.add_argument("--max-iter", type=int, default=60)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import visibility_heuristic_path_planner_b200 as vhp
    from visibility_heuristic_path_planner_b200.giant import StripPlanner
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.size
    g = np.random.default_rng(8192)
    occ = np.ones((n, n), dtype=np.uint8)
    for _ in range(int(6000 * (n / 8192) ** 2)):
        x, y = int(g.integers(1, n)), int(g.integers(1, n))
        w, h = int(g.integers(8, 65)), int(g.integers(8, 65))
        occ[y:y + h, x:x + w] = 0
    free = np.argwhere(occ != 0)
    picks = free[g.integers(0, len(free), 2 * args.queries)]
    queries = [((int(a[1]), int(a[0])), (int(b[1]), int(b[0]))) for a, b in zip(picks[::2], picks[1::2])]
    sp = StripPlanner(occ, world, device=local, dist=dist if world > 1 else None)
    ref_ctx = vhp.Context(local) if rank == 0 else None
    out = []
    for start, end in queries:
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        r = sp.solve(start, end, args.thr, args.max_iter)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        vg = sp.gather_field("vg")
        came = sp.gather_field("came")
        rec = dict(start=start, end=end, status=r["status"], nb=r["nb_of_sources"],
                   path_length=r["path_length"], ms=dt * 1e3, halo_bytes=sp.halo_bytes)
        if rank == 0:
            t0 = time.perf_counter()
            ref = ref_ctx.planner_batch(occ, [start + end], threshold=args.thr, max_iter=args.max_iter)
            rec["single_gpu_ms"] = (time.perf_counter() - t0) * 1e3
            nb = int(ref["nb_sources"][0])
            same = (int(ref["status"][0]) == r["status"] and nb == r["nb_of_sources"]
                    and np.array_equal(ref["light_sources"][0][: nb + 1], r["light_sources"])
                    and float(ref["path_len"][0]) == r["path_length"]
                    and np.array_equal(ref["vg"][0], vg) and np.array_equal(ref["came"][0], came))
            rec["equal_single_gpu"] = bool(same)
            out.append(rec)
    if rank == 0:
        print(json.dumps(dict(n=n, world=world, queries=out)), flush=True)
        assert all(q["equal_single_gpu"] for q in out)
    sp.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
