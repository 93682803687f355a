#!/usr/bin/env python
"""benchmarkSeries() (reference :295-374) on this box: the 60 log-spaced empty grids 50 ... 5000,
source in the centre, ONE un-warmed call each -- the GPU library through vhp_solver_benchmark_series
and, beside it, the reference's computeVisibility() / ray-casting loop on one host core (ray
casting only up to --cpu-ray-max cells per side: it is O(N^3)).  Writes a JSON with both curves.

    python tools/benchmark_series.py [--out gpurun_out/benchmark_series.json] [--cpu-ray-max 1300]
"""
import argparse
import ctypes as C
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "benchmark_series.json"))
    ap.add_argument("--cpu-ray-max", type=int, default=1300)
    ap.add_argument("--sizes", type=int, default=60)
    args = ap.parse_args()
    import visibility_heuristic_path_planner_b200 as vhp
    from oracle_py import Ref
    lib = vhp.load_library()
    cfg = vhp.Config()
    lib.vhp_config_default(C.byref(cfg))
    cfg.ncols = cfg.nrows = 64
    cfg.start_x = cfg.start_y = 5
    cfg.silent, cfg.save_results = 1, 0
    ctx = vhp.Context(0)
    h = C.c_void_p()
    occ8 = np.ones((64, 64), np.uint8)
    lib.vhp_solver_create.argtypes = [C.c_void_p, C.POINTER(vhp.Config), C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    assert lib.vhp_solver_create(ctx.h, C.byref(cfg), occ8.ctypes.data, 64, 64, C.byref(h)) == 0
    lib.vhp_solver_benchmark_series.argtypes = [C.c_void_p, C.c_int]
    work = tempfile.mkdtemp()
    cwd = os.getcwd()
    os.chdir(work)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    os.dup2(devnull, 1)
    try:
        runs = []
        for rep in range(3):  # the reference's sample file holds 20 runs; three here
            assert lib.vhp_solver_benchmark_series(h, args.sizes) == 0
    finally:
        os.dup2(saved, 1)
        os.chdir(cwd)
    lines = [l.split() for l in open(os.path.join(work, "output", "benchmark_results.txt")).read().splitlines()]
    n = args.sizes
    gpu = []
    for k in range(n):
        rows = [lines[r * n + k] for r in range(3)]
        gpu.append({"size": int(rows[0][3].split("x")[0]),
                    "t_vis_us": [float(r[0]) for r in rows], "t_ray_us": [float(r[1]) for r in rows]})
    cpu = []
    if Ref.available("fast"):
        ref = Ref("fast")
        for g in gpu:
            s = g["size"]
            occ = np.ones((s, s))
            src = np.array([[s // 2, s // 2]], np.int32)
            rec = {"size": s, "t_vis_us": ref.time_compute_visibility(occ, src, nthreads=1) * 1e6}
            if s <= args.cpu_ray_max:
                t0 = time.perf_counter()
                ref.raycast_all(occ, s // 2, s // 2)
                rec["t_ray_us"] = (time.perf_counter() - t0) * 1e6
            cpu.append(rec)
        flags = ref.flags()
    else:
        flags = None
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"protocol": "benchmarkSeries(): empty NxN grid, source at the centre, one un-warmed computeVisibility and "
                           "one all-targets ray casting per size (reference :320-343)",
               "gpu": gpu, "cpu_reference": cpu, "cpu_flags": flags, "cpu_cores": 1}, open(args.out, "w"), indent=1)
    print("wrote", args.out, "sizes", n)


if __name__ == "__main__":
    main()
