#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled in
oracle/_ref (strict build: -O2 -ffp-contract=off).

TEST INFRASTRUCTURE.  Run in the dev container (needs /root/reference for the
build and for images/maze_5.png):

    python oracle/gen_golden.py

The reference ships no tests or golden vectors, so every
fixture here is an output of the reference's own code on a stated input.  Big
fields (1000x1000) are stored as SHA-256 of the raw little-endian bytes; small
ones in full.  All arrays are (ny, nx), index = x + y*nx.
"""
from __future__ import annotations

import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from oracle_py import Ref, build  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
REF_SRC = os.environ.get("VHP_REF_SRC", "/root/reference")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rect_map(nx, ny, nobs, seed, lo=3, hi=9):
    """Deterministic test maps (numpy PCG64; documented so tests can rebuild)."""
    g = np.random.default_rng(seed)
    occ = np.ones((ny, nx))
    for _ in range(nobs):
        x = int(g.integers(0, nx)); y = int(g.integers(0, ny))
        w = int(g.integers(lo, hi + 1)); h = int(g.integers(lo, hi + 1))
        occ[y:y + h, x:x + w] = 0
    return occ


def pack(occ):
    return np.packbits(occ.astype(np.uint8), axis=None)


def main():
    build(ref=True)
    ref = Ref("strict")
    fast = Ref("fast")
    os.makedirs(OUT, exist_ok=True)

    # ---- 1. standalone sweeps (computeVisibility) on small maps ------------
    sweep = {}
    cases = [(64, 48, 12, 3, 20, 30), (48, 64, 10, 5, 40, 5),
             (101, 101, 25, 7, 50, 50), (33, 77, 8, 9, 0, 0),
             (77, 33, 8, 11, 76, 32), (60, 60, 0, 1, 13, 47),
             (129, 65, 30, 2, 1, 63), (1, 1, 0, 0, 0, 0), (5, 1, 0, 0, 2, 0),
             (1, 7, 0, 0, 0, 3), (2, 2, 0, 0, 1, 1), (40, 40, 6, 4, 39, 0),
             (40, 40, 6, 4, 0, 39)]
    sweep["cases"] = np.array(cases, dtype=np.int32)
    for k, (nx, ny, nobs, seed, sx, sy) in enumerate(cases):
        occ = rect_map(nx, ny, nobs, seed)
        sweep[f"occ_{k}"] = occ.astype(np.uint8)
        sweep[f"vis_{k}"] = ref.compute_visibility(occ, sx, sy)
        sweep[f"ray_{k}"] = ref.raycast_all(occ, sx, sy)
    # diagonal + border KAT: 9x9, obstacle (4,3), source (2,2)
    occ = np.ones((9, 9)); occ[3, 4] = 0
    sweep["kat_diag_occ"] = occ.astype(np.uint8)
    sweep["kat_diag_vis"] = ref.compute_visibility(occ, 2, 2)
    # occupied source: everything dark
    occ = np.ones((12, 10)); occ[5, 4] = 0
    sweep["kat_occsrc_vis"] = ref.compute_visibility(occ, 4, 5)
    # no-reset semantics: stale content survives on column 0 / row 0
    init = np.arange(63, dtype=np.float64).reshape(7, 9) / 64.0
    sweep["kat_stale_init"] = init
    sweep["kat_stale_vis"] = ref.compute_visibility(np.ones((7, 9)), 4, 3, init)
    # strict vs -Ofast: a cell whose thr-0.5 decision flips under FMA contraction
    flip = None
    for seed in range(1, 40):
        occ = ref.generate_environment(101, 101, 10, 10, 20, 10, 20, seed)
        a = ref.compute_visibility(occ, 50, 50)
        b = fast.compute_visibility(occ, 50, 50)
        d = np.argwhere((a >= 0.5) != (b >= 0.5))
        if len(d):
            y, x = map(int, d[0])
            flip = (seed, x, y, a[y, x], b[y, x])
            sweep["kat_flip"] = np.array([seed, x, y], dtype=np.int32)
            sweep["kat_flip_vals"] = np.array([a[y, x], b[y, x]])
            sweep["kat_flip_vis"] = a
            break
    print("flip cell (seed,x,y,strict,fast):", flip)
    np.savez_compressed(os.path.join(OUT, "sweep.npz"), **sweep)

    # ---- 2. planner: one sweep tie-break KAT -------------------------------
    plan = {}
    occ = np.ones((101, 101))
    occ[48:53, :] = 0
    occ[48:53, 18:23] = 1
    occ[48:53, 78:83] = 1
    vg = np.zeros((101, 101)); came = np.full((101, 101), 10**15, dtype=np.uint64)
    came[10, 50] = 0
    ls = np.array([[50, 10]], dtype=np.int32)
    vis, top, h, pushes = ref.update_visibility(occ, (50, 10), (50, 90), 0.5, vg,
                                                came, ls, 0)
    plan["tie_occ"] = occ.astype(np.uint8)
    plan["tie_top"] = np.array(top, dtype=np.int32)
    plan["tie_h"] = np.array([h])
    plan["tie_pushes"] = np.array([pushes])
    plan["tie_vg"] = vg
    plan["tie_came"] = came
    print("tie KAT: top", top, "h", repr(h), "pushes", pushes)

    # ---- 3. planner: full solves on reference-generated 101^2 maps ----------
    for seed in (1, 2, 3, 4, 5, 6):
        occ = ref.generate_environment(101, 101, 10, 10, 20, 10, 20, seed)
        r = ref.solve(occ, (5, 5), (95, 95), 0.25, 100)
        plan[f"s101_{seed}_status"] = np.array([r["status"], r["nb_of_sources"]])
        plan[f"s101_{seed}_ls"] = r["light_sources"]
        plan[f"s101_{seed}_path"] = r["path"]
        plan[f"s101_{seed}_len"] = np.array([r["path_length"]])
        plan[f"s101_{seed}_vg"] = r["vg"]
        plan[f"s101_{seed}_vis"] = r["vis"]
        plan[f"s101_{seed}_came"] = r["came"]
        print("101^2 seed", seed, "status", r["status"], "nb", r["nb_of_sources"],
              "len", repr(r["path_length"]))
    # other shapes / thresholds incl. a stall to max_iter
    extra = [(160, 120, 40, 5, (3, 3), (150, 110), 0.5, 60),
             (200, 200, 60, 11, (10, 190), (190, 10), 0.2, 60),
             (120, 160, 45, 8, (110, 150), (4, 6), 0.3, 25),
             (90, 70, 20, 21, (2, 2), (85, 66), 0.0, 10),
             (90, 70, 20, 21, (2, 2), (85, 66), 1.0, 30)]
    plan["extra_cases"] = np.array(
        [(a, b, c, d, e[0], e[1], f[0], f[1], m) for a, b, c, d, e, f, g, m in extra],
        dtype=np.int32)
    plan["extra_thr"] = np.array([c[6] for c in extra])
    for k, (nx, ny, nobs, seed, st, en, thr, mi) in enumerate(extra):
        occ = rect_map(nx, ny, nobs, seed, 4, 14)
        occ[st[1], st[0]] = 1; occ[en[1], en[0]] = 1
        r = ref.solve(occ, st, en, thr, mi)
        plan[f"extra_{k}_occ"] = occ.astype(np.uint8)
        plan[f"extra_{k}_status"] = np.array([r["status"], r["nb_of_sources"]])
        plan[f"extra_{k}_ls"] = r["light_sources"]
        plan[f"extra_{k}_path"] = r["path"]
        plan[f"extra_{k}_len"] = np.array([r["path_length"]])
        plan[f"extra_{k}_vg"] = r["vg"]
        plan[f"extra_{k}_came"] = r["came"]
        print("extra", k, "status", r["status"], "nb", r["nb_of_sources"],
              "len", repr(r["path_length"]))
    np.savez_compressed(os.path.join(OUT, "planner.npz"), **plan)

    # ---- 4. shipped 1000^2 config (config/settings.config, randomSeed=0) ----
    big = {}
    for seed in (1, 2, 3, 25):
        occ = ref.generate_environment(1000, 1000, 15, 100, 200, 100, 200, seed)
        r = ref.solve(occ, (50, 50), (990, 990), 0.25, 250)
        big[f"seed{seed}_status"] = np.array([r["status"], r["nb_of_sources"]])
        big[f"seed{seed}_ls"] = r["light_sources"]
        big[f"seed{seed}_path"] = r["path"]
        big[f"seed{seed}_len"] = np.array([r["path_length"]])
        big[f"seed{seed}_sha"] = np.array(
            [sha(occ.astype(np.uint8)), sha(r["vg"]), sha(r["came"]), sha(r["vis"])])
        big[f"seed{seed}_density"] = np.array([100.0 * (occ == 0).mean()])
        if seed == 1:
            v = ref.compute_visibility(occ, 50, 50)
            big["seed1_cv_sha"] = np.array([sha(v), sha(v.astype(np.float32)),
                                            sha((v >= 0.25).astype(np.uint8))])
        print("1000^2 seed", seed, "status", r["status"], "nb", r["nb_of_sources"],
              "len", repr(r["path_length"]))
    # empty 1000^2, centre source (the benchmarkSeries case)
    v = ref.compute_visibility(np.ones((1000, 1000)), 500, 500)
    big["empty_cv_sha"] = np.array([sha(v)])
    big["empty_cv_sum"] = np.array([v.sum()])
    np.savez_compressed(os.path.join(OUT, "shipped1000.npz"), **big)

    # ---- 5. maze_5.png (BASELINE config 3) ---------------------------------
    from PIL import Image
    img = np.array(Image.open(os.path.join(REF_SRC, "images", "maze_5.png")).convert("RGBA"))
    occ = (img[:, :, 0] == 255).astype(np.float64)  # src/environment.cpp:198-209
    ny, nx = occ.shape
    # start {118,317}, end {123,10} in settings.config's bottom-left frame,
    # flipped like solve() does in mode 2 (:83-86)
    start = (118, ny - 1 - 317); end = (123, ny - 1 - 10)
    mz = {"occ_bits": pack(occ), "shape": np.array([ny, nx]),
          "start": np.array(start), "end": np.array(end)}
    for thr, tag in ((0.2, "thr020"), (0.3, "thr030"), (0.25, "thr025")):
        r = ref.solve(occ, start, end, thr, 250)
        mz[f"{tag}_status"] = np.array([r["status"], r["nb_of_sources"]])
        mz[f"{tag}_ls"] = r["light_sources"]
        mz[f"{tag}_path"] = r["path"]
        mz[f"{tag}_len"] = np.array([r["path_length"]])
        mz[f"{tag}_sha"] = np.array([sha(r["vg"]), sha(r["came"]), sha(r["vis"])])
        print("maze_5 thr", thr, "status", r["status"], "nb", r["nb_of_sources"],
              "len", repr(r["path_length"]), "points", len(r["path"]))
    np.savez_compressed(os.path.join(OUT, "maze5.npz"), **mz)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
