#!/usr/bin/env python
"""Generate tests/golden/queue.npz from the UNMODIFIED reference compiled in oracle/_ref:
outputs of computeVisibilityUsingQueue() (src/visibilityBasedSolver.cpp:701-893) on a family
of small random maps.  TEST INFRASTRUCTURE; run in the dev container:

    python oracle/gen_golden_queue.py

The reference's function is a FIFO search whose result can depend on the queue order (see
oracle/vhp_oracle.h, vhp_oracle_visibility_cutoff).  The fixture therefore keeps, for every case
of the family, the reference's field and a flag telling whether the order-free rule reproduces
it bit for bit ("agree"), so that the tests pin the rule to the reference where they coincide
and document how often they do not.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from gen_golden import rect_map  # noqa: E402
from oracle_py import Oracle, Ref, build  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def family(n):
    """(nx, ny, nobs, seed, sx, sy) of case t: PCG64 seed 0 stream, as the tests rebuild it."""
    rng = np.random.default_rng(0)
    for t in range(n):
        nx, ny = int(rng.integers(8, 70)), int(rng.integers(8, 70))
        nobs = int(rng.integers(0, 25))
        sx, sy = int(rng.integers(0, nx)), int(rng.integers(0, ny))
        yield t, nx, ny, nobs, sx, sy


def main():
    build(ref=True)
    ref, ora = Ref("strict"), Oracle()
    out = {}
    agree, stats = [], []
    for t, nx, ny, nobs, sx, sy in family(400):
        occ = rect_map(nx, ny, nobs, t, 1, 9)
        a = ref.compute_visibility_queue(occ, sx, sy)
        b = ora.visibility_cutoff(occ, sx, sy)
        same = bool(np.array_equal(a, b))
        agree.append(same)
        stats.append([t, nx, ny, nobs, sx, sy, int((a != b).sum()), int(occ[sy, sx] != 0)])
        if t < 40:  # full fields of the first 40 cases
            out[f"ref_{t}"] = a
    out["cases"] = np.array(stats, dtype=np.int64)
    out["agree"] = np.array(agree)
    np.savez_compressed(os.path.join(OUT, "queue.npz"), **out)
    print("cases", len(agree), "order-free rule == reference BFS:", int(sum(agree)))


if __name__ == "__main__":
    main()
