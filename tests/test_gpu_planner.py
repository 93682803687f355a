"""K2/K3/K5 (device-side planner loop) through the C-ABI against the oracle and the
golden vectors generated from the reference: parents, light sources, path and
thresholded visibility bit-exact; fp64 fields bit-exact; path length exact."""
import hashlib

import numpy as np
import pytest

from conftest import load_golden, rect_map

pytestmark = pytest.mark.gpu
NO_PARENT_U64 = 10**15


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def vhp():
    import visibility_heuristic_path_planner_b200 as m
    return m


@pytest.fixture(scope="module", params=["cta", "first", "grid"])
def ctx(vhp, request):
    """Every test runs on all planner routes: one persistent CTA per problem (batches); the same
    with the first sweep of every problem done batch-wide before it (what batches of many problems
    take by default; forced here for every size); and one problem at a time on the whole GPU
    (grid-mode sweep + strip epilogue + control kernels, what large single problems take by default;
    forced here for every size)."""
    import os
    if request.param == "first":
        os.environ["VHP_PLANNER_FIRST"] = "2"
        os.environ["VHP_PLANNER_ROUNDS"] = "3"
    c = vhp.Context(0)
    os.environ.pop("VHP_PLANNER_FIRST", None)
    os.environ.pop("VHP_PLANNER_ROUNDS", None)
    c.set_grid_sweep(2 if request.param == "grid" else 0)
    yield c
    c.close()


def came_u64(came_i32):
    out = came_i32.astype(np.int64)
    out[out < 0] = NO_PARENT_U64
    return out.astype(np.uint64)


def check(r, k, ref, full=True):
    """r: batch result dict, k: problem index, ref: oracle/golden dict."""
    st, nb = int(r["status"][k]), int(r["nb_sources"][k])
    assert (st, nb) == (ref["status"], ref["nb_of_sources"]), (st, nb, ref["status"], ref["nb_of_sources"])
    if st in (0, 5):
        assert np.array_equal(r["light_sources"][k][: nb + 1], ref["light_sources"])
    n = int(r["path_n"][k])
    assert np.array_equal(r["path"][k][:n], ref["path"])
    assert r["path_len"][k] == ref["path_length"]
    if full:
        assert np.array_equal(r["vg"][k], ref["vg"])
        assert np.array_equal(came_u64(r["came"][k]), ref["came"])
        assert np.array_equal(r["vis"][k], ref["vis"])


def test_solve_101_goldens(ctx, vhp, oracle):
    g = load_golden("planner.npz")
    maps = np.stack([oracle.generate_environment(101, 101, 10, 10, 20, 10, 20, s) for s in range(1, 7)])
    se = np.tile(np.array([5, 5, 95, 95], np.int32), (6, 1))
    r = ctx.planner_batch(maps, se, prob_map=np.arange(6), threshold=0.25, max_iter=100)
    for k, seed in enumerate(range(1, 7)):
        ref = dict(status=int(g[f"s101_{seed}_status"][0]), nb_of_sources=int(g[f"s101_{seed}_status"][1]),
                   light_sources=g[f"s101_{seed}_ls"], path=g[f"s101_{seed}_path"],
                   path_length=g[f"s101_{seed}_len"][0], vg=g[f"s101_{seed}_vg"],
                   came=g[f"s101_{seed}_came"], vis=g[f"s101_{seed}_vis"])
        check(r, k, ref)
    assert int(r["status"][0]) == 4  # seed 1: "End point is not valid (occupied)"


def test_extra_shapes_and_thresholds(ctx, vhp, oracle):
    g = load_golden("planner.npz")
    for k, c in enumerate(g["extra_cases"]):
        occ = g[f"extra_{k}_occ"]
        thr, mi = float(g["extra_thr"][k]), int(c[8])
        r = ctx.planner_batch(occ, [(c[4], c[5], c[6], c[7])], threshold=thr, max_iter=mi)
        ref = oracle.solve(occ.astype(np.float64), (c[4], c[5]), (c[6], c[7]), thr, mi)
        check(r, 0, ref)
        assert int(r["status"][0]) == int(g[f"extra_{k}_status"][0])
        assert r["path_len"][0] == g[f"extra_{k}_len"][0]


def test_tie_break_first_pushed(ctx, vhp):
    """Mirror-symmetric map: two cells tie on h; the heap returns the first pushed."""
    g = load_golden("planner.npz")
    r = ctx.planner_batch(g["tie_occ"], [(50, 10, 50, 90)], threshold=0.5, max_iter=0)
    # max_iter = 0: one sweep, then nb = 1 > max_iter -> MAX_ITER, lightSources_[1] = heap top
    assert int(r["status"][0]) == 5 and int(r["nb_sources"][0]) == 1
    assert tuple(r["light_sources"][0][1]) == tuple(g["tie_top"]) == (82, 54)
    assert np.array_equal(r["vg"][0], g["tie_vg"])
    assert np.array_equal(came_u64(r["came"][0]), g["tie_came"])


def test_validation_statuses(ctx, vhp):
    occ = np.ones((20, 30)); occ[5, 5] = 0
    se = [(-1, 0, 3, 3), (30, 0, 3, 3), (3, 3, 0, 20), (5, 5, 3, 3), (3, 3, 5, 5), (3, 3, 3, 3)]
    r = ctx.planner_batch(occ, se, threshold=0.5, max_iter=10)
    assert list(r["status"]) == [1, 1, 2, 3, 4, 0]
    # start == end: one sweep, path = [start, start]? the reference walks cameFrom once
    assert int(r["nb_sources"][5]) == 1


def test_maze5(ctx, vhp):
    """BASELINE config 3: images/maze_5.png, start {118,317}, end {123,10}."""
    g = load_golden("maze5.npz")
    ny, nx = g["shape"]
    occ = np.unpackbits(g["occ_bits"])[: ny * nx].reshape(ny, nx)
    se = [tuple(g["start"]) + tuple(g["end"])]
    for tag, thr in (("thr020", 0.2), ("thr030", 0.3), ("thr025", 0.25)):
        r = ctx.planner_batch(occ, se, threshold=thr, max_iter=250)
        st, nb = g[f"{tag}_status"]
        assert (int(r["status"][0]), int(r["nb_sources"][0])) == (st, nb)
        assert np.array_equal(r["light_sources"][0][: nb + 1], g[f"{tag}_ls"])
        n = int(r["path_n"][0])
        assert np.array_equal(r["path"][0][:n], g[f"{tag}_path"])
        assert r["path_len"][0] == g[f"{tag}_len"][0]
        assert [sha(r["vg"][0]), sha(came_u64(r["came"][0])), sha(r["vis"][0])] == list(g[f"{tag}_sha"])
    assert g["thr020_len"][0] == 1341.7118586874171


def test_shipped_1000_config(ctx, vhp, oracle):
    g = load_golden("shipped1000.npz")
    seeds = (1, 2, 3, 25)
    maps = np.stack([oracle.generate_environment(1000, 1000, 15, 100, 200, 100, 200, s) for s in seeds])
    se = np.tile(np.array([50, 50, 990, 990], np.int32), (len(seeds), 1))
    r = ctx.planner_batch(maps, se, prob_map=np.arange(len(seeds)), threshold=0.25, max_iter=250)
    for k, seed in enumerate(seeds):
        st, nb = g[f"seed{seed}_status"]
        assert (int(r["status"][k]), int(r["nb_sources"][k])) == (st, nb)
        if st == 0:
            assert np.array_equal(r["light_sources"][k][: nb + 1], g[f"seed{seed}_ls"])
            assert np.array_equal(r["path"][k][: int(r["path_n"][k])], g[f"seed{seed}_path"])
        assert r["path_len"][k] == g[f"seed{seed}_len"][0]
        s = g[f"seed{seed}_sha"]
        assert [sha(r["vg"][k]), sha(came_u64(r["came"][k])), sha(r["vis"][k])] == list(s[1:])


def test_random_problems_vs_oracle(ctx, vhp, oracle):
    g = np.random.default_rng(77)
    for trial in range(6):
        nx, ny = int(g.integers(20, 200)), int(g.integers(20, 200))
        occ = rect_map(nx, ny, int(g.integers(0, 30)), 500 + trial, 2, 14)
        se = np.stack([g.integers(0, nx, 8), g.integers(0, ny, 8), g.integers(0, nx, 8),
                       g.integers(0, ny, 8)], axis=1).astype(np.int32)
        thr = float(g.choice([0.0, 0.2, 0.5, 0.9, 1.0]))
        r = ctx.planner_batch(occ, se, threshold=thr, max_iter=12)
        for k in range(8):
            ref = oracle.solve(occ, se[k][:2], se[k][2:], thr, 12)
            check(r, k, ref)


def test_large_map_planner_vs_oracle(ctx, vhp, oracle):
    """A map well beyond the BASELINE planner sizes (1536 x 1200, dense obstacles): the
    single-CTA device loop against the CPU oracle, all fields bit-exact."""
    nx, ny = 1536, 1200
    occ = oracle.generate_environment(nx, ny, 220, 12, 90, 12, 90, 17)
    free = np.argwhere(occ != 0)
    s, e = free[7], free[-9]
    se = np.array([[s[1], s[0], e[1], e[0]]], dtype=np.int32)
    r = ctx.planner_batch(occ, se, threshold=0.3, max_iter=40)
    ref = oracle.solve(occ, se[0][:2], se[0][2:], 0.3, 40)
    check(r, 0, ref)


def test_fixed_point_fast_forward_matches_reference_stall(ctx, vhp, oracle):
    """maze_5 at the shipped threshold 0.25 re-selects the same cell
    until max_iter.  The kernel detects the fixed point and skips the identical
    iterations; status, nb_of_sources, the light-source list and every field must still be
    what the reference ends with."""
    g = load_golden("maze5.npz")
    ny, nx = g["shape"]
    occ = np.unpackbits(g["occ_bits"])[: ny * nx].reshape(ny, nx)
    se = np.array([tuple(g["start"]) + tuple(g["end"])], dtype=np.int32)
    for max_iter in (150, 197):
        r = ctx.planner_batch(occ, se, threshold=0.25, max_iter=max_iter)
        ref = oracle.solve(occ, se[0][:2], se[0][2:], 0.25, max_iter)
        assert ref["status"] == 5 and ref["nb_of_sources"] == max_iter + 1
        check(r, 0, ref)


def test_f32_export(ctx, vhp, oracle):
    occ = oracle.generate_environment(101, 101, 10, 10, 20, 10, 20, 2)
    r64 = ctx.planner_batch(occ, [(5, 5, 95, 95)], threshold=0.25, max_iter=100, dtype=vhp.F64)
    r32 = ctx.planner_batch(occ, [(5, 5, 95, 95)], threshold=0.25, max_iter=100, dtype=vhp.F32)
    assert np.array_equal(r32["vg"][0], r64["vg"][0].astype(np.float32))
    assert np.array_equal(r32["vis"][0], r64["vis"][0].astype(np.float32))
    assert np.array_equal(r32["came"], r64["came"]) and r32["path_len"][0] == r64["path_len"][0]


def test_stalling_problems_of_the_bench_batch(ctx, vhp, oracle):
    """The problems of bench.py's planner batch that stall until max_iter (the reference has no
    closed set: once the heap's pick is the current source it is picked again forever) and a
    few that solve: the fast-forward must be invisible in every output -- statuses, the full
    light-source list, parents, fields and the LAST sweep's visibility -- for two values of
    max_iter."""
    import sys
    sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))
    from bench import planner_workload
    maps, se, pmap = planner_workload(0)
    r = ctx.planner_batch(maps, se, prob_map=pmap, threshold=0.5, max_iter=100, fields=False)
    stall = [int(k) for k in np.flatnonzero(r["status"] == 5)][:4]
    assert stall, "the batch is expected to contain problems that hit max_iter"
    picks = stall + [int(k) for k in np.flatnonzero(r["status"] == 0)[:2]]
    for max_iter in (100, 37):
        sel = np.array(picks)
        rr = ctx.planner_batch(maps, se[sel], prob_map=pmap[sel], threshold=0.5, max_iter=max_iter)
        for i, k in enumerate(picks):
            ref = oracle.solve(maps[pmap[k]].astype(np.float64), se[k][:2], se[k][2:], 0.5, max_iter)
            check(rr, i, ref)


def test_wide_map_takes_the_grid_route(vhp, oracle):
    """A map too wide for the persistent single-CTA kernel (its boundary rows do not fit shared
    memory) is solved on the many-CTA route by default, and refused only if that route is
    switched off."""
    nx, ny = 12000, 48
    occ = rect_map(nx, ny, 300, 12, 4, 30)
    occ[5, 5] = occ[40, 11990] = occ[20, 6000] = 1
    se = np.array([[5, 5, 11990, 40], [6000, 20, 5, 5]], np.int32)
    c = vhp.Context(0)
    r = c.planner_batch(occ, se, threshold=0.3, max_iter=12)
    for k in range(2):
        check(r, k, oracle.solve(occ, se[k][:2], se[k][2:], 0.3, 12))
    c.set_grid_sweep(0)
    with pytest.raises(vhp.VhpError):
        c.planner_batch(occ, se, threshold=0.3, max_iter=12)
    c.close()
