"""Strip-partitioned planner on one map: the window sweep, the halo-row plan and
the strip epilogue / arg-min, exercised in ONE process that owns every strip (the
multi-rank exchange itself is covered by tests/test_giant_cpu.py with gloo and by
tools/giant_multi_gpu.py on >= 2 GPUs).  Everything must equal the single-CTA planner kernel
and the CPU oracle bit for bit."""
import numpy as np
import pytest

from conftest import rect_map

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import visibility_heuristic_path_planner_b200 as vhp
    yield vhp, None


def run_case(vhp, ctx, oracle, occ, start, end, thr, max_iter, nstrips, grid_sweep=None):
    from visibility_heuristic_path_planner_b200.giant import StripPlanner
    sp = StripPlanner(occ, nstrips, device=0, grid_sweep=grid_sweep)
    r = sp.solve(start, end, thr, max_iter)
    ref = oracle.solve(occ, start, end, thr, max_iter)
    assert (r["status"], r["nb_of_sources"]) == (ref["status"], ref["nb_of_sources"])
    if ref["status"] in (0, 5):
        assert np.array_equal(r["light_sources"], ref["light_sources"])
    assert np.array_equal(r["path"], ref["path"])
    assert r["path_length"] == ref["path_length"]
    assert np.array_equal(sp.gather_field("vg"), ref["vg"])
    assert np.array_equal(sp.gather_field("vis"), ref["vis"])
    came = sp.gather_field("came").astype(np.int64)
    came[came < 0] = 1000000000000000
    assert np.array_equal(came.astype(np.uint64), ref["came"])
    sp.close()
    return r


@pytest.mark.parametrize("grid_sweep", [0, 2])  # single-CTA window kernel / many-CTA grid kernels
@pytest.mark.parametrize("nstrips", [1, 2, 3, 5])
def test_strips_equal_oracle_small(env, oracle, nstrips, grid_sweep):
    vhp, ctx = env
    occ = rect_map(200, 170, 30, 91, 3, 16)
    free = np.argwhere(occ != 0)
    g = np.random.default_rng(nstrips)
    for _ in range(3):
        a, b = free[g.integers(0, len(free))], free[g.integers(0, len(free))]
        run_case(vhp, ctx, oracle, occ, (int(a[1]), int(a[0])), (int(b[1]), int(b[0])), 0.4, 30, nstrips,
                 grid_sweep)


def test_strips_source_on_boundaries_and_stall(env, oracle):
    vhp, ctx = env
    nx, ny = 300, 256
    occ = rect_map(nx, ny, 45, 5, 4, 20)
    occ[0, :] = 1; occ[:, 0] = 1; occ[63:65, 100:140] = 1; occ[ny - 1, :] = 1
    for start, end in (((0, 0), (nx - 1, ny - 1)), ((120, 63), (10, 250)), ((120, 64), (290, 3)),
                       ((nx - 1, ny - 1), (0, 0))):
        for thr, grid in ((0.2, 0), (0.7, 2)):
            run_case(vhp, ctx, oracle, occ, start, end, thr, 14, 4, grid)
    # invalid problems keep the reference's status codes
    from visibility_heuristic_path_planner_b200.giant import StripPlanner
    sp = StripPlanner(occ, 4, device=0)
    assert sp.solve((nx, 0), (1, 1), 0.5, 10)["status"] == 1
    assert sp.solve((1, 1), (0, ny), 0.5, 10)["status"] == 2
    sp.close()


def test_strips_1000_map(env, oracle):
    """The shipped 1000 x 1000 configuration (seed 1) on 8 strips of 125 rows."""
    vhp, ctx = env
    occ = oracle.generate_environment(1000, 1000, 15, 100, 200, 100, 200, 1)
    r = run_case(vhp, ctx, oracle, occ, (50, 50), (990, 990), 0.25, 250, 8)
    assert r["status"] == 0 and abs(r["path_length"] - 1346.71) < 0.01
    # the whole map as one window, swept by many CTAs (grid mode is the default at this size)
    r = run_case(vhp, ctx, oracle, occ, (50, 50), (990, 990), 0.25, 250, 1)
    assert r["status"] == 0 and abs(r["path_length"] - 1346.71) < 0.01
    r = run_case(vhp, ctx, oracle, occ, (500, 500), (990, 10), 0.25, 250, 2, grid_sweep=2)


def test_grid_sweep_equals_single_cta_4096(env):
    """One 4096 x 4096 map, dense random obstacles: every strip field of the many-CTA grid
    sweep equals the single-CTA window kernel bit for bit (too large for the CPU oracle)."""
    from visibility_heuristic_path_planner_b200.giant import StripPlanner
    n = 4096
    occ = rect_map(n, n, 1500, 17, 8, 64)
    free = np.argwhere(occ != 0)
    g = np.random.default_rng(5)
    a, b = free[g.integers(0, len(free))], free[g.integers(0, len(free))]
    start, end = (int(a[1]), int(a[0])), (int(b[1]), int(b[0]))
    res = []
    for grid in (0, 2):
        sp = StripPlanner(occ, 2, device=0, grid_sweep=grid)
        r = sp.solve(start, end, 0.3, 12)
        res.append((r, sp.gather_field("vis"), sp.gather_field("vg"), sp.gather_field("came")))
        sp.close()
    (r0, vis0, vg0, came0), (r1, vis1, vg1, came1) = res
    assert r0["status"] == r1["status"] and r0["nb_of_sources"] == r1["nb_of_sources"] > 1
    assert np.array_equal(r0["light_sources"], r1["light_sources"]) and r0["path_length"] == r1["path_length"]
    assert np.array_equal(vis0, vis1) and np.array_equal(vg0, vg1) and np.array_equal(came0, came1)


def test_halo_plan_consistency(env):
    """Both sides of a strip boundary derive the same halo rows, and they always lie inside
    the neighbouring strip."""
    vhp, ctx = env
    from visibility_heuristic_path_planner_b200.giant import halo_rows, strip_layout, sweep_schedule
    lib = vhp.load_library()
    nx, ny = 777, 640
    strips = strip_layout(ny, 6)
    g = np.random.default_rng(0)
    for _ in range(200):
        sx, sy = int(g.integers(0, nx)), int(g.integers(0, ny))
        for k, nb in sweep_schedule(strips, sy):
            rows = halo_rows(lib, nx, ny, sx, sy, *strips[k])
            if nb is None:
                assert rows == [-1, -1, -1, -1]
            else:
                # (no row at all when the strip still starts inside the first tile row)
                lo, hi = strips[nb]
                assert all(lo <= r < hi for r in rows if r >= 0), (sx, sy, k, rows)
                upper = k > nb
                assert all(rows[q] < 0 for q in ((2, 3) if upper else (0, 1))), (sx, sy, k, rows)
