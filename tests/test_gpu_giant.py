"""Strip-partitioned planner on one map through the C-ABI `vhp_giant_*`: ONE process
that owns every strip (world 1, several strips per rank), so the window sweeps with the +y / -y
quadrant split, the halo-row hand-over, the strip epilogue / arg-min, the device-side loop control
(CUDA-graph WHILE node, batches, read-back per iteration) and reconstructPath are all exercised
without a second GPU.  The NCCL exchange between ranks is covered by
tests/test_gpu_giant_multirank.py (>= 2 GPUs) and by bench.py --gpus N.  Everything must equal the
CPU oracle bit for bit."""
import numpy as np
import pytest

from conftest import rect_map

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import visibility_heuristic_path_planner_b200 as vhp
    ctx = vhp.Context(0)
    yield vhp, ctx
    ctx.close()


def run_case(vhp, ctx, oracle, occ, start, end, thr, max_iter, nstrips, grid_sweep=None, loop=0, batch=0,
             planner=None):
    from visibility_heuristic_path_planner_b200.giant import GiantPlanner
    ctx.set_grid_sweep(1 if grid_sweep is None else grid_sweep)
    gp = planner or GiantPlanner(occ, ctx=ctx, strips_per_rank=nstrips)
    gp.set_loop_mode(loop, batch)
    r = gp.solve(start, end, thr, max_iter)
    ref = oracle.solve(occ, start, end, thr, max_iter)
    assert (r["status"], r["nb_of_sources"]) == (ref["status"], ref["nb_of_sources"])
    if ref["status"] in (0, 5):
        assert np.array_equal(r["light_sources"], ref["light_sources"])
    assert np.array_equal(r["path"], ref["path"])
    assert r["path_length"] == ref["path_length"]
    assert r["rows"] == (0, occ.shape[0])
    assert np.array_equal(r["vg"], ref["vg"])
    assert np.array_equal(r["vis"], ref["vis"])
    came = r["came"].astype(np.int64)
    came[came < 0] = 1000000000000000
    assert np.array_equal(came.astype(np.uint64), ref["came"])
    if planner is None:
        gp.close()
    ctx.set_grid_sweep(1)
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


@pytest.mark.parametrize("loop,batch", [(1, 0), (2, 1), (2, 3), (2, 8), (3, 0)])
def test_loop_modes_agree(env, oracle, loop, batch):
    """CUDA-graph WHILE node, batches of iterations one snapshot behind, read-back per iteration:
    same results; the handle (and its captured graph) is reused for several queries."""
    vhp, ctx = env
    from visibility_heuristic_path_planner_b200.giant import GiantPlanner
    occ = rect_map(260, 230, 40, 17, 3, 18)
    free = np.argwhere(occ != 0)
    g = np.random.default_rng(7)
    gp = GiantPlanner(occ, ctx=ctx, strips_per_rank=3)
    for thr, max_iter in ((0.3, 40), (0.6, 9), (0.3, 40)):
        a, b = free[g.integers(0, len(free))], free[g.integers(0, len(free))]
        r = run_case(vhp, ctx, oracle, occ, (int(a[1]), int(a[0])), (int(b[1]), int(b[0])), thr, max_iter, 3,
                     loop=loop, batch=batch, planner=gp)
        assert r["stats"]["loop_mode"] == loop
    gp.close()


def test_strips_source_on_boundaries_and_stall(env, oracle):
    vhp, ctx = env
    nx, ny = 300, 256
    occ = rect_map(nx, ny, 45, 5, 4, 20)
    occ[0, :] = 1; occ[:, 0] = 1; occ[63:65, 100:140] = 1; occ[ny - 1, :] = 1
    for start, end in (((0, 0), (nx - 1, ny - 1)), ((120, 63), (10, 250)), ((120, 64), (290, 3)),
                       ((nx - 1, ny - 1), (0, 0))):
        for thr, grid in ((0.2, 0), (0.7, 2)):
            run_case(vhp, ctx, oracle, occ, start, end, thr, 14, 4, grid)
    # a source whose first tile row is one cell high (sx % 32 == 31 in the +x quadrants, sx % 32 == 0
    # in the -x ones): the -y halo row is the source's own row, which the +y quadrants store
    for sx in (31, 63, 64, 96):
        occ[100, sx] = 1
        run_case(vhp, ctx, oracle, occ, (sx, 100), (5, 5), 0.3, 10, 7, 2)
        run_case(vhp, ctx, oracle, occ, (sx, 100), (5, 5), 0.3, 10, 7, 0)
    # invalid problems keep the reference's status codes
    from visibility_heuristic_path_planner_b200.giant import GiantPlanner
    gp = GiantPlanner(occ, ctx=ctx, strips_per_rank=4)
    assert gp.solve((nx, 0), (1, 1), 0.5, 10)["status"] == 1
    assert gp.solve((1, 1), (0, ny), 0.5, 10)["status"] == 2
    r = gp.solve((1, 1), (0, ny), 0.5, 10)
    assert not r["vg"].any() and not r["vis"].any() and (r["came"] == -1).all()
    gp.close()


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
    sweep equals the single-CTA window kernel bit for bit, and both equal the batch planner
    (too large for the CPU oracle)."""
    vhp, ctx = env
    from visibility_heuristic_path_planner_b200.giant import GiantPlanner
    n = 4096
    occ = rect_map(n, n, 1500, 17, 8, 64)
    free = np.argwhere(occ != 0)
    g = np.random.default_rng(5)
    a, b = free[g.integers(0, len(free))], free[g.integers(0, len(free))]
    start, end = (int(a[1]), int(a[0])), (int(b[1]), int(b[0]))
    res = []
    for grid in (0, 2):
        ctx.set_grid_sweep(grid)
        gp = GiantPlanner(occ, ctx=ctx, strips_per_rank=2)
        res.append(gp.solve(start, end, 0.3, 12))
        gp.close()
    ctx.set_grid_sweep(1)
    r0, r1 = res
    assert r0["status"] == r1["status"] and r0["nb_of_sources"] == r1["nb_of_sources"] > 1
    assert np.array_equal(r0["light_sources"], r1["light_sources"]) and r0["path_length"] == r1["path_length"]
    for k in ("vis", "vg", "came"):
        assert np.array_equal(r0[k], r1[k])
    ref = ctx.planner_batch(occ, [start + end], threshold=0.3, max_iter=12)
    assert int(ref["status"][0]) == r0["status"] and int(ref["nb_sources"][0]) == r0["nb_of_sources"]
    assert float(ref["path_len"][0]) == r0["path_length"]
    assert np.array_equal(ref["vg"][0], r0["vg"]) and np.array_equal(ref["came"][0], r0["came"])


def test_halo_plan_consistency(env):
    """Both sides of a strip boundary derive the same halo rows, and they always lie inside
    the neighbouring strip (vhp_strip_halo_rows is the host view of the device-side plan)."""
    vhp, ctx = env
    from visibility_heuristic_path_planner_b200.giant import halo_rows, strip_layout
    lib = vhp.load_library()
    nx, ny = 777, 640
    strips = strip_layout(ny, 6)
    g = np.random.default_rng(0)
    for _ in range(200):
        sx, sy = int(g.integers(0, nx)), int(g.integers(0, ny))
        for k, (lo, hi) in enumerate(strips):
            rows = halo_rows(lib, nx, ny, sx, sy, lo, hi)
            if lo <= sy < hi:
                assert rows == [-1, -1, -1, -1]
                continue
            upper = lo > sy
            nlo, nhi = strips[k - 1] if upper else strips[k + 1]
            assert all(nlo <= r < nhi for r in rows if r >= 0), (sx, sy, k, rows)
            assert all(rows[q] < 0 for q in ((2, 3) if upper else (0, 1))), (sx, sy, k, rows)
