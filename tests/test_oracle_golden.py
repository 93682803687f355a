"""The CPU oracle (oracle/vhp_oracle.c) against the golden vectors generated from
the unmodified reference (oracle/gen_golden.py) and, where oracle/_ref is
present, against the reference itself.  Everything here is bit-exact."""
import hashlib

import numpy as np
import pytest

from conftest import load_golden, rect_map


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_sweep_goldens(oracle):
    g = load_golden("sweep.npz")
    for k, (nx, ny, nobs, seed, sx, sy) in enumerate(g["cases"]):
        occ = g[f"occ_{k}"].astype(np.float64)
        assert occ.shape == (ny, nx)
        assert np.array_equal(occ, rect_map(nx, ny, nobs, seed))
        assert np.array_equal(oracle.compute_visibility(occ, sx, sy), g[f"vis_{k}"]), k
        assert np.array_equal(oracle.raycast_all(occ, sx, sy), g[f"ray_{k}"]), k


def test_quirk_kats(oracle):
    g = load_golden("sweep.npz")
    occ = g["kat_diag_occ"].astype(np.float64)
    v = oracle.compute_visibility(occ, 2, 2)
    assert np.array_equal(v, g["kat_diag_vis"])
    # quirk 1: the diagonal copies its y-predecessor, so the obstacle at (4,3)
    # darkens (4,4) while (3,3) stays lit
    assert v[4, 4] == 0.0 and v[3, 3] == 1.0
    # quirk 2: column 0 / row 0 are never written when the source is off them
    assert not v[0, :].any() and not v[:, 0].any()
    # occupied source -> all dark
    occ = np.ones((12, 10)); occ[5, 4] = 0
    assert np.array_equal(oracle.compute_visibility(occ, 4, 5), g["kat_occsrc_vis"])
    assert not g["kat_occsrc_vis"].any()
    # computeVisibility() does not reset visibility_: stale borders survive
    out = oracle.compute_visibility(np.ones((7, 9)), 4, 3, g["kat_stale_init"])
    assert np.array_equal(out, g["kat_stale_vis"])
    assert np.array_equal(out[0, :], g["kat_stale_init"][0, :])


def test_strict_ieee_flip_cell(oracle):
    """The strict build and the -Ofast build of the reference disagree on a
    threshold-0.5 decision; the oracle follows the strict one."""
    g = load_golden("sweep.npz")
    seed, x, y = map(int, g["kat_flip"])
    occ = oracle.generate_environment(101, 101, 10, 10, 20, 10, 20, seed)
    v = oracle.compute_visibility(occ, 50, 50)
    assert np.array_equal(v, g["kat_flip_vis"])
    strict, fast = g["kat_flip_vals"]
    assert v[y, x] == strict and (strict >= 0.5) != (fast >= 0.5)


def test_tie_break_kat(oracle):
    g = load_golden("planner.npz")
    occ = g["tie_occ"].astype(np.float64)
    vg = np.zeros((101, 101)); came = np.full((101, 101), 10**15, dtype=np.uint64)
    came[10, 50] = 0
    ls = np.array([[50, 10]], dtype=np.int32)
    _, top, h, pushes = oracle.update_visibility(occ, (50, 10), (50, 90), 0.5, vg, came, ls, 0)
    assert top == tuple(g["tie_top"]) == (82, 54)
    assert h == g["tie_h"][0] and pushes == g["tie_pushes"][0]
    assert np.array_equal(vg, g["tie_vg"]) and np.array_equal(came, g["tie_came"])


def _check_solve(r, g, tag, full=True):
    st, nb = g[f"{tag}_status"]
    assert (r["status"], r["nb_of_sources"]) == (st, nb)
    if st in (0, 5):
        assert np.array_equal(r["light_sources"], g[f"{tag}_ls"])
    assert np.array_equal(r["path"], g[f"{tag}_path"])
    assert r["path_length"] == g[f"{tag}_len"][0]
    if full:
        assert np.array_equal(r["vg"], g[f"{tag}_vg"])
        assert np.array_equal(r["came"], g[f"{tag}_came"])


def test_solve_101(oracle):
    g = load_golden("planner.npz")
    for seed in (1, 2, 3, 4, 5, 6):
        occ = oracle.generate_environment(101, 101, 10, 10, 20, 10, 20, seed)
        r = oracle.solve(occ, (5, 5), (95, 95), 0.25, 100)
        _check_solve(r, g, f"s101_{seed}")
        assert np.array_equal(r["vis"], g[f"s101_{seed}_vis"])
    assert g["s101_1_status"][0] == 4          # "End point is not valid (occupied)"
    assert abs(g["s101_2_len"][0] - 139.535) < 5e-4   # SURVEY 8c printed values
    assert abs(g["s101_3_len"][0] - 138.743) < 5e-4


def test_solve_extra_shapes(oracle):
    g = load_golden("planner.npz")
    for k, c in enumerate(g["extra_cases"]):
        occ = g[f"extra_{k}_occ"].astype(np.float64)
        r = oracle.solve(occ, (c[4], c[5]), (c[6], c[7]), g["extra_thr"][k], int(c[8]))
        _check_solve(r, g, f"extra_{k}")
    assert g["extra_4_status"][0] == 5 and g["extra_4_status"][1] == 31  # stall -> max_iter+1


def test_shipped_1000(oracle):
    g = load_golden("shipped1000.npz")
    for seed in (1, 2, 25):
        occ = oracle.generate_environment(1000, 1000, 15, 100, 200, 100, 200, seed)
        r = oracle.solve(occ, (50, 50), (990, 990), 0.25, 250)
        _check_solve(r, g, f"seed{seed}", full=False)
        s = g[f"seed{seed}_sha"]
        assert [sha(occ.astype(np.uint8)), sha(r["vg"]), sha(r["came"]), sha(r["vis"])] == list(s)
    assert abs(g["seed1_density"][0] - 20.3243) < 1e-4
    v = oracle.compute_visibility(np.ones((1000, 1000)), 500, 500)
    assert sha(v) == g["empty_cv_sha"][0]


def test_maze5(oracle):
    g = load_golden("maze5.npz")
    ny, nx = g["shape"]
    occ = np.unpackbits(g["occ_bits"])[: ny * nx].reshape(ny, nx).astype(np.float64)
    for tag, thr in (("thr020", 0.2), ("thr025", 0.25)):
        r = oracle.solve(occ, tuple(g["start"]), tuple(g["end"]), thr, 250)
        _check_solve(r, g, tag, full=False)
        assert [sha(r["vg"]), sha(r["came"]), sha(r["vis"])] == list(g[f"{tag}_sha"])
    assert g["thr020_status"][1] == 112 and len(g["thr020_path"]) == 42
    assert g["thr020_len"][0] == 1341.7118586874171
    assert g["thr025_status"][0] == 5   # the shipped threshold stalls: "Max iters hit"


# ---- the oracle against the compiled reference itself (dev container only) ----
@pytest.mark.parametrize("seed", range(12))
def test_oracle_vs_reference_random(oracle, ref_strict, seed):
    g = np.random.default_rng(1000 + seed)
    nx, ny = int(g.integers(2, 90)), int(g.integers(2, 90))
    occ = rect_map(nx, ny, int(g.integers(0, 25)), 77 + seed, 2, 12)
    sx, sy = int(g.integers(0, nx)), int(g.integers(0, ny))
    init = g.random((ny, nx))
    assert np.array_equal(oracle.compute_visibility(occ, sx, sy, init),
                          ref_strict.compute_visibility(occ, sx, sy, init))
    assert np.array_equal(oracle.raycast_all(occ, sx, sy), ref_strict.raycast_all(occ, sx, sy))
    ex, ey = int(g.integers(0, nx)), int(g.integers(0, ny))
    thr = float(g.choice([0.0, 0.1, 0.25, 0.5, 0.9, 1.0]))
    a = oracle.solve(occ, (sx, sy), (ex, ey), thr, 20)
    b = ref_strict.solve(occ, (sx, sy), (ex, ey), thr, 20)
    assert a["status"] == b["status"] and a["nb_of_sources"] == b["nb_of_sources"]
    for key in ("vis", "vg", "came", "path"):
        assert np.array_equal(a[key], b[key]), key
    if a["status"] in (0, 5):
        assert np.array_equal(a["light_sources"], b["light_sources"])
    assert a["path_length"] == b["path_length"]


def test_environment_generator_vs_reference(oracle, ref_strict):
    for args in [(101, 101, 10, 10, 20, 10, 20, 2), (64, 200, 30, 1, 5, 2, 40, 9),
                 (1000, 1000, 15, 100, 200, 100, 200, 25)]:
        assert np.array_equal(oracle.generate_environment(*args),
                              ref_strict.generate_environment(*args))


def test_counter_environment_generator(oracle):
    """The batch generator's definition (include/vhp.h: SplitMix64 finaliser over a counter,
    rectangle rule of src/environment.cpp:57-79), restated here with Python integers, equals the
    C restatement the GPU test checks the device kernel against."""
    M = (1 << 64) - 1

    def draw(seed, m, o, d):
        z = (seed + 0x9E3779B97F4A7C15 * ((m * 0x100000001B3 + o * 4 + d + 1) & M)) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return ((z ^ (z >> 31)) >> 33)

    for seed, m, o, d in ((0, 0, 0, 0), (4321, 3, 2, 1), (2**63 + 5, 2**40, 77, 3), (7, 16383, 5999, 2)):
        assert oracle.lib.vhp_oracle_env_draw(seed, m, o, d) == draw(seed, m, o, d) < 2**31
    for (nx, ny, nb, lw, hw, lh, hh, seed, m) in ((64, 48, 12, 3, 9, 2, 11, 5, 9), (101, 101, 10, 10, 20, 10, 20, 2, 0)):
        occ = np.ones((ny, nx))
        for o in range(nb):
            c1 = 1 + draw(seed, m, o, 0) % (nx + 1)
            c2 = c1 + lw + draw(seed, m, o, 1) % (hw - lw + 1)
            r1 = 1 + draw(seed, m, o, 2) % (ny + 1)
            r2 = r1 + lh + draw(seed, m, o, 3) % (hh - lh + 1)
            c1, c2, r1, r2 = min(c1, nx - 1), min(c2, nx - 1), min(r1, ny - 1), min(r2, ny - 1)
            occ[r1:r2, c1:c2] = 0
        assert np.array_equal(occ, oracle.generate_environment_counter(nx, ny, nb, lw, hw, lh, hh, seed, m))


# ---- sweep variants -----------------------------------------------------
def _queue_family(n):
    rng = np.random.default_rng(0)
    for t in range(n):
        nx, ny = int(rng.integers(8, 70)), int(rng.integers(8, 70))
        nobs = int(rng.integers(0, 25))
        sx, sy = int(rng.integers(0, nx)), int(rng.integers(0, ny))
        yield t, nx, ny, nobs, sx, sy


def test_queue_rule_against_reference_goldens(oracle):
    """vhp_oracle_visibility_cutoff (the order-free statement of computeVisibilityUsingQueue,
    :701-893) reproduces the compiled reference's FIFO search bit for bit on every case the
    fixture marks "agree" (304 of 400 random maps); the others are the order-dependent ones."""
    g = load_golden("queue.npz")
    agree = g["agree"]
    assert int(agree.sum()) == 304 and len(agree) == 400
    checked = 0
    for t, nx, ny, nobs, sx, sy in _queue_family(40):
        row = g["cases"][t]
        assert tuple(row[:6]) == (t, nx, ny, nobs, sx, sy)
        mine = oracle.visibility_cutoff(rect_map(nx, ny, nobs, t, 1, 9), sx, sy)
        ndiff = int((mine != g[f"ref_{t}"]).sum())
        assert ndiff == int(row[6]) and (ndiff == 0) == bool(agree[t]), t
        checked += bool(agree[t])
    assert checked >= 25


def test_queue_rule_against_live_reference(oracle, ref_strict):
    g = load_golden("queue.npz")
    for t, nx, ny, nobs, sx, sy in _queue_family(120):
        occ = rect_map(nx, ny, nobs, t, 1, 9)
        same = np.array_equal(oracle.visibility_cutoff(occ, sx, sy), ref_strict.compute_visibility_queue(occ, sx, sy))
        assert same == bool(g["agree"][t]), t


def test_queue_rule_properties(oracle):
    # empty map: everything lit, border included (no never-written column / row in this variant)
    assert (oracle.visibility_cutoff(np.ones((40, 57)), 20, 9) == 1.0).all()
    # an occupied source still shines (:707 stores lightStrength_ unconditionally)
    occ = np.ones((20, 20)); occ[5, 5] = 0
    v = oracle.visibility_cutoff(occ, 5, 5)
    assert v[5, 5] == 1.0 and (v == 1.0).all()
    # cutoff 0 never terminates early: free cells follow the DP with the MATLAB diagonal
    occ = rect_map(50, 44, 9, 3, 2, 7)
    a = oracle.visibility_cutoff(occ, 10, 12, cutoff=0.0)
    b = oracle.accessibility_map(occ, 10, 12)
    lit = a > 0
    assert np.array_equal(a[lit], b[lit])
    # the early termination only removes light: cutoff 0.001 <= cutoff 0 everywhere
    assert (oracle.visibility_cutoff(occ, 10, 12) <= a).all()


def test_accessibility_map_properties(oracle):
    """getAccessibilityMap.m restated (UNPINNED: no MATLAB here) -- properties that follow from
    the .m file alone."""
    # empty map, fac = 1: v = alpha^max(|dx|, |dy|) as the running product alpha * alpha * ...
    nx, ny, sx, sy, alpha = 37, 29, 11, 20, 0.93
    v = oracle.accessibility_map(np.ones((ny, nx)), sx, sy, alpha=alpha)
    pw = np.multiply.accumulate(np.concatenate([[1.0], np.full(40, alpha)]))
    yy, xx = np.mgrid[0:ny, 0:nx]
    assert np.array_equal(v, pw[np.maximum(np.abs(xx - sx), np.abs(yy - sy))])
    # alpha = fac = 1 with obstacles: equal to the C++ sweep wherever no diagonal cell and no
    # never-written border cell lies upstream -- in particular on the two axes through the source
    occ = rect_map(60, 50, 12, 5, 3, 8)
    occ[25, :] = 1; occ[:, 30] = 1
    m, c = oracle.accessibility_map(occ, 30, 25), oracle.compute_visibility(occ, 30, 25)
    assert np.array_equal(m[25, 1:], c[25, 1:]) and np.array_equal(m[1:, 30], c[1:, 30])
    assert m[25, 0] == 1.0 and c[25, 0] == 0.0  # MATLAB's loops reach the border, the C++ port's do not
    # light strength scales linearly (power of two: exact), occupied cells are 0 for any fac
    for fac in (0.5, 1.0, 1.7, 3.0):
        a1 = oracle.accessibility_map(occ, 7, 9, alpha=0.99, fac=fac)
        a2 = oracle.accessibility_map(occ, 7, 9, alpha=0.99, fac=fac, light_strength=0.5)
        assert np.array_equal(a2, 0.5 * a1) and not a1[occ == 0].any() and a1.max() <= 1.0
