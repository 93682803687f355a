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
