"""Drop-in surface on the GPU: the solver handle / CLI write the same ./output/*.txt
as the reference executable (oracle/_ref/ref_cli_strict, prebuilt where the
reference checkout exists) and print the same result lines."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden, rect_map

pytestmark = pytest.mark.gpu
CLI = os.path.join(ROOT, "visibility_heuristic_path_planner_b200", "bin", "visibility_heuristic_planner")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "ref_cli_strict")

CONFIG = """mode={mode}
ncols={n}
nrows={n}
nb_of_obstacles=10
minWidth=10
maxWidth=20
minHeight=10
maxHeight=20
randomSeed=0
seedValue={seed}
imagePath={image}
start={{{sx},{sy}}}
end={{{ex},{ey}}}
max_iter=250
visibilityThreshold={thr}
lightStrength=1
timer=1
saveResults=1
saveCameFrom=1
saveLightSources=1
saveGlobalVisibility=1
saveLocalVisibility=1
saveVisibilityField=1
silent=0
ballRadius=5
"""


def run(exe, workdir, cfg):
    os.makedirs(os.path.join(workdir, "config"), exist_ok=True)
    with open(os.path.join(workdir, "config", "settings.config"), "w") as f:
        f.write(cfg)
    p = subprocess.run([exe], cwd=workdir, capture_output=True, text=True, timeout=600)
    return p.returncode, p.stdout


FILES = ["cameFrom.txt", "lightSources.txt", "VisibilityMap.txt", "LocalVisibilityMap.txt", "visibilityField.txt"]


@pytest.mark.parametrize("seed,thr", [(2, 0.25), (3, 0.25), (5, 0.5)])
def test_cli_outputs_match_reference_executable(tmp_path, seed, thr):
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/ref_cli_strict not prebuilt")
    cfg = CONFIG.format(mode=1, n=101, seed=seed, image="none", sx=5, sy=5, ex=95, ey=95, thr=thr)
    a, b = str(tmp_path / "ours"), str(tmp_path / "ref")
    rc_a, out_a = run(CLI, a, cfg)
    rc_b, out_b = run(REF_CLI, b, cfg)
    assert rc_a == 0 and rc_b == 0, (out_a[-500:], out_b[-500:])
    for f in FILES:
        assert open(os.path.join(a, "output", f), "rb").read() == open(os.path.join(b, "output", f), "rb").read(), f
    for pat in (r"Path length: (\S+)", r"Density of the occupancy grid: (\S+)"):
        assert re.search(pat, out_a).group(1) == re.search(pat, out_b).group(1)
    assert "Visibility computation time in us:" in out_a and "Raycasting computation time in us:" in out_a
    # the renderings: ours are PNGs, the reference build writes binary PPMs under the same names
    # (SFML stub); every pixel must agree
    from PIL import Image
    for f in ("ResultingPath.png", "standAloneVisibility.png", "rayCastingVisibility.png"):
        ours = np.asarray(Image.open(os.path.join(a, "output", f)).convert("RGB"))
        ref = np.asarray(Image.open(os.path.join(b, "output", f)).convert("RGB"))
        assert ours.shape == ref.shape == (101, 101, 3), f
        assert np.array_equal(ours, ref), (f, np.argwhere((ours != ref).any(axis=2))[:5].tolist())


def test_cli_error_messages(tmp_path):
    cfg = CONFIG.format(mode=1, n=101, seed=1, image="none", sx=5, sy=5, ex=95, ey=95, thr=0.25)
    rc, out = run(CLI, str(tmp_path / "occ"), cfg)
    assert rc == 0 and "End point is not valid (occupied)" in out   # seed 1 (golden)
    cfg = CONFIG.format(mode=1, n=101, seed=2, image="none", sx=500, sy=5, ex=95, ey=95, thr=0.25)
    rc, out = run(CLI, str(tmp_path / "oob"), cfg)
    assert "Start point is out of bounds." in out


def test_cli_image_mode_maze5(tmp_path):
    """BASELINE config 3 through the CLI: mode 2, bottom-left start/end, thr 0.2."""
    from PIL import Image
    g = load_golden("maze5.npz")
    ny, nx = map(int, g["shape"])
    occ = np.unpackbits(g["occ_bits"])[: ny * nx].reshape(ny, nx)
    img = np.zeros((ny, nx, 4), dtype=np.uint8)
    img[..., 0] = np.where(occ, 255, 0); img[..., 3] = 255
    png = str(tmp_path / "maze_5.png")
    Image.fromarray(img, "RGBA").save(png)
    cfg = CONFIG.format(mode=2, n=100, seed=1, image=png, sx=118, sy=317, ex=123, ey=10, thr=0.2)
    work = str(tmp_path / "maze")
    rc, out = run(CLI, work, cfg)
    assert rc == 0, out[-800:]
    assert "Loaded image of dimensions 242x322 successfully" in out
    assert "Path length: 1341.71" in out
    ls = np.loadtxt(os.path.join(work, "output", "lightSources.txt"), dtype=int)
    assert len(ls) == 112 and tuple(ls[0]) == (118, 317)          # written back in the config frame
    ref_ls = g["thr020_ls"][:112].copy(); ref_ls[:, 1] = ny - 1 - ref_ls[:, 1]
    assert np.array_equal(ls, ref_ls)
    came = np.loadtxt(os.path.join(work, "output", "cameFrom.txt"), dtype=np.uint64)
    assert came.shape == (ny, nx) and came.max() == 10**15


def test_solver_handle_fields(tmp_path):
    """vhp_solver_* through ctypes: fields, light sources and path of a solve."""
    import ctypes as C
    import visibility_heuristic_path_planner_b200 as vhp
    from oracle_py import Oracle
    lib = vhp.load_library()
    ora = Oracle()
    occ = ora.generate_environment(101, 101, 10, 10, 20, 10, 20, 4)
    ref = ora.solve(occ, (5, 5), (95, 95), 0.25, 100)
    cfg = vhp.Config()
    lib.vhp_config_default(C.byref(cfg))
    cfg.ncols = cfg.nrows = 101
    cfg.start_x, cfg.start_y, cfg.end_x, cfg.end_y = 5, 5, 95, 95
    cfg.visibility_threshold, cfg.max_iter, cfg.silent, cfg.save_results = 0.25, 100, 1, 0
    for k in ("save_came_from", "save_light_sources", "save_global_visibility", "save_local_visibility",
              "save_visibility_field"):
        setattr(cfg, k, 0)
    ctx = vhp.Context(0)
    h = C.c_void_p()
    occ8 = np.ascontiguousarray(occ != 0, dtype=np.uint8)
    lib.vhp_solver_create.argtypes = [C.c_void_p, C.POINTER(vhp.Config), C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    assert lib.vhp_solver_create(ctx.h, C.byref(cfg), occ8.ctypes.data, 101, 101, C.byref(h)) == 0
    cwd = os.getcwd(); os.chdir(tmp_path)
    try:
        lib.vhp_solver_solve.argtypes = [C.c_void_p]
        assert lib.vhp_solver_solve(h) == 0
    finally:
        os.chdir(cwd)
    lib.vhp_solver_get_field.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    vg = np.zeros((101, 101)); came = np.zeros((101, 101), np.int32)
    lib.vhp_solver_get_field(h, 1, vg.ctypes.data); lib.vhp_solver_get_field(h, 2, came.ctypes.data)
    assert np.array_equal(vg, ref["vg"])
    u = came.astype(np.int64); u[u < 0] = 10**15
    assert np.array_equal(u.astype(np.uint64), ref["came"])
    lib.vhp_solver_nb_of_sources.argtypes = [C.c_void_p]; lib.vhp_solver_nb_of_sources.restype = C.c_int64
    assert lib.vhp_solver_nb_of_sources(h) == ref["nb_of_sources"]
    lib.vhp_solver_path.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_double)]
    lib.vhp_solver_path.restype = C.c_int64
    xy = np.zeros((64, 2), np.int32); ln = C.c_double(0)
    n = lib.vhp_solver_path(h, xy.ctypes.data, 64, C.byref(ln))
    assert np.array_equal(xy[:n], ref["path"]) and ln.value == ref["path_length"]
    lib.vhp_solver_destroy.argtypes = [C.c_void_p]
    lib.vhp_solver_destroy(h)
    ctx.close()


def test_device_environment_batch_matches_oracle(oracle):
    """a batch of random-rectangle maps generated on the device equals the CPU
    restatement of the same counter-based generator, map by map; the maps feed a sweep without
    ever touching the host."""
    import torch
    import visibility_heuristic_path_planner_b200 as vhp
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(dev)
    ctx = vhp.torch_context(0, stream)
    lib = ctx.lib
    for (nx, ny, nb, lo_w, hi_w, lo_h, hi_h, seed, first, nmaps) in (
            (256, 256, 12, 8, 40, 8, 40, 4321, 0, 40), (101, 77, 10, 10, 20, 5, 9, 7, 1000, 9),
            (33, 200, 0, 1, 1, 1, 1, 1, 5, 3), (64, 64, 50, 0, 70, 0, 3, 99, 2**33, 4)):
        occ_t = torch.zeros((nmaps, ny, nx), dtype=torch.uint8, device=dev)
        with torch.cuda.stream(stream):
            ctx.generate_environments_dev(occ_t, nb, lo_w, hi_w, lo_h, hi_h, seed, first)
        stream.synchronize()
        occ = occ_t.cpu().numpy()
        for k in range(nmaps):
            ref = oracle.generate_environment_counter(nx, ny, nb, lo_w, hi_w, lo_h, hi_h, seed, first + k)
            assert np.array_equal(occ[k], ref.astype(np.uint8)), (nx, ny, k)
        assert set(np.unique(occ)) <= {0, 1}
    assert lib.vhp_environment_draw(4321, 3, 2, 1) == oracle.lib.vhp_oracle_env_draw(4321, 3, 2, 1)
    # device-resident maps straight into a sweep: equal to the oracle on the downloaded map
    nx = ny = 256
    occ_t = torch.empty((8, ny, nx), dtype=torch.uint8, device=dev)
    src = torch.tensor([[5, 5]] * 8, dtype=torch.int32, device=dev)
    smap = torch.arange(8, dtype=torch.int32, device=dev)
    out = torch.empty((8, ny, nx), dtype=torch.float64, device=dev)
    with torch.cuda.stream(stream):
        ctx.generate_environments_dev(occ_t, 12, 8, 40, 8, 40, 11)
        occ_t[:, 5, 5] = 1
        ctx.visibility_batch_dev(occ_t, src, out, smap)
    stream.synchronize()
    occ = occ_t.cpu().numpy()
    for k in (0, 7):
        assert np.array_equal(out[k].cpu().numpy(), oracle.compute_visibility(occ[k].astype(np.float64), 5, 5))
    ctx.close()


def test_benchmark_series_file(tmp_path, capfd):
    """benchmarkSeries() (:295-374): the first sizes of the 60 log-spaced grids 50 ... 5000, one
    un-warmed sweep and one all-targets ray casting each, appended to output/benchmark_results.txt
    as `t_vis_us t_ray_us ratio NxN` (the format of the reference's :368-373)."""
    import ctypes as C
    import visibility_heuristic_path_planner_b200 as vhp
    lib = vhp.load_library()
    cfg = vhp.Config()
    lib.vhp_config_default(C.byref(cfg))
    cfg.ncols = cfg.nrows = 64
    cfg.start_x = cfg.start_y = 5
    cfg.end_x = cfg.end_y = 60
    cfg.silent, cfg.save_results = 1, 0
    ctx = vhp.Context(0)
    h = C.c_void_p()
    occ8 = np.ones((64, 64), np.uint8)
    lib.vhp_solver_create.argtypes = [C.c_void_p, C.POINTER(vhp.Config), C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    assert lib.vhp_solver_create(ctx.h, C.byref(cfg), occ8.ctypes.data, 64, 64, C.byref(h)) == 0
    lib.vhp_solver_benchmark_series.argtypes = [C.c_void_p, C.c_int]
    cwd = os.getcwd(); os.chdir(tmp_path)
    try:
        assert lib.vhp_solver_benchmark_series(h, 14) == 0
        assert lib.vhp_solver_benchmark_series(h, 3) == 0      # appends
    finally:
        os.chdir(cwd)
    lib.vhp_solver_destroy.argtypes = [C.c_void_p]
    lib.vhp_solver_destroy(h)
    # a sweep after the series still sees its own map (the series' temporary maps are released)
    occ = rect_map(70, 60, 10, 3)
    from oracle_py import Oracle
    assert np.array_equal(ctx.visibility_batch(occ, [(9, 9)])[0], Oracle().compute_visibility(occ, 9, 9))
    ctx.close()
    lines = open(tmp_path / "output" / "benchmark_results.txt").read().splitlines()
    assert len(lines) == 17
    sizes = [int(round(50 * np.exp(np.log(100) * i / 59))) for i in range(60)]   # :311-315
    for k, line in enumerate(lines):
        tv, tr, ratio, dims = line.split()
        n = sizes[k if k < 14 else k - 14]
        assert dims == f"{n}x{n}" and float(tv) > 0 and float(tr) > 0
        assert abs(float(ratio) - float(tr) / float(tv)) <= 1e-3 * float(ratio)
    out = capfd.readouterr().out
    assert "For grid size: 50x50" in out and "Ratios:" in out
